// NVLS pools (KK_FANOUT_NVLS): per-device VMM allocations bound to one NVSwitch multicast object.  See kk_nvls.cpp.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "kk_common.hpp"

namespace kk {

class NvlsPools {
 public:
  NvlsPools();
  ~NvlsPools();
  NvlsPools(const NvlsPools&) = delete;
  NvlsPools& operator=(const NvlsPools&) = delete;

  // Every listed device reports CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED and the driver exports the multicast entry points.
  static bool supported(const std::vector<int>& ordinals, std::string* why);
  // One allocation of >= bytes on every device, all bound at offset 0 of one multicast object, all mapped.  Throws kk::Error
  // (KK_EUNSUPPORTED when the host does not expose NVLS, KK_ENOMEM, KK_ECUDA); a partially built object cleans up in its destructor.
  void create(const std::vector<int>& ordinals, uint64_t bytes);

  uint8_t* pool(size_t i) const;   // unicast address of device i's allocation (readable / writable from every device)
  uint8_t* multicast() const;      // multicast address: a multimem.st here lands at the same offset of every pool
  uint64_t bytes() const;          // size of each allocation after rounding to the multicast granularity

 private:
  struct Impl;
  Impl* p_;
};

}  // namespace kk
