// VMM-backed pools (KK_CFG_VMM_POOLS): device memory allocated with cuMemCreate so that it can be exported as a POSIX file descriptor and
// mapped by another process READ-ONLY.  See kk_vmm.cpp.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "kk_common.hpp"

namespace kk {

// One physical allocation on one device, mapped read-write into this process for `access` devices.
class VmmAlloc {
 public:
  VmmAlloc() = default;
  ~VmmAlloc();
  VmmAlloc(const VmmAlloc&) = delete;
  VmmAlloc& operator=(const VmmAlloc&) = delete;

  // >= bytes on CUDA device `ordinal`, exportable as a POSIX fd; read-write for every device in `access` (the owner is always included).
  // Throws kk::Error (KK_EUNSUPPORTED when the driver lacks VMM / fd handles, KK_ENOMEM, KK_ECUDA).
  void create(int ordinal, uint64_t bytes, const std::vector<int>& access);
  uint8_t* ptr() const { return reinterpret_cast<uint8_t*>(va_); }
  uint64_t bytes() const { return size_; }
  // A new file descriptor referring to the allocation (the caller owns and closes it).  Whoever holds it can map the memory with the
  // protection it chooses — hand it only to processes that may at least read the weights; kk_import_fd maps it read-only.
  int export_fd() const;

 private:
  unsigned long long handle_ = 0;
  unsigned long long va_ = 0;
  uint64_t size_ = 0;
  bool have_handle_ = false, mapped_ = false;
};

// Consumer side: map an exported allocation into this process on `ordinal`.  readonly = CU_MEM_ACCESS_FLAGS_PROT_READ.
struct VmmImport {
  unsigned long long handle = 0, va = 0;
  uint64_t size = 0;
};
VmmImport vmm_import_fd(int fd, int ordinal, uint64_t bytes, bool readonly);
void vmm_import_close(VmmImport& im);

}  // namespace kk
