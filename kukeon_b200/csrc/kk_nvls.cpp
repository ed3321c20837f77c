// NVLS (NVSwitch multicast) pools for KK_FANOUT_NVLS: one physical allocation per device bound to one multicast object, so that a
// single multimem.st from the convert kernel lands in every device's pool.
//
// Everything here goes through the driver's virtual-memory-management and multicast entry points, fetched at run time with
// cudaGetDriverEntryPoint — the library keeps linking only the static runtime and still loads on a machine without libcuda (the CPU
// test tier checks exactly that).  One process owning all devices only; pools allocated this way cannot be exported with
// cudaIpcGetMemHandle (they are VMM allocations), which is why this path is the comparison the north_star names, not the default
// (DESIGN.md §3.1).
#include "kk_nvls.hpp"

#include <cuda.h>

#include <cstring>

namespace kk {

namespace {

struct Drv {
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
};

template <class F>
void load(F& fn, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    fail(KK_EUNSUPPORTED, "fan-out NVLS: the CUDA driver does not export %s", name);
  }
  fn = reinterpret_cast<F>(p);
}

const Drv& drv() {
  static const Drv d = [] {
    Drv x;
    load(x.MulticastCreate, "cuMulticastCreate");
    load(x.MulticastAddDevice, "cuMulticastAddDevice");
    load(x.MulticastBindMem, "cuMulticastBindMem");
    load(x.MulticastUnbind, "cuMulticastUnbind");
    load(x.MulticastGetGranularity, "cuMulticastGetGranularity");
    load(x.MemCreate, "cuMemCreate");
    load(x.MemRelease, "cuMemRelease");
    load(x.MemAddressReserve, "cuMemAddressReserve");
    load(x.MemAddressFree, "cuMemAddressFree");
    load(x.MemMap, "cuMemMap");
    load(x.MemUnmap, "cuMemUnmap");
    load(x.MemSetAccess, "cuMemSetAccess");
    load(x.MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
    load(x.DeviceGet, "cuDeviceGet");
    load(x.DeviceGetAttribute, "cuDeviceGetAttribute");
    load(x.GetErrorString, "cuGetErrorString");
    return x;
  }();
  return d;
}

void check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  drv().GetErrorString(r, &s);
  // NOT_SUPPORTED / NOT_PERMITTED / SYSTEM_NOT_READY mean "this host does not expose NVLS (right now)", not a bug on our side
  const int code = (r == CUDA_ERROR_NOT_SUPPORTED || r == CUDA_ERROR_NOT_PERMITTED || r == CUDA_ERROR_SYSTEM_NOT_READY) ? KK_EUNSUPPORTED
                   : r == CUDA_ERROR_OUT_OF_MEMORY ? KK_ENOMEM : KK_ECUDA;
  fail(code, "fan-out NVLS: %s: %s (%d)", what, s ? s : "?", (int)r);
}

}  // namespace

struct NvlsPools::Impl {
  std::vector<int> ordinals;
  std::vector<CUdevice> devs;
  size_t size = 0;
  CUmemGenericAllocationHandle mc = 0;
  bool have_mc = false;
  std::vector<CUmemGenericAllocationHandle> mem;  // one physical allocation per device
  std::vector<bool> bound;
  std::vector<CUdeviceptr> uc;                    // unicast mapping of each device's allocation (accessible from every device)
  std::vector<bool> uc_mapped;
  CUdeviceptr mc_va = 0;
  bool mc_mapped = false;
};

NvlsPools::NvlsPools() : p_(new Impl) {}

NvlsPools::~NvlsPools() {
  Impl& s = *p_;
  // best effort, in reverse order of creation; errors during teardown are not reportable
  if (s.mc_va) {
    if (s.mc_mapped) drv().MemUnmap(s.mc_va, s.size);
    drv().MemAddressFree(s.mc_va, s.size);
  }
  for (size_t i = 0; i < s.uc.size(); ++i) {
    if (!s.uc[i]) continue;
    if (s.uc_mapped[i]) drv().MemUnmap(s.uc[i], s.size);
    drv().MemAddressFree(s.uc[i], s.size);
  }
  for (size_t i = 0; i < s.mem.size(); ++i) {
    if (!s.mem[i]) continue;
    if (s.bound[i]) drv().MulticastUnbind(s.mc, s.devs[i], 0, s.size);
    drv().MemRelease(s.mem[i]);
  }
  if (s.have_mc) drv().MemRelease(s.mc);
  delete p_;
}

bool NvlsPools::supported(const std::vector<int>& ordinals, std::string* why) {
  try {
    const Drv& d = drv();
    for (int o : ordinals) {
      CUdevice dev;
      check(d.DeviceGet(&dev, o), "cuDeviceGet");
      int v = 0;
      check(d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev), "cuDeviceGetAttribute(MULTICAST_SUPPORTED)");
      if (!v) {
        if (why) *why = "device " + std::to_string(o) + " does not support multicast (no NVSwitch / NVLS on this host)";
        return false;
      }
    }
    return true;
  } catch (const Error& e) {
    if (why) *why = e.what();
    return false;
  }
}

void NvlsPools::create(const std::vector<int>& ordinals, uint64_t bytes) {
  Impl& s = *p_;
  const Drv& d = drv();
  const size_t n = ordinals.size();
  if (n < 2) fail(KK_EINVAL, "fan-out NVLS needs at least two devices");
  s.ordinals = ordinals;
  s.devs.resize(n);
  for (size_t i = 0; i < n; ++i) check(d.DeviceGet(&s.devs[i], ordinals[i]), "cuDeviceGet");

  CUmulticastObjectProp mp;
  memset(&mp, 0, sizeof mp);
  mp.numDevices = (unsigned)n;
  mp.handleTypes = 0;  // not shared with other processes
  mp.flags = 0;
  mp.size = (size_t)bytes;
  size_t gran = 0, agran = 0;
  check(d.MulticastGetGranularity(&gran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof ap);
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = ordinals[0];
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_NONE;
  check(d.MemGetAllocationGranularity(&agran, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  if (agran > gran) gran = agran;
  if (gran == 0) gran = 2u << 20;
  s.size = (size_t)((bytes + gran - 1) / gran * gran);
  mp.size = s.size;

  check(d.MulticastCreate(&s.mc, &mp), "cuMulticastCreate");
  s.have_mc = true;
  for (size_t i = 0; i < n; ++i) check(d.MulticastAddDevice(s.mc, s.devs[i]), "cuMulticastAddDevice");  // all devices before any bind

  s.mem.assign(n, 0);
  s.bound.assign(n, false);
  s.uc.assign(n, 0);
  s.uc_mapped.assign(n, false);
  std::vector<CUmemAccessDesc> acc(n);
  for (size_t i = 0; i < n; ++i) {
    acc[i].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[i].location.id = ordinals[i];
    acc[i].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  for (size_t i = 0; i < n; ++i) {
    ap.location.id = ordinals[i];
    check(d.MemCreate(&s.mem[i], s.size, &ap, 0), "cuMemCreate");
    check(d.MulticastBindMem(s.mc, 0, s.mem[i], 0, s.size, 0), "cuMulticastBindMem");
    s.bound[i] = true;
    check(d.MemAddressReserve(&s.uc[i], s.size, gran, 0, 0), "cuMemAddressReserve");
    check(d.MemMap(s.uc[i], s.size, 0, s.mem[i], 0), "cuMemMap");
    s.uc_mapped[i] = true;
    check(d.MemSetAccess(s.uc[i], s.size, acc.data(), n), "cuMemSetAccess");
  }
  check(d.MemAddressReserve(&s.mc_va, s.size, gran, 0, 0), "cuMemAddressReserve(multicast)");
  check(d.MemMap(s.mc_va, s.size, 0, s.mc, 0), "cuMemMap(multicast)");
  s.mc_mapped = true;
  check(d.MemSetAccess(s.mc_va, s.size, acc.data(), n), "cuMemSetAccess(multicast)");
}

uint8_t* NvlsPools::pool(size_t i) const { return reinterpret_cast<uint8_t*>(p_->uc[i]); }
uint8_t* NvlsPools::multicast() const { return reinterpret_cast<uint8_t*>(p_->mc_va); }
uint64_t NvlsPools::bytes() const { return p_->size; }

}  // namespace kk
