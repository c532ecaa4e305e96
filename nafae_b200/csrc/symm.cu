// Symmetric multicast memory for the NVLS (in-switch) gradient all-reduce.
//
// New functionality (the reference is single-GPU).  One process per GPU: the
// root creates an NVSwitch multicast object (cuMulticastCreate), exports it as a POSIX file
// descriptor that the host code hands to the other ranks (SCM_RIGHTS, nafae_b200/parallel.py);
// every rank adds its device, allocates `bytes` of physical memory (cuMemCreate), binds it to the
// object and maps BOTH views: `uc` (this rank's own memory, ordinary loads/stores) and `mc` (the
// multicast address: a store is replicated into every rank's memory by the switch, a
// multimem.ld_reduce returns the SUM of every rank's copy, reduced inside the switch).
//
// The driver API is resolved through cudaGetDriverEntryPoint (static cudart), so the library has no
// link-time dependency on libcuda.so and still loads on a box without a driver (CPU-side tests).
#include <cuda.h>
#include <string.h>
#include <unistd.h>

#include "common.cuh"

namespace nafae {
namespace {

struct Driver {
  bool ok = false;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*,
                        unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*,
                                          CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                         unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*,
                                           CUmemAllocationHandleType) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t,
                               size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*,
                                      CUmulticastGranularity_flags) = nullptr;
};

template <typename F>
bool resolve(const char* name, F* out) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || fn == nullptr ||
      q != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(fn);
  return true;
}

Driver& driver() {
  static Driver d;
  static bool tried = false;
  if (tried) return d;
  tried = true;
  bool ok = cudaFree(nullptr) == cudaSuccess;  // runtime + primary context up
  ok = ok && resolve("cuGetErrorString", &d.GetErrorString);
  ok = ok && resolve("cuDeviceGet", &d.DeviceGet);
  ok = ok && resolve("cuDeviceGetAttribute", &d.DeviceGetAttribute);
  ok = ok && resolve("cuMemCreate", &d.MemCreate);
  ok = ok && resolve("cuMemRelease", &d.MemRelease);
  ok = ok && resolve("cuMemAddressReserve", &d.MemAddressReserve);
  ok = ok && resolve("cuMemAddressFree", &d.MemAddressFree);
  ok = ok && resolve("cuMemMap", &d.MemMap);
  ok = ok && resolve("cuMemUnmap", &d.MemUnmap);
  ok = ok && resolve("cuMemSetAccess", &d.MemSetAccess);
  ok = ok && resolve("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity);
  ok = ok && resolve("cuMemExportToShareableHandle", &d.MemExportToShareableHandle);
  ok = ok && resolve("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle);
  ok = ok && resolve("cuMulticastCreate", &d.MulticastCreate);
  ok = ok && resolve("cuMulticastAddDevice", &d.MulticastAddDevice);
  ok = ok && resolve("cuMulticastBindMem", &d.MulticastBindMem);
  ok = ok && resolve("cuMulticastUnbind", &d.MulticastUnbind);
  ok = ok && resolve("cuMulticastGetGranularity", &d.MulticastGetGranularity);
  d.ok = ok;
  return d;
}

const char* cu_err(CUresult r) {
  const char* s = nullptr;
  if (driver().GetErrorString && driver().GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown CUDA driver error";
}

#define NAFAE_CU(call, what)                                       \
  do {                                                             \
    CUresult r_ = (call);                                          \
    if (r_ != CUDA_SUCCESS) {                                      \
      set_error("%s: %s (CUresult %d)", what, cu_err(r_), (int)r_); \
      return -(int)cudaErrorUnknown;                               \
    }                                                              \
  } while (0)

struct Symm {
  CUmemGenericAllocationHandle mc = 0, mem = 0;
  CUdeviceptr uc_va = 0, mc_va = 0;
  size_t size = 0, gran = 0;
  int world = 0, dev = 0;
  bool have_mc = false, have_mem = false, uc_mapped = false, mc_mapped = false, bound = false;
};

CUmulticastObjectProp mc_prop(int world, size_t size) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)world;
  p.size = size;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.flags = 0;
  return p;
}

// same rounded size on every rank: a multiple of the multicast object's MINIMUM granularity and of
// the device allocation granularity (the recommended multicast granularity can be hundreds of MB,
// which only matters for page-table reach on far larger buffers than a 9 MB gradient bucket)
int symm_size(int world, size_t bytes, size_t* out, size_t* gran_out) {
  Driver& d = driver();
  CUmulticastObjectProp p = mc_prop(world, 0);
  size_t g_mc = 0, g_mem = 0;
  NAFAE_CU(d.MulticastGetGranularity(&g_mc, &p, CU_MULTICAST_GRANULARITY_MINIMUM),
           "cuMulticastGetGranularity");
  int dev = 0;
  cudaGetDevice(&dev);
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  NAFAE_CU(d.MemGetAllocationGranularity(&g_mem, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
           "cuMemGetAllocationGranularity");
  size_t g = g_mc > g_mem ? g_mc : g_mem;
  if (g == 0) g = (size_t)2 << 20;
  *out = (bytes + g - 1) / g * g;
  *gran_out = g;
  return 1;
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_API int nafae_mc_supported(void) {
  Driver& d = driver();
  if (!d.ok) return 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  CUdevice cd;
  if (d.DeviceGet(&cd, dev) != CUDA_SUCCESS) return 0;
  int v = 0;
  if (d.DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd) != CUDA_SUCCESS) return 0;
  return v ? 1 : 0;
}

static int symm_new(int world, size_t bytes, Symm** out) {
  NAFAE_REQUIRE(out != nullptr && world >= 2 && world <= 64 && bytes > 0, "mc: bad arguments");
  NAFAE_REQUIRE(driver().ok, "mc: CUDA driver entry points unavailable");
  NAFAE_REQUIRE(nafae_mc_supported() == 1, "mc: device does not support NVSwitch multicast");
  Symm* s = new Symm();
  s->world = world;
  cudaGetDevice(&s->dev);
  int st = symm_size(world, bytes, &s->size, &s->gran);
  if (st != 1) {
    delete s;
    return st;
  }
  *out = s;
  return 1;
}

NAFAE_API int nafae_mc_create(int world, size_t bytes, void** handle, int* fd_out) {
  NAFAE_REQUIRE(handle && fd_out, "mc_create: NULL argument");
  Symm* s = nullptr;
  int st = symm_new(world, bytes, &s);
  if (st != 1) return st;
  Driver& d = driver();
  CUmulticastObjectProp p = mc_prop(world, s->size);
  CUresult r = d.MulticastCreate(&s->mc, &p);
  if (r != CUDA_SUCCESS) {
    set_error("cuMulticastCreate(%d devices, %zu bytes): %s", world, s->size, cu_err(r));
    delete s;
    return -(int)cudaErrorUnknown;
  }
  s->have_mc = true;
  int fd = -1;
  r = d.MemExportToShareableHandle(&fd, s->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    set_error("cuMemExportToShareableHandle(multicast): %s", cu_err(r));
    d.MemRelease(s->mc);
    delete s;
    return -(int)cudaErrorUnknown;
  }
  *fd_out = fd;
  *handle = s;
  return 1;
}

NAFAE_API int nafae_mc_import(int fd, int world, size_t bytes, void** handle) {
  NAFAE_REQUIRE(handle && fd >= 0, "mc_import: bad argument");
  Symm* s = nullptr;
  int st = symm_new(world, bytes, &s);
  if (st != 1) return st;
  CUresult r = driver().MemImportFromShareableHandle(&s->mc, (void*)(uintptr_t)fd,
                                                     CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) {
    set_error("cuMemImportFromShareableHandle(multicast fd %d): %s", fd, cu_err(r));
    delete s;
    return -(int)cudaErrorUnknown;
  }
  s->have_mc = true;
  *handle = s;
  return 1;
}

NAFAE_API int nafae_mc_add_device(void* handle) {
  NAFAE_REQUIRE(handle, "mc_add_device: NULL handle");
  Symm* s = static_cast<Symm*>(handle);
  CUdevice cd;
  NAFAE_CU(driver().DeviceGet(&cd, s->dev), "cuDeviceGet");
  NAFAE_CU(driver().MulticastAddDevice(s->mc, cd), "cuMulticastAddDevice");
  return 1;
}

// Call after EVERY rank has returned from nafae_mc_add_device (host-side barrier in between).
NAFAE_API int nafae_mc_bind(void* handle, void** uc_ptr, void** mc_ptr) {
  NAFAE_REQUIRE(handle && uc_ptr && mc_ptr, "mc_bind: NULL argument");
  Symm* s = static_cast<Symm*>(handle);
  Driver& d = driver();
  CUmemAllocationProp ap;
  memset(&ap, 0, sizeof(ap));
  ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  ap.location.id = s->dev;
  ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  NAFAE_CU(d.MemCreate(&s->mem, s->size, &ap, 0), "cuMemCreate");
  s->have_mem = true;
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = s->dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  NAFAE_CU(d.MemAddressReserve(&s->uc_va, s->size, s->gran, 0, 0),
           "cuMemAddressReserve(uc)");
  NAFAE_CU(d.MemMap(s->uc_va, s->size, 0, s->mem, 0), "cuMemMap(uc)");
  s->uc_mapped = true;
  NAFAE_CU(d.MemSetAccess(s->uc_va, s->size, &acc, 1), "cuMemSetAccess(uc)");
  if (cudaMemset(reinterpret_cast<void*>(s->uc_va), 0, s->size) != cudaSuccess ||
      cudaDeviceSynchronize() != cudaSuccess) {
    set_error("mc_bind: cudaMemset of the new buffer failed");
    return -(int)cudaErrorUnknown;
  }
  NAFAE_CU(d.MulticastBindMem(s->mc, 0, s->mem, 0, s->size, 0), "cuMulticastBindMem");
  s->bound = true;
  NAFAE_CU(d.MemAddressReserve(&s->mc_va, s->size, s->gran, 0, 0), "cuMemAddressReserve(mc)");
  NAFAE_CU(d.MemMap(s->mc_va, s->size, 0, s->mc, 0), "cuMemMap(mc)");
  s->mc_mapped = true;
  NAFAE_CU(d.MemSetAccess(s->mc_va, s->size, &acc, 1), "cuMemSetAccess(mc)");
  *uc_ptr = reinterpret_cast<void*>(s->uc_va);
  *mc_ptr = reinterpret_cast<void*>(s->mc_va);
  return 1;
}

NAFAE_API size_t nafae_mc_size(void* handle) {
  return handle ? static_cast<Symm*>(handle)->size : 0;
}

NAFAE_API int nafae_mc_free(void* handle) {
  if (!handle) return 1;
  Symm* s = static_cast<Symm*>(handle);
  Driver& d = driver();
  cudaDeviceSynchronize();
  if (s->mc_mapped) d.MemUnmap(s->mc_va, s->size);
  if (s->mc_va) d.MemAddressFree(s->mc_va, s->size);
  if (s->bound) {
    CUdevice cd;
    if (d.DeviceGet(&cd, s->dev) == CUDA_SUCCESS) d.MulticastUnbind(s->mc, cd, 0, s->size);
  }
  if (s->uc_mapped) d.MemUnmap(s->uc_va, s->size);
  if (s->uc_va) d.MemAddressFree(s->uc_va, s->size);
  if (s->have_mem) d.MemRelease(s->mem);
  if (s->have_mc) d.MemRelease(s->mc);
  delete s;
  return 1;
}
