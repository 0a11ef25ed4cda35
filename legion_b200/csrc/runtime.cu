// runtime.cu — error reporting and the thin device plumbing a non-CUDA host needs
// (storage/storage_management.cu:5-23,100-115; engine/ipc_service.cu:163-169;
//  training_backend/ipc_cuda_kernel.cu:62-68).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include <sys/mman.h>
#include <unistd.h>
#include <mutex>
#include <vector>
#include <stdlib.h>

static thread_local char g_err[1024] = "";

int lg_l2_hints() {  // see common.cuh; read once
  static int v = [] {
    const char* e = getenv("LG_L2_HINTS");
    return e ? atoi(e) : 4;  // default: keep the small random-access arrays (profiles/r01b_l2_hints.md)
  }();
  return v;
}

int lg_pdl() {  // see common.cuh; read once
  static int v = [] {
    const char* e = getenv("LG_PDL");
    return e ? atoi(e) : -1;
  }();
  return v;
}

int lg_chain_carveout() {  // see common.cuh; read once
  static int v = [] {
    const char* e = getenv("LG_CARVEOUT");
    return e ? atoi(e) : -1;
  }();
  return v;
}
void lg_apply_carveout(const void* kernel) {  // once per kernel
  static std::mutex mu;
  static std::vector<const void*> done;
  std::lock_guard<std::mutex> g(mu);
  for (const void* k : done)
    if (k == kernel) return;
  done.push_back(kernel);
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, lg_chain_carveout());
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void lg_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long lg_debug_launch_count(int32_t reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int lg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

extern "C" const char* lg_last_error(void) { return g_err; }
extern "C" int lg_version(void) { return 100; }

extern "C" int lg_device_count(int32_t* n) {
  LG_REQUIRE(n, "null");
  int c = 0;
  LG_CUDA(cudaGetDeviceCount(&c));
  *n = c;
  return 0;
}
extern "C" int lg_set_device(int32_t device) {
  LG_CUDA(cudaSetDevice(device));
  return 0;
}
extern "C" int lg_enable_peer_access(int32_t n_devices) {
  int cur = 0;
  LG_CUDA(cudaGetDevice(&cur));
  for (int i = 0; i < n_devices; i++) {
    LG_CUDA(cudaSetDevice(i));
    for (int j = 0; j < n_devices; j++) {
      if (i == j) continue;
      int ok = 0;
      LG_CUDA(cudaDeviceCanAccessPeer(&ok, i, j));
      if (!ok) continue;
      cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        continue;
      }
      LG_CUDA(e);
    }
  }
  LG_CUDA(cudaSetDevice(cur));
  return 0;
}
// ---- VMM allocations for cache shards that other processes map (one process per GPU) ----
// A legacy cudaIpcOpenMemHandle mapping is read through small pages on the importing GPU: random 512-byte rows out of
// a 2.8 GB peer shard reach 95 GB/s, out of a 0.5 GB shard 630 GB/s (profiles/r01b_peer_mapping.md).  cuMemCreate +
// cuMemExportToShareableHandle (POSIX fd) + cuMemMap on the importer keeps the allocation's 2 MB pages.  The driver
// entry points are fetched through cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda
// (it must still load on a box without a GPU).
#include <cuda.h>
namespace {
struct Vmm {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  bool ok = false;
};
int vmm_load(Vmm* v) {
  static Vmm cached;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaFree(0);  // make sure the primary context exists
    struct {
      const char* name;
      void** fn;
    } tab[] = {{"cuMemCreate", (void**)&cached.MemCreate},
               {"cuMemRelease", (void**)&cached.MemRelease},
               {"cuMemAddressReserve", (void**)&cached.MemAddressReserve},
               {"cuMemAddressFree", (void**)&cached.MemAddressFree},
               {"cuMemMap", (void**)&cached.MemMap},
               {"cuMemUnmap", (void**)&cached.MemUnmap},
               {"cuMemSetAccess", (void**)&cached.MemSetAccess},
               {"cuMemExportToShareableHandle", (void**)&cached.MemExport},
               {"cuMemImportFromShareableHandle", (void**)&cached.MemImport},
               {"cuMemGetAllocationGranularity", (void**)&cached.MemGranularity}};
    bool all = true;
    for (auto& t : tab) {
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint(t.name, t.fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*t.fn)
        all = false;
    }
    cudaGetLastError();
    cached.ok = all;
  }
  *v = cached;
  return cached.ok ? 0 : lg_set_error("CUDA VMM driver entry points are not available");
}
struct VmmRegion {
  void* ptr;
  size_t bytes;
  CUmemGenericAllocationHandle handle;
};
std::mutex g_vmm_mu;
std::vector<VmmRegion> g_vmm;
CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
int vmm_map(const Vmm& v, CUmemGenericAllocationHandle h, size_t bytes, int access_device, void** out) {
  CUdeviceptr va = 0;
  CUresult r = v.MemAddressReserve(&va, bytes, 0, 0, 0);
  if (r != CUDA_SUCCESS) return lg_set_error("cuMemAddressReserve(%zu) -> %d", bytes, (int)r);
  r = v.MemMap(va, bytes, 0, h, 0);
  if (r != CUDA_SUCCESS) {
    v.MemAddressFree(va, bytes);
    return lg_set_error("cuMemMap(%zu) -> %d", bytes, (int)r);
  }
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = access_device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = v.MemSetAccess(va, bytes, &acc, 1);
  if (r != CUDA_SUCCESS) {
    v.MemUnmap(va, bytes);
    v.MemAddressFree(va, bytes);
    return lg_set_error("cuMemSetAccess(device %d) -> %d", access_device, (int)r);
  }
  *out = (void*)va;
  std::lock_guard<std::mutex> g(g_vmm_mu);
  g_vmm.push_back({(void*)va, bytes, h});
  return 0;
}
}  // namespace

// the length an allocation of `bytes` occupies: a multiple of the driver's RECOMMENDED granularity for device memory on the
// current device (2 MB on B200) — the exporter and the importer both derive the mapped length from it, so they agree
// whatever the granularity is
static size_t vmm_granularity(Vmm& v, int dev) {
  CUmemAllocationProp prop = vmm_prop(dev);
  size_t gran = 0;
  if (!v.ok || v.MemGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) gran = 2u << 20;
  return gran;
}
extern "C" int64_t lg_vmm_round_up(int64_t bytes) {
  Vmm v;
  int dev = 0;
  size_t g = 2u << 20;
  if (vmm_load(&v) == 0 && cudaGetDevice(&dev) == cudaSuccess) g = vmm_granularity(v, dev);
  return (int64_t)(((size_t)bytes + g - 1) / g * g);
}

extern "C" int lg_vmm_alloc(int64_t bytes, void** ptr, int32_t* shareable_fd) {
  LG_REQUIRE(ptr && shareable_fd && bytes > 0, "lg_vmm_alloc: bad argument");
  Vmm v;
  int rc = vmm_load(&v);
  if (rc) return rc;
  int dev = 0;
  LG_CUDA(cudaGetDevice(&dev));
  CUmemAllocationProp prop = vmm_prop(dev);
  const size_t gran = vmm_granularity(v, dev);
  const size_t len = ((size_t)bytes + gran - 1) / gran * gran;
  CUmemGenericAllocationHandle h;
  CUresult r = v.MemCreate(&h, len, &prop, 0);
  if (r != CUDA_SUCCESS) return lg_set_error("cuMemCreate(%zu bytes on device %d) -> %d", len, dev, (int)r);
  int fd = -1;
  r = v.MemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    v.MemRelease(h);
    return lg_set_error("cuMemExportToShareableHandle -> %d", (int)r);
  }
  rc = vmm_map(v, h, len, dev, ptr);
  if (rc) {
    close(fd);
    v.MemRelease(h);
    return rc;
  }
  *shareable_fd = fd;
  return 0;
}

extern "C" int lg_vmm_import(int32_t shareable_fd, int64_t bytes, void** ptr) {
  LG_REQUIRE(ptr && shareable_fd >= 0 && bytes > 0, "lg_vmm_import: bad argument");
  Vmm v;
  int rc = vmm_load(&v);
  if (rc) return rc;
  int dev = 0;
  LG_CUDA(cudaGetDevice(&dev));
  CUmemGenericAllocationHandle h;
  CUresult r = v.MemImport(&h, (void*)(uintptr_t)shareable_fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (r != CUDA_SUCCESS) return lg_set_error("cuMemImportFromShareableHandle(fd %d) -> %d", shareable_fd, (int)r);
  const size_t gran = vmm_granularity(v, dev);  // same rule as lg_vmm_alloc on the exporting GPU (same architecture)
  const size_t len = ((size_t)bytes + gran - 1) / gran * gran;
  rc = vmm_map(v, h, len, dev, ptr);
  if (rc) v.MemRelease(h);
  return rc;
}

extern "C" int lg_vmm_free(void* ptr) {
  Vmm v;
  int rc = vmm_load(&v);
  if (rc) return rc;
  VmmRegion reg{nullptr, 0, 0};
  {
    std::lock_guard<std::mutex> g(g_vmm_mu);
    for (size_t i = 0; i < g_vmm.size(); i++)
      if (g_vmm[i].ptr == ptr) {
        reg = g_vmm[i];
        g_vmm.erase(g_vmm.begin() + i);
        break;
      }
  }
  LG_REQUIRE(reg.ptr, "lg_vmm_free: %p is not a VMM allocation of this library", ptr);
  v.MemUnmap((CUdeviceptr)reg.ptr, reg.bytes);
  v.MemAddressFree((CUdeviceptr)reg.ptr, reg.bytes);
  v.MemRelease(reg.handle);
  return 0;
}

extern "C" int lg_device_alloc(void** ptr, int64_t bytes) {
  LG_REQUIRE(ptr && bytes >= 0, "lg_device_alloc: bad argument");
  LG_CUDA(cudaMalloc(ptr, (size_t)(bytes > 0 ? bytes : 1)));
  return 0;
}
extern "C" int lg_device_free(void* ptr) {
  LG_CUDA(cudaFree(ptr));
  return 0;
}
// Pinned, device-mapped host memory (storage/storage_management.cu:108-109 uses cudaHostAllocMapped).
// LG_HOST_HUGEPAGES=1: large regions come from an anonymous mapping advised to transparent huge pages and are then
// registered (cudaHostRegisterMapped) — fewer host page-table / IOMMU entries behind the random UVA reads of the
// host tiers.  Registered regions are remembered so lg_host_free can undo the right way.
namespace {
struct HugeRegion {
  void* p;
  size_t bytes;
};
std::mutex g_huge_mu;
std::vector<HugeRegion> g_huge;
}  // namespace
extern "C" int lg_host_alloc_mapped(void** host_ptr, void** device_ptr, int64_t bytes) {
  LG_REQUIRE(host_ptr && bytes >= 0, "lg_host_alloc_mapped: bad argument");
  static const bool huge = [] {
    const char* e = getenv("LG_HOST_HUGEPAGES");
    return e && atoi(e) != 0;
  }();
  if (huge && bytes >= (64ll << 20)) {
    const size_t two_mb = 2u << 20;
    const size_t len = ((size_t)bytes + two_mb - 1) / two_mb * two_mb;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    LG_REQUIRE(p != MAP_FAILED, "lg_host_alloc_mapped: mmap of %zu bytes failed", len);
    madvise(p, len, MADV_HUGEPAGE);
    for (size_t off = 0; off < len; off += two_mb) ((volatile char*)p)[off] = 0;  // fault the huge pages in
    cudaError_t e = cudaHostRegister(p, len, cudaHostRegisterMapped | cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      munmap(p, len);
      return lg_set_error("cudaHostRegister(%zu bytes) -> %s", len, cudaGetErrorString(e));
    }
    *host_ptr = p;
    {
      std::lock_guard<std::mutex> g(g_huge_mu);
      g_huge.push_back({p, len});
    }
  } else {
    LG_CUDA(cudaHostAlloc(host_ptr, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocMapped | cudaHostAllocPortable));
  }
  if (device_ptr) LG_CUDA(cudaHostGetDevicePointer(device_ptr, *host_ptr, 0));
  return 0;
}
// Large regions are registered in chunks (one cudaHostRegister of tens of GB of shared-memory pages fails on some
// kernels / hypervisors); with unified addressing a registered host pointer IS its device pointer, so the chunks still
// form one contiguous device range — checked chunk by chunk.
namespace {
struct RegRegion {
  void* p;
  size_t bytes, chunk;
};
std::mutex g_reg_mu;
std::vector<RegRegion> g_reg;
}  // namespace
extern "C" int lg_host_register(void* host_ptr, int64_t bytes, void** device_ptr) {
  LG_REQUIRE(host_ptr && bytes > 0, "lg_host_register: bad argument");
  size_t chunk = (size_t)1 << 30;
  if (const char* e = getenv("LG_HOST_REGISTER_CHUNK_MB")) chunk = (size_t)atoll(e) << 20;
  if (chunk == 0 || chunk > (size_t)bytes) chunk = (size_t)bytes;
  char* base = (char*)host_ptr;
  for (size_t off = 0; off < (size_t)bytes; off += chunk) {
    const size_t n = off + chunk <= (size_t)bytes ? chunk : (size_t)bytes - off;
    cudaError_t e = cudaHostRegister(base + off, n, cudaHostRegisterPortable | cudaHostRegisterMapped);
    void* d = nullptr;
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&d, base + off, 0);
    if (e == cudaSuccess && d != (void*)(base + off)) e = cudaErrorNotSupported;  // no unified host/device addressing
    if (e != cudaSuccess) {
      for (size_t o2 = 0; o2 < off; o2 += chunk) cudaHostUnregister(base + o2);
      if (d) cudaHostUnregister(base + off);
      cudaGetLastError();
      return lg_set_error("lg_host_register: chunk at +%zu MB of %lld MB -> %s", off >> 20, (long long)(bytes >> 20),
                          cudaGetErrorString(e));
    }
  }
  {
    std::lock_guard<std::mutex> g(g_reg_mu);
    g_reg.push_back({host_ptr, (size_t)bytes, chunk});
  }
  if (device_ptr) *device_ptr = host_ptr;
  return 0;
}
extern "C" int lg_host_unregister(void* host_ptr) {
  LG_REQUIRE(host_ptr, "lg_host_unregister: null");
  RegRegion r{nullptr, 0, 0};
  {
    std::lock_guard<std::mutex> g(g_reg_mu);
    for (size_t i = 0; i < g_reg.size(); i++)
      if (g_reg[i].p == host_ptr) {
        r = g_reg[i];
        g_reg.erase(g_reg.begin() + i);
        break;
      }
  }
  LG_REQUIRE(r.p, "lg_host_unregister: pointer was not registered through lg_host_register");
  for (size_t off = 0; off < r.bytes; off += r.chunk) cudaHostUnregister((char*)host_ptr + off);
  cudaGetLastError();
  return 0;
}
extern "C" int lg_host_free(void* host_ptr) {
  {
    std::lock_guard<std::mutex> g(g_huge_mu);
    for (size_t i = 0; i < g_huge.size(); i++)
      if (g_huge[i].p == host_ptr) {
        cudaHostUnregister(host_ptr);
        munmap(host_ptr, g_huge[i].bytes);
        g_huge.erase(g_huge.begin() + i);
        return 0;
      }
  }
  LG_CUDA(cudaFreeHost(host_ptr));
  return 0;
}
extern "C" int lg_ipc_export(const void* device_ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes on the wire (shmStruct)");
  LG_REQUIRE(device_ptr && handle, "lg_ipc_export: null argument");
  cudaIpcMemHandle_t h;
  LG_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
  memcpy(handle, &h, 64);
  return 0;
}
extern "C" int lg_ipc_open(const unsigned char handle[64], void** device_ptr) {
  LG_REQUIRE(device_ptr && handle, "lg_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  LG_CUDA(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int lg_ipc_close(void* device_ptr) {
  LG_CUDA(cudaIpcCloseMemHandle(device_ptr));
  return 0;
}
extern "C" int lg_stream_create(lg_stream_t* stream) {
  LG_REQUIRE(stream, "null");
  cudaStream_t s;
  LG_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (lg_stream_t)s;
  return 0;
}
extern "C" int lg_stream_destroy(lg_stream_t stream) {
  LG_CUDA(cudaStreamDestroy((cudaStream_t)stream));
  return 0;
}
extern "C" int lg_stream_synchronize(lg_stream_t stream) {
  LG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
extern "C" int lg_memcpy_h2d(void* dst, const void* src, int64_t bytes, lg_stream_t stream) {
  LG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}
extern "C" int lg_memcpy_d2h(void* dst, const void* src, int64_t bytes, lg_stream_t stream) {
  LG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}
extern "C" int lg_memcpy_d2d(void* dst, const void* src, int64_t bytes, lg_stream_t stream) {
  LG_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}
extern "C" int lg_memset_async(void* dst, int32_t byte_value, int64_t bytes, lg_stream_t stream) {
  LG_CUDA(cudaMemsetAsync(dst, byte_value, (size_t)bytes, (cudaStream_t)stream));
  return 0;
}
extern "C" int lg_event_create(lg_event_t* ev) {
  LG_REQUIRE(ev, "null");
  cudaEvent_t e;
  LG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *ev = (lg_event_t)e;
  return 0;
}
extern "C" int lg_event_destroy(lg_event_t ev) {
  LG_CUDA(cudaEventDestroy((cudaEvent_t)ev));
  return 0;
}
extern "C" int lg_event_record(lg_event_t ev, lg_stream_t stream) {
  LG_CUDA(cudaEventRecord((cudaEvent_t)ev, (cudaStream_t)stream));
  return 0;
}
extern "C" int lg_event_query(lg_event_t ev, int32_t* ready) {
  LG_REQUIRE(ready, "null");
  cudaError_t e = cudaEventQuery((cudaEvent_t)ev);
  if (e == cudaErrorNotReady) {
    *ready = 0;
    return 0;
  }
  LG_CUDA(e);
  *ready = 1;
  return 0;
}
extern "C" int lg_event_synchronize(lg_event_t ev) {
  LG_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
  return 0;
}
extern "C" int lg_stream_wait_event(lg_stream_t stream, lg_event_t ev) {
  LG_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)ev, 0));
  return 0;
}
extern "C" int lg_device_mem_info(int64_t* free_bytes, int64_t* total_bytes) {
  size_t f = 0, t = 0;
  LG_CUDA(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = (int64_t)f;
  if (total_bytes) *total_bytes = (int64_t)t;
  return 0;
}

// diagnostics: a kernel that only spins (no memory traffic) — scripts/overlap_probe.py uses it to tell SM-side
// contention from memory-side contention next to the gather
__global__ void lg_spin_kernel(long long cycles, int* sink) {
  asm volatile("griddepcontrol.wait;" ::: "memory");  // no-op unless launched with the PDL attribute
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long t0 = clock64();
  int x = threadIdx.x;
  while (clock64() - t0 < cycles) x = x * 1664525 + 1013904223;
  if (x == 0x7fffffff && sink) *sink = x;
}
// LG_SPIN_PDL=1: the spinner is launched with programmatic stream serialization (probe: does a PDL launch disturb the
// gather less than a plain one?)
extern "C" int lg_debug_spin(lg_stream_t stream, int32_t ctas, int32_t threads, int64_t cycles) {
  static int carve = [] {
    const char* e = getenv("LG_SPIN_CARVEOUT");
    int v = e ? atoi(e) : -1;
    if (v >= 0) cudaFuncSetAttribute(lg_spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, v);
    return v;
  }();
  (void)carve;
  static const bool pdl = [] {
    const char* e = getenv("LG_SPIN_PDL");
    return e && atoi(e) != 0;
  }();
  LG_CUDA(lg_launch_opt(pdl, lg_spin_kernel, ctas, threads, 0, (cudaStream_t)stream, (long long)cycles, (int*)nullptr));
  return 0;
}
