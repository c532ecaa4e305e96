// Error reporting and small host-side utilities shared by the entry points.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace nafae {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 1;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return -(int)e;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// per device (one process may drive several GPUs); relaxed atomics: a plain configuration value
static std::atomic<int> g_reserved_sms[64];

static int current_device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return dev;
}

int persistent_grid() {
  int n = sm_count() - g_reserved_sms[current_device_slot()].load(std::memory_order_relaxed);
  return n < 1 ? 1 : n;
}

// gate layout (ints): [0] arrivals of the current launch, [1] epoch, [2 + slot] epochs seen by slot
__global__ void gate_wait_kernel(int* gate, int slot) {
  NAFAE_CTA_TRACE(cta_trace, 6);
  if (threadIdx.x == 0) {
    const int seen = gate[2 + slot];
    int epoch;
    do {
      asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(epoch) : "l"(gate + 1) : "memory");
      if (epoch == seen) __nanosleep(200);
    } while (epoch == seen);
    gate[2 + slot] = epoch;
  }
}

__global__ void gate_open_kernel(int* gate) {  // paths that do not run the persistent kernel
  if (threadIdx.x == 0) atomicAdd(gate + 1, 1);
}

// every slot has "seen" every launch so far: the next wait of any slot blocks until the NEXT open
__global__ void gate_sync_kernel(int* gate) {
  if (threadIdx.x < 6) gate[2 + threadIdx.x] = gate[1];
}

// diagnostics: CTAs that hold `smem` bytes of shared memory (i.e. a whole SM each when large) and
// spin for a fixed wall time -- lets a test squeeze the other kernels onto the few remaining SMs
__global__ void occupy_kernel(unsigned long long ns) {
  extern __shared__ unsigned char occupy_smem[];
  if (threadIdx.x == 0) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    occupy_smem[0] = 1;
    do {
      __nanosleep(1000);
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
  }
}

void gate_open(void* gate, cudaStream_t stream) {
  gate_open_kernel<<<1, 32, 0, stream>>>(static_cast<int*>(gate));
}

}  // namespace nafae

NAFAE_CTA_TRACE_READER(nafae_debug_cta_trace_runtime)

NAFAE_API int nafae_gate_wait(void* gate, int slot, cudaStream_t stream) {
  NAFAE_REQUIRE(gate != nullptr && slot >= 0 && slot < 6, "gate_wait: bad gate/slot");
  nafae::gate_wait_kernel<<<1, 32, 0, stream>>>(static_cast<int*>(gate), slot);
  return nafae::launch_status("gate_wait_kernel");
}

NAFAE_API int nafae_debug_occupy_sms(int num_ctas, int smem_bytes, unsigned long long nanoseconds,
                                     cudaStream_t stream) {
  NAFAE_REQUIRE(num_ctas >= 1 && num_ctas <= 1024 && smem_bytes >= 0 && smem_bytes <= 227 * 1024 &&
                    nanoseconds <= 2000000000ull,
                "debug_occupy: bad arguments (at most 1024 CTAs, 227 KB, 2 s)");
  if (smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nafae::occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes);
    if (e != cudaSuccess) {
      nafae::set_error("debug_occupy: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  nafae::occupy_kernel<<<num_ctas, 32, smem_bytes, stream>>>(nanoseconds);
  return nafae::launch_status("occupy_kernel");
}

NAFAE_API int nafae_gate_sync(void* gate, cudaStream_t stream) {
  NAFAE_REQUIRE(gate != nullptr, "gate_sync: NULL gate");
  nafae::gate_sync_kernel<<<1, 32, 0, stream>>>(static_cast<int*>(gate));
  return nafae::launch_status("gate_sync_kernel");
}

NAFAE_API int nafae_set_reserved_sms(int n) {
  return nafae::g_reserved_sms[nafae::current_device_slot()].exchange(n < 0 ? 0 : n,
                                                                      std::memory_order_relaxed);
}

NAFAE_API int nafae_abi_version(void) { return NAFAE_B200_ABI_VERSION; }
NAFAE_API const char* nafae_last_error(void) { return nafae::g_err; }
