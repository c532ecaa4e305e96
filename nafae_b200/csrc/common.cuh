// Shared helpers for the sm_100a kernels of libnafae_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "nafae_b200.h"

#define NAFAE_API extern "C" __attribute__((visibility("default")))

namespace nafae {

// status codes of the C ABI: 1 ok, 0 invalid argument, <0 = -cudaError_t
void set_error(const char* fmt, ...);
int launch_status(const char* what);  // cudaGetLastError() -> 1 or -(err), records message

#define NAFAE_REQUIRE(cond, ...)         \
  do {                                   \
    if (!(cond)) {                       \
      ::nafae::set_error(__VA_ARGS__);   \
      return 0;                          \
    }                                    \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int sm_count();  // cached multiprocessor count of the current device
int persistent_grid();  // sm_count() minus the SMs reserved for concurrent collectives
void gate_open(void* gate, cudaStream_t stream);  // bump a residency gate's epoch (see header)

// ---------------------------------------------------------------- device side helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (TMA engine, no tensor map), completion on an mbarrier.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 1-D bulk async copy shared -> global (bulk-group completion).
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// "Last CTA to arrive continues" ticket: ONE thread publishes the whole CTA's prior global writes
// (callers do __syncthreads() first; release is cumulative over the barrier) and acquires the
// other CTAs' -- much cheaper than every thread executing __threadfence() (fence.sc.gpu).
__device__ __forceinline__ int ticket_acq_rel(int* counter) {
  int old;
  asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(counter) : "memory");
  return old;
}

// ---------------------------------------------------------------- CTA timeline (debug) ----
// -DNAFAE_TRACE builds (libnafae_b200_trace.so, tools/timeline.py) record every CTA's start / end
// (%globaltimer, ns), SM id and two kernel-defined counters; one buffer per translation unit.
#ifdef NAFAE_TRACE
struct CtaRec {
  unsigned long long t0, t1;
  int kernel, cta, smid, a, b, pad;
};
constexpr int kCtaRecMax = 1 << 15;
static __device__ CtaRec g_cta_rec[kCtaRecMax];
static __device__ int g_cta_rec_n;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
struct CtaTrace {
  int idx;
  __device__ __forceinline__ explicit CtaTrace(int kernel) : idx(-1) {
    if (threadIdx.x == 0) {
      idx = atomicAdd(&g_cta_rec_n, 1);
      if (idx < kCtaRecMax) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_cta_rec[idx].t0 = global_ns();
        g_cta_rec[idx].kernel = kernel;
        g_cta_rec[idx].cta = blockIdx.x + gridDim.x * blockIdx.y;
        g_cta_rec[idx].smid = (int)smid;
        g_cta_rec[idx].a = g_cta_rec[idx].b = 0;
      }
    }
  }
  // any ONE thread of the CTA may count (thread 0 publishes the slot through shared memory)
  __device__ __forceinline__ void publish(int* smem_slot) {
    if (threadIdx.x == 0) *smem_slot = idx;
  }
  static __device__ __forceinline__ void count(int slot, int a, int b) {
    if (slot >= 0 && slot < kCtaRecMax) {
      g_cta_rec[slot].a += a;
      g_cta_rec[slot].b += b;
    }
  }
  // thread 0: phase stamps, ns since the CTA started, into the record's a / b counters
  __device__ __forceinline__ void mark_a() {
    if (idx >= 0 && idx < kCtaRecMax) g_cta_rec[idx].a = (int)(global_ns() - g_cta_rec[idx].t0);
  }
  __device__ __forceinline__ void mark_b() {
    if (idx >= 0 && idx < kCtaRecMax) g_cta_rec[idx].b = (int)(global_ns() - g_cta_rec[idx].t0);
  }
  __device__ __forceinline__ ~CtaTrace() {
    if (idx >= 0 && idx < kCtaRecMax) g_cta_rec[idx].t1 = global_ns();
  }
};
#define NAFAE_CTA_TRACE(name, kernel) ::nafae::CtaTrace name(kernel)
#define NAFAE_CTA_TRACE_READER(fn)                                                          \
  NAFAE_API int fn(void* host_out, int max_recs, int reset) {                                \
    int n = 0;                                                                               \
    cudaMemcpyFromSymbol(&n, ::nafae::g_cta_rec_n, sizeof(int));                             \
    if (n > ::nafae::kCtaRecMax) n = ::nafae::kCtaRecMax;                                    \
    if (n > max_recs) n = max_recs;                                                          \
    if (n > 0) cudaMemcpyFromSymbol(host_out, ::nafae::g_cta_rec, sizeof(::nafae::CtaRec) * n); \
    if (reset) {                                                                             \
      const int z = 0;                                                                       \
      cudaMemcpyToSymbol(::nafae::g_cta_rec_n, &z, sizeof(int));                             \
    }                                                                                        \
    return n;                                                                                \
  }
#else
struct CtaTrace {
  __device__ __forceinline__ void publish(int*) {}
  static __device__ __forceinline__ void count(int, int, int) {}
  __device__ __forceinline__ void mark_a() {}
  __device__ __forceinline__ void mark_b() {}
};
#define NAFAE_CTA_TRACE(name, kernel) [[maybe_unused]] ::nafae::CtaTrace name
#define NAFAE_CTA_TRACE_READER(fn)
#endif

template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace nafae
