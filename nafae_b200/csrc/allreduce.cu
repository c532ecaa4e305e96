// Gradient all-reduce (AVG) over NVLink 5 / NVSwitch peer memory for the data-parallel step.
//
// The reference is single-GPU ("--mGPUs parsed, never read"); data parallelism
// is new functionality and its only collective is the per-step all-reduce of the flat fp32 bucket of
// trainable gradients (8.8 MB).  At ~60 us per step an NCCL all-reduce (58 us alone with 32 channels,
// ~140 us when it shares the GPU with the HBM-bound RoIAlign kernel) is the critical path, so the
// bucket lives in a symmetric buffer that every rank maps through CUDA IPC and ONE kernel does a
// two-shot all-reduce with plain peer loads:
//
//   barrier 0   every rank's bucket is complete (its producer ran earlier on the same stream)
//   phase 1     rank r reduces slice r: reads it from all `world` buffers in rank order (so every
//               element is summed in ONE fixed order: replicas end up bit-identical), scales by
//               1/world, writes it back into its own buffer
//   barrier 1   all slices reduced
//   phase 2     every rank copies the other ranks' reduced slices into its own buffer
//   barrier 2   nobody still reads my slice: the next step may overwrite the bucket
//
// CTA b of every rank handles chunk b of every slice and synchronises only with CTA b of the other
// ranks through per-CTA flag words in peer memory (st.release.sys / ld.acquire.sys): no intra-GPU
// grid barrier, and since CTAs are dispatched in index order on every rank the lowest unfinished
// chunk is always resident everywhere (no deadlock even when few SMs are free).  CTAs are small
// (no shared memory, <= 64 registers) and come in two sizes: 256 threads -- four fit on each SM the
// persistent RoIAlign kernel leaves free (nafae_set_reserved_sms) -- and 128 threads, 8 K registers:
// one of those fits BESIDE a resident RoIAlign CTA (544 threads x 96 registers leave 13 K), so the
// collective needs no SMs of its own.  In both cases launch it behind the RoIAlign kernel's residency
// gate (nafae_gate_wait): started at the same instant, its CTAs would land on every SM first and
// keep the 210 KB persistent CTAs out until the whole all-reduce has finished.
// The kernel has no host-side state (the epoch lives in the buffer): it is CUDA-graph capturable.
#include <string.h>

#include "common.cuh"

namespace nafae {
namespace {

constexpr int kArThreadsMax = 256;
constexpr int kArMaxWorld = 8;
constexpr int kArMaxCtas = 256;
// header layout (bytes) at the start of every rank's symmetric buffer
constexpr size_t kArFlagsBytes = (size_t)3 * kArMaxCtas * kArMaxWorld * sizeof(unsigned);  // [stage][cta][rank]
constexpr size_t kArHeaderBytes = kArFlagsBytes + 256;  // + epoch word, finish ticket

struct ArParams {
  char* bufs[kArMaxWorld];  // every rank's buffer base (own one included), as mapped in THIS process
  int rank, world;
  long long count;  // floats
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Plain (weak) 128-bit load.  Peer data is only read after this thread has passed a barrier whose
// flag loads are acquire.sys (causality orders it after the peers' writes); no address is read
// twice in one launch and L1 is invalidated at launch, so a stale line cannot be hit.
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

__device__ __forceinline__ unsigned* flag_ptr(char* base, int stage, int cta, int rank) {
  return reinterpret_cast<unsigned*>(base) + ((size_t)stage * kArMaxCtas + cta) * kArMaxWorld + rank;
}

// CTA-level barrier with the same CTA index on every rank
__device__ __forceinline__ void cta_barrier_all_ranks(const ArParams& p, int stage, unsigned epoch) {
  // every thread's data writes happen-before the barrier; the release.sys stores below are
  // cumulative over it, so no per-thread system fence is needed
  __syncthreads();
  const int tid = threadIdx.x;
  if (tid < p.world) st_release_sys(flag_ptr(p.bufs[tid], stage, blockIdx.x, p.rank), epoch);
  if (tid < p.world) {
    const unsigned* f = flag_ptr(p.bufs[p.rank], stage, blockIdx.x, tid);
    // bounded (2 s): a dead peer must not hang the GPU; the error word is sticky
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if ((++spins & 1023u) == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        if (t - t0 > 2000000000ull) {
          *reinterpret_cast<unsigned*>(p.bufs[p.rank] + kArFlagsBytes + 128) = 1u;
          break;
        }
      }
    }
  }
  __syncthreads();
}

// W = compile-time bound on the world size (buffers beyond p.world are skipped), U = unroll
template <int W, int U, int kArThreads>
__device__ __forceinline__ void reduce_slice(const ArParams& p, long long base, long long c_begin,
                                             long long c_end, float scale, float4* out) {
  const int tid = threadIdx.x;
  for (long long i0 = c_begin; i0 < c_end; i0 += (long long)U * kArThreads) {
    float4 v[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * kArThreads + tid;
#pragma unroll
      for (int r = 0; r < W; ++r)
        if (r < p.world && i < c_end)
          v[u][r] = ld_peer(reinterpret_cast<const float4*>(p.bufs[r] + kArHeaderBytes) + base + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * kArThreads + tid;
      if (i < c_end) {
        float4 acc = v[u][0];
#pragma unroll
        for (int r = 1; r < W; ++r)
          if (r < p.world) {
            acc.x += v[u][r].x;
            acc.y += v[u][r].y;
            acc.z += v[u][r].z;
            acc.w += v[u][r].w;
          }
        acc.x *= scale;
        acc.y *= scale;
        acc.z *= scale;
        acc.w *= scale;
        out[i] = acc;
      }
    }
  }
}

template <int kArThreads>
__global__ void __launch_bounds__(kArThreads, 1024 / kArThreads) allreduce_avg_kernel(const ArParams p) {
  NAFAE_CTA_TRACE(cta_trace, 5);
  char* mine = p.bufs[p.rank];
  unsigned* epoch_word = reinterpret_cast<unsigned*>(mine + kArFlagsBytes);
  int* finished = reinterpret_cast<int*>(mine + kArFlagsBytes + 64);
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(epoch_word) + 1u;
  const int tid = threadIdx.x;
  const long long n4 = p.count >> 2;  // float4 elements (count is padded to a multiple of 4*world)
  const long long slice4 = n4 / p.world;
  const long long per_cta = (slice4 + gridDim.x - 1) / gridDim.x;
  const long long c_begin = min(slice4, (long long)blockIdx.x * per_cta);
  const long long c_end = min(slice4, c_begin + per_cta);
  const float scale = 1.f / (float)p.world;

  cta_barrier_all_ranks(p, 0, epoch);

  // phase 1: reduce my slice (fixed rank order => identical bits on every replica).
  // ~8 x 16 B peer loads in flight per thread whatever the world size (peer latency ~2-3 us).
  {
    const long long base = (long long)p.rank * slice4;
    float4* out = reinterpret_cast<float4*>(mine + kArHeaderBytes) + base;
    if (p.world == 2)
      reduce_slice<2, 4, kArThreads>(p, base, c_begin, c_end, scale, out);
    else if (p.world <= 4)
      reduce_slice<4, 2, kArThreads>(p, base, c_begin, c_end, scale, out);
    else
      reduce_slice<8, 1, kArThreads>(p, base, c_begin, c_end, scale, out);
  }

  cta_barrier_all_ranks(p, 1, epoch);

  // phase 2: gather the other ranks' reduced slices
  for (int q = 1; q < p.world; ++q) {
    const int r = (p.rank + q) % p.world;  // start at different peers to spread the links
    const long long base = (long long)r * slice4;
    const float4* src = reinterpret_cast<const float4*>(p.bufs[r] + kArHeaderBytes) + base;
    float4* dst = reinterpret_cast<float4*>(mine + kArHeaderBytes) + base;
    for (long long i0 = c_begin; i0 < c_end; i0 += 8 * kArThreads) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + u * kArThreads + tid;
        if (i < c_end) v[u] = ld_peer(src + i);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long i = i0 + u * kArThreads + tid;
        if (i < c_end) dst[i] = v[u];
      }
    }
  }

  cta_barrier_all_ranks(p, 2, epoch);

  // the last CTA of this rank publishes the new epoch for the next launch
  __shared__ int s_ticket;
  if (tid == 0) s_ticket = atomicAdd(finished, 1);
  __syncthreads();
  if (s_ticket == (int)gridDim.x - 1 && tid == 0) {
    *finished = 0;
    *epoch_word = epoch;
  }
}


// ------------------------------------------------------------- bulk-copy (TMA) variant ----
// Same two-shot data flow, fused into ONE pass and driven by the bulk-copy engine instead of
// per-thread 16-byte loads:
//   barrier 0   every rank's bucket is complete
//   per chunk   cp.async.bulk the chunk of my slice from all `world` buffers into shared memory
//               (peer buffers over NVLink), add them in rank order, scale, and cp.async.bulk the
//               result from shared memory into EVERY rank's buffer (the all-gather is a push)
//   barrier 2   all pushes into my buffer have landed / nobody still reads my input
// A pull with plain loads needs (NVLink bandwidth x ~3 us latency) ~ 2 MB of loads in flight, i.e.
// registers and LSU slots of ~100 SMs; here ONE thread per CTA keeps ~90 KB of peer data in flight
// per SM (4-slot ring, the own copy goes through registers), so the few SMs left to the collective
// carry the traffic, one phase and one cross-GPU barrier disappear, and there is no gather pass
// re-reading the reduced slices.
constexpr int kTmaThreads = 256;
constexpr int kTmaRingBytes = 128 * 1024;

// V = 0: the configuration measured in round 1 (4-slot ring of 128 KB, one thread issues every bulk
//        copy).  At world 8 that is 4 KB chunks: 17 iterations x 15 bulk copies per CTA.
// V = 1: experiment for larger worlds (NAFAE_AR_VARIANT=1; not yet on hardware): 3-slot ring with
//        twice the chunk size at W = 8 / 4 (half the iterations and barriers per CTA), and lane q of
//        warp 0 issues the copies that involve peer (rank + q) % world, so the 15 copies of a chunk
//        leave in parallel instead of one after the other.
template <int W, int V>
struct TmaCfg {  // W = compile-time bound on the world size
  static constexpr int kStages = V == 0 ? 4 : 3;
  // per peer per stage, a multiple of 4 KB (16 bytes per thread per pass of the 256 threads)
  static constexpr int kChunkBytes =
      V == 0 ? kTmaRingBytes / (kStages * (W - 1)) / 4096 * 4096 : (W == 8 ? 8192 : W == 4 ? 16384 : 32768);
  static constexpr int kChunk4 = kChunkBytes / 16;
  static constexpr int kOwn = kChunk4 / kTmaThreads;  // float4 of the own copy per thread
  static constexpr size_t kRing = (size_t)kStages * (W - 1) * kChunkBytes;
  static constexpr size_t kSmem = kRing + 2 * (size_t)kChunkBytes;
  static_assert(kSmem <= 200 * 1024 && kOwn >= 1, "all-reduce ring does not fit");
};

template <int W, int V>
__global__ void __launch_bounds__(kTmaThreads, 1) allreduce_tma_kernel(const ArParams p) {
  NAFAE_CTA_TRACE(cta_trace, 5);
  using Cfg = TmaCfg<W, V>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunk4 = Cfg::kChunk4;
  constexpr int kOwn = Cfg::kOwn;
  extern __shared__ __align__(128) unsigned char ar_smem[];
  float4* in = reinterpret_cast<float4*>(ar_smem);                    // [stage][W-1][kChunk4]
  float4* out = reinterpret_cast<float4*>(ar_smem + Cfg::kRing);      // [2][kChunk4]
  __shared__ __align__(8) uint64_t full[kStages];

  char* mine = p.bufs[p.rank];
  unsigned* epoch_word = reinterpret_cast<unsigned*>(mine + kArFlagsBytes);
  int* finished = reinterpret_cast<int*>(mine + kArFlagsBytes + 64);
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(epoch_word) + 1u;
  const int tid = threadIdx.x, lane = tid & 31;
  const bool issuer = V == 0 ? tid == 0 : tid < p.world;  // threads that own bulk-copy groups
  const long long n4 = p.count >> 2;
  const long long slice4 = n4 / p.world;
  const long long per_cta = (slice4 + gridDim.x - 1) / gridDim.x;
  const long long c_begin = min(slice4, (long long)blockIdx.x * per_cta);
  const long long c_end = min(slice4, c_begin + per_cta);
  const long long base4 = (long long)p.rank * slice4;  // my slice inside every buffer
  const float scale = 1.f / (float)p.world;
  const int nchunks = (int)((c_end - c_begin + kChunk4 - 1) / kChunk4);
  const float4* own = reinterpret_cast<const float4*>(mine + kArHeaderBytes) + base4 + c_begin;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  cta_barrier_all_ranks(p, 0, epoch);  // includes __syncthreads on both sides
  cta_trace.mark_a();

  auto chunk_len4 = [&](int i) { return (int)min((long long)kChunk4, c_end - (c_begin + (long long)i * kChunk4)); };
  // the peers' copies of chunk i -> ring slot i % stages.  V = 0: thread 0 alone; V = 1: all of warp 0
  // calls this (lane 0 arms the barrier, lane q >= 1 pulls from peer (rank + q) % world)
  auto issue_loads = [&](int i) {
    const int s = i % kStages;
    const uint32_t bytes = (uint32_t)chunk_len4(i) * 16u;
    const long long off4 = base4 + c_begin + (long long)i * kChunk4;
    auto pull = [&](int q) {
      const int r = (p.rank + q) % p.world;  // start at different peers to spread the links
      const int slot = r < p.rank ? r : r - 1;
      bulk_g2s(in + ((size_t)s * (W - 1) + slot) * kChunk4,
               reinterpret_cast<const float4*>(p.bufs[r] + kArHeaderBytes) + off4, bytes, &full[s]);
    };
    if (V == 0) {
      mbar_arrive_expect_tx(&full[s], bytes * (uint32_t)(p.world - 1));
      for (int q = 1; q < p.world; ++q) pull(q);
    } else {
      if (lane == 0) mbar_arrive_expect_tx(&full[s], bytes * (uint32_t)(p.world - 1));
      __syncwarp();
      if (lane >= 1 && lane < p.world) pull(lane);
    }
  };
  const bool loader = V == 0 ? tid == 0 : tid < 32;  // who calls issue_loads (warp-uniform for V = 1)
  if (loader) {
    asm volatile("fence.proxy.async;" ::: "memory");
    for (int i = 0; i < kStages - 1 && i < nchunks; ++i) issue_loads(i);
  }
  for (int i = 0; i < nchunks; ++i) {
    const int s = i % kStages, so = i & 1;
    const int len4 = chunk_len4(i);
    // slot of chunk i-1 is free since the closing __syncthreads of the previous iteration
    if (loader && i + kStages - 1 < nchunks) issue_loads(i + kStages - 1);
    // own copy -> registers, in flight while the peers' data arrives
    float4 mine4[kOwn];
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
      const int k = tid + j * kTmaThreads;
      mine4[j] = k < len4 ? own[(size_t)i * kChunk4 + k] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mbar_wait(&full[s], (uint32_t)(i / kStages) & 1u);
    // the bulk stores that read out[so] two chunks ago must have drained it
    if (issuer && i >= 2) bulk_wait_read<1>();
    __syncthreads();
    const float4* src = in + (size_t)s * (W - 1) * kChunk4;
    float4* dst = out + (size_t)so * kChunk4;
#pragma unroll
    for (int j = 0; j < kOwn; ++j) {
      const int k = tid + j * kTmaThreads;
      if (k < len4) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < W; ++r)  // rank order; only the slice owner adds: replicas get its bits
          if (r < p.world) {
            const float4 v = r == p.rank ? mine4[j] : src[(size_t)(r < p.rank ? r : r - 1) * kChunk4 + k];
            if (r == 0) {
              acc = v;
            } else {
              acc.x += v.x;
              acc.y += v.y;
              acc.z += v.z;
              acc.w += v.w;
            }
          }
        acc.x *= scale;
        acc.y *= scale;
        acc.z *= scale;
        acc.w *= scale;
        dst[k] = acc;
      }
    }
    fence_proxy_async_smem();  // generic-proxy writes of out[so] -> visible to the bulk-copy engine
    __syncthreads();           // ... and everybody is done reading ring slot s
    if (issuer) {
      const long long off4 = base4 + c_begin + (long long)i * kChunk4;
      if (V == 0) {
        for (int q = 0; q < p.world; ++q) {
          const int r = (p.rank + q) % p.world;  // own copy first, then spread over the links
          bulk_s2g(reinterpret_cast<float4*>(p.bufs[r] + kArHeaderBytes) + off4, dst, (uint32_t)len4 * 16u);
        }
      } else {  // thread q pushes to rank (rank + q) % world; every issuer owns its own bulk groups
        const int r = (p.rank + tid) % p.world;
        bulk_s2g(reinterpret_cast<float4*>(p.bufs[r] + kArHeaderBytes) + off4, dst, (uint32_t)len4 * 16u);
      }
      bulk_commit();
    }
  }
  cta_trace.mark_b();
  if (issuer) {
    bulk_wait_all<0>();  // every push has been performed
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  cta_barrier_all_ranks(p, 2, epoch);

  __shared__ int s_ticket;
  if (tid == 0) s_ticket = atomicAdd(finished, 1);
  __syncthreads();
  if (s_ticket == (int)gridDim.x - 1 && tid == 0) {
    *finished = 0;
    *epoch_word = epoch;
  }
}

// ------------------------------------------------------------- NVLS (multimem) variant ----
// The bucket lives in memory bound to an NVSwitch multicast object (csrc/symm.cu): `mc` is the
// multicast address of the same offsets `uc` addresses in this rank's own copy.
//   barrier 0   every rank's bucket is complete (multimem.red +1 on a counter replicated into
//               every rank's header, each rank spins on its LOCAL copy)
//   data        rank r owns slice r: multimem.ld_reduce returns the sum of all ranks' copies,
//               added INSIDE the switch; scale by 1/world; multimem.st replicates the result into
//               every rank's bucket.  One reducer per element => replicas are bit-identical.
//   barrier 2   every slice has landed everywhere / nobody still reads my input
// Per rank the SMs move count/world floats in and out (1.1 MB each way for the 8.8 MB bucket at
// world 8) instead of pulling and pushing 7/8 of the bucket, so a handful of small CTAs saturate
// the collective and the ring of shared-memory staging buffers disappears.
// Barrier counters are monotonic (launch e waits for e*world arrivals): nothing to reset, and a
// rank that runs ahead can only over-satisfy a slower rank's wait.  Spins are bounded (2 s): a
// dead peer sets the error word instead of hanging the GPU.
struct McParams {
  char* uc;  // this rank's own copy (header + bucket)
  char* mc;  // multicast view of the same offsets on every rank
  int rank, world;
  long long count;  // floats
};

__device__ __forceinline__ float4 mm_ld_reduce_f4(const void* mc_addr) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc_addr)
               : "memory");
  return v;
}
__device__ __forceinline__ void mm_st_f4(void* mc_addr, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mm_red_release_add(void* mc_addr, unsigned v) {
  asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ar_global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) allreduce_mc_kernel(const McParams p) {
  NAFAE_CTA_TRACE(cta_trace, 5);
  constexpr int kU = 8;  // 16-byte multimem loads in flight per thread
  unsigned* epoch_word = reinterpret_cast<unsigned*>(p.uc + kArFlagsBytes);
  int* finished = reinterpret_cast<int*>(p.uc + kArFlagsBytes + 64);
  unsigned* err_word = reinterpret_cast<unsigned*>(p.uc + kArFlagsBytes + 128);
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(epoch_word) + 1u;
  const unsigned target = epoch * (unsigned)p.world;
  const int tid = threadIdx.x;
  const long long n4 = p.count >> 2;
  const long long slice4 = n4 / p.world;
  const long long per_cta = (slice4 + gridDim.x - 1) / gridDim.x;
  const long long c_begin = min(slice4, (long long)blockIdx.x * per_cta);
  const long long c_end = min(slice4, c_begin + per_cta);
  const float scale = 1.f / (float)p.world;

  auto barrier_all_ranks = [&](int stage) {
    __syncthreads();
    if (tid == 0) {
      const size_t off = ((size_t)stage * kArMaxCtas + blockIdx.x) * sizeof(unsigned);
      mm_red_release_add(p.mc + off, 1u);
      const unsigned* f = reinterpret_cast<const unsigned*>(p.uc + off);
      const unsigned long long t0 = ar_global_ns();
      while ((int)(ld_acquire_sys(f) - target) < 0) {
        if (ar_global_ns() - t0 > 2000000000ull) {
          *err_word = 1u;
          break;
        }
      }
    }
    __syncthreads();
  };

  barrier_all_ranks(0);
  cta_trace.mark_a();

  {
    char* data = p.mc + kArHeaderBytes + ((size_t)p.rank * (size_t)slice4) * sizeof(float4);
    for (long long i0 = c_begin; i0 < c_end; i0 += (long long)kU * kThreads) {
      float4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long long i = i0 + u * kThreads + tid;
        if (i < c_end) v[u] = mm_ld_reduce_f4(data + (size_t)i * sizeof(float4));
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long long i = i0 + u * kThreads + tid;
        if (i < c_end) {
          v[u].x *= scale;
          v[u].y *= scale;
          v[u].z *= scale;
          v[u].w *= scale;
          mm_st_f4(data + (size_t)i * sizeof(float4), v[u]);
        }
      }
    }
  }
  cta_trace.mark_b();
  __threadfence_system();  // my multicast stores are performed everywhere before I arrive
  barrier_all_ranks(2);

  __shared__ int s_ticket;
  if (tid == 0) s_ticket = atomicAdd(finished, 1);
  __syncthreads();
  if (s_ticket == (int)gridDim.x - 1 && tid == 0) {
    *finished = 0;
    *epoch_word = epoch;
  }
}

template <int W, int V>
int launch_tma(const ArParams& p, int num_ctas, cudaStream_t stream) {
  const size_t smem = TmaCfg<W, V>::kSmem;
  auto kern = allreduce_tma_kernel<W, V>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("allreduce: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  kern<<<num_ctas, kTmaThreads, smem, stream>>>(p);
  return launch_status("allreduce_tma_kernel");
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_CTA_TRACE_READER(nafae_debug_cta_trace_allreduce)

NAFAE_API size_t nafae_ar_buffer_bytes(size_t count_floats, int world) {
  if (world < 1) world = 1;
  const size_t pad = (size_t)4 * world;
  const size_t padded = (count_floats + pad - 1) / pad * pad;
  return kArHeaderBytes + padded * sizeof(float);
}

NAFAE_API size_t nafae_ar_data_offset(void) { return kArHeaderBytes; }

NAFAE_API int nafae_ar_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64) {
  NAFAE_REQUIRE(dev_ptr && handle64 && bytes >= kArHeaderBytes, "ar_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("ar_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return -(int)e;
  }
  cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("ar_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    cudaFree(p);
    return -(int)e;
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  cudaDeviceSynchronize();
  return 1;
}

NAFAE_API int nafae_ar_open(const unsigned char* handle64, void** peer_ptr) {
  NAFAE_REQUIRE(handle64 && peer_ptr, "ar_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("ar_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  return 1;
}

NAFAE_API int nafae_ar_close(void* peer_ptr) {
  return cudaIpcCloseMemHandle(peer_ptr) == cudaSuccess ? 1 : 0;
}

NAFAE_API int nafae_ar_free(void* dev_ptr) { return cudaFree(dev_ptr) == cudaSuccess ? 1 : 0; }

NAFAE_API int nafae_allreduce_avg(void* const* bufs, int rank, int world, size_t count_floats,
                                  int num_ctas, int cta_threads, unsigned flags, cudaStream_t stream) {
  NAFAE_REQUIRE(cta_threads == 0 || cta_threads == 128 || cta_threads == 256,
                "allreduce: cta_threads must be 0 (bulk-copy kernel), 128 or 256");
  NAFAE_REQUIRE(bufs && world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world,
                "allreduce: bad rank/world (world <= %d)", kArMaxWorld);
  NAFAE_REQUIRE(num_ctas >= 1 && num_ctas <= kArMaxCtas, "allreduce: num_ctas must be in [1, %d]",
                kArMaxCtas);
  NAFAE_REQUIRE(count_floats % ((size_t)4 * world) == 0,
                "allreduce: count must be a multiple of 4*world (use nafae_ar_buffer_bytes)");
  if (world == 1 || count_floats == 0) return 1;
  ArParams p;
  for (int r = 0; r < kArMaxWorld; ++r) p.bufs[r] = r < world ? static_cast<char*>(bufs[r]) : nullptr;
  for (int r = 0; r < world; ++r) NAFAE_REQUIRE(p.bufs[r], "allreduce: NULL buffer for rank %d", r);
  p.rank = rank;
  p.world = world;
  p.count = (long long)count_floats;
  static_assert(kArThreadsMax == 256, "flag layout");
  if (cta_threads == 0) {
    // template bound on the world size: the smallest instantiation that holds `world` (larger
    // chunks per peer); NAFAE_AR_WIDTH(w) in `flags` forces a wider one (tests run <4> and <8>
    // on a 2-GPU box that way)
    int width = world == 2 ? 2 : (world <= 4 ? 4 : 8);
    const int forced = (int)((flags >> 8) & 0xffu);
    if (forced != 0) {
      NAFAE_REQUIRE((forced == 2 || forced == 4 || forced == 8) && forced >= world,
                    "allreduce: forced width %d invalid for world %d", forced, world);
      width = forced;
    }
    const int variant = (int)(flags & 0xfu);
    NAFAE_REQUIRE(variant == 0 || variant == 1, "allreduce: unknown variant %d", variant);
    if (variant == 1) {  // 3-slot ring, larger chunks, parallel issue (see TmaCfg)
      if (width == 2) return launch_tma<2, 1>(p, num_ctas, stream);
      if (width == 4) return launch_tma<4, 1>(p, num_ctas, stream);
      return launch_tma<8, 1>(p, num_ctas, stream);
    }
    if (width == 2) return launch_tma<2, 0>(p, num_ctas, stream);
    if (width == 4) return launch_tma<4, 0>(p, num_ctas, stream);
    return launch_tma<8, 0>(p, num_ctas, stream);
  }
  if (cta_threads == 128)
    allreduce_avg_kernel<128><<<num_ctas, 128, 0, stream>>>(p);
  else
    allreduce_avg_kernel<256><<<num_ctas, 256, 0, stream>>>(p);
  return launch_status("allreduce_avg_kernel");
}

// ------------------------------------------------------------- NVLS (multimem) variant ----
NAFAE_API size_t nafae_mc_buffer_bytes(size_t count_floats, int world) {
  return nafae_ar_buffer_bytes(count_floats, world);
}

NAFAE_API int nafae_allreduce_mc(void* uc_base, void* mc_base, int rank, int world,
                                 size_t count_floats, int num_ctas, int cta_threads,
                                 cudaStream_t stream) {
  NAFAE_REQUIRE(uc_base && mc_base, "allreduce_mc: NULL buffer");
  NAFAE_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "allreduce_mc: bad rank/world");
  NAFAE_REQUIRE(num_ctas >= 1 && num_ctas <= kArMaxCtas, "allreduce_mc: num_ctas must be in [1, %d]",
                kArMaxCtas);
  NAFAE_REQUIRE(count_floats % ((size_t)4 * world) == 0,
                "allreduce_mc: count must be a multiple of 4*world (use nafae_mc_buffer_bytes)");
  if (world == 1 || count_floats == 0) return 1;
  McParams p;
  p.uc = static_cast<char*>(uc_base);
  p.mc = static_cast<char*>(mc_base);
  p.rank = rank;
  p.world = world;
  p.count = (long long)count_floats;
  if (cta_threads == 0) cta_threads = 512;
  if (cta_threads == 256)
    allreduce_mc_kernel<256><<<num_ctas, 256, 0, stream>>>(p);
  else if (cta_threads == 512)
    allreduce_mc_kernel<512><<<num_ctas, 512, 0, stream>>>(p);
  else if (cta_threads == 1024)
    allreduce_mc_kernel<1024><<<num_ctas, 1024, 0, stream>>>(p);
  else
    NAFAE_REQUIRE(false, "allreduce_mc: cta_threads must be 0, 256, 512 or 1024");
  return launch_status("allreduce_mc_kernel");
}

// Host-synchronising check (tests, teardown): 0 = no cross-GPU barrier of this buffer ever timed
// out, 1 = one did (the results of that launch are undefined), <0 = CUDA error.
NAFAE_API int nafae_allreduce_mc_error(void* uc_base) {
  NAFAE_REQUIRE(uc_base, "allreduce_mc_error: NULL buffer");
  unsigned v = 0;
  cudaError_t e = cudaMemcpy(&v, static_cast<char*>(uc_base) + kArFlagsBytes + 128, sizeof(v),
                             cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    set_error("allreduce_mc_error: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  return v ? 1 : 0;
}
