// Fused region x query similarity, masks, max-over-boxes picks, frame-weighted ranking loss and
// visual clustering loss -- forward and backward -- for sm_100a.
//
// Replaces reference model.py:517-614 (DVSA.forward: ~60 un-fused ATen ops, numpy mask build +
// H2D per step, Na*Ns*Ne scalar device writes in Python loops, a .nonzero() host sync) and the
// backward autograd derives from it (model.py:772), plus postprocess (model.py:457-474).
//
// Forward = ONE kernel, three phases chained by "last CTA to arrive continues" counters (no
// cooperative launch, no host sync, graph-capturable):
//   P1  grid (frames, column chunks): S_ = vis @ word^T tile for one frame, column mask, max /
//       first-argmax over the Nb boxes -> D_sim, D_ind.  The (R x Na*Ne) similarity matrix is
//       never written to HBM; masked columns are not computed at all.
//   P2  (last P1 CTA of a segment) per-column min-max frame attention over the segment's frames,
//       S*S_att summed over entities -> Sf[a,:,:]; clustering loss partials of the segment.
//   P3  (last P2) hinge terms, frame_score, margin_loss; dL/dSf saved for the backward.
// Backward = ONE kernel with a CTA per frame (dL/dvis rows, dense overwrite), a CTA per query
// column (dL/dword) and (train) a CTA per entity for the clustering gradient: dL/dS_ has at most
// one non-zero per (frame, column) -- the argmax box -- so all are gather-scale-accumulate sweeps,
// not GEMMs.
//
// Arithmetic: fp32 FMA, fp32 accumulate.  The contraction is 85 MFLOP at the benchmark shape and
// latency bound; see DESIGN.md for why it runs on the FMA pipe rather than tcgen05 tiles.
#include "tc05.cuh"

namespace nafae {
namespace {

constexpr float kEps = 1e-5f;  // model.py:33

#ifdef NAFAE_TRACE
__device__ unsigned long long g_trace[256];
__device__ __forceinline__ void trace(int slot) {
  // SM-local cycle counter: cheap (globaltimer reads cost ~0.5 us each and distort the trace);
  // only differences taken on the same CTA are meaningful
  g_trace[slot] = (unsigned long long)clock64();
}
#define TRACE(slot) do { if (threadIdx.x == 0) trace(slot); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif
constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kColsPerCta = 8;   // live columns per CTA: 2 column groups x 4 row groups of warps
constexpr int kLiveMax = 2048;   // Na * Ne upper bound for the compacted column list
constexpr int kRowTile = 32;     // boxes staged in shared memory at a time
constexpr int kKS = 16;          // k-slices per lane per chunk: 512 features per chunk

struct Dims {
  int Na, Ns, Nb, Ne, D;
  int F, NQ;
};

// workspace layout (4-byte words)
struct Ws {
  int* seg_cnt;     // [Na]   arrival counters (zero between launches)
  int* done_cnt;    // [1]
  int* vcnt;        // [Na]   nonzero count of the segment's Gram blocks
  float* vsum;      // [Na]   sum of the segment's Gram blocks
  float* scal;      // [8]    0: vis_loss, 1: dem (as float), 2: mean frame_score
  float* Sf;        // [Na*Ns*Na]
  float* hgrad;     // [Na*Ns*Na]  d(margin_loss)/dSf
  float* clus;      // [Nb*D]      clustering-loss gradient accumulators (zero between launches)
};

__host__ __device__ inline size_t ws_words(const Dims& d) {
  return (size_t)64 + 2 * (size_t)d.Na /*seg_cnt,vcnt*/ + d.Na /*vsum*/ + 8 +
         2 * (size_t)d.Na * d.Ns * d.Na + (size_t)d.Nb * d.D + 64;
}

__host__ __device__ inline Ws ws_carve(void* base, const Dims& d) {
  Ws w;
  int* p = static_cast<int*>(base);
  w.done_cnt = p;
  p += 16;
  w.seg_cnt = p;
  p += d.Na;
  w.vcnt = p;
  p += d.Na;
  float* q = reinterpret_cast<float*>(p);
  w.vsum = q;
  q += d.Na;
  w.scal = q;
  q += 8;
  w.Sf = q;
  q += (size_t)d.Na * d.Ns * d.Na;
  w.hgrad = q;
  q += (size_t)d.Na * d.Ns * d.Na;
  // 16-byte align the accumulator block
  uintptr_t u = reinterpret_cast<uintptr_t>(q);
  u = (u + 15) & ~(uintptr_t)15;
  w.clus = reinterpret_cast<float*>(u);
  return w;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Warp-cooperative dot product of two global rows with the loads issued in independent batches
// of 8 per lane (a plain k-loop with a runtime trip count serialises on L2 latency).
// x / y may point to global OR shared memory (generic loads).
__device__ __forceinline__ float warp_dot(const float* x, const float* y, int D, int lane) {
  float acc = 0.f;
  for (int k0 = 0; k0 < D; k0 += 256) {
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + lane + 32 * j;
      a[j] = k < D ? x[k] : 0.f;
      b[j] = k < D ? y[k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(a[j], b[j], acc);
  }
  return warp_sum(acc);
}

// per-column frame statistics over the Ns frames of one segment (model.py:583-588)
struct ColStat {
  float mn, mx, den;
  int s_mn, s_mx;  // first index of the minimum / maximum (torch.min/max(dim) tie rule)
};

// ------------------------------------------------------------------------ forward ----
struct FwdParams {
  const float* vis;
  const float* word;
  const int* lens;
  long long* D_ind;
  float* D_sim;
  float* loss;
  void* ws;
  Dims d;
  float Delta, vis_lam;
  int train;
  int col_chunks;
  int groups;            // independent batches (gridDim.y); each has its own inputs, outputs, workspace
  size_t ws_group_bytes;
};

// 16 per-lane partial sums -> lanes 2*i and 2*i+1 both hold the warp total of value i.
// 16 shuffles, dependency depth 5 (a per-value butterfly would be 80 shuffles, depth 5 each).
__device__ __forceinline__ float reduce16(const float (&v)[16], int lane) {
  float k8[8], k4[4], k2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h16 ? v[8 + i] : v[i], send = h16 ? v[i] : v[8 + i];
    k8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h8 ? k8[4 + i] : k8[i], send = h8 ? k8[i] : k8[4 + i];
    k4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h4 ? k4[2 + i] : k4[i], send = h4 ? k4[i] : k4[2 + i];
    k2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float keep = h2 ? k2[1] : k2[0], send = h2 ? k2[0] : k2[1];
  float k1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  k1 += __shfl_xor_sync(0xffffffffu, k1, 1);
  return k1;  // value index = lane >> 1
}

__device__ void phase2_segment(const FwdParams& p, const Ws& w, int a);
__device__ void phase3_final(const FwdParams& p, const Ws& w);

__global__ void __launch_bounds__(kFwdThreads, 1) ground_fwd_kernel(const FwdParams p_in) {
  NAFAE_CTA_TRACE(cta_trace, 3);
  // independent groups (evaluation sweep: Na = 1 per segment, many segments per launch): the group
  // index only offsets the pointers, everything below is unchanged
  FwdParams p = p_in;
  if (p_in.groups > 1) {
    const size_t g = blockIdx.y;
    p.vis += g * (size_t)p_in.d.F * p_in.d.Nb * p_in.d.D;
    p.word += g * (size_t)p_in.d.NQ * p_in.d.D;
    p.lens += g * (size_t)p_in.d.Na;
    p.D_ind += g * (size_t)p_in.d.F * p_in.d.NQ;
    p.D_sim += g * (size_t)p_in.d.F * p_in.d.NQ;
    p.loss += g;
    p.ws = static_cast<char*>(p_in.ws) + g * p_in.ws_group_bytes;
  }
  extern __shared__ __align__(16) float sm[];  // kRowTile * D floats (vis rows of the frame)
  __shared__ int s_ticket;
  __shared__ int s_live[kLiveMax];       // compacted list of live (unmasked) columns
  __shared__ int s_nlive;
  __shared__ float s_best[kColsPerCta][16];
  __shared__ int s_besti[kColsPerCta][16];
  __shared__ __align__(8) uint64_t s_bar;
  const Dims& d = p.d;
  const Ws w = ws_carve(p.ws, d);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.x / p.col_chunks, chunk = blockIdx.x % p.col_chunks;
  const int a = f / d.Ns;
  if (blockIdx.x == 0) TRACE(0);

  // ---- P1: one frame x 8 LIVE columns per CTA pass.  Masked columns (e >= len[a'],
  // model.py:535-536) are never computed: the column list is compacted first.
  // warp = (column group of 4: warp & 1) x (row group: warp >> 1).
  const float* vis_f = p.vis + (size_t)f * d.Nb * d.D;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
    // first row tile: one bulk async copy (TMA engine), in flight while the columns are compacted
    const uint32_t bytes = (uint32_t)min(kRowTile, d.Nb) * d.D * 4u;
    mbar_arrive_expect_tx(&s_bar, bytes);
    bulk_g2s(sm, vis_f, bytes, &s_bar);
  }
  if (warp == 0) {  // ordered compaction of the live columns
    int n = 0;
    for (int c0 = 0; c0 < d.NQ; c0 += 32) {
      const int c = c0 + lane;
      const bool lv = c < d.NQ && (c % d.Ne) < __ldg(p.lens + c / d.Ne);
      const unsigned bal = __ballot_sync(0xffffffffu, lv);
      const int pos = n + __popc(bal & ((1u << lane) - 1u));
      if (lv && pos < kLiveMax) s_live[pos] = c;
      n += __popc(bal);
    }
    if (lane == 0) s_nlive = n;
  }
  __syncthreads();
  const int nlive = s_nlive;

  if (chunk == 0) {
    // masked column: every entry is 0 after masked_fill_ -> max 0 at index 0 (model.py:551,612)
    for (int c = tid; c < d.NQ; c += kFwdThreads)
      if ((c % d.Ne) >= __ldg(p.lens + c / d.Ne)) {
        p.D_sim[(size_t)f * d.NQ + c] = 0.f;
        p.D_ind[(size_t)f * d.NQ + c] = 0ll;
      }
  }

  // this CTA serves live-column chunks chunk, chunk + col_chunks, ... of frame f
  {
    const int cgp = warp & 1, rg = warp >> 1;
    uint32_t parity = 0;
    int tile_in_smem = 0;  // row tile currently staged (the prologue loaded tile 0)
    bool pending = true;   // a bulk copy has been issued and not yet waited for
    for (int first = chunk * kColsPerCta; first < nlive; first += p.col_chunks * kColsPerCta) {
      int col[4];
      bool live[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int li = first + cgp * 4 + j;
        live[j] = li < nlive;
        col[j] = s_live[min(li, nlive - 1)];
      }
      // best (value, row) of this lane's (row slot, column) = (lane>>3, (lane>>1)&3)
      float best = -INFINITY;
      int best_r = 0x7fffffff;
      for (int r0 = 0; r0 < d.Nb; r0 += kRowTile) {
        const int rows = min(kRowTile, d.Nb - r0);
        if (tile_in_smem != r0) {
          __syncthreads();  // previous tile consumed
          if (tid == 0) {
            const uint32_t bytes = (uint32_t)rows * d.D * 4u;
            mbar_arrive_expect_tx(&s_bar, bytes);
            bulk_g2s(sm, vis_f + (size_t)r0 * d.D, bytes, &s_bar);
          }
          tile_in_smem = r0;
          pending = true;
        }
        // this warp's rows inside the tile: contiguous block of <= 8 rows, two batches of 4
        const int per = (rows + 3) >> 2;
        const int rb = rg * per, re = min(rows, rb + per);
        float acc[2][16];
#pragma unroll
        for (int bt = 0; bt < 2; ++bt)
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[bt][i] = 0.f;
        for (int k0 = 0; k0 < d.D; k0 += 32 * kKS) {
          // this chunk's word slices for the warp's 4 columns (overlaps the tile copy)
          float wv[4][kKS];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* wr = p.word + (size_t)col[j] * d.D + k0;
#pragma unroll
            for (int i = 0; i < kKS; ++i) {
              const int k = lane + 32 * i;
              wv[j][i] = (live[j] && k0 + k < d.D) ? __ldg(wr + k) : 0.f;
            }
          }
          if (pending) {
            mbar_wait(&s_bar, parity);
            parity ^= 1u;
            pending = false;
          }
          if (!live[0]) break;
#pragma unroll
          for (int bt = 0; bt < 2; ++bt) {
            if (rb + 4 * bt >= re) break;
            const float* v0 = sm + (size_t)min(rb + 4 * bt + 0, rows - 1) * d.D + k0;
            const float* v1 = sm + (size_t)min(rb + 4 * bt + 1, rows - 1) * d.D + k0;
            const float* v2 = sm + (size_t)min(rb + 4 * bt + 2, rows - 1) * d.D + k0;
            const float* v3 = sm + (size_t)min(rb + 4 * bt + 3, rows - 1) * d.D + k0;
#pragma unroll
            for (int i = 0; i < kKS; ++i) {
              const int k = lane + 32 * i;
              const bool ok = k0 + k < d.D;
              const float x0 = ok ? v0[k] : 0.f, x1 = ok ? v1[k] : 0.f;
              const float x2 = ok ? v2[k] : 0.f, x3 = ok ? v3[k] : 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc[bt][0 + j] = fmaf(x0, wv[j][i], acc[bt][0 + j]);
                acc[bt][4 + j] = fmaf(x1, wv[j][i], acc[bt][4 + j]);
                acc[bt][8 + j] = fmaf(x2, wv[j][i], acc[bt][8 + j]);
                acc[bt][12 + j] = fmaf(x3, wv[j][i], acc[bt][12 + j]);
              }
            }
          }
        }
        if (live[0]) {
#pragma unroll
          for (int bt = 0; bt < 2; ++bt) {
            if (rb + 4 * bt >= re) break;
            const float tot = reduce16(acc[bt], lane);
            const int r = rb + 4 * bt + (lane >> 3);  // row slot = value index >> 2
            // ascending rows per lane: strict '>' keeps the first maximal box (torch.max(dim))
            if (r < re && tot > best) {
              best = tot;
              best_r = r0 + r;
            }
          }
        }
      }
      // combine the 4 row groups x 4 row slots of every column: max value, ties -> lowest row
      if ((lane & 1) == 0) {
        const int j = (lane >> 1) & 3, slot = lane >> 3;
        s_best[cgp * 4 + j][rg * 4 + slot] = best;
        s_besti[cgp * 4 + j][rg * 4 + slot] = best_r;
      }
      __syncthreads();
      if (tid < kColsPerCta && first + tid < nlive) {
        float bv = s_best[tid][0];
        int bi = s_besti[tid][0];
#pragma unroll
        for (int g = 1; g < 16; ++g) {
          const float v = s_best[tid][g];
          const int vi = s_besti[tid][g];
          if (v > bv || (v == bv && vi < bi)) {
            bv = v;
            bi = vi;
          }
        }
        const int c = s_live[first + tid];
        p.D_sim[(size_t)f * d.NQ + c] = bv;
        p.D_ind[(size_t)f * d.NQ + c] = (long long)bi;
      }
      __syncthreads();  // s_best is reused by the next chunk
    }
    if (pending) mbar_wait(&s_bar, parity);  // never leave a bulk copy in flight
  }

  // ---- chain: last CTA of the segment runs P2, last segment runs P3
  if (blockIdx.x == 0) TRACE(1);
  __syncthreads();
  if (tid == 0) s_ticket = ticket_acq_rel(w.seg_cnt + a);
  __syncthreads();
  if (blockIdx.x == 0) TRACE(2);
  if (s_ticket != d.Ns * p.col_chunks - 1) return;
  TRACE(16 + a * 8 + 0);
  phase2_segment(p, w, a);
  TRACE(16 + a * 8 + 6);
  __syncthreads();
  if (tid == 0) {
    w.seg_cnt[a] = 0;  // leave the counter clean for the next launch
    s_ticket = ticket_acq_rel(w.done_cnt);
  }
  __syncthreads();
  TRACE(16 + a * 8 + 7);
  if (s_ticket != d.Na - 1) return;
  TRACE(120);
  phase3_final(p, w);
  TRACE(125);
  if (tid == 0) *w.done_cnt = 0;
}

// P2 / P3 alone (one CTA per segment): the tail of the forward when P1 ran as the tensor-core kernel
__global__ void __launch_bounds__(kFwdThreads, 1) ground_p23_kernel(const FwdParams p) {
  NAFAE_CTA_TRACE(cta_trace, 3);
  __shared__ int s_ticket;
  const Dims& d = p.d;
  const Ws w = ws_carve(p.ws, d);
  const int a = blockIdx.x;
  phase2_segment(p, w, a);
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = ticket_acq_rel(w.done_cnt);
  __syncthreads();
  if (s_ticket != d.Na - 1) return;
  phase3_final(p, w);
  if (threadIdx.x == 0) *w.done_cnt = 0;
}

// ---------------------------------------------------- P1 on the tensor cores (tcgen05) ----
// S_ = vis_feats @ word_feats^T (model.py:548) as 128-row tiles of tcgen05.mma, fp32 inputs fed as
// THREE tf32 products per K step (x = hi + lo with hi = x truncated to tf32, lo = x - hi exactly):
//     S += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo          (the dropped lo.lo term is ~2^-22 relative)
// so the accumulator agrees with the fp32 FMA path to ~1e-6 relative instead of tf32's 1e-3, and the
// argmax over boxes is re-checked in exact fp32 whenever the two best candidates are closer than
// kTcTieTol -- D_ind stays the fp32 pick.
//   grid  (row tiles, column tiles): a row tile = floor(128 / Nb) whole frames (so every frame's max
//         is tile-local), a column tile = up to 128 query columns (TMEM columns of one accumulator)
//   warp 0      TMA producer: 128 x 32 and NQt x 32 fp32 boxes (128-byte swizzle), 3-stage ring
//   warps 2-9   split hi / lo in place (element-wise: layout agnostic), then warps 2-5 read the
//               accumulator from TMEM into a shared S tile and all of them reduce per (frame, column)
//   warp 1      TMEM allocation + MMA issue (12 tcgen05.mma per stage), tcgen05.commit
constexpr int kTcThreads = 320;
constexpr int kTcSplitThreads = 256;
constexpr int kTcStages = 3;
constexpr int kTcBK = 32;             // fp32 elements per stage row = 128 bytes
constexpr float kTcTieTol = 6e-5f;    // absolute gap below which the fp32 re-check decides

struct P1TcParams {
  const float* vis;
  const float* word;
  const int* lens;
  long long* D_ind;
  float* D_sim;
  Dims d;
  int fpt;    // frames per row tile
  int nqt;    // columns per column tile (multiple of 16, <= 128)
};

__device__ __forceinline__ float exact_dot(const float* __restrict__ x, const float* __restrict__ y, int D) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k = 0; k < D; k += 4) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + k));
    const float4 yv = __ldg(reinterpret_cast<const float4*>(y + k));
    a0 = fmaf(xv.x, yv.x, a0);
    a1 = fmaf(xv.y, yv.y, a1);
    a2 = fmaf(xv.z, yv.z, a2);
    a3 = fmaf(xv.w, yv.w, a3);
  }
  return (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(kTcThreads, 1)
ground_p1_tc_kernel(const __grid_constant__ CUtensorMap map_vis, const __grid_constant__ CUtensorMap map_word,
                    const P1TcParams p) {
  NAFAE_CTA_TRACE(cta_trace, 3);
  extern __shared__ unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  const Dims& d = p.d;
  const int a_bytes = 128 * 128, b_bytes = p.nqt * 128;
  const int stage_bytes = 2 * (a_bytes + b_bytes);  // [A hi | A lo | B hi | B lo]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)kTcStages * stage_bytes);
  uint64_t* ready = full + kTcStages;
  uint64_t* empty = ready + kTcStages;
  uint64_t* acc_ready = empty + kTcStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);
  float* S = reinterpret_cast<float*>(smem);  // epilogue: [128][nqt + 1], reuses the ring

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f0 = blockIdx.x * p.fpt;                       // first frame of the tile
  const int nf = min(p.fpt, d.F - f0);                     // frames in the tile
  const int m0 = f0 * d.Nb;                                // first vis row
  const int c0 = blockIdx.y * p.nqt;                       // first query column
  const int nc = min(p.nqt, d.NQ - c0);
  const int nkb = (d.D + kTcBK - 1) / kTcBK;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < p.nqt) tmem_cols <<= 1;

  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], kTcSplitThreads);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_ready, 1);
    fence_mbar_init();
    tc05::tma_prefetch_desc(&map_vis);
    tc05::tma_prefetch_desc(&map_word);
  }
  if (warp == 1) tc05::tmem_alloc(tmem_slot, tmem_cols);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kTcStages;
        mbar_wait(&empty[s], ((uint32_t)(kb / kTcStages) & 1u) ^ 1u);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        mbar_arrive_expect_tx(&full[s], (uint32_t)(a_bytes + b_bytes));
        tc05::tma_load_2d(st, &map_vis, &full[s], kb * kTcBK, m0);
        tc05::tma_load_2d(st + 2 * a_bytes, &map_word, &full[s], kb * kTcBK, c0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc05::instr_desc(tc05::kFmtTF32, 128, p.nqt);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kTcStages;
        mbar_wait(&ready[s], (uint32_t)(kb / kTcStages) & 1u);
        tc05::fence_after_sync();
        const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
        const uint64_t a_hi = tc05::smem_desc_sw128(st), a_lo = tc05::smem_desc_sw128(st + a_bytes);
        const uint64_t b_hi = tc05::smem_desc_sw128(st + 2 * a_bytes);
        const uint64_t b_lo = tc05::smem_desc_sw128(st + 2 * a_bytes + b_bytes);
#pragma unroll
        for (int k = 0; k < kTcBK / 8; ++k) {  // 8 tf32 = 32 bytes of K per instruction
          const uint64_t o = (uint64_t)(2 * k);
          tc05::mma_tf32(tmem_base, a_hi + o, b_hi + o, idesc, (kb | k) != 0);
          tc05::mma_tf32(tmem_base, a_lo + o, b_hi + o, idesc, 1u);
          tc05::mma_tf32(tmem_base, a_hi + o, b_lo + o, idesc, 1u);
        }
        tc05::commit(&empty[s]);
      }
      tc05::commit(acc_ready);
    }
  } else {
    // ---- hi / lo split of every landed stage (256 threads), element-wise in place
    const int t = tid - 64;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kTcStages;
      mbar_wait(&full[s], (uint32_t)(kb / kTcStages) & 1u);
      unsigned char* st = smem + (size_t)s * stage_bytes;
      auto split = [&](unsigned char* hi_base, int bytes) {
        uint4* hi = reinterpret_cast<uint4*>(hi_base);
        uint4* lo = reinterpret_cast<uint4*>(hi_base + bytes);
        for (int i = t; i < bytes / 16; i += kTcSplitThreads) {
          const uint4 x = hi[i];
          uint4 h, l;
          h.x = x.x & 0xffffe000u; h.y = x.y & 0xffffe000u; h.z = x.z & 0xffffe000u; h.w = x.w & 0xffffe000u;
          l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w));
          hi[i] = h;
          lo[i] = l;
        }
      };
      split(st, a_bytes);
      split(st + 2 * a_bytes, b_bytes);
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async reads
      mbar_arrive(&ready[s]);
    }
    // ---- epilogue: accumulator -> shared S tile (warps 2..5 own TMEM lane quarters 2, 3, 0, 1)
    mbar_wait(acc_ready, 0);
    tc05::fence_after_sync();
    const int ld = p.nqt + 1;
    if (warp < 6) {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      for (int cc = 0; cc < p.nqt; cc += 32) {
        uint32_t v[32];
        tc05::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)cc, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cc + j < p.nqt) S[row * ld + cc + j] = __uint_as_float(v[j]);
      }
    }
    asm volatile("bar.sync 2, %0;" ::"n"(kTcSplitThreads) : "memory");  // the 8 split / epilogue warps
    // ---- per (frame, column): max / first argmax over the Nb boxes (model.py:608-612), masked
    // columns are all-zero after masked_fill_ (model.py:551): max 0 at index 0
    for (int i = t; i < nf * nc; i += kTcSplitThreads) {
      const int fl = i / nc, c = c0 + i % nc;
      const bool live = (c % d.Ne) < __ldg(p.lens + c / d.Ne);
      float best = 0.f;
      int bi = 0;
      if (live) {
        const float* col = S + (size_t)(fl * d.Nb) * ld + (c - c0);
        best = -INFINITY;
        float second = -INFINITY;
        for (int r = 0; r < d.Nb; ++r) {
          const float x = col[(size_t)r * ld];
          if (x > best) {
            second = best;
            best = x;
            bi = r;
          } else if (x > second) {
            second = x;
          }
        }
        if (best - second < kTcTieTol) {  // too close for the tf32x3 accumulator: decide in fp32
          const float* wrow = p.word + (size_t)c * d.D;
          float eb = -INFINITY;
          int ei = 0;
          for (int r = 0; r < d.Nb; ++r)
            if (best - col[(size_t)r * ld] < 2.f * kTcTieTol) {
              const float e = exact_dot(p.vis + (size_t)(m0 + fl * d.Nb + r) * d.D, wrow, d.D);
              if (e > eb) {
                eb = e;
                ei = r;
              }
            }
          best = eb;
          bi = ei;
        }
      }
      p.D_sim[(size_t)(f0 + fl) * d.NQ + c] = best;
      p.D_ind[(size_t)(f0 + fl) * d.NQ + c] = (long long)bi;
    }
  }
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc05::tmem_dealloc(tmem_base, tmem_cols);
}

// Column statistics of one segment from a shared-memory copy of its D_sim block.
// Sblk is [Ns][NQ]; first-index tie rule like torch.min/max(dim).
__device__ __forceinline__ ColStat col_stat_smem(const float* __restrict__ Sblk, int c, int Ns,
                                                 int NQ) {
  ColStat st;
  st.mn = INFINITY;
  st.mx = -INFINITY;
  st.s_mn = 0;
  st.s_mx = 0;
  for (int s = 0; s < Ns; ++s) {
    const float x = Sblk[s * NQ + c];
    if (x < st.mn) {
      st.mn = x;
      st.s_mn = s;
    }
    if (x > st.mx) {
      st.mx = x;
      st.s_mx = s;
    }
  }
  st.den = (st.mx - st.mn) + kEps;
  return st;
}

// 4x4 block of the Gram matrix of the staged rows: M[ra+i][rb+j] = x_{ra+i} . x_{rb+j}.
// Lanes split k (128-bit LDS), 16 independent accumulators, one reduce16.
__device__ __forceinline__ void gram_block(const float* rows_sm, int D, int nrows, int ra, int rb,
                                           float* M, int lane) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const int n4 = D >> 2;
  const float4* pa[4];
  const float4* pb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    pa[i] = reinterpret_cast<const float4*>(rows_sm + (size_t)min(ra + i, nrows - 1) * D);
    pb[i] = reinterpret_cast<const float4*>(rows_sm + (size_t)min(rb + i, nrows - 1) * D);
  }
  for (int k0 = 0; k0 < n4; k0 += 64) {  // 2 float4 per lane per row per step
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = k0 + lane + 32 * u;
      float4 xa[4], xb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xa[i] = k < n4 ? pa[i][k] : make_float4(0.f, 0.f, 0.f, 0.f);
        xb[i] = k < n4 ? pb[i][k] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float t = acc[i * 4 + j];
          t = fmaf(xa[i].x, xb[j].x, t);
          t = fmaf(xa[i].y, xb[j].y, t);
          t = fmaf(xa[i].z, xb[j].z, t);
          t = fmaf(xa[i].w, xb[j].w, t);
          acc[i * 4 + j] = t;
        }
    }
  }
  const float tot = reduce16(acc, lane);
  if ((lane & 1) == 0) {
    const int i = lane >> 3, j = (lane >> 1) & 3;
    if (ra + i < nrows && rb + j < nrows) {
      M[(ra + i) * kRowTile + rb + j] = tot;
      M[(rb + j) * kRowTile + ra + i] = tot;
    }
  }
}

// P2: frame attention + Sf for segment a; clustering partials (train).
// Everything the phase needs from other CTAs (the segment's Ns x NQ block of D_sim / D_ind) is
// pulled into shared memory with ONE round of independent L2 loads, overlapped with the bulk copy
// of the vis rows the clustering loss reads.
__device__ void phase2_segment(const FwdParams& p, const Ws& w, int a) {
  // same dynamic shared memory as the kernel (declared here so that loads compile to LDS)
  extern __shared__ __align__(16) float sm[];
  const Dims& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = d.Ns * d.NQ;
  const int staged = p.train ? min(d.Nb, kRowTile) : 0;  // vis rows 0..staged-1 live in smem
  const bool all_staged = staged == d.Nb;
  float* rows_sm = sm;                                // [staged][D]
  float* Sblk = sm + (size_t)staged * d.D;            // [Ns][NQ]
  int* Iblk = reinterpret_cast<int*>(Sblk + blk);     // [Ns][NQ]
  float* c_mn = Sblk + 2 * blk;                       // [NQ]
  float* c_iden = c_mn + d.NQ;                        // [NQ] 1/den
  float* inv = c_iden + d.NQ;                         // [max(Nb, kRowTile)] 1/(||x_r||+eps)
  float* M = inv + max(d.Nb, kRowTile);               // [kRowTile][kRowTile] Gram of staged rows
  __shared__ __align__(8) uint64_t s_bar2;
  __shared__ int s_len[256];  // entities_length[a'] (first 256 segments; others re-read)
  __shared__ float s_red[kFwdWarps];
  __shared__ int s_redi[kFwdWarps];
  __syncthreads();
  if (tid == 0 && staged > 0) {
    // The clustering loss gathers rows WITHOUT their (segment, frame) offset (SURVEY.md fact
    // 0.7): only rows 0..Nb-1 of vis_feats are ever read.  One bulk copy brings them on chip.
    mbar_init(&s_bar2, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)staged * d.D * 4u;
    mbar_arrive_expect_tx(&s_bar2, bytes);
    bulk_g2s(rows_sm, p.vis, bytes, &s_bar2);
  }
  for (int i = tid; i < min(d.Na, 256); i += kFwdThreads) s_len[i] = __ldg(p.lens + i);
  const float* gS = p.D_sim + (size_t)a * blk;
  const long long* gI = p.D_ind + (size_t)a * blk;
  for (int i0 = 0; i0 < blk; i0 += 4 * kFwdThreads) {  // 4 independent loads per thread in flight
    float sv[4];
    int iv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kFwdThreads + tid;
      sv[j] = i < blk ? __ldcg(gS + i) : 0.f;
      iv[j] = i < blk ? (int)__ldcg(gI + i) : 0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kFwdThreads + tid;
      if (i < blk) {
        Sblk[i] = sv[j];
        Iblk[i] = iv[j];
      }
    }
  }
  __syncthreads();
  TRACE(16 + a * 8 + 1);
  for (int c = tid; c < d.NQ; c += kFwdThreads) {
    const ColStat st = col_stat_smem(Sblk, c, d.Ns, d.NQ);
    c_mn[c] = st.mn;
    c_iden[c] = 1.f / st.den;
  }
  if (p.train && all_staged) {
    // Gram matrix of the Nb candidate rows, 4x4 blocks dealt to the warps (upper triangle)
    mbar_wait(&s_bar2, 0);
    const int nb4 = (staged + 3) >> 2;
    int idx = 0;
    for (int bi = 0; bi < nb4; ++bi)
      for (int bj = bi; bj < nb4; ++bj, ++idx)
        if ((idx & (kFwdWarps - 1)) == warp) gram_block(rows_sm, d.D, staged, bi * 4, bj * 4, M, lane);
  }
  __syncthreads();
  TRACE(16 + a * 8 + 2);
  // Sf[a,s,a'] = sum_e S*S_att / div[a']   (model.py:583-593); thread per (s, a')
  for (int i = tid; i < d.Ns * d.Na; i += kFwdThreads) {
    const int s = i / d.Na, a2 = i % d.Na;
    const int len = a2 < 256 ? s_len[a2] : __ldg(p.lens + a2);
    // columns saturate at Ne like the reference's mask slicing (model.py:535-536); only the
    // divisor uses the raw length (model.py:538)
    const int ncol = min(len, d.Ne);
    float acc = 0.f;
    for (int e = 0; e < ncol; ++e) {
      const int c = a2 * d.Ne + e;
      const float x = Sblk[s * d.NQ + c];
      acc += x * ((x - c_mn[c]) * c_iden[c]);
    }
    w.Sf[((size_t)a * d.Ns + s) * d.Na + a2] = acc / (float)(len == 0 ? 1 : len);
  }
  if (!p.train) return;

  // clustering loss of segment a (model.py:553-577)
  const int len = min(a < 256 ? s_len[a] : __ldg(p.lens + a), d.Ne);
  const int pairs = d.Ns * (d.Ns - 1) / 2;
  float gsum = 0.f;
  int gcnt = 0;
  if (all_staged) {
    for (int r = tid; r < d.Nb; r += kFwdThreads) inv[r] = 1.f / (sqrtf(M[r * kRowTile + r]) + kEps);
    __syncthreads();
    TRACE(16 + a * 8 + 4);
    // Gram entries s<t of every entity (G is symmetric: the ordered sum / count double);
    // thread per (entity, pair), everything is a table lookup now
    for (int i = tid; i < len * pairs; i += kFwdThreads) {
      const int e = i / pairs;
      int q = i % pairs, s = 0;
      while (q >= d.Ns - 1 - s) {
        q -= d.Ns - 1 - s;
        ++s;
      }
      const int t = s + 1 + q;
      const int c = a * d.Ne + e;
      const int rs = Iblk[s * d.NQ + c], rt = Iblk[t * d.NQ + c];
      const float ss = (Sblk[s * d.NQ + c] - c_mn[c]) * c_iden[c];
      const float st = (Sblk[t * d.NQ + c] - c_mn[c]) * c_iden[c];
      const float g = 1.f - M[rs * kRowTile + rt] * (ss * inv[rs]) * (st * inv[rt]);
      gsum += 2.f * g;
      gcnt += (g != 0.f) ? 2 : 0;
    }
  } else {
    // Nb > kRowTile: rows beyond the staged tile are read from global memory, warp per dot
    if (staged > 0) mbar_wait(&s_bar2, 0);
    auto row_ptr = [&](int r) -> const float* {
      return r < staged ? rows_sm + (size_t)r * d.D : p.vis + (size_t)r * d.D;
    };
    for (int r = warp; r < d.Nb; r += kFwdWarps) {
      const float acc = warp_dot(row_ptr(r), row_ptr(r), d.D, lane);
      if (lane == 0) inv[r] = 1.f / (sqrtf(acc) + kEps);
    }
    __syncthreads();
    for (int i = warp; i < len * pairs; i += kFwdWarps) {
      const int e = i / pairs;
      int q = i % pairs, s = 0;
      while (q >= d.Ns - 1 - s) {
        q -= d.Ns - 1 - s;
        ++s;
      }
      const int t = s + 1 + q;
      const int c = a * d.Ne + e;
      const int rs = Iblk[s * d.NQ + c], rt = Iblk[t * d.NQ + c];
      const float acc = warp_dot(row_ptr(rs), row_ptr(rt), d.D, lane);
      const float ss = (Sblk[s * d.NQ + c] - c_mn[c]) * c_iden[c];
      const float st = (Sblk[t * d.NQ + c] - c_mn[c]) * c_iden[c];
      const float g = 1.f - acc * (ss * inv[rs]) * (st * inv[rt]);
      if (lane == 0) {
        gsum += 2.f * g;
        gcnt += (g != 0.f) ? 2 : 0;
      }
    }
  }
  TRACE(16 + a * 8 + 5);
  gsum = warp_sum(gsum);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gcnt += __shfl_xor_sync(0xffffffffu, gcnt, o);
  if (lane == 0) {
    s_red[warp] = gsum;
    s_redi[warp] = gcnt;
  }
  __syncthreads();
  if (tid == 0) {
    float ts = 0.f;
    int tc = 0;
    for (int k = 0; k < kFwdWarps; ++k) {
      ts += s_red[k];
      tc += s_redi[k];
    }
    w.vsum[a] = ts;
    w.vcnt[a] = tc;
  }
}

// P3: hinge ranking loss over all segments (model.py:594-606) and its gradient w.r.t. Sf.
// Gather form, thread per Sf element (p, s, q): no atomics, one round of L2 loads.
__device__ void phase3_final(const FwdParams& p, const Ws& w) {
  extern __shared__ __align__(16) float sm[];
  const Dims& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = d.Na * d.Ns * d.Na;
  float* Sf = sm;  // [Na][Ns][Na]
  __shared__ float s_part[kFwdWarps];
  __shared__ float s_vs[kFwdWarps];
  __shared__ int s_vc[kFwdWarps];
  __syncthreads();
  float vs = 0.f;
  int vc = 0;
  if (p.train)
    for (int a = tid; a < d.Na; a += kFwdThreads) {
      vs += __ldcg(w.vsum + a);
      vc += __ldcg(w.vcnt + a);
    }
  for (int i0 = 0; i0 < n; i0 += 4 * kFwdThreads) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kFwdThreads + tid;
      v[j] = i < n ? __ldcg(w.Sf + i) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + j * kFwdThreads + tid;
      if (i < n) Sf[i] = v[j];
    }
  }
  vs = warp_sum(vs);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vc += __shfl_xor_sync(0xffffffffu, vc, o);
  if (lane == 0) {
    s_vs[warp] = vs;
    s_vc[warp] = vc;
  }
  __syncthreads();
  TRACE(121);
  // frame_score[a,s] = mean_o relu(Sf[o,s,a] - d[a,s] + Delta) + mean_o relu(Sf[a,s,o] - d[a,s] + Delta)
  // Thread per (pp, s): its row Sf[pp,s,:] holds the "text o as a negative for video pp" hinges and
  // column pp of frame slot s the "video o as a negative for text pp" hinges; both use d[pp,s].
  float part = 0.f;
  const float inv_na = 1.f / (float)d.Na;
  const float gscale = 10.f / (float)(d.Na * d.Ns) * inv_na;  // d(margin)/d(relu term)
  for (int i = tid; i < d.Na * d.Ns; i += kFwdThreads) {
    const int pp = i / d.Ns, s = i - pp * d.Ns;
    const float* row = Sf + (size_t)(pp * d.Ns + s) * d.Na;
    const float dp = row[pp];
    int cnt = 0;  // active hinges that contain -d[pp,s]
    for (int o = 0; o < d.Na; ++o) {
      const float x = row[o];
      const float vB = (x - dp) + p.Delta;                                       // text o vs video pp
      const float vA = (x - Sf[(size_t)(o * d.Ns + s) * d.Na + o]) + p.Delta;    // video pp vs text o
      const float vC = (Sf[(size_t)(o * d.Ns + s) * d.Na + pp] - dp) + p.Delta;  // video o vs text pp
      part += fmaxf(vB, 0.f) + fmaxf(vC, 0.f);
      cnt += (vB > 0.f ? 1 : 0) + (vC > 0.f ? 1 : 0);
      if (o != pp)
        w.hgrad[(size_t)(pp * d.Ns + s) * d.Na + o] =
            gscale * ((vA > 0.f ? 1.f : 0.f) + (vB > 0.f ? 1.f : 0.f));
    }
    // diagonal element: its own two hinges are relu(Delta) (+2 if Delta > 0) minus every active hinge
    w.hgrad[(size_t)(pp * d.Ns + s) * d.Na + pp] = gscale * ((p.Delta > 0.f ? 2.f : 0.f) - (float)cnt);
  }
  part = warp_sum(part);
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  TRACE(122);
  if (tid == 0) {
    float fs = 0.f, ts = 0.f;
    int tc = 0;
    for (int k = 0; k < kFwdWarps; ++k) {
      fs += s_part[k];
      ts += s_vs[k];
      tc += s_vc[k];
    }
    const float mean_fs = fs * inv_na / (float)(d.Na * d.Ns);
    float loss = mean_fs * 10.f;
    float vis_loss = 0.f, dem = 0.f;
    if (p.train) {
      dem = (float)tc;
      vis_loss = ts / dem;  // NaN when nothing is unmasked, like the reference (model.py:576-577)
      loss = (mean_fs + p.vis_lam * vis_loss) * 10.f;
    }
    w.scal[0] = vis_loss;
    w.scal[1] = dem;
    w.scal[2] = mean_fs;
    // d|loss|/dloss for the L1Loss(margin_loss, 0) of the step wrapper (model.py:771): the backward
    // uses it when the caller passes no upstream gradient
    w.scal[3] = loss > 0.f ? 1.f : (loss < 0.f ? -1.f : loss);
    *p.loss = loss;
  }
}

// ----------------------------------------------------------------------- backward ----
struct BwdParams {
  const float* gout;
  const float* vis;
  const float* word;
  const int* lens;
  const long long* D_ind;
  const float* D_sim;
  float* gvis;
  float* gword;
  void* ws;
  Dims d;
  float vis_lam;
  int train;
};

#ifndef NAFAE_BWD_THREADS
#define NAFAE_BWD_THREADS 512
#endif
constexpr int kBwdThreads = NAFAE_BWD_THREADS;
constexpr int kMaxNsLocal = 64;  // frames per segment handled by the fused backward

// d(margin_loss)/dS[a,s,c] for all s of one (segment, column), from shared-memory copies:
// x[s*xs] = S[a,s,c], h[s*hs] = d(margin)/dSf[a,s,a2] (A12 in SURVEY.md).  out[s*os].
__device__ __forceinline__ void col_grad_smem(const float* __restrict__ x, int xs,
                                              const float* __restrict__ h, int hs, int Ns,
                                              float scale /* gout / div */, float* out, int os) {
  float mn = INFINITY, mx = -INFINITY;
  int s_mn = 0, s_mx = 0;
  for (int s = 0; s < Ns; ++s) {
    const float v = x[s * xs];
    if (v < mn) {
      mn = v;
      s_mn = s;
    }
    if (v > mx) {
      mx = v;
      s_mx = s;
    }
  }
  const float iden = 1.f / ((mx - mn) + kEps);
  float g_mn = 0.f, g_mx = 0.f;
  for (int s = 0; s < Ns; ++s) {
    const float v = x[s * xs];
    const float q = scale * h[s * hs];
    out[s * os] = q * ((v - mn) * iden + v * iden);
    const float qx = q * v * iden * iden;
    g_mn += qx * ((v - mx) - kEps);
    g_mx -= qx * (v - mn);
  }
  out[s_mn * os] += g_mn;
  out[s_mx * os] += g_mx;
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ONE backward kernel, three kinds of CTAs (512 threads):
//   blockIdx <  F          : dL/dvis rows of frame f   (dense overwrite, Nb x D)
//   blockIdx <  F + NQ     : dL/dword row of column c
//   blockIdx <  F + 2 NQ   : (train) clustering-loss gradient of entity (a, e), added atomically
//                            into grad_vis rows 0..Nb-1 -- the rows the reference's un-offset
//                            index_select gathers -- AFTER block 0 (frame 0)
//                            has written them: block 0 publishes a flag, these CTAs acquire it.
//                            CTAs are dispatched in blockIdx order, so block 0 is always running
//                            or finished when a later block spins (decoupled-look-back argument).
// dL/dS_ has at most one non-zero per (frame, column) -- the argmax box -- so every sweep is a
// gather-scale-accumulate over <= F*NQ pairs, not a GEMM.  All cross-CTA inputs are staged into
// shared memory with one round of independent loads; gather loops issue loads in batches.
// min 2 CTAs/SM: caps the kernel at 64 registers (66 without the bound = ONE 512-thread CTA per SM by
// registers, i.e. 16 resident CTAs on the 16 SMs the pipelined step leaves to the head: 15 waves)
__global__ void __launch_bounds__(kBwdThreads, 1024 / kBwdThreads) ground_bwd_kernel(const BwdParams p) {
  NAFAE_CTA_TRACE(cta_trace, 4);
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_nlive;
  __shared__ int s_ticket;
  __shared__ __align__(8) uint64_t s_bar;
  const Dims& d = p.d;
  const Ws w = ws_carve(p.ws, d);
  int* flag = w.done_cnt + 1;      // frame-0 rows written
  int* finished = w.done_cnt + 2;  // CTAs that are done (the last one resets the scratch words)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float gout = p.gout != nullptr ? __ldg(p.gout) : __ldcg(w.scal + 3);
  const int bid = blockIdx.x;

  if (bid < d.F) {
    // ------------------------------------------------------------------ dL/dvis ----
    const int f = bid, a = f / d.Ns, s_me = f % d.Ns;
    const int blk = d.Ns * d.NQ;
    float* Sblk = sm;                                  // [Ns][NQ]
    float* hg = Sblk + blk;                            // [Ns][Na]
    float* gme = hg + d.Ns * d.Na;                     // [NQ] gradient of this frame's S row
    int* ridx = reinterpret_cast<int*>(gme + d.NQ);    // [NQ]
    int* live = ridx + d.NQ;                           // [NQ] compact list of live columns
    // [kRowTile][D], 16-byte aligned for the float4 write-out
    float* acc = sm + (((size_t)blk + (size_t)d.Ns * d.Na + 3 * (size_t)d.NQ + 3) & ~(size_t)3);
    for (int i = tid; i < blk; i += kBwdThreads) Sblk[i] = __ldg(p.D_sim + (size_t)a * blk + i);
    for (int i = tid; i < d.Ns * d.Na; i += kBwdThreads)
      hg[i] = __ldg(w.hgrad + (size_t)a * d.Ns * d.Na + i);
    for (int c = tid; c < d.NQ; c += kBwdThreads) {
      const bool lv = (c % d.Ne) < __ldg(p.lens + c / d.Ne);
      ridx[c] = lv ? (int)__ldg(p.D_ind + (size_t)f * d.NQ + c) : -1;
    }
    __syncthreads();
    for (int c = tid; c < d.NQ; c += kBwdThreads) {
      if (ridx[c] >= 0) {
        const int a2 = c / d.Ne;
        const int len = __ldg(p.lens + a2);
        float tmp[kMaxNsLocal];
        col_grad_smem(Sblk + c, d.NQ, hg + a2, d.Na, d.Ns, gout / (float)(len == 0 ? 1 : len), tmp, 1);
        gme[c] = tmp[s_me];
      }
    }
    if (tid < 32) {  // ordered compaction of the live columns (warp 0)
      int n = 0;
      for (int c0 = 0; c0 < d.NQ; c0 += 32) {
        const int c = c0 + lane;
        const bool lv = c < d.NQ && ridx[c] >= 0;
        const unsigned bal = __ballot_sync(0xffffffffu, lv);
        if (lv) live[n + __popc(bal & ((1u << lane) - 1u))] = c;
        n += __popc(bal);
      }
      if (lane == 0) s_nlive = n;
    }
    __syncthreads();
    const int nlive = s_nlive;
    for (int r0 = 0; r0 < d.Nb; r0 += kRowTile) {
      const int rows = min(kRowTile, d.Nb - r0);
      {
        float4* z = reinterpret_cast<float4*>(acc);
        for (int i = tid; i < rows * d.D / 4; i += kBwdThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      for (int k = tid; k < d.D; k += kBwdThreads) {  // thread owns feature k of every row
        for (int j0 = 0; j0 < nlive; j0 += 16) {
          float wv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            wv[j] = __ldg(p.word + (size_t)live[min(j0 + j, nlive - 1)] * d.D + k);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j0 + j < nlive) {
              const int c = live[j0 + j];
              const int r = ridx[c] - r0;
              if (r >= 0 && r < rows) acc[r * d.D + k] = fmaf(gme[c], wv[j], acc[r * d.D + k]);
            }
          }
        }
      }
      __syncthreads();
      float4* dst = reinterpret_cast<float4*>(p.gvis + ((size_t)f * d.Nb + r0) * d.D);
      const float4* src = reinterpret_cast<const float4*>(acc);
      for (int i = tid; i < rows * d.D / 4; i += kBwdThreads) dst[i] = src[i];
      __syncthreads();
    }
    if (bid == 0 && tid == 0) st_release(flag, 1);  // ordered after the CTA's stores by the barrier
  } else {
    // Column roles are dealt by LIVE rank, so that the CTAs with real work are dispatched first
    // (blockIdx order) and the rest -- zero-fill of a masked dL/dword row, or nothing -- trail:
    //   [F, F+nlive)            dL/dword of live column #k
    //   [F+nlive, F+2 nlive)    (train) clustering gradient of live column #k
    //   then                    one masked column each: dL/dword row = 0
    __shared__ int s_col;   // column this CTA serves
    __shared__ int s_role;  // 0 word, 1 cluster, 2 zero-fill, 3 nothing
    if (tid < 32) {
      const int want = bid - d.F;
      int nlive = 0, ndead = 0, col_live = -1, col_dead = -1;
      // first pass: count; the k-th live / dead column is located in the same sweep
      int nl_tot = 0;
      for (int c0 = 0; c0 < d.NQ; c0 += 32) {
        const int c = c0 + lane;
        const bool lv = c < d.NQ && (c % d.Ne) < __ldg(p.lens + c / d.Ne);
        nl_tot += __popc(__ballot_sync(0xffffffffu, lv));
      }
      const int k_live = want < nl_tot ? want : (p.train && want < 2 * nl_tot ? want - nl_tot : -1);
      const int k_dead = want - (p.train ? 2 : 1) * nl_tot;
      for (int c0 = 0; c0 < d.NQ; c0 += 32) {
        const int c = c0 + lane;
        const bool in = c < d.NQ;
        const bool lv = in && (c % d.Ne) < __ldg(p.lens + c / d.Ne);
        const unsigned bl = __ballot_sync(0xffffffffu, lv), bd = __ballot_sync(0xffffffffu, in && !lv);
        const int pl = nlive + __popc(bl & ((1u << lane) - 1u)), pd = ndead + __popc(bd & ((1u << lane) - 1u));
        if (lv && pl == k_live) col_live = c;
        if (in && !lv && pd == k_dead) col_dead = c;
        nlive += __popc(bl);
        ndead += __popc(bd);
      }
      // exactly one lane (or none) found its column: reduce with max
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        col_live = max(col_live, __shfl_xor_sync(0xffffffffu, col_live, o));
        col_dead = max(col_dead, __shfl_xor_sync(0xffffffffu, col_dead, o));
      }
      if (lane == 0) {
        if (k_live >= 0) {
          s_col = col_live;
          s_role = want < nl_tot ? 0 : 1;
        } else {
          s_col = col_dead;
          s_role = col_dead >= 0 ? 2 : 3;
        }
      }
    }
    __syncthreads();
    const int role = s_role;
    if (role == 2) {
      float* dst = p.gword + (size_t)s_col * d.D;
      for (int k = tid; k < d.D; k += kBwdThreads) dst[k] = 0.f;
    } else if (role == 0) {
    // ----------------------------------------------------------------- dL/dword ----
    const int c = s_col, a2 = c / d.Ne;
    float* dst = p.gword + (size_t)c * d.D;
    const int len = __ldg(p.lens + a2);
    {
      float* xcol = sm;                                  // [F]
      float* hcol = xcol + d.F;                          // [F]
      float* g = hcol + d.F;                             // [F]
      int* ridx = reinterpret_cast<int*>(g + d.F);       // [F] global vis row of the picked box
      for (int f = tid; f < d.F; f += kBwdThreads) {
        xcol[f] = __ldg(p.D_sim + (size_t)f * d.NQ + c);
        hcol[f] = __ldg(w.hgrad + (size_t)f * d.Na + a2);
        ridx[f] = f * d.Nb + (int)__ldg(p.D_ind + (size_t)f * d.NQ + c);
      }
      __syncthreads();
      for (int a = tid; a < d.Na; a += kBwdThreads)
        col_grad_smem(xcol + a * d.Ns, 1, hcol + a * d.Ns, 1, d.Ns, gout / (float)len, g + a * d.Ns, 1);
      __syncthreads();
      for (int k = tid; k < d.D; k += kBwdThreads) {
        float accv = 0.f;
        for (int f0 = 0; f0 < d.F; f0 += 20) {
          float v[20];
#pragma unroll
          for (int j = 0; j < 20; ++j) v[j] = __ldg(p.vis + (size_t)ridx[min(f0 + j, d.F - 1)] * d.D + k);
#pragma unroll
          for (int j = 0; j < 20; ++j)
            if (f0 + j < d.F) accv = fmaf(g[f0 + j], v[j], accv);
        }
        dst[k] = accv;
      }
    }
    } else if (role == 1) {
    // ------------------------------------------------ clustering-loss gradient ----
    const int c = s_col, a = c / d.Ne;
    const float dem = __ldg(w.scal + 1);
    if (dem > 0.f) {
      float* rows = sm;                       // [Ns][D] the picked rows of this entity
      float* Vsum = rows + (size_t)d.Ns * d.D;
      float* simn = Vsum + d.D;
      float* inv = simn + d.Ns;
      float* dots = inv + d.Ns;
      int* row = reinterpret_cast<int*>(dots + d.Ns);
      for (int s = tid; s < d.Ns; s += kBwdThreads) {
        simn[s] = __ldg(p.D_sim + (size_t)(a * d.Ns + s) * d.NQ + c);
        row[s] = (int)__ldg(p.D_ind + (size_t)(a * d.Ns + s) * d.NQ + c);
      }
      if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
      }
      __syncthreads();
      if (tid == 0) {  // one 1-D bulk copy per picked row (rows differ per frame)
        mbar_arrive_expect_tx(&s_bar, (uint32_t)d.Ns * d.D * 4u);
        for (int s = 0; s < d.Ns; ++s)
          bulk_g2s(rows + (size_t)s * d.D, p.vis + (size_t)row[s] * d.D, (uint32_t)d.D * 4u, &s_bar);
      }
      float mn = INFINITY, mx = -INFINITY;
      for (int s = 0; s < d.Ns; ++s) {
        mn = fminf(mn, simn[s]);
        mx = fmaxf(mx, simn[s]);
      }
      const float iden = 1.f / ((mx - mn) + kEps);
      __syncthreads();
      for (int s = tid; s < d.Ns; s += kBwdThreads) simn[s] = (simn[s] - mn) * iden;
      mbar_wait(&s_bar, 0);
      for (int s = warp; s < d.Ns; s += kBwdThreads / 32) {
        const float* x = rows + (size_t)s * d.D;
        float a0 = 0.f, a1 = 0.f;
        for (int k = lane; k < d.D; k += 64) {
          const float v0 = x[k], v1 = k + 32 < d.D ? x[k + 32] : 0.f;
          a0 = fmaf(v0, v0, a0);
          a1 = fmaf(v1, v1, a1);
        }
        const float acc = warp_sum(a0 + a1);
        if (lane == 0) inv[s] = 1.f / (sqrtf(acc) + kEps);
      }
      __syncthreads();
      for (int k = tid; k < d.D; k += kBwdThreads) {
        float acc = 0.f;
        for (int s = 0; s < d.Ns; ++s) acc = fmaf(simn[s] * inv[s], rows[(size_t)s * d.D + k], acc);
        Vsum[k] = acc;
      }
      __syncthreads();
      // vis_loss = sum(G)/dem, G[s,t] = 1 - V_s.V_t (s != t): d/dV_s = -2 (Vsum - V_s) / dem
      const float coef = -2.f * (10.f * p.vis_lam * gout) / dem;
      for (int s = warp; s < d.Ns; s += kBwdThreads / 32) {  // dots[s] = x_s . gu_s
        const float* x = rows + (size_t)s * d.D;
        const float sc = simn[s] * inv[s];
        float a0 = 0.f, a1 = 0.f;
        for (int k = lane; k < d.D; k += 64) {
          const float v0 = x[k];
          a0 = fmaf(v0, Vsum[k] - sc * v0, a0);
          if (k + 32 < d.D) {
            const float v1 = x[k + 32];
            a1 = fmaf(v1, Vsum[k + 32] - sc * v1, a1);
          }
        }
        const float acc = warp_sum(a0 + a1);
        if (lane == 0) dots[s] = acc * simn[s] * coef;
      }
      __syncthreads();
      if (tid == 0)
        while (ld_acquire(flag) == 0) {
        }
      __syncthreads();
      for (int k = tid; k < d.D; k += kBwdThreads) {
        for (int s = 0; s < d.Ns; ++s) {
          const float iv = inv[s];             // 1/(n+eps)
          const float nrm = 1.f / iv - kEps;   // n
          const float sc = simn[s] * iv;
          // dL/dx = gu/(n+eps) - x (x.gu) / (n (n+eps)^2);  gu = simn*coef*(Vsum - V_s)
          const float c2 = nrm > 0.f ? dots[s] * iv * iv / nrm : 0.f;
          const float xv = rows[(size_t)s * d.D + k];
          const float gu = simn[s] * coef * (Vsum[k] - sc * xv);
          atomicAdd(p.gvis + (size_t)row[s] * d.D + k, gu * iv - xv * c2);
        }
      }
    }
    }
  }
  // last CTA out resets the scratch words for the next launch
  __syncthreads();
  if (tid == 0) s_ticket = ticket_acq_rel(finished);
  __syncthreads();
  if (s_ticket == (int)gridDim.x - 1 && tid == 0) {
    *flag = 0;
    *finished = 0;
  }
}

__global__ void postprocess_kernel(const long long* __restrict__ D_ind,
                                   const float* __restrict__ D_sim, int Na, int Ns, int Nb, int Ne,
                                   long long* __restrict__ out_ind, float* __restrict__ out_sim) {
  const int n = Na * Ns * Ne;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i % Ne, s = (i / Ne) % Ns, a = i / (Ne * Ns);
    const size_t src = ((size_t)(a * Ns + s) * Na + a) * Ne + e;
    out_ind[i] = D_ind[src] + (long long)a * Ns * Nb + (long long)s * Nb;  // model.py:471
    out_sim[i] = D_sim[src];
  }
}

// Evaluation sweep, one thread per (segment g, frame s, entity slot e) of `G` independent segments
// (D_ind / D_sim in the batched Na = 1 layout (G, Ns, Ne)):
//   model.py:457-474 postprocess : global box row (g*Ns + s)*Nb + D_ind
//   model.py:477-487 record_det  : box / confidence / image id of every real entity (e < len[g])
//   youcook_eval.py:241-336      : box accuracy against ONE ground-truth box per (image, label) --
//        overlap with the +1 pixel convention in float64, detection area in the boxes' own fp32
//        (the dtypes NumPy uses there, :206-221), hit when overlap >= thr; per-class counters
// Slots e >= len[g] are written as invalid (image id -1).
__global__ void eval_record_kernel(const long long* __restrict__ D_ind, const float* __restrict__ D_sim,
                                   const int* __restrict__ lens, const float* __restrict__ rois, int G,
                                   int Ns, int Nb, int Ne, long long img_base,
                                   long long* __restrict__ out_img, long long* __restrict__ out_row,
                                   float* __restrict__ out_box, float* __restrict__ out_conf,
                                   const double* __restrict__ gt_box, const int* __restrict__ gt_cls,
                                   float gt_thr, int n_cls, int* __restrict__ class_match,
                                   int* __restrict__ class_count) {
  const int n = G * Ns * Ne;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i % Ne, s = (i / Ne) % Ns, g = i / (Ne * Ns);
    const bool live = e < __ldg(lens + g);
    const long long row = (long long)(g * Ns + s) * Nb + (live ? D_ind[i] : 0);
    const float* r = rois + row * 5;
    const float x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
    if (out_img) out_img[i] = live ? img_base + (long long)g * Ns + s : -1;
    if (out_row) out_row[i] = live ? row : -1;
    if (out_box) {
      out_box[(size_t)i * 4 + 0] = live ? x1 : 0.f;
      out_box[(size_t)i * 4 + 1] = live ? y1 : 0.f;
      out_box[(size_t)i * 4 + 2] = live ? x2 : 0.f;
      out_box[(size_t)i * 4 + 3] = live ? y2 : 0.f;
    }
    if (out_conf) out_conf[i] = live ? D_sim[i] : 0.f;
    if (gt_box != nullptr && live) {
      const int cls = __ldg(gt_cls + g * Ne + e);
      if (cls >= 0 && cls < n_cls) {
        const double* gb = gt_box + (size_t)i * 4;
        const double left = fmax((double)x1, gb[0]), top = fmax((double)y1, gb[1]);
        const double right = fmin((double)x2, gb[2]), bottom = fmin((double)y2, gb[3]);
        const double iw = __dadd_rn(__dsub_rn(right, left), 1.), ih = __dadd_rn(__dsub_rn(bottom, top), 1.);
        const float darea = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
        const double garea = __dmul_rn(__dadd_rn(__dsub_rn(gb[2], gb[0]), 1.), __dadd_rn(__dsub_rn(gb[3], gb[1]), 1.));
        const double inter = __dmul_rn(iw, ih);
        const double uni = __dsub_rn(__dadd_rn((double)darea, garea), inter);
        const bool hit = iw > 0. && ih > 0. && __ddiv_rn(inter, uni) >= (double)gt_thr;
        atomicAdd(class_count + cls, 1);
        if (hit) atomicAdd(class_match + cls, 1);
      }
    }
  }
}

bool make_dims(int Na, int Ns, int Nb, int Ne, int D, Dims* d) {
  if (Na <= 0 || Ns <= 0 || Nb <= 0 || Ne <= 0 || D <= 0) return false;
  d->Na = Na;
  d->Ns = Ns;
  d->Nb = Nb;
  d->Ne = Ne;
  d->D = D;
  d->F = Na * Ns;
  d->NQ = Na * Ne;
  return true;
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_CTA_TRACE_READER(nafae_debug_cta_trace_ground)
#ifdef NAFAE_TRACE
NAFAE_API int nafae_debug_read_trace(unsigned long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_trace, sizeof(unsigned long long) * 256);
}
#endif

NAFAE_API size_t nafae_ground_workspace_bytes(int Na, int Ns, int Nb, int Ne, int D) {
  Dims d;
  if (!make_dims(Na, Ns, Nb, Ne, D, &d)) return 0;
  return align_up(ws_words(d) * 4, 256);
}

static int check_ground_args(const Dims& d, const void* ws, size_t ws_bytes, int train) {
  NAFAE_REQUIRE(d.D % 4 == 0, "ground: D must be a multiple of 4, got %d", d.D);
  NAFAE_REQUIRE(d.Ne <= 16, "ground: max_ent_len %d > 16 not supported", d.Ne);
  NAFAE_REQUIRE(d.NQ <= kLiveMax, "ground: Na*Ne = %d exceeds %d", d.NQ, kLiveMax);
  NAFAE_REQUIRE(!train || d.Ns <= kMaxNsLocal,
                "ground: train phase supports at most %d frames per segment", kMaxNsLocal);
  NAFAE_REQUIRE((long long)d.F * d.Nb * (long long)d.D < (1ll << 31), "ground: vis_feats too large");
  NAFAE_REQUIRE(ws && ws_bytes >= nafae_ground_workspace_bytes(d.Na, d.Ns, d.Nb, d.Ne, d.D),
                "ground: workspace too small");
  NAFAE_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "ground: workspace must be 16B aligned");
  return 1;
}

NAFAE_API int nafae_ground_forward(const float* vis_feats, const float* word_feats,
                                   const int* entities_length, int Na, int Ns, int Nb, int Ne,
                                   int D, float Delta, float vis_lam, int train, int64_t* D_ind,
                                   float* D_sim, float* margin_loss, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
  return nafae_ground_forward_batched(vis_feats, word_feats, entities_length, 1, Na, Ns, Nb, Ne, D, Delta,
                                      vis_lam, train, D_ind, D_sim, margin_loss, workspace,
                                      workspace_bytes, stream);
}

NAFAE_API int nafae_ground_forward_batched(const float* vis_feats, const float* word_feats,
                                           const int* entities_length, int groups, int Na, int Ns,
                                           int Nb, int Ne, int D, float Delta, float vis_lam, int train,
                                           int64_t* D_ind, float* D_sim, float* margin_loss,
                                           void* workspace, size_t workspace_bytes,
                                           cudaStream_t stream) {
  Dims d;
  NAFAE_REQUIRE(make_dims(Na, Ns, Nb, Ne, D, &d), "ground: sizes must be positive");
  NAFAE_REQUIRE(groups >= 1 && groups <= 65535, "ground: groups must be in [1, 65535]");
  const size_t ws_group = nafae_ground_workspace_bytes(Na, Ns, Nb, Ne, D);
  NAFAE_REQUIRE(workspace_bytes / (size_t)groups >= ws_group, "ground: workspace too small for %d groups",
                groups);
  if (check_ground_args(d, workspace, ws_group, train) != 1) return 0;
  NAFAE_REQUIRE(vis_feats && word_feats && entities_length && D_ind && D_sim && margin_loss,
                "ground: NULL buffer");
  NAFAE_REQUIRE((reinterpret_cast<uintptr_t>(vis_feats) & 15) == 0,
                "ground: vis_feats must be 16-byte aligned");
  FwdParams p;
  p.vis = vis_feats;
  p.word = word_feats;
  p.lens = entities_length;
  p.D_ind = reinterpret_cast<long long*>(D_ind);
  p.D_sim = D_sim;
  p.loss = margin_loss;
  p.ws = workspace;
  p.d = d;
  p.Delta = Delta;
  p.vis_lam = vis_lam;
  p.train = train ? 1 : 0;
  // CTAs per frame: enough to fill the GPU once, never more than the column chunks there can be
  p.groups = groups;
  p.ws_group_bytes = ws_group;
  {
    int g = ceil_div(sm_count(), d.F * groups);
    const int gmax = ceil_div(d.NQ, kColsPerCta);
    p.col_chunks = g < 1 ? 1 : (g > gmax ? gmax : g);
  }
  // shared memory: row tile (+ parked partial sums when D > 512), reused by P2's small tables
  size_t smem = (size_t)kRowTile * d.D * 4;
  const size_t p2 = ((size_t)(train ? (d.Nb < kRowTile ? d.Nb : kRowTile) : 0) * d.D +
                     (size_t)2 * d.Ns * d.NQ + 2 * (size_t)d.NQ + (size_t)(d.Nb > kRowTile ? d.Nb : kRowTile) +
                     (size_t)kRowTile * kRowTile) * 4;
  const size_t p3 = (size_t)d.Na * d.Ns * d.Na * 4;
  if (p2 > smem) smem = p2;
  if (p3 > smem) smem = p3;
  NAFAE_REQUIRE(smem <= 200 * 1024, "ground: D=%d / Ns=%d need too much shared memory", d.D, d.Ns);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ground_fwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("ground: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  ground_fwd_kernel<<<dim3(d.F * p.col_chunks, groups), kFwdThreads, smem, stream>>>(p);
  return launch_status("ground_fwd_kernel");
}

NAFAE_API int nafae_ground_backward(const float* grad_margin_loss, const float* vis_feats,
                                    const float* word_feats, const int* entities_length, int Na,
                                    int Ns, int Nb, int Ne, int D, float Delta, float vis_lam,
                                    int train, const int64_t* D_ind, const float* D_sim,
                                    float* grad_vis, float* grad_word, void* workspace,
                                    size_t workspace_bytes, cudaStream_t stream) {
  (void)Delta;
  Dims d;
  NAFAE_REQUIRE(make_dims(Na, Ns, Nb, Ne, D, &d), "ground: sizes must be positive");
  if (check_ground_args(d, workspace, workspace_bytes, 1) != 1) return 0;
  NAFAE_REQUIRE(vis_feats && word_feats && entities_length && D_ind && D_sim && grad_vis && grad_word,
                "ground: NULL buffer");
  BwdParams p;
  p.gout = grad_margin_loss;
  p.vis = vis_feats;
  p.word = word_feats;
  p.lens = entities_length;
  p.D_ind = reinterpret_cast<const long long*>(D_ind);
  p.D_sim = D_sim;
  p.gvis = grad_vis;
  p.gword = grad_word;
  p.ws = workspace;
  p.d = d;
  p.vis_lam = vis_lam;
  p.train = train ? 1 : 0;
  // shared memory: the largest of the three CTA roles
  // (the accumulator tile holds min(Nb, kRowTile) rows: 40 KB at Nb = 20, so that several CTAs fit an SM)
  size_t smem = ((size_t)d.Ns * d.NQ + (size_t)d.Ns * d.Na + 3 * (size_t)d.NQ +
                 (size_t)(d.Nb < kRowTile ? d.Nb : kRowTile) * d.D + 4) * 4;
  const size_t smem_w = (size_t)4 * d.F * 4;
  const size_t smem_c = ((size_t)d.Ns * d.D + (size_t)d.D + 4 * (size_t)d.Ns) * 4;
  if (smem_w > smem) smem = smem_w;
  if (p.train && smem_c > smem) smem = smem_c;
  NAFAE_REQUIRE(smem <= 200 * 1024, "ground backward: sizes need too much shared memory");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ground_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      set_error("ground: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  const int grid = d.F + d.NQ + (p.train ? d.NQ : 0);
  ground_bwd_kernel<<<grid, kBwdThreads, smem, stream>>>(p);
  return launch_status("ground_bwd_kernel");
}

NAFAE_API int nafae_ground_postprocess(const int64_t* D_ind, const float* D_sim, int Na, int Ns,
                                       int Nb, int Ne, int64_t* out_ind, float* out_sim,
                                       cudaStream_t stream) {
  NAFAE_REQUIRE(Na > 0 && Ns > 0 && Nb > 0 && Ne > 0, "postprocess: sizes must be positive");
  NAFAE_REQUIRE(D_ind && D_sim && out_ind && out_sim, "postprocess: NULL buffer");
  const int n = Na * Ns * Ne;
  postprocess_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(
      reinterpret_cast<const long long*>(D_ind), D_sim, Na, Ns, Nb, Ne,
      reinterpret_cast<long long*>(out_ind), out_sim);
  return launch_status("postprocess_kernel");
}

NAFAE_API int nafae_eval_record(const int64_t* D_ind, const float* D_sim, const int* entities_length,
                                const float* rois, int num_segments, int Ns, int Nb, int Ne,
                                int64_t image_id_base, int64_t* out_image_ids, int64_t* out_box_rows,
                                float* out_boxes, float* out_confs, const double* gt_boxes,
                                const int* gt_classes, float gt_thr, int num_classes, int* class_match,
                                int* class_count, cudaStream_t stream) {
  NAFAE_REQUIRE(num_segments > 0 && Ns > 0 && Nb > 0 && Ne > 0, "eval_record: sizes must be positive");
  NAFAE_REQUIRE(D_ind && D_sim && entities_length && rois, "eval_record: NULL input");
  NAFAE_REQUIRE((long long)num_segments * Ns * Ne < (1ll << 31), "eval_record: too many slots");
  NAFAE_REQUIRE(gt_boxes == nullptr || (gt_classes && class_match && class_count && num_classes > 0),
                "eval_record: ground truth needs gt_classes, class_match, class_count, num_classes");
  const int n = num_segments * Ns * Ne;
  int grid = ceil_div(n, 256);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  eval_record_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const long long*>(D_ind), D_sim, entities_length, rois, num_segments, Ns, Nb, Ne,
      (long long)image_id_base, reinterpret_cast<long long*>(out_image_ids),
      reinterpret_cast<long long*>(out_box_rows), out_boxes, out_confs, gt_boxes, gt_classes, gt_thr,
      num_classes, class_match, class_count);
  return launch_status("eval_record_kernel");
}

// Same contract as nafae_ground_forward; the region x query contraction runs on the tensor cores
// (tf32 x 3, fp32-rechecked picks), P2 / P3 follow in a second launch.
NAFAE_API int nafae_ground_forward_tc(const float* vis_feats, const float* word_feats,
                                      const int* entities_length, int Na, int Ns, int Nb, int Ne, int D,
                                      float Delta, float vis_lam, int train, int64_t* D_ind, float* D_sim,
                                      float* margin_loss, void* workspace, size_t workspace_bytes,
                                      cudaStream_t stream) {
  Dims d;
  NAFAE_REQUIRE(make_dims(Na, Ns, Nb, Ne, D, &d), "ground: sizes must be positive");
  if (check_ground_args(d, workspace, workspace_bytes, train) != 1) return 0;
  NAFAE_REQUIRE(vis_feats && word_feats && entities_length && D_ind && D_sim && margin_loss,
                "ground: NULL buffer");
  NAFAE_REQUIRE(((reinterpret_cast<uintptr_t>(vis_feats) | reinterpret_cast<uintptr_t>(word_feats)) & 15) == 0,
                "ground_tc: vis_feats / word_feats must be 16-byte aligned");
  NAFAE_REQUIRE(Nb <= 128, "ground_tc: at most 128 boxes per frame (got %d); use nafae_ground_forward", Nb);
  NAFAE_REQUIRE(D % 4 == 0, "ground_tc: D must be a multiple of 4");
  P1TcParams q;
  q.vis = vis_feats;
  q.word = word_feats;
  q.lens = entities_length;
  q.D_ind = reinterpret_cast<long long*>(D_ind);
  q.D_sim = D_sim;
  q.d = d;
  q.fpt = 128 / Nb;
  int nqt = (d.NQ + 15) / 16 * 16;  // columns per tile: 3 ring stages of [A hi|A lo|B hi|B lo] must fit
  if (nqt > 128) nqt = 128;
  q.nqt = nqt;
  CUtensorMap mv, mw;
  if (!tc05::make_tensor_map_2d(&mv, vis_feats, 4, false, (long long)d.F * Nb, D, 128)) return 0;
  if (!tc05::make_tensor_map_2d(&mw, word_feats, 4, false, d.NQ, D, nqt)) return 0;
  const size_t ring = (size_t)kTcStages * 2 * (128 * 128 + (size_t)nqt * 128);
  const size_t stile = (size_t)128 * (nqt + 1) * 4;
  const size_t smem_tc = 1024 + (ring > stile ? ring : stile) + 256;
  NAFAE_REQUIRE(smem_tc <= 227 * 1024, "ground_tc: shared memory");
  cudaError_t e = cudaFuncSetAttribute(ground_p1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_tc);
  if (e != cudaSuccess) {
    set_error("ground_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return -(int)e;
  }
  dim3 grid(ceil_div(d.F, q.fpt), ceil_div(d.NQ, nqt));
  ground_p1_tc_kernel<<<grid, kTcThreads, smem_tc, stream>>>(mv, mw, q);
  int st = launch_status("ground_p1_tc_kernel");
  if (st != 1) return st;
  FwdParams p;
  p.vis = vis_feats;
  p.word = word_feats;
  p.lens = entities_length;
  p.D_ind = reinterpret_cast<long long*>(D_ind);
  p.D_sim = D_sim;
  p.loss = margin_loss;
  p.ws = workspace;
  p.d = d;
  p.Delta = Delta;
  p.vis_lam = vis_lam;
  p.train = train ? 1 : 0;
  p.col_chunks = 1;
  p.groups = 1;
  p.ws_group_bytes = 0;
  size_t smem = (size_t)kRowTile * d.D * 4;
  const size_t p2 = ((size_t)(train ? (d.Nb < kRowTile ? d.Nb : kRowTile) : 0) * d.D +
                     (size_t)2 * d.Ns * d.NQ + 2 * (size_t)d.NQ + (size_t)(d.Nb > kRowTile ? d.Nb : kRowTile) +
                     (size_t)kRowTile * kRowTile) * 4;
  const size_t p3 = (size_t)d.Na * d.Ns * d.Na * 4;
  if (p2 > smem) smem = p2;
  if (p3 > smem) smem = p3;
  NAFAE_REQUIRE(smem <= 200 * 1024, "ground: D=%d / Ns=%d need too much shared memory", d.D, d.Ns);
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(ground_p23_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("ground: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  ground_p23_kernel<<<d.Na, kFwdThreads, smem, stream>>>(p);
  return launch_status("ground_p23_kernel");
}
