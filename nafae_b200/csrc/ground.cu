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
// Backward = a clustering-gradient kernel (train only) + one kernel with a CTA per frame
// (dL/dvis rows, dense overwrite) and a CTA per query column (dL/dword): dL/dS_ has at most one
// non-zero per (frame, column) -- the argmax box -- so both are gather-scale-accumulate sweeps,
// not GEMMs.
//
// Arithmetic: fp32 FMA, fp32 accumulate.  The contraction is 85 MFLOP at the benchmark shape and
// latency bound; see DESIGN.md for why it runs on the FMA pipe rather than tcgen05 tiles.
#include "common.cuh"

namespace nafae {
namespace {

constexpr float kEps = 1e-5f;  // model.py:33
constexpr int kFwdThreads = 256;
constexpr int kFwdWarps = kFwdThreads / 32;
constexpr int kColsPerCta = 32;  // 8 warps x 4 columns
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
__device__ __forceinline__ float warp_dot(const float* __restrict__ x, const float* __restrict__ y,
                                          int D, int lane) {
  float acc = 0.f;
  for (int k0 = 0; k0 < D; k0 += 256) {
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + lane + 32 * j;
      a[j] = k < D ? __ldg(x + k) : 0.f;
      b[j] = k < D ? __ldg(y + k) : 0.f;
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
};

// 4 per-lane partial sums -> lane group (lane>>3) holds the total of column (lane>>3)
__device__ __forceinline__ float reduce4(float a0, float a1, float a2, float a3, int lane) {
  // xor 16: lanes 0-15 keep columns 0,1; lanes 16-31 keep columns 2,3
  const bool hi = lane & 16;
  float k0 = hi ? a2 : a0, k1 = hi ? a3 : a1;
  float s0 = hi ? a0 : a2, s1 = hi ? a1 : a3;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  // xor 8: within each half, lanes with bit3 clear keep the first column
  const bool hi2 = lane & 8;
  float k = hi2 ? k1 : k0, s = hi2 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, s, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;  // column index = (lane >> 3): 0,1 for lanes 0-15 ; 2,3 for lanes 16-31
}

__device__ void phase2_segment(const FwdParams& p, const Ws& w, int a, float* sm);
__device__ void phase3_final(const FwdParams& p, const Ws& w, float* sm);

__global__ void __launch_bounds__(kFwdThreads) ground_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(16) float sm[];  // kRowTile * D floats (vis rows of the frame)
  __shared__ int s_ticket;
  const Dims& d = p.d;
  const Ws w = ws_carve(p.ws, d);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int f = blockIdx.x / p.col_chunks, chunk = blockIdx.x % p.col_chunks;
  const int a = f / d.Ns;

  // ---- P1: one frame x up to 32 columns; each warp owns 4 consecutive columns
  const int c_base = chunk * kColsPerCta + warp * 4;
  bool live[4];
  bool any_live = false;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c_base + j;
    live[j] = c < d.NQ && (c % d.Ne) < __ldg(p.lens + c / d.Ne);  // model.py:535-536
    any_live |= live[j];
  }
  const int my_col = c_base + (lane >> 3);  // column this lane reports after reduce4
  float best = -INFINITY;
  int best_r = 0;
  const float* vis_f = p.vis + (size_t)f * d.Nb * d.D;

  __shared__ __align__(8) uint64_t s_bar;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  uint32_t parity = 0;
  for (int r0 = 0; r0 < d.Nb; r0 += kRowTile) {
    const int rows = min(kRowTile, d.Nb - r0);
    __syncthreads();  // previous tile consumed (and the barrier init is visible)
    if (tid == 0) {   // one bulk async copy (TMA engine) stages the whole row tile
      const uint32_t bytes = (uint32_t)rows * d.D * 4u;
      mbar_arrive_expect_tx(&s_bar, bytes);
      bulk_g2s(sm, vis_f + (size_t)r0 * d.D, bytes, &s_bar);
    }
    bool waited = false;
    for (int k0 = 0; k0 < d.D; k0 += 32 * kKS) {
      // this chunk's word slices for the warp's 4 columns (overlaps the tile copy)
      float wv[4][kKS];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* wr = p.word + (size_t)min(c_base + j, d.NQ - 1) * d.D + k0;
#pragma unroll
        for (int i = 0; i < kKS; ++i) {
          const int k = lane + 32 * i;
          wv[j][i] = (live[j] && k0 + k < d.D) ? __ldg(wr + k) : 0.f;
        }
      }
      if (!waited) {
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        waited = true;
      }
      if (!any_live) break;
      const bool last_chunk = k0 + 32 * kKS >= d.D;
      for (int r = 0; r < rows; ++r) {
        const float* vr = sm + (size_t)r * d.D + k0;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < kKS; ++i) {
          const int k = lane + 32 * i;
          const float v = (k0 + k < d.D) ? vr[k] : 0.f;
          a0 = fmaf(v, wv[0][i], a0);
          a1 = fmaf(v, wv[1][i], a1);
          a2 = fmaf(v, wv[2][i], a2);
          a3 = fmaf(v, wv[3][i], a3);
        }
        float tot = reduce4(a0, a1, a2, a3, lane);
        if (d.D > 32 * kKS) {
          // multi-chunk D: partial sums are parked in shared memory after the row tile
          float* part = sm + (size_t)kRowTile * d.D + (size_t)(warp * 4 + (lane >> 3)) * kRowTile + r;
          if ((lane & 7) == 0) {
            if (k0 > 0) tot += *part;
            *part = tot;
          }
          __syncwarp();
          if (!last_chunk) continue;
          tot = *part;
        }
        if (tot > best) {  // strict '>': first maximal box wins (torch.max(dim) tie rule)
          best = tot;
          best_r = r0 + r;
        }
      }
    }
  }
  {
    const int c = my_col;
    if ((lane & 7) == 0 && c < d.NQ) {
      const bool is_live = (c % d.Ne) < __ldg(p.lens + c / d.Ne);
      // masked column: every entry is 0 after masked_fill_ -> max 0 at index 0 (model.py:551,612)
      p.D_sim[(size_t)f * d.NQ + c] = is_live ? best : 0.f;
      p.D_ind[(size_t)f * d.NQ + c] = is_live ? (long long)best_r : 0ll;
    }
  }

  // ---- chain: last CTA of the segment runs P2, last segment runs P3
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(w.seg_cnt + a, 1);
  __syncthreads();
  if (s_ticket != d.Ns * p.col_chunks - 1) return;
  __threadfence();
  phase2_segment(p, w, a, sm);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    w.seg_cnt[a] = 0;  // leave the counter clean for the next launch
    s_ticket = atomicAdd(w.done_cnt, 1);
  }
  __syncthreads();
  if (s_ticket != d.Na - 1) return;
  __threadfence();
  phase3_final(p, w, sm);
  if (tid == 0) *w.done_cnt = 0;
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

// P2: frame attention + Sf for segment a; clustering partials (train).
// Everything the phase needs from other CTAs (the segment's Ns x NQ block of D_sim / D_ind) is
// pulled into shared memory with ONE round of independent L2 loads.
__device__ void phase2_segment(const FwdParams& p, const Ws& w, int a, float* sm) {
  const Dims& d = p.d;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int blk = d.Ns * d.NQ;
  float* Sblk = sm;                                   // [Ns][NQ]
  int* Iblk = reinterpret_cast<int*>(sm + blk);       // [Ns][NQ]
  float* c_mn = sm + 2 * blk;                         // [NQ]
  float* c_iden = c_mn + d.NQ;                        // [NQ] 1/den
  float* inv = c_iden + d.NQ;                         // [Ne][Ns] 1/(||x||+eps) of the picked rows
  __syncthreads();
  const float* gS = p.D_sim + (size_t)a * blk;
  const long long* gI = p.D_ind + (size_t)a * blk;
  for (int i = tid; i < blk; i += kFwdThreads) {
    Sblk[i] = __ldcg(gS + i);
    Iblk[i] = (int)__ldcg(gI + i);
  }
  __syncthreads();
  for (int c = tid; c < d.NQ; c += kFwdThreads) {
    const ColStat st = col_stat_smem(Sblk, c, d.Ns, d.NQ);
    c_mn[c] = st.mn;
    c_iden[c] = 1.f / st.den;
  }
  __syncthreads();
  // Sf[a,s,a'] = sum_e S*S_att / div[a']   (model.py:583-593); thread per (s, a')
  for (int i = tid; i < d.Ns * d.Na; i += kFwdThreads) {
    const int s = i / d.Na, a2 = i % d.Na;
    const int len = __ldg(p.lens + a2);
    float acc = 0.f;
    for (int e = 0; e < len; ++e) {
      const int c = a2 * d.Ne + e;
      const float x = Sblk[s * d.NQ + c];
      acc += x * ((x - c_mn[c]) * c_iden[c]);
    }
    w.Sf[((size_t)a * d.Ns + s) * d.Na + a2] = acc / (float)(len == 0 ? 1 : len);
  }
  if (!p.train) return;

  // clustering loss of segment a (model.py:553-577).  Rows come from frame 0 of segment 0: the
  // box index is used without its (segment, frame) offset.
  const int len = __ldg(p.lens + a);
  for (int i = warp; i < len * d.Ns; i += kFwdWarps) {  // norms, warp per picked row
    const int e = i / d.Ns, s = i % d.Ns;
    const float* x = p.vis + (size_t)Iblk[s * d.NQ + a * d.Ne + e] * d.D;
    const float acc = warp_dot(x, x, d.D, lane);
    if (lane == 0) inv[i] = 1.f / (sqrtf(acc) + kEps);
  }
  __syncthreads();
  // Gram entries s<t of every entity; G is symmetric so the ordered sum / count double
  float gsum = 0.f;
  int gcnt = 0;
  const int pairs = d.Ns * (d.Ns - 1) / 2;
  for (int i = warp; i < len * pairs; i += kFwdWarps) {
    const int e = i / pairs;
    int q = i % pairs, s = 0;
    while (q >= d.Ns - 1 - s) {
      q -= d.Ns - 1 - s;
      ++s;
    }
    const int t = s + 1 + q;
    const int c = a * d.Ne + e;
    const float* xs = p.vis + (size_t)Iblk[s * d.NQ + c] * d.D;
    const float* xt = p.vis + (size_t)Iblk[t * d.NQ + c] * d.D;
    const float acc = warp_dot(xs, xt, d.D, lane);
    const float ss = (Sblk[s * d.NQ + c] - c_mn[c]) * c_iden[c];
    const float st = (Sblk[t * d.NQ + c] - c_mn[c]) * c_iden[c];
    const float dot = acc * (ss * inv[e * d.Ns + s]) * (st * inv[e * d.Ns + t]);
    const float g = 1.f - dot;
    if (lane == 0) {
      gsum += 2.f * g;
      gcnt += (g != 0.f) ? 2 : 0;
    }
  }
  __shared__ float s_gsum[kFwdWarps];
  __shared__ int s_gcnt[kFwdWarps];
  if (lane == 0) {
    s_gsum[warp] = gsum;
    s_gcnt[warp] = gcnt;
  }
  __syncthreads();
  if (tid == 0) {
    float ts = 0.f;
    int tc = 0;
    for (int k = 0; k < kFwdWarps; ++k) {
      ts += s_gsum[k];
      tc += s_gcnt[k];
    }
    w.vsum[a] = ts;
    w.vcnt[a] = tc;
  }
}

// P3: hinge ranking loss over all segments (model.py:594-606) and its gradient w.r.t. Sf
__device__ void phase3_final(const FwdParams& p, const Ws& w, float* sm) {
  const Dims& d = p.d;
  const int tid = threadIdx.x;
  const int n = d.Na * d.Ns * d.Na;
  __shared__ float s_part[kFwdThreads];
  float* Sf = sm;       // [Na][Ns][Na]
  float* hg = sm + n;   // [Na][Ns][Na]
  __syncthreads();
  for (int i = tid; i < n; i += kFwdThreads) {
    Sf[i] = __ldcg(w.Sf + i);
    hg[i] = 0.f;
  }
  __syncthreads();
  // frame_score[a,s] = mean_a'' relu(Sf[a'',s,a] - d[a,s] + Delta) + mean_a' relu(Sf[a,s,a'] - d[a,s] + Delta)
  float part = 0.f;
  const float inv_na = 1.f / (float)d.Na;
  const float gscale = 10.f / (float)(d.Na * d.Ns) * inv_na;  // d(margin)/d(relu term)
  for (int i = tid; i < d.Na * d.Ns; i += kFwdThreads) {
    const int a = i / d.Ns, s = i % d.Ns;
    const float dg = Sf[(a * d.Ns + s) * d.Na + a];
    float t1 = 0.f, t2 = 0.f;
    float gd = 0.f;  // gradient reaching d[a,s]
    for (int o = 0; o < d.Na; ++o) {
      const float v1 = (Sf[(o * d.Ns + s) * d.Na + a] - dg) + p.Delta;
      const float v2 = (Sf[(a * d.Ns + s) * d.Na + o] - dg) + p.Delta;
      if (v1 > 0.f) {
        t1 += v1;
        atomicAdd(hg + (o * d.Ns + s) * d.Na + a, gscale);
        gd -= gscale;
      }
      if (v2 > 0.f) {
        t2 += v2;
        atomicAdd(hg + (a * d.Ns + s) * d.Na + o, gscale);
        gd -= gscale;
      }
    }
    atomicAdd(hg + (a * d.Ns + s) * d.Na + a, gd);
    part += t1 * inv_na + t2 * inv_na;
  }
  s_part[tid] = part;
  __syncthreads();
  for (int i = tid; i < n; i += kFwdThreads) w.hgrad[i] = hg[i];
  if (tid < 32) {  // deterministic tree over the 256 partials
    float v = 0.f;
    for (int k = tid; k < kFwdThreads; k += 32) v += s_part[k];
    v = warp_sum(v);
    if (tid == 0) {
      const float mean_fs = v / (float)(d.Na * d.Ns);
      float loss = mean_fs * 10.f;
      float vis_loss = 0.f, dem = 0.f;
      if (p.train) {
        float ts = 0.f;
        int tc = 0;
        for (int a = 0; a < d.Na; ++a) {
          ts += __ldcg(w.vsum + a);
          tc += __ldcg(w.vcnt + a);
        }
        dem = (float)tc;
        vis_loss = ts / dem;  // NaN when nothing is unmasked, like the reference (model.py:576-577)
        loss = (mean_fs + p.vis_lam * vis_loss) * 10.f;
      }
      w.scal[0] = vis_loss;
      w.scal[1] = dem;
      w.scal[2] = mean_fs;
      *p.loss = loss;
    }
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

constexpr int kBwdThreads = 256;
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

// grid F + NQ.  blockIdx < F: dL/dvis rows of frame f (dense overwrite, Nb x D).
//               else        : dL/dword row of column c.
// Every cross-CTA input is staged into shared memory with one round of independent loads; the
// gather loops issue their global loads in batches so they overlap instead of serialising.
__global__ void __launch_bounds__(kBwdThreads) ground_bwd_main_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int s_nlive;
  const Dims& d = p.d;
  const Ws w = ws_carve(p.ws, d);
  const int tid = threadIdx.x, lane = tid & 31;
  const float gout = __ldg(p.gout);
  if ((int)blockIdx.x < d.F) {
    const int f = blockIdx.x, a = f / d.Ns, s_me = f % d.Ns;
    const int blk = d.Ns * d.NQ;
    float* Sblk = sm;                                  // [Ns][NQ]
    float* hg = Sblk + blk;                            // [Ns][Na]
    float* gtmp = hg + d.Ns * d.Na;                    // [Ns][NQ] per-column gradients
    int* ridx = reinterpret_cast<int*>(gtmp + blk);    // [NQ]
    int* live = ridx + d.NQ;                           // [NQ] compact list of live columns
    // [kRowTile][D], 16-byte aligned for the float4 write-out
    float* acc = sm + ((2 * (size_t)blk + (size_t)d.Ns * d.Na + 2 * (size_t)d.NQ + 3) & ~(size_t)3);
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
        col_grad_smem(Sblk + c, d.NQ, hg + a2, d.Na, d.Ns, gout / (float)(len == 0 ? 1 : len),
                      gtmp + c, d.NQ);
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
    const float* g = gtmp + s_me * d.NQ;
    for (int r0 = 0; r0 < d.Nb; r0 += kRowTile) {
      const int rows = min(kRowTile, d.Nb - r0);
      for (int i = tid; i < rows * d.D; i += kBwdThreads) acc[i] = 0.f;
      __syncthreads();
      for (int k = tid; k < d.D; k += kBwdThreads) {  // thread owns feature k of every row
        for (int j0 = 0; j0 < nlive; j0 += 8) {
          float wv[8];
          int rr[8];
          float gg[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = live[min(j0 + j, nlive - 1)];
            rr[j] = (j0 + j < nlive) ? ridx[c] - r0 : -1;
            gg[j] = g[c];
            wv[j] = __ldg(p.word + (size_t)c * d.D + k);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (rr[j] >= 0 && rr[j] < rows) acc[rr[j] * d.D + k] = fmaf(gg[j], wv[j], acc[rr[j] * d.D + k]);
        }
      }
      __syncthreads();
      float4* dst = reinterpret_cast<float4*>(p.gvis + ((size_t)f * d.Nb + r0) * d.D);
      const float4* src = reinterpret_cast<const float4*>(acc);
      for (int i = tid; i < rows * d.D / 4; i += kBwdThreads) dst[i] = src[i];
      __syncthreads();
    }
  } else {
    const int c = blockIdx.x - d.F, a2 = c / d.Ne;
    float* dst = p.gword + (size_t)c * d.D;
    const int len = __ldg(p.lens + a2);
    if ((c % d.Ne) >= len) {
      for (int k = tid; k < d.D; k += kBwdThreads) dst[k] = 0.f;
      return;
    }
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
      for (int f0 = 0; f0 < d.F; f0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(p.vis + (size_t)ridx[min(f0 + j, d.F - 1)] * d.D + k);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (f0 + j < d.F) accv = fmaf(g[f0 + j], v[j], accv);
      }
      dst[k] = accv;
    }
  }
}

// Clustering-loss gradient, launched AFTER ground_bwd_main_kernel: adds into grad_vis rows
// 0..Nb-1 (the rows the reference's un-offset index_select gathers).
// grid NQ, CTA per (a, e); masked entities exit immediately.
__global__ void __launch_bounds__(kBwdThreads) ground_bwd_cluster_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float sm[];  // Vsum[D] | simn[Ns] | inv[Ns] | dots[Ns] | row[Ns]
  const Dims& d = p.d;
  const Ws w = ws_carve(p.ws, d);
  const int c = blockIdx.x, a = c / d.Ne, e = c % d.Ne;
  if (e >= __ldg(p.lens + a)) return;
  const float dem = __ldg(w.scal + 1);
  if (!(dem > 0.f)) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* Vsum = sm;
  float* simn = sm + d.D;
  float* inv = simn + d.Ns;
  float* dots = inv + d.Ns;
  int* row = reinterpret_cast<int*>(dots + d.Ns);
  for (int s = tid; s < d.Ns; s += kBwdThreads) {
    simn[s] = __ldg(p.D_sim + (size_t)(a * d.Ns + s) * d.NQ + c);
    row[s] = (int)__ldg(p.D_ind + (size_t)(a * d.Ns + s) * d.NQ + c);
  }
  __syncthreads();
  float mn = INFINITY, mx = -INFINITY;
  for (int s = 0; s < d.Ns; ++s) {
    mn = fminf(mn, simn[s]);
    mx = fmaxf(mx, simn[s]);
  }
  const float iden = 1.f / ((mx - mn) + kEps);
  __syncthreads();
  for (int s = tid; s < d.Ns; s += kBwdThreads) simn[s] = (simn[s] - mn) * iden;
  for (int s = warp; s < d.Ns; s += kBwdThreads / 32) {
    const float* x = p.vis + (size_t)row[s] * d.D;
    const float acc = warp_dot(x, x, d.D, lane);
    if (lane == 0) inv[s] = 1.f / (sqrtf(acc) + kEps);
  }
  __syncthreads();
  for (int k = tid; k < d.D; k += kBwdThreads) {
    float acc = 0.f;
    for (int s0 = 0; s0 < d.Ns; s0 += 8) {
      float xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xv[j] = __ldg(p.vis + (size_t)row[min(s0 + j, d.Ns - 1)] * d.D + k);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (s0 + j < d.Ns) acc += simn[s0 + j] * inv[s0 + j] * xv[j];
    }
    Vsum[k] = acc;
  }
  __syncthreads();
  // vis_loss = sum(G)/dem, G[s,t] = 1 - V_s.V_t (s != t): d/dV_s = -2 (Vsum - V_s) / dem
  const float coef = -2.f * (10.f * p.vis_lam * __ldg(p.gout)) / dem;
  for (int s = warp; s < d.Ns; s += kBwdThreads / 32) {  // dots[s] = x_s . gu_s
    const float* x = p.vis + (size_t)row[s] * d.D;
    const float sc = simn[s] * inv[s];
    float acc = 0.f;
    for (int k0 = 0; k0 < d.D; k0 += 256) {
      float xv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + lane + 32 * j;
        xv[j] = k < d.D ? __ldg(x + k) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + lane + 32 * j;
        if (k < d.D) acc = fmaf(xv[j], Vsum[k] - sc * xv[j], acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) dots[s] = acc * simn[s] * coef;
  }
  __syncthreads();
  for (int k = tid; k < d.D; k += kBwdThreads) {
    for (int s0 = 0; s0 < d.Ns; s0 += 8) {
      float xs[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xs[j] = __ldg(p.vis + (size_t)row[min(s0 + j, d.Ns - 1)] * d.D + k);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int s = s0 + j;
        if (s >= d.Ns) break;
        const float iv = inv[s];             // 1/(n+eps)
        const float nrm = 1.f / iv - kEps;   // n
        const float sc = simn[s] * iv;
        // dL/dx = gu/(n+eps) - x (x.gu) / (n (n+eps)^2);  gu = simn*coef*(Vsum - V_s)
        const float c2 = nrm > 0.f ? dots[s] * iv * iv / nrm : 0.f;
        const float gu = simn[s] * coef * (Vsum[k] - sc * xs[j]);
        atomicAdd(p.gvis + (size_t)row[s] * d.D + k, gu * iv - xs[j] * c2);
      }
    }
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

NAFAE_API size_t nafae_ground_workspace_bytes(int Na, int Ns, int Nb, int Ne, int D) {
  Dims d;
  if (!make_dims(Na, Ns, Nb, Ne, D, &d)) return 0;
  return align_up(ws_words(d) * 4, 256);
}

static int check_ground_args(const Dims& d, const void* ws, size_t ws_bytes, int train) {
  NAFAE_REQUIRE(d.D % 4 == 0, "ground: D must be a multiple of 4, got %d", d.D);
  NAFAE_REQUIRE(d.Ne <= 16, "ground: max_ent_len %d > 16 not supported", d.Ne);
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
  Dims d;
  NAFAE_REQUIRE(make_dims(Na, Ns, Nb, Ne, D, &d), "ground: sizes must be positive");
  if (check_ground_args(d, workspace, workspace_bytes, train) != 1) return 0;
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
  p.col_chunks = ceil_div(d.NQ, kColsPerCta);
  // shared memory: row tile (+ parked partial sums when D > 512), reused by P2's small tables
  size_t smem = (size_t)kRowTile * d.D * 4;
  if (d.D > 32 * kKS) smem += (size_t)kFwdWarps * 4 * kRowTile * 4;
  const size_t p2 = ((size_t)2 * d.Ns * d.NQ + 2 * (size_t)d.NQ + (size_t)d.Ne * d.Ns) * 4;
  const size_t p3 = (size_t)2 * d.Na * d.Ns * d.Na * 4;
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
  ground_fwd_kernel<<<d.F * p.col_chunks, kFwdThreads, smem, stream>>>(p);
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
  NAFAE_REQUIRE(grad_margin_loss && vis_feats && word_feats && entities_length && D_ind && D_sim &&
                    grad_vis && grad_word,
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
  size_t smem = ((size_t)2 * d.Ns * d.NQ + (size_t)d.Ns * d.Na + 2 * (size_t)d.NQ +
                 (size_t)kRowTile * d.D + 4) * 4;
  const size_t smem_w = (size_t)4 * d.F * 4;
  if (smem_w > smem) smem = smem_w;
  NAFAE_REQUIRE(smem <= 200 * 1024, "ground backward: sizes need too much shared memory");
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ground_bwd_main_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("ground: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  ground_bwd_main_kernel<<<d.F + d.NQ, kBwdThreads, smem, stream>>>(p);
  int st = launch_status("ground_bwd_main_kernel");
  if (st != 1 || !p.train) return st;
  const size_t smem_c = ((size_t)d.D + 4 * d.Ns) * 4;
  ground_bwd_cluster_kernel<<<d.NQ, kBwdThreads, smem_c, stream>>>(p);
  return launch_status("ground_bwd_cluster_kernel");
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
