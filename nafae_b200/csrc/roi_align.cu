// Corner-grid RoIAlign (+ fused 2x2/stride-1 avg or max post pool) for sm_100a.
//
// Replaces (reference tree paths):
//   lib/model/roi_align/src/roi_align_kernel.cu:15-70    ROIAlignForward
//   lib/model/roi_align/src/roi_align_kernel.cu:94-143   ROIAlignBackward
//   lib/model/roi_align/src/roi_align_kernel.cu:73-91,145-162  launchers (same symbols exported)
//   lib/model/roi_align/modules/roi_align.py:26-29,39-42 RoIAlignAvg / RoIAlignMax (= kernel +
//                                                        avg_pool2d / max_pool2d(2, 1))
//
// Semantics that are kept (they differ from torchvision's roi_align): the sampled points form
// a corner grid with bin = max(end-start+1, 0)/(aligned-1); a sample outside [0,H)x[0,W) is 0;
// cells are hstart = min(floor(h), H-2), so for h in [H-1, H) the weight h-hstart is in [1,2):
// linear EXTRAPOLATION, not clamping (roi_align_kernel.cu:48-49, 57-67).
//
// Two arithmetic modes share one geometry routine (bit-identical in/out decisions and cells):
//   exact : the reference's mixed fp32/fp64 expression exactly as nvcc compiles it (checked in
//           SASS: which partial products are float, which DFMAs are fused) -> bit-identical output
//   fast  : fp32 FMAs on the same taps -> |err| ~ 1e-7 relative, used by the bandwidth kernel
//
// Kernels:
//   align_fwd_generic / align_bwd_generic   any shape, exact or fast; the *Laucher symbols
//   align_pool_fwd_slab                     the hot path (7x7 out, 8x8 samples, avg or max):
//       persistent CTAs, one (frame, channel-group) slab of the NCHW map at a time, staged into
//       shared memory by 1-D bulk async copies (TMA engine, cp.async.bulk + mbarrier) in a
//       3-stage ring; every feature byte is read from HBM exactly once, all RoIs of the frame
//       are served from shared memory, the (R,C,8,8) intermediate never exists, and the 2x2
//       pool is a register/shuffle epilogue.
#include <cuda_bf16.h>

#include <type_traits>

#include "roi_geom.cuh"

namespace nafae {
namespace {

// roi_align_kernel.cu:64-67 as compiled (see header comment)
__device__ __forceinline__ float interp_exact(float ul, float ur, float dl, float dr, float hr,
                                              float wr) {
  const double omh = __dsub_rn(1., (double)hr);
  const double omw = __dsub_rn(1., (double)wr);
  const double t2 = __dmul_rn(__dmul_rn((double)ur, omh), (double)wr);
  double s = __fma_rn(__dmul_rn((double)ul, omh), omw, t2);
  s = __fma_rn(omw, (double)__fmul_rn(dl, hr), s);
  s = __dadd_rn(s, (double)__fmul_rn(__fmul_rn(dr, hr), wr));
  return __double2float_rn(s);
}

__device__ __forceinline__ float interp_fast(float ul, float ur, float dl, float dr, float hr,
                                             float wr) {
  const float omw = 1.f - wr;
  const float top = fmaf(ur, wr, ul * omw);
  const float bot = fmaf(dr, wr, dl * omw);
  return fmaf(bot, hr, top * (1.f - hr));
}

template <bool EXACT>
__device__ __forceinline__ float sample_point(const float* __restrict__ plane, int W,
                                              const RoiGeom& g, int ph, int pw, int H) {
  int hc, wc;
  float hr, wr;
  const bool okh = axis_sample(g.start_h, g.bin_h, ph, H, &hc, &hr);
  const bool okw = axis_sample(g.start_w, g.bin_w, pw, W, &wc, &wr);
  if (!(okh && okw)) return 0.f;
  const float* p = plane + hc * W + wc;
  const float ul = __ldg(p), ur = __ldg(p + 1), dl = __ldg(p + W), dr = __ldg(p + W + 1);
  return EXACT ? interp_exact(ul, ur, dl, dr, hr, wr) : interp_fast(ul, ur, dl, dr, hr, wr);
}

// ATen's 2x2 window reductions, row-major order (avg: sum from 0 then /4; max: v > m || isnan)
__device__ __forceinline__ float pool4(int mode, float a, float b, float c, float d) {
  if (mode == NAFAE_POOL_AVG)
    return __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d), 0.25f);
  float m = -INFINITY;
  if (a > m || a != a) m = a;
  if (b > m || b != b) m = b;
  if (c > m || c != c) m = c;
  if (d > m || d != d) m = d;
  return m;
}
// index (0..3) ATen's max_pool2d reports for the window: first maximum, NaN wins, default 0
__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
  float m = -INFINITY;
  int k = 0;
  if (a > m || a != a) { m = a; k = 0; }
  if (b > m || b != b) { m = b; k = 1; }
  if (c > m || c != c) { m = c; k = 2; }
  if (d > m || d != d) { m = d; k = 3; }
  return k;
}

// ------------------------------------------------------------------ generic kernels ----
// One thread per output element (n, c, oh, ow).  POOL none: one sample.  avg/max: the 2x2 window
// of the (oh+1)x(ow+1) sample grid.  RoIs whose batch index is outside [0, B) produce zeros.
template <bool EXACT>
__global__ void __launch_bounds__(256)
align_fwd_generic(const float* __restrict__ bottom, float scale, int B, long long total, int H,
                  int W, int C, int oh_n, int ow_n, int pool, const float* __restrict__ rois,
                  float* __restrict__ top) {
  const int sh = pool ? oh_n + 1 : oh_n, sw = pool ? ow_n + 1 : ow_n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(idx % ow_n);
    const int oh = (int)((idx / ow_n) % oh_n);
    const int c = (int)((idx / ow_n / oh_n) % C);
    const int n = (int)(idx / ow_n / oh_n / C);
    const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, sh, sw);
    float v = 0.f;
    if (g.batch >= 0 && g.batch < B) {
      const float* plane = bottom + ((size_t)g.batch * C + c) * H * W;
      if (pool == NAFAE_POOL_NONE) {
        v = sample_point<EXACT>(plane, W, g, oh, ow, H);
      } else {
        const float a = sample_point<EXACT>(plane, W, g, oh, ow, H);
        const float b = sample_point<EXACT>(plane, W, g, oh, ow + 1, H);
        const float cc = sample_point<EXACT>(plane, W, g, oh + 1, ow, H);
        const float d = sample_point<EXACT>(plane, W, g, oh + 1, ow + 1, H);
        v = pool4(pool, a, b, cc, d);
      }
    }
    top[idx] = v;
  }
}

// One thread per SAMPLE-grid element (n, c, ph, pw): gathers the gradient that reaches the sample
// through the pool (avg: every containing window's g/4, ascending window order like ATen's
// avg_pool2d backward; max: windows whose ATen argmax is this sample), then scatters it to the
// four taps with the reference's weights (roi_align_kernel.cu:137-140 as compiled).
template <bool EXACT>
__global__ void __launch_bounds__(256)
align_bwd_generic(const float* __restrict__ top_diff, const float* __restrict__ bottom,
                  float scale, int B, long long total, int H, int W, int C, int oh_n, int ow_n,
                  int pool, const float* __restrict__ rois, float* __restrict__ bottom_diff) {
  const int sh = pool ? oh_n + 1 : oh_n, sw = pool ? ow_n + 1 : ow_n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % sw);
    const int ph = (int)((idx / sw) % sh);
    const int c = (int)((idx / sw / sh) % C);
    const int n = (int)(idx / sw / sh / C);
    const RoiGeom g = roi_geom(rois + (size_t)n * 5, scale, sh, sw);
    if (g.batch < 0 || g.batch >= B) continue;
    int hc, wc;
    float hr, wr;
    const bool okh = axis_sample(g.start_h, g.bin_h, ph, H, &hc, &hr);
    const bool okw = axis_sample(g.start_w, g.bin_w, pw, W, &wc, &wr);
    if (!(okh && okw)) continue;
    const float* td = top_diff + ((size_t)n * C + c) * oh_n * ow_n;
    float gs = 0.f;
    if (pool == NAFAE_POOL_NONE) {
      gs = td[ph * ow_n + pw];
    } else {
      const float* plane = bottom ? bottom + ((size_t)g.batch * C + c) * H * W : nullptr;
      for (int i = max(ph - 1, 0); i <= min(ph, oh_n - 1); ++i)
        for (int j = max(pw - 1, 0); j <= min(pw, ow_n - 1); ++j) {
          const float gy = td[i * ow_n + j];
          if (pool == NAFAE_POOL_AVG) {
            gs = __fadd_rn(gs, __fmul_rn(gy, 0.25f));
          } else {
            const float a = sample_point<EXACT>(plane, W, g, i, j, H);
            const float b = sample_point<EXACT>(plane, W, g, i, j + 1, H);
            const float cc = sample_point<EXACT>(plane, W, g, i + 1, j, H);
            const float d = sample_point<EXACT>(plane, W, g, i + 1, j + 1, H);
            const int k = argmax4(a, b, cc, d);
            if (i + (k >> 1) == ph && j + (k & 1) == pw) gs = __fadd_rn(gs, gy);
          }
        }
    }
    float* p = bottom_diff + ((size_t)g.batch * C + c) * H * W + hc * W + wc;
    float g1, g2, g3, g4;
    if (EXACT) {
      const double a = __dmul_rn((double)gs, __dsub_rn(1., (double)hr));
      const float omw = __fsub_rn(1.f, wr);  // "(1 - w_ratio)" is int - float
      g1 = __double2float_rn(__dmul_rn(a, (double)omw));
      g2 = __double2float_rn(__dmul_rn(a, (double)wr));
      const float bq = __fmul_rn(gs, hr);
      g3 = __fmul_rn(bq, omw);
      g4 = __fmul_rn(bq, wr);
    } else {
      const float a = gs * (1.f - hr), bq = gs * hr, omw = 1.f - wr;
      g1 = a * omw;
      g2 = a * wr;
      g3 = bq * omw;
      g4 = bq * wr;
    }
    atomicAdd(p, g1);
    atomicAdd(p + 1, g2);
    atomicAdd(p + W, g3);
    atomicAdd(p + W + 1, g4);
  }
}

// --------------------------------------------------------------- bandwidth kernel ----
// 7x7 output from an 8x8 sample grid (RoIAlignAvg/Max(7, 7, s): the only configuration the
// reference instantiates, faster_rcnn/rpn.py:34).
//
// Persistent CTAs (one per SM).  Work unit = (frame, group of cg channels): its NCHW slab is
// contiguous in HBM and is staged into shared memory by 1-D bulk async copies (TMA engine) through
// a ring of `stages` buffers with full/empty mbarriers: a dedicated producer warp keeps the ring
// full, 16 consumer warps drain it and never meet at a CTA-wide barrier except when the frame
// (and with it the RoI table) changes.  Pass-groups (RoI x 4*CPL channels) are dealt round-robin
// to the consumer warps across unit boundaries, so a 20-RoI frame keeps 16 warps evenly busy.
// Inside a pass a lane owns one sample column (pw) of CPL channels: the RoI's row offsets and
// weights are warp-uniform table reads, the taps are LDS with compile-time offsets (W and the
// padded channel stride are template constants for the two production map sizes), the 2x2 pool is
// one shuffle per sample row.  For POOL_AVG the 1/4 is folded into the column weights.
// The kernel is bound by the shared-memory / L1 data stage, not by arithmetic, so two things keep
// the number of wavefronts down: (a) a sample row whose cell rows equal (or follow by one) the
// previous sample row's reuses the horizontally interpolated rows it already holds (RoIs shorter
// than 7 cells -- half of them -- need ~6 row loads instead of 16); (b) the pooled 7x7 blocks of a
// pass (4*CPL consecutive channels of one RoI = 784*CPL contiguous bytes of the output) are staged
// in a per-warp shared buffer and leave with ONE bulk store (cp.async.bulk shared -> global)
// instead of 7*CPL scattered 28-byte-per-channel stores.
constexpr int kOut = 7;
constexpr int kS = 8;               // sample grid side
#ifndef NAFAE_CONS_WARPS
#define NAFAE_CONS_WARPS 16
#endif
#ifndef NAFAE_SLAB_PREDICATED
#define NAFAE_SLAB_PREDICATED 0
#endif
constexpr int kConsWarps = NAFAE_CONS_WARPS;
constexpr int kConsThreads = kConsWarps * 32;
constexpr int kSlabThreads = kConsThreads + 32;  // + producer warp
constexpr int kMaxRoiTable = 104;   // RoIs of one frame resident in the table at a time
constexpr int kStagesMax = 4;
constexpr int kPreFrames = 4;      // frames whose RoI tables a CTA may hold at once (prebuilt)

struct __align__(16) RoiEntry {  // 192 B: everything a pass needs about one RoI
  int hoff_b[kS];   // byte offset of row hstart (hstart*W*4); 0 if the sample row is outside
  float h0[kS];     // 1-h_ratio, 0 if outside
  float h1[kS];     // h_ratio,   0 if outside
  int woff_b[kS];   // byte offset of column wstart; 0 if the sample column is outside
  float w0[kS];     // (1-w_ratio) [* 1/4 for avg], 0 if outside
  float w1[kS];     // w_ratio     [* 1/4 for avg], 0 if outside
};

struct SlabParams {
  const float* bottom;
  const float* rois;
  void* top;       // (R, C, 7, 7) fp32, or bf16 (NAFAE_FLAG_OUT_BF16)
  float scale;
  int B, R, H, W, C;
  int cg;          // channels per slab
  int groups;      // C / cg
  int hw;          // H*W
  int hwp;         // padded per-channel stride in shared memory (floats), hwp % 32 == 8
  int stages;
  int units;       // B * groups
  int* gate;       // optional residency gate (nafae_gate_wait): [0] arrivals, [1] epoch
};

__device__ __forceinline__ void cons_barrier() {  // the 16 consumer warps only
  asm volatile("bar.sync 1, %0;" ::"n"(kConsThreads) : "memory");
}

template <int POOL, int W_CT, int HWP_CT, int CPL, int NBLK_CT, bool BF16>
__global__ void __launch_bounds__(kSlabThreads, 1) align_pool_fwd_slab(const SlabParams p) {
  using OutT = typename std::conditional<BF16, __nv_bfloat16, float>::type;
  auto to_out = [](float v) -> OutT {
    if constexpr (BF16) return __float2bfloat16_rn(v);
    else return v;
  };
  OutT* const top = static_cast<OutT*>(p.top);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [stages][cg][hwp] floats | RoiEntry[kMaxRoiTable] | int roi_id[kMaxRoiTable] | bars |
  //         per-warp output staging
  float* slabs = reinterpret_cast<float*>(smem_raw);
  const int W = W_CT ? W_CT : p.W;
  const int hwp = HWP_CT ? HWP_CT : p.hwp;
  const size_t stage_floats = (size_t)p.cg * hwp;
  RoiEntry* table = reinterpret_cast<RoiEntry*>(slabs + stage_floats * p.stages);
  int* roi_id = reinterpret_cast<int*>(table + kMaxRoiTable);
  uint64_t* full = reinterpret_cast<uint64_t*>(roi_id + kMaxRoiTable);
  uint64_t* empty = full + kStagesMax;
  OutT* out_stage = reinterpret_cast<OutT*>(empty + kStagesMax);  // [kConsWarps][4*CPL][7][7]
  __shared__ int s_nroi, s_next, s_warp_cnt[kConsWarps], s_fbeg[kPreFrames + 1];

  const int tid = threadIdx.x, lane = tid & 31;
  // warp index through a lane-0 broadcast: the compiler then knows it is warp-uniform, and loops whose
  // bounds depend on it need no collective-reconvergence scaffolding (WARPSYNC.COLLECTIVE / VOTE /
  // ENDCOLLECTIVE, ~3 instructions per shuffle) around the pooling shuffles
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  NAFAE_CTA_TRACE(cta_trace, 1);  // debug builds only
  // contiguous unit range per CTA: it touches 1-2 frames, whose RoI tables are built once
  const int u_begin = (int)((long long)p.units * blockIdx.x / gridDim.x);
  const int u_end = (int)((long long)p.units * (blockIdx.x + 1) / gridDim.x);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsWarps);
    }
    fence_mbar_init();
    // residency gate: the last CTA to become resident opens it for the concurrent branches
    if (p.gate != nullptr && atomicAdd(p.gate, 1) == (int)gridDim.x - 1) {
      p.gate[0] = 0;
      __threadfence();
      atomicAdd(p.gate + 1, 1);
    }
  }
  __syncthreads();

  if (warp == kConsWarps) {
    // ------------------------------------------------------------ producer warp ----
    if (lane == 0) {
      const uint32_t chan_bytes = (uint32_t)p.hw * 4u;
      for (int u = u_begin; u < u_end; ++u) {
        const int it = u - u_begin, stage = it % p.stages;
        if (it >= p.stages) mbar_wait(&empty[stage], (uint32_t)(it / p.stages - 1) & 1u);
        const int f = u / p.groups, gidx = u % p.groups;
        const float* src = p.bottom + ((size_t)f * p.C + (size_t)gidx * p.cg) * p.hw;
        float* dst = slabs + stage_floats * stage;
        mbar_arrive_expect_tx(&full[stage], chan_bytes * p.cg);
        for (int c = 0; c < p.cg; ++c)
          bulk_g2s(dst + (size_t)c * hwp, src + (size_t)c * p.hw, chan_bytes, &full[stage]);
      }
    }
    return;
  }

  // -------------------------------------------------------------- consumer warps ----
  // Frame ids of RoIs tid, tid + 512, ... stay in registers for the whole kernel: ONE L2 round
  // trip (in flight while the first slab streams in); every later table scan is register /
  // shared-memory work only.  (R > kFrCache * 512: scans fall back to global loads.)
  constexpr int kFrCache = 4;
  constexpr int kNoFrame = -2147483647 - 1;
  const bool fr_cached = p.R <= kFrCache * kConsThreads;
  int fr[kFrCache];
#pragma unroll
  for (int i = 0; i < kFrCache; ++i) {
    const int r = tid + i * kConsThreads;
    fr[i] = (fr_cached && r < p.R) ? (int)__ldg(p.rois + (size_t)r * 5) : kNoFrame;
  }

  // RoIs whose batch index is outside [0, B): defined as all-zero rows (CTA 0 writes them)
  if (blockIdx.x == 0) {
    auto zero_bad = [&](int base, int b) {  // b = batch index of RoI base + tid
      unsigned bad = __ballot_sync(0xffffffffu, base + tid < p.R && (b < 0 || b >= p.B));
      while (bad) {
        const int rr = base + warp * 32 + __ffs(bad) - 1;
        bad &= bad - 1;
        OutT* o = top + (size_t)rr * p.C * (kOut * kOut);
        for (int i = lane; i < p.C * kOut * kOut; i += 32) o[i] = to_out(0.f);
      }
    };
    if (fr_cached) {
#pragma unroll
      for (int i = 0; i < kFrCache; ++i)
        if (i * kConsThreads < p.R) zero_bad(i * kConsThreads, fr[i]);
    } else {
      for (int base = 0; base < p.R; base += kConsThreads)
        zero_bad(base, base + tid < p.R ? (int)__ldg(p.rois + (size_t)(base + tid) * 5) : 0);
    }
  }

  // One 512-RoI step of a table scan: appends the RoIs base + tid (index >= r_start) of frame f
  // in ascending index order; true when the table is full.  All consumer warps together.
  auto scan_step = [&](int f, int r_start, int base, int frv) -> bool {
    const int r = base + tid;
    const bool hit = r < p.R && r >= r_start && frv == f;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    cons_barrier();
    const int have = s_nroi;
    int before = have, tot = 0;
    for (int w = 0; w < kConsWarps; ++w) {
      const int cw = s_warp_cnt[w];
      if (w < warp) before += cw;
      tot += cw;
    }
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (hit && pos < kMaxRoiTable) roi_id[pos] = r;
    if (hit && pos == kMaxRoiTable) s_next = r;  // first RoI that did not fit (unique thread)
    cons_barrier();
    if (have + tot >= kMaxRoiTable) {  // uniform
      if (tid == 0) {
        s_nroi = kMaxRoiTable;
        if (have + tot == kMaxRoiTable) s_next = min(base + kConsThreads, p.R);
      }
      return true;
    }
    if (tid == 0) s_nroi = have + tot;
    return false;
  };
  // Appends (from table position pos0) the RoIs of frame f whose index is >= r_start, at most up
  // to kMaxRoiTable entries; afterwards s_nroi = entries in the table, s_next = index where the
  // following chunk starts (>= R when the frame is exhausted).
  auto scan_frame = [&](int f, int r_start, int pos0) {
    cons_barrier();  // every warp is done with the previous table contents / counters
    if (tid == 0) {
      s_nroi = pos0;
      s_next = p.R;
    }
    cons_barrier();
    if (fr_cached) {
#pragma unroll
      for (int i = 0; i < kFrCache; ++i) {
        const int base = i * kConsThreads;
        if (base >= p.R) break;
        if (base + kConsThreads <= r_start) continue;
        if (scan_step(f, r_start, base, fr[i])) break;
      }
    } else {
      for (int base = r_start / kConsThreads * kConsThreads; base < p.R; base += kConsThreads) {
        const int r = base + tid;
        if (scan_step(f, r_start, base, r < p.R ? (int)p.rois[(size_t)r * 5] : kNoFrame)) break;
      }
    }
    cons_barrier();
  };
  // geometry of table entries [j_lo, j_hi): one thread per (RoI, axis, sample index)
  auto geometry = [&](int j_lo, int j_hi) {
    const float wscale = POOL == NAFAE_POOL_AVG ? 0.25f : 1.f;
    for (int e = j_lo * 2 * kS + tid; e < j_hi * 2 * kS; e += kConsThreads) {
      const int j = e / (2 * kS), k = e % (2 * kS);
      const RoiGeom g = roi_geom(p.rois + (size_t)roi_id[j] * 5, p.scale, kS, kS);
      int cell;
      float ratio;
      if (k < kS) {
        const bool ok = axis_sample(g.start_h, g.bin_h, k, p.H, &cell, &ratio);
        table[j].hoff_b[k] = ok ? cell * W * 4 : 0;
        table[j].h0[k] = ok ? 1.f - ratio : 0.f;
        table[j].h1[k] = ok ? ratio : 0.f;
      } else {
        const bool ok = axis_sample(g.start_w, g.bin_w, k - kS, p.W, &cell, &ratio);
        table[j].woff_b[k - kS] = ok ? cell * 4 : 0;
        table[j].w0[k - kS] = ok ? wscale * (1.f - ratio) : 0.f;
        table[j].w1[k - kS] = ok ? wscale * ratio : 0.f;
      }
    }
    cons_barrier();
  };
  auto build_table = [&](int f, int r_start) {  // one (frame, chunk) at a time
    scan_frame(f, r_start, 0);
    geometry(0, s_nroi);
  };

  // Normal case: the RoIs of ALL frames this CTA touches (contiguous unit range: 1-2 frames) fit in
  // the table together -> built once, here, while the first slab is still in flight; no table work
  // and no CTA-wide barrier inside the unit loop.  Otherwise: per-(frame, chunk) tables as needed.
  const int f_first = u_begin / p.groups, f_last = (u_end - 1) / p.groups;
  bool pre = fr_cached && f_last - f_first < kPreFrames;
  if (pre) {
    int pos = 0;
    for (int f = f_first; f <= f_last; ++f) {
      if (tid == 0) s_fbeg[f - f_first] = pos;
      scan_frame(f, 0, pos);
      if (s_next < p.R) {  // uniform: does not fit
        pre = false;
        break;
      }
      pos = s_nroi;
    }
    if (pre) {
      if (tid == 0) s_fbeg[f_last - f_first + 1] = pos;
      geometry(0, pos);
    }
  }

  int cur_f = -1, cur_start = -1;  // which (frame, chunk start) the table holds
  // channel blocks per RoI inside a unit (compile-time for the two production shapes: the
  // idx / nblk split below is a 20-instruction integer division otherwise)
  const int nblk = NBLK_CT ? NBLK_CT : p.cg / (4 * CPL);
  const int cq = lane >> 3, pw = lane & 7;
  int g_base = 0;  // running pass-group counter (uniform): deals groups round-robin to warps
  OutT* my_stage = out_stage + warp * (4 * CPL * kOut * kOut);
  for (int u = u_begin; u < u_end; ++u) {
    const int it = u - u_begin;
    const int stage = it % p.stages;
    const uint32_t parity = (uint32_t)(it / p.stages) & 1u;
    const int f = u / p.groups, gidx = u % p.groups;
    const unsigned char* slab = reinterpret_cast<const unsigned char*>(slabs + stage_floats * stage);
    bool waited = false;
    int r_next = 0;
    do {
      int j0 = 0, nroi, next_after;
      if (pre) {
        j0 = s_fbeg[f - f_first];
        nroi = s_fbeg[f - f_first + 1] - j0;
        next_after = p.R;
      } else {
        if (!(cur_f == f && cur_start == r_next)) {
          build_table(f, r_next);
          cur_f = f;
          cur_start = r_next;
        }
        nroi = s_nroi;
        next_after = s_next;
      }
      if (!waited) {
        mbar_wait(&full[stage], parity);
        waited = true;
      }
      const int ng = nroi * nblk;
      for (int idx = warp >= g_base ? warp - g_base : warp - g_base + kConsWarps; idx < ng; idx += kConsWarps) {
        const int jr = idx / nblk, blk = idx - jr * nblk, j = j0 + jr;
        const RoiEntry& e = table[j];
        const float w0 = e.w0[pw], w1 = e.w1[pw];
        const int ch0 = blk * (4 * CPL) + cq;  // first channel of this lane inside the slab
        const unsigned char* lane_base = slab + (size_t)ch0 * hwp * 4 + e.woff_b[pw];
        int hoff[kS];
        float h0[kS], h1[kS];
#pragma unroll
        for (int v = 0; v < kS; v += 4) {
          const int4 o4 = *reinterpret_cast<const int4*>(&e.hoff_b[v]);
          const float4 a4 = *reinterpret_cast<const float4*>(&e.h0[v]);
          const float4 b4 = *reinterpret_cast<const float4*>(&e.h1[v]);
          hoff[v] = o4.x; hoff[v + 1] = o4.y; hoff[v + 2] = o4.z; hoff[v + 3] = o4.w;
          h0[v] = a4.x; h0[v + 1] = a4.y; h0[v + 2] = a4.z; h0[v + 3] = a4.w;
          h1[v] = b4.x; h1[v + 1] = b4.y; h1[v + 2] = b4.z; h1[v + 3] = b4.w;
        }
        // T0/T1: horizontally interpolated cell rows (hstart, hstart+1) of the current sample row;
        // kept across sample rows while the cell rows repeat or advance by one (warp-uniform tests)
        float s[CPL][kS], T0[CPL], T1[CPL];
#if NAFAE_SLAB_PREDICATED
        // experiment (make pred): the same reuse rule as predicated loads + selects instead of
        // branches -- no reconvergence / collective-shuffle scaffolding in the pass loop, every
        // instruction slot is spent but a skipped load costs no shared-memory wavefront
#pragma unroll
        for (int k = 0; k < CPL; ++k) T0[k] = T1[k] = 0.f;
#pragma unroll
        for (int ph = 0; ph < kS; ++ph) {
          const int d = ph ? hoff[ph] - hoff[ph - 1] : -1;
          const bool ld_bot = d != 0;                  // the cell rows changed at all
          const bool ld_top = ld_bot && d != W * 4;    // ... and not just by one row (top = old bottom)
          const unsigned char* t = lane_base + hoff[ph];
#pragma unroll
          for (int k = 0; k < CPL; ++k) {
            const float* q = reinterpret_cast<const float*>(t + (size_t)k * 4 * hwp * 4);
            float top = ld_bot ? T1[k] : T0[k];
            if (ld_top) top = fmaf(q[1], w1, q[0] * w0);
            float bot = T1[k];
            if (ld_bot) bot = fmaf(q[W + 1], w1, q[W] * w0);
            T0[k] = top;
            T1[k] = bot;
            s[k][ph] = fmaf(bot, h1[ph], top * h0[ph]);
          }
        }
#else
#pragma unroll
        for (int ph = 0; ph < kS; ++ph) {
          const int d = ph ? hoff[ph] - hoff[ph - 1] : -1;
          if (d != 0) {
            const unsigned char* t = lane_base + hoff[ph];
            if (d == W * 4) {
#pragma unroll
              for (int k = 0; k < CPL; ++k) T0[k] = T1[k];
            } else {
#pragma unroll
              for (int k = 0; k < CPL; ++k) {
                const float* q = reinterpret_cast<const float*>(t + (size_t)k * 4 * hwp * 4);
                T0[k] = fmaf(q[1], w1, q[0] * w0);
              }
            }
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
              const float* q = reinterpret_cast<const float*>(t + (size_t)k * 4 * hwp * 4);
              T1[k] = fmaf(q[W + 1], w1, q[W] * w0);
            }
          }
#pragma unroll
          for (int k = 0; k < CPL; ++k) s[k][ph] = fmaf(T1[k], h1[ph], T0[k] * h0[ph]);
        }
#endif
        // pooled block -> per-warp staging (channel-major like the output) -> one bulk store
        if (lane == 0) bulk_wait_read<0>();  // the previous pass's store has drained the buffer
        __syncwarp();
        OutT* o = my_stage + cq * (kOut * kOut) + pw;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          OutT* ok = o + k * 4 * (kOut * kOut);
          if (POOL == NAFAE_POOL_AVG) {
            float hs_prev = s[k][0] + __shfl_down_sync(0xffffffffu, s[k][0], 1);
#pragma unroll
            for (int i = 0; i < kOut; ++i) {
              const float hs = s[k][i + 1] + __shfl_down_sync(0xffffffffu, s[k][i + 1], 1);
              if (pw < kOut) ok[i * kOut] = to_out(hs_prev + hs);
              hs_prev = hs;
            }
          } else {
            float right_prev = __shfl_down_sync(0xffffffffu, s[k][0], 1);
#pragma unroll
            for (int i = 0; i < kOut; ++i) {
              const float right_next = __shfl_down_sync(0xffffffffu, s[k][i + 1], 1);
              const float v = pool4(POOL, s[k][i], right_prev, s[k][i + 1], right_next);
              if (pw < kOut) ok[i * kOut] = to_out(v);
              right_prev = right_next;
            }
          }
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0) {
          const int c_first = gidx * p.cg + blk * (4 * CPL);
          bulk_s2g(top + ((size_t)roi_id[j] * p.C + c_first) * (kOut * kOut), my_stage,
                   (uint32_t)(4 * CPL * kOut * kOut * sizeof(OutT)));
          bulk_commit();
        }
      }
      g_base = (g_base + ng) % kConsWarps;  // warp that takes the next pass-group
      r_next = next_after;
    } while (r_next < p.R);

    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);  // this warp is done with the stage
  }
  if (lane == 0) bulk_wait_all<0>();  // staging buffer must outlive the last bulk store
}

int smem_optin_limit() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
      v = 48 * 1024;
    cached = v;
  }
  return cached;
}

// Persistent grid for `units` equal work units: the units are split statically, so the kernel ends
// when the CTAs with ceil(units / grid) units end -- use the SMALLEST grid with that same maximum
// (2560 units on 132 SMs: 20 per CTA either way, but 128 CTAs instead of 132) and leave the other
// SMs to whatever runs concurrently.
int slab_grid(int units) {
  int grid = persistent_grid();
  if (grid > units) grid = units;
  if (grid < 1) return 1;
  const int per_cta = (units + grid - 1) / grid;
  return (units + per_cta - 1) / per_cta;
}

template <int W_CT, int HWP_CT, int CPL, int NBLK_CT>
int launch_slab(const SlabParams& p, int pool, bool bf16, size_t smem, cudaStream_t stream) {
  auto kern = pool == NAFAE_POOL_AVG
                  ? (bf16 ? align_pool_fwd_slab<NAFAE_POOL_AVG, W_CT, HWP_CT, CPL, NBLK_CT, true>
                          : align_pool_fwd_slab<NAFAE_POOL_AVG, W_CT, HWP_CT, CPL, NBLK_CT, false>)
                  : (bf16 ? align_pool_fwd_slab<NAFAE_POOL_MAX, W_CT, HWP_CT, CPL, NBLK_CT, true>
                          : align_pool_fwd_slab<NAFAE_POOL_MAX, W_CT, HWP_CT, CPL, NBLK_CT, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("roi_align: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  const int grid = slab_grid(p.units);
  kern<<<grid, kSlabThreads, smem, stream>>>(p);
  return launch_status("align_pool_fwd_slab");
}

// Returns 1 if the slab kernel was launched, 0 if the shape is not eligible (caller falls back),
// <0 on a launch error.
int try_launch_slab(const float* bottom, float scale, int B, int R, int H, int W, int C, int pool,
                    const float* rois, void* top, bool bf16, int* gate, cudaStream_t stream) {
  const int hw = H * W;
  if (hw % 4 != 0 || C % 8 != 0 || H < 2 || W < 2) return 0;
  if ((reinterpret_cast<uintptr_t>(bottom) & 15) != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(top) & 15) != 0) return 0;  // bulk stores of the pooled blocks
  if ((long long)R * 5 >= (1ll << 31)) return 0;
  int hwp = hw;
  while (hwp % 32 != 8) hwp += 4;
  const bool small_map = W == 14 && hwp == 200 && C % 32 == 0;  // 14x14: 16-channel passes (CPL 4)
  const size_t staging = (size_t)kConsWarps * 4 * (small_map ? 4 : 2) * kOut * kOut * sizeof(float);
  const size_t fixed = sizeof(RoiEntry) * kMaxRoiTable + sizeof(int) * kMaxRoiTable +
                       sizeof(uint64_t) * 2 * kStagesMax + staging + 128;
  const size_t budget = (size_t)smem_optin_limit() - 1024;  // static smem + slack
  int cg = 0, stages = 0;
  // prefer >= 3 stages with a slab of <= 64 KB
  for (int cand : {32, 16, 8}) {
    if (C % cand) continue;
    const size_t stage_bytes = (size_t)cand * hwp * 4;
    if (stage_bytes > 64 * 1024 && cand > 8) continue;
    int st = (int)((budget - fixed) / stage_bytes);
    if (st > kStagesMax) st = kStagesMax;
    if (st >= 2) {
      cg = cand;
      stages = st;
      break;
    }
  }
  if (cg == 0) return 0;
  if ((size_t)hw * 4 * cg >= (1u << 20)) return 0;  // mbarrier tx-count range
  SlabParams p;
  p.bottom = bottom;
  p.rois = rois;
  p.top = top;
  p.scale = scale;
  p.B = B;
  p.R = R;
  p.H = H;
  p.W = W;
  p.C = C;
  p.cg = cg;
  p.groups = C / cg;
  p.hw = hw;
  p.hwp = hwp;
  p.stages = stages;
  p.units = B * p.groups;
  p.gate = gate;
  const size_t smem = (size_t)cg * hwp * 4 * stages + fixed;
  if (W == 50 && hwp == 1928 && cg == 8) return launch_slab<50, 1928, 2, 1>(p, pool, bf16, smem, stream);  // 38x50
  if (small_map && cg == 32) return launch_slab<14, 200, 4, 2>(p, pool, bf16, smem, stream);                // 14x14
  return launch_slab<0, 0, 2, 0>(p, pool, bf16, smem, stream);
}

}  // namespace
// roi_align_bwd.cu: cell-gather backward without atomics (1 launched, 0 not eligible, < 0 error)
int try_launch_avg_bwd_scatter(const float* top_diff, float scale, int B, int R, int H, int W, int C,
                               const float* rois, float* bottom_diff, bool accumulate, cudaStream_t stream);
int try_launch_avg_bwd_gather(const float* top_diff, float scale, int B, int R, int H, int W, int C,
                              const float* rois, float* bottom_diff, cudaStream_t stream);
namespace {

int grid_for(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_CTA_TRACE_READER(nafae_debug_cta_trace_roi_align)

NAFAE_API int nafae_roi_align_persistent_ctas(int num_units) { return slab_grid(num_units); }

NAFAE_API size_t nafae_roi_align_workspace_bytes(int batch_size, int num_rois) {
  (void)batch_size;
  (void)num_rois;
  // optional: the residency gate (the RoI tables live in shared memory)
  return NAFAE_ROI_ALIGN_WS_BYTES;
}

NAFAE_API int nafae_roi_align_forward(const float* bottom_data, float spatial_scale, int batch_size,
                                      int num_rois, int height, int width, int channels,
                                      int out_height, int out_width, int pool_mode,
                                      const float* bottom_rois, void* top_data, unsigned flags,
                                      void* workspace, size_t workspace_bytes,
                                      cudaStream_t stream) {
  NAFAE_REQUIRE(workspace == nullptr || workspace_bytes == 0 ||
                    workspace_bytes >= NAFAE_ROI_ALIGN_WS_BYTES,
                "roi_align: workspace must be NULL or >= %d bytes", NAFAE_ROI_ALIGN_WS_BYTES);
  int* gate = workspace != nullptr && workspace_bytes > 0 && !(flags & NAFAE_FLAG_NO_GATE)
                  ? static_cast<int*>(workspace)
                  : nullptr;
  NAFAE_REQUIRE(batch_size >= 0 && num_rois >= 0 && channels >= 0, "roi_align: negative sizes");
  NAFAE_REQUIRE(pool_mode >= NAFAE_POOL_NONE && pool_mode <= NAFAE_POOL_MAX,
                "roi_align: bad pool_mode %d", pool_mode);
  NAFAE_REQUIRE(out_height >= 1 && out_width >= 1, "roi_align: bad output size %dx%d", out_height,
                out_width);
  const long long total = (long long)num_rois * channels * out_height * out_width;
  if (total == 0) {
    if (gate) gate_open(gate, stream);
    return 1;
  }
  NAFAE_REQUIRE(height >= 2 && width >= 2, "roi_align: feature map must be at least 2x2");
  NAFAE_REQUIRE(bottom_data && bottom_rois && top_data, "roi_align: NULL buffer");
  const bool exact = (flags & NAFAE_FLAG_EXACT) != 0;
  const bool bf16 = (flags & NAFAE_FLAG_OUT_BF16) != 0;
  NAFAE_REQUIRE(!(bf16 && exact), "roi_align: NAFAE_FLAG_OUT_BF16 and NAFAE_FLAG_EXACT exclude each other");
  if (!exact && pool_mode != NAFAE_POOL_NONE && out_height == kOut && out_width == kOut &&
      batch_size > 0) {
    const int st = try_launch_slab(bottom_data, spatial_scale, batch_size, num_rois, height, width,
                                   channels, pool_mode, bottom_rois, top_data, bf16, gate, stream);
    if (st != 0) return st;
  }
  NAFAE_REQUIRE(!bf16, "roi_align: bf16 output needs the bandwidth kernel's shapes (7x7 avg / max pooling, "
                       "H*W %% 4 == 0, C %% 8 == 0, 16-byte aligned buffers)");
  if (gate) gate_open(gate, stream);  // no persistent kernel on this path: nothing to wait for
  const int grid = grid_for(total);
  if (exact)
    align_fwd_generic<true><<<grid, 256, 0, stream>>>(bottom_data, spatial_scale, batch_size, total,
                                                      height, width, channels, out_height,
                                                      out_width, pool_mode, bottom_rois, static_cast<float*>(top_data));
  else
    align_fwd_generic<false><<<grid, 256, 0, stream>>>(bottom_data, spatial_scale, batch_size,
                                                       total, height, width, channels, out_height,
                                                       out_width, pool_mode, bottom_rois, static_cast<float*>(top_data));
  return launch_status("align_fwd_generic");
}

NAFAE_API int nafae_roi_align_backward(const float* top_diff, const float* bottom_data,
                                       float spatial_scale, int batch_size, int num_rois,
                                       int height, int width, int channels, int out_height,
                                       int out_width, int pool_mode, const float* bottom_rois,
                                       float* bottom_diff, unsigned flags, cudaStream_t stream) {
  NAFAE_REQUIRE(batch_size >= 0 && num_rois >= 0 && channels >= 0, "roi_align: negative sizes");
  NAFAE_REQUIRE(pool_mode >= NAFAE_POOL_NONE && pool_mode <= NAFAE_POOL_MAX,
                "roi_align: bad pool_mode %d", pool_mode);
  NAFAE_REQUIRE(out_height >= 1 && out_width >= 1, "roi_align: bad output size");
  NAFAE_REQUIRE(pool_mode != NAFAE_POOL_MAX || bottom_data,
                "roi_align backward: max pooling needs bottom_data");
  const int sh = pool_mode ? out_height + 1 : out_height, sw = pool_mode ? out_width + 1 : out_width;
  const long long total = (long long)num_rois * channels * sh * sw;
  const bool overwrite = (flags & NAFAE_FLAG_OVERWRITE) != 0;
  const bool fast_avg = total > 0 && batch_size > 0 && !(flags & NAFAE_FLAG_EXACT) && pool_mode == NAFAE_POOL_AVG &&
                        out_height == kOut && out_width == kOut && top_diff && bottom_rois && bottom_diff &&
                        height >= 2 && width >= 2;
  if (fast_avg) {
    int st = 0;
    if (overwrite && (flags & NAFAE_FLAG_DETERMINISTIC))
      st = try_launch_avg_bwd_gather(top_diff, spatial_scale, batch_size, num_rois, height, width, channels,
                                     bottom_rois, bottom_diff, stream);
    if (st == 0)
      st = try_launch_avg_bwd_scatter(top_diff, spatial_scale, batch_size, num_rois, height, width, channels,
                                      bottom_rois, bottom_diff, !overwrite, stream);
    if (st == 0 && overwrite)
      st = try_launch_avg_bwd_gather(top_diff, spatial_scale, batch_size, num_rois, height, width, channels,
                                     bottom_rois, bottom_diff, stream);
    if (st != 0) return st;
  }
  if (overwrite && batch_size > 0 && channels > 0) {
    NAFAE_REQUIRE(bottom_diff && height >= 1 && width >= 1, "roi_align backward: NULL / empty bottom_diff");
    cudaError_t e = cudaMemsetAsync(bottom_diff, 0, sizeof(float) * (size_t)batch_size * channels * height * width,
                                    stream);
    if (e != cudaSuccess) {
      set_error("roi_align backward: cudaMemsetAsync: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  if (total == 0 || batch_size == 0) return 1;
  NAFAE_REQUIRE(height >= 2 && width >= 2, "roi_align: feature map must be at least 2x2");
  NAFAE_REQUIRE(top_diff && bottom_rois && bottom_diff, "roi_align: NULL buffer");
  const int grid = grid_for(total);
  if (flags & NAFAE_FLAG_EXACT)
    align_bwd_generic<true><<<grid, 256, 0, stream>>>(top_diff, bottom_data, spatial_scale,
                                                      batch_size, total, height, width, channels,
                                                      out_height, out_width, pool_mode,
                                                      bottom_rois, bottom_diff);
  else
    align_bwd_generic<false><<<grid, 256, 0, stream>>>(top_diff, bottom_data, spatial_scale,
                                                       batch_size, total, height, width, channels,
                                                       out_height, out_width, pool_mode,
                                                       bottom_rois, bottom_diff);
  return launch_status("align_bwd_generic");
}

// Reference-named launchers (roi_align_kernel.h:13-27).  The reference never checks the batch
// index; here it must lie in [0, 2^20) (rows outside the real batch are the caller's bug there too).
NAFAE_API int ROIAlignForwardLaucher(const float* bottom_data, const float spatial_scale,
                                     const int num_rois, const int height, const int width,
                                     const int channels, const int aligned_height,
                                     const int aligned_width, const float* bottom_rois,
                                     float* top_data, cudaStream_t stream) {
  return nafae_roi_align_forward(bottom_data, spatial_scale, 1 << 20, num_rois, height, width,
                                 channels, aligned_height, aligned_width, NAFAE_POOL_NONE,
                                 bottom_rois, top_data, NAFAE_FLAG_EXACT, nullptr, 0, stream);
}

NAFAE_API int ROIAlignBackwardLaucher(const float* top_diff, const float spatial_scale,
                                      const int batch_size, const int num_rois, const int height,
                                      const int width, const int channels,
                                      const int aligned_height, const int aligned_width,
                                      const float* bottom_rois, float* bottom_diff,
                                      cudaStream_t stream) {
  return nafae_roi_align_backward(top_diff, nullptr, spatial_scale, batch_size, num_rois, height,
                                  width, channels, aligned_height, aligned_width, NAFAE_POOL_NONE,
                                  bottom_rois, bottom_diff, NAFAE_FLAG_EXACT, stream);
}
