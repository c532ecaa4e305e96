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
#include "common.cuh"

namespace nafae {
namespace {

// ------------------------------------------------------------------------ geometry ----
struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int batch;
};

// roi_align_kernel.cu:33-43 as compiled: end-start is fma(x2, s, -RN(x1*s)); "+ 1." in double
// then fmaxf's float conversion is an exact float add; bin is a double division rounded to float.
__device__ __forceinline__ RoiGeom roi_geom(const float* __restrict__ roi, float scale, int sh,
                                            int sw) {
  RoiGeom g;
  g.batch = (int)roi[0];
  g.start_w = __fmul_rn(roi[1], scale);
  g.start_h = __fmul_rn(roi[2], scale);
  const float rw = fmaxf(__fadd_rn(__fmaf_rn(roi[3], scale, -g.start_w), 1.f), 0.f);
  const float rh = fmaxf(__fadd_rn(__fmaf_rn(roi[4], scale, -g.start_h), 1.f), 0.f);
  g.bin_h = __double2float_rn(__ddiv_rn((double)rh, __dsub_rn((double)sh, 1.)));
  g.bin_w = __double2float_rn(__ddiv_rn((double)rw, __dsub_rn((double)sw, 1.)));
  return g;
}

// one axis of a sample point (roi_align_kernel.cu:45-49,54,58-59): returns false if outside
__device__ __forceinline__ bool axis_sample(float start, float bin, int p, int extent, int* cell,
                                            float* ratio) {
  const float x = __fmaf_rn((float)p, bin, start);
  if (x < 0.f || x >= (float)extent || x != x) {
    *cell = 0;
    *ratio = 0.f;
    // NaN: the reference's comparisons are all false -> it would read out of bounds; we
    // define the sample as outside instead.
    return false;
  }
  const int c = (int)fminf(floorf(x), (float)(extent - 2));
  *cell = c;
  *ratio = __fsub_rn(x, (float)c);
  return true;
}

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
// a ring of `stages` buffers with full/empty mbarriers.
//
// PRODUCER WARP (all 32 lanes) -- scheduling, loads and RoI tables:
//   * units are split into contiguous ranges, one per CTA (a range touches 1-2 frames, so RoI
//     tables are rebuilt once or twice per CTA).  A CTA works through the PREFIX of its range without
//     any coordination; the last kTail units of every range are a TAIL that is claimed through a
//     counter in the workspace: the owner claims its own tail first (the atomic is issued one item
//     ahead, its latency never delays a bulk copy), and a CTA that runs out of work early STEALS
//     tail units from the nearest range that still has some.  A CTA that starts late (a concurrent
//     kernel still holds its SM) or drew expensive frames sheds up to kTail units instead of
//     setting the kernel's end time (round 1: avg 73.0 k vs max 84.5 k active cycles per SM).
//     Without a workspace the split is purely static.
//   * the RoI table of a frame (cells + weights of its 8x8 sample grid per RoI) is built by the
//     producer warp into one of two half buffers while the slab is in flight; consumer warps never
//     build tables and never meet at a CTA-wide barrier.  Frames with more than half / all of the
//     table capacity use the whole buffer / several chunks (the slab is then issued once per chunk).
//   * lane 0 issues the bulk copies; a stage is published (second arrival on its full barrier) with
//     a small descriptor (frame, channel group, table range).
// CONSUMER WARPS (16): pass-groups (RoI x 4*CPL channels) are dealt round-robin to the warps across
// unit boundaries, so a 20-RoI frame keeps 16 warps evenly busy.  Inside a pass a lane owns one
// sample column (pw) of CPL channels: the RoI's row offsets and weights are warp-uniform table
// reads, the taps are LDS with compile-time offsets (W and the padded channel stride are template
// constants for the two production map sizes), the 2x2 pool is one shuffle per sample row.  For
// POOL_AVG the 1/4 is folded into the column weights.
// The kernel is bound by instruction issue and the shared-memory data stage, not by arithmetic, so:
// (a) a sample row whose cell rows equal (or follow by one) the previous sample row's reuses the
// horizontally interpolated rows it already holds (RoIs shorter than 7 cells -- half of them --
// need ~6 row loads instead of 16); (b) the pooled 7x7 blocks of a pass (4*CPL consecutive channels
// of one RoI = 784*CPL contiguous bytes of the output) are staged in a per-warp shared buffer and
// leave with ONE bulk store (cp.async.bulk shared -> global).
constexpr int kOut = 7;
constexpr int kS = 8;               // sample grid side
#ifndef NAFAE_CONS_WARPS
#define NAFAE_CONS_WARPS 16
#endif
constexpr int kConsWarps = NAFAE_CONS_WARPS;
constexpr int kConsThreads = kConsWarps * 32;
constexpr int kSlabThreads = kConsThreads + 32;  // + producer warp
constexpr int kTabCap = 104;        // RoI table entries in shared memory
constexpr int kTabHalf = kTabCap / 2;
constexpr int kStagesMax = 4;
#ifndef NAFAE_SLAB_SCAV_DEPTH
#define NAFAE_SLAB_SCAV_DEPTH 2
#endif
constexpr int kScavDepth = NAFAE_SLAB_SCAV_DEPTH;  // items in the ring while stealing from other ranges
#ifndef NAFAE_SLAB_TAIL
#define NAFAE_SLAB_TAIL 3
#endif
constexpr int kTail = NAFAE_SLAB_TAIL;             // stealable units at the end of every CTA's range
// workspace words (ints): [0] residency arrivals, [1] gate epoch, [2..7] epochs seen per waiting slot,
// [8] exit ticket, [16 + b] tail units claimed from the range of CTA b
constexpr int kMaxSlabCtas = 1024;  // tail counters in the workspace (persistent grid <= SM count)
constexpr int kWsExit = 8;
constexpr int kWsSched = 16;

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
  float* top;
  float scale;
  int B, R, H, W, C;
  int cg;          // channels per slab
  int groups;      // C / cg
  int hw;          // H*W
  int hwp;         // padded per-channel stride in shared memory (floats), hwp % 32 == 8
  int stages;
  int units;       // B * groups
  int* gate;       // optional workspace: residency gate (nafae_gate_wait) ...
  int* sched;      // ... and per-frame claim counters (NULL: static contiguous split)
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one axis of a RoI's sample grid (roi_geom restricted to one axis: one double division)
__device__ __forceinline__ void roi_axis(const float* __restrict__ roi, float scale, bool is_w,
                                         float* start, float* bin) {
  const float lo = __ldg(roi + (is_w ? 1 : 2)), hi = __ldg(roi + (is_w ? 3 : 4));
  const float st = __fmul_rn(lo, scale);
  const float ext = fmaxf(__fadd_rn(__fmaf_rn(hi, scale, -st), 1.f), 0.f);
  *start = st;
  *bin = __double2float_rn(__ddiv_rn((double)ext, __dsub_rn((double)kS, 1.)));
}

template <int POOL, int W_CT, int HWP_CT, int CPL, int NBLK_CT>
__global__ void __launch_bounds__(kSlabThreads, 1) align_pool_fwd_slab(const SlabParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [stages][cg][hwp] floats | RoiEntry[kTabCap] | int roi_id[kTabCap] | int scan[kTabCap] |
  //         full/empty barriers | per-warp output staging
  float* slabs = reinterpret_cast<float*>(smem_raw);
  const int W = W_CT ? W_CT : p.W;
  const int hwp = HWP_CT ? HWP_CT : p.hwp;
  const size_t stage_floats = (size_t)p.cg * hwp;
  RoiEntry* table = reinterpret_cast<RoiEntry*>(slabs + stage_floats * p.stages);
  int* roi_id = reinterpret_cast<int*>(table + kTabCap);
  int* scan_list = roi_id + kTabCap;
  uint64_t* full = reinterpret_cast<uint64_t*>(scan_list + kTabCap);
  uint64_t* empty = full + kStagesMax;
  float* out_stage = reinterpret_cast<float*>(empty + kStagesMax);  // [kConsWarps][4*CPL][7][7]
  __shared__ int4 s_desc[kStagesMax];  // per stage: x = frame (< 0: no more work), y = channel group,
                                       //            z = first table entry, w = number of RoIs

  const int tid = threadIdx.x, lane = tid & 31;
  // warp index through a lane-0 broadcast: the compiler then knows it is warp-uniform, and loops whose
  // bounds depend on it need no collective-reconvergence scaffolding around the pooling shuffles
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  NAFAE_CTA_TRACE(cta_trace, 1);  // debug builds only

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 2);  // the copies' expect_tx arrival + the "table ready" arrival
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
    // ============================================================ producer warp ====
    const uint32_t chan_bytes = (uint32_t)p.hw * 4u;
    const bool dynamic = p.sched != nullptr && kTail > 0;
    // contiguous range of this CTA; with a workspace its last kTail units are claimed, not owned
    auto range_begin = [&](int b) { return (int)((long long)p.units * b / gridDim.x); };
    auto tail_of = [&](int b, int* first) {  // number of tail units of range b, *first = the first one
      const int lo = range_begin(b), hi = range_begin(b + 1);
      const int t = hi - lo < kTail ? hi - lo : kTail;
      *first = hi - t;
      return t;
    };
    int u_next = range_begin(blockIdx.x);
    const int u_end = range_begin(blockIdx.x + 1);
    int tail_first = u_end;
    const int my_tail = dynamic ? tail_of(blockIdx.x, &tail_first) : 0;
    const int prefix_end = u_end - my_tail;
    int it = 0;                     // items published so far
    int half_f[2] = {-1, -1};       // frame whose complete table sits in half h (-2: part of a whole-buffer table)
    int half_n[2] = {0, 0};
    int half_use[2] = {-1, -1};     // last item that reads half h

    // all consumer warps are done with `item`.  Items older than it - stages have been waited for
    // already (their stage has been re-armed since: the parity test would alias), so only the
    // latest user of a stage is ever waited on.
    auto wait_released = [&](int item) {
      if (item >= 0 && item >= it - p.stages)
        mbar_wait(&empty[item % p.stages], (uint32_t)(item / p.stages) & 1u);
    };
    // RoIs r >= r0 of frame f, ascending, at most kTabCap of them -> scan_list; returns the count,
    // *next = where the following chunk starts (>= R: none left).  Frame ids come from global memory
    // (L1 / L2 hits after the first scan), eight 32-RoI loads in flight.
    auto scan = [&](int f, int r0, int* next) -> int {
      int n = 0;
      *next = p.R;
      bool done = false;
      for (int base0 = r0 / 32 * 32; base0 < p.R && !done; base0 += 256) {
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = base0 + 32 * j + lane;
          v[j] = r < p.R ? (int)__ldg(p.rois + (size_t)r * 5) : -1;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int base = base0 + 32 * j;
          if (done || base >= p.R) continue;
          const int r = base + lane;
          const bool hit = r < p.R && r >= r0 && v[j] == f;
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          const int cnt = __popc(bal);
          const int pos = n + __popc(bal & ((1u << lane) - 1u));
          if (hit && pos < kTabCap) scan_list[pos] = r;
          if (n + cnt > kTabCap) {
            *next = base + (int)__fns(bal, 0, kTabCap - n + 1);  // first hit that did not fit
            n = kTabCap;
            done = true;
          } else {
            n += cnt;
            if (n == kTabCap) {
              *next = base + 32;
              done = true;
            }
          }
        }
      }
      __syncwarp();
      return n;
    };
    // table entries [j0, j0 + n) from scan_list[0, n): one lane per (RoI, axis)
    auto geometry = [&](int j0, int n) {
      const float wscale = POOL == NAFAE_POOL_AVG ? 0.25f : 1.f;
      for (int e = lane; e < 2 * n; e += 32) {
        const int j = e >> 1;
        const bool is_w = e & 1;
        const int r = scan_list[j];
        float start, bin;
        roi_axis(p.rois + (size_t)r * 5, p.scale, is_w, &start, &bin);
        RoiEntry& t = table[j0 + j];
        if (!is_w) roi_id[j0 + j] = r;
#pragma unroll
        for (int k = 0; k < kS; ++k) {
          int cell;
          float ratio;
          const bool ok = axis_sample(start, bin, k, is_w ? p.W : p.H, &cell, &ratio);
          if (is_w) {
            t.woff_b[k] = ok ? cell * 4 : 0;
            t.w0[k] = ok ? wscale * (1.f - ratio) : 0.f;
            t.w1[k] = ok ? wscale * ratio : 0.f;
          } else {
            t.hoff_b[k] = ok ? cell * W * 4 : 0;
            t.h0[k] = ok ? 1.f - ratio : 0.f;
            t.h1[k] = ok ? ratio : 0.f;
          }
        }
      }
      __syncwarp();
    };
    // Tail claiming.  The owner's claim on its own tail is issued one item ahead (start_claim) and
    // only read when the unit is needed.  Stealing is synchronous and keeps only one unit ahead of the
    // one being processed, so that nobody sits on stolen work while others are idle.
    int pend = 0;           // lane 0: result of the own-tail claim in flight
    bool own_tail = true;   // still claiming from the own range's tail
    bool stealing = false;
    int victim = (int)blockIdx.x;
    auto start_claim = [&]() {
      if (lane == 0) pend = atomicAdd(p.sched + blockIdx.x, 1);
    };
    auto steal = [&]() -> int {  // -1: no tail unit left anywhere
      while (victim >= 0) {
        int found = -1;
        for (int base = 1; base < (int)gridDim.x && found < 0; base += 32) {
          const int k = base + lane;  // distance from the last victim
          int b = victim + k;
          if (b >= (int)gridDim.x) b -= gridDim.x;
          int first;
          const bool open = k < (int)gridDim.x && __ldcg(p.sched + b) < tail_of(b, &first);
          const unsigned bal = __ballot_sync(0xffffffffu, open);
          if (bal) {
            found = victim + base + __ffs(bal) - 1;
            if (found >= (int)gridDim.x) found -= gridDim.x;
          }
        }
        victim = found;
        if (found < 0) break;
        int g = 0;
        if (lane == 0) g = atomicAdd(p.sched + found, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        int first;
        if (g < tail_of(found, &first)) return first + g;
      }
      return -1;
    };
    // next unit of this CTA, -1 when there is none (warp-uniform)
    auto claim = [&]() -> int {
      if (u_next < prefix_end) return u_next++;
      if (!dynamic) return -1;
      if (own_tail) {
        const int g = __shfl_sync(0xffffffffu, pend, 0);
        if (g < my_tail) return tail_first + g;
        own_tail = false;
        stealing = true;
      }
      return steal();
    };

    if (dynamic && u_next >= prefix_end) start_claim();  // a range without a prefix starts on its tail
    for (;;) {
      const int u = claim();  // known before the ring has room: the copy below leaves at once
      if (u < 0) break;
      // the next unit comes from the own tail: get its claim going now, hidden behind this item
      if (dynamic && own_tail && u_next >= prefix_end) start_claim();
      const int f = u / p.groups, gidx = u - f * p.groups;
      int r_next = 0, n = 0;
      bool first_chunk = true;
      do {  // one item per table chunk (a single one unless the frame has > kTabCap RoIs)
        // room for one more item?  `stages` deep on the own range, one unit ahead of the one being
        // processed while stealing
        const int depth = !stealing ? p.stages : (p.stages < kScavDepth ? p.stages : kScavDepth);
        wait_released(it - depth);
        const int stage = it % p.stages;
        if (lane == 0) {  // the slab is in flight while the table is looked up / built
          const float* src = p.bottom + ((size_t)f * p.C + (size_t)gidx * p.cg) * p.hw;
          float* dst = slabs + stage_floats * stage;
          mbar_arrive_expect_tx(&full[stage], chan_bytes * p.cg);
          for (int c = 0; c < p.cg; ++c)
            bulk_g2s(dst + (size_t)c * hwp, src + (size_t)c * p.hw, chan_bytes, &full[stage]);
        }
        int h = first_chunk ? (half_f[0] == f ? 0 : (half_f[1] == f ? 1 : -1)) : -1;
        if (h >= 0) {  // table already on chip
          n = half_n[h];
          r_next = p.R;
          half_use[h] = it;
          if (half_f[1] == -2) half_use[1] = it;  // whole-buffer table
        } else {
          n = scan(f, r_next, &r_next);
          const bool whole = n > kTabHalf;
          if (whole || half_f[1] == -2) {  // both halves are involved: both must drain
            wait_released(half_use[0]);
            wait_released(half_use[1]);
            half_f[0] = half_f[1] = -1;
            h = 0;
          } else {  // the half that is not serving the most recent item
            h = half_use[0] <= half_use[1] ? 0 : 1;
            wait_released(half_use[h]);
          }
          geometry(h * kTabHalf, n);
          // only a table that holds ALL RoIs of the frame can be reused by later units of that frame
          half_f[h] = (r_next >= p.R && first_chunk) ? f : -1;
          half_n[h] = n;
          half_use[h] = it;
          if (whole) {
            half_f[1] = -2;
            half_use[1] = it;
          }
        }
        first_chunk = false;
        __syncwarp();  // table / roi_id stores of all lanes are ordered before lane 0's release-arrive
        if (lane == 0) {
          s_desc[stage] = make_int4(f, gidx, h * kTabHalf, n);  // n may be 0: an empty item
          mbar_arrive(&full[stage]);
        }
        ++it;
      } while (r_next < p.R);
    }
    // no more work: tell the consumers
    wait_released(it - p.stages);
    {
      const int stage = it % p.stages;
      if (lane == 0) {
        s_desc[stage] = make_int4(-1, 0, 0, 0);
        mbar_arrive(&full[stage]);
        mbar_arrive(&full[stage]);
      }
    }
    // last CTA out leaves the claim counters zeroed for the next launch
    if (dynamic && lane == 0) {
      __threadfence();
      int* exit_ticket = p.sched - kWsSched + kWsExit;
      if (atomicAdd(exit_ticket, 1) == (int)gridDim.x - 1) {
        for (unsigned b = 0; b < gridDim.x; ++b) p.sched[b] = 0;
        *exit_ticket = 0;
      }
    }
    return;
  }

  // ============================================================== consumer warps ====
  // RoIs whose batch index is outside [0, B): defined as all-zero rows (CTA 0 writes them)
  if (blockIdx.x == 0) {
    for (int base = 0; base < p.R; base += kConsThreads) {
      const int r = base + tid;
      const int b = r < p.R ? (int)__ldg(p.rois + (size_t)r * 5) : 0;
      unsigned bad = __ballot_sync(0xffffffffu, r < p.R && (b < 0 || b >= p.B));
      while (bad) {
        const int rr = base + warp * 32 + __ffs(bad) - 1;
        bad &= bad - 1;
        float* o = p.top + (size_t)rr * p.C * (kOut * kOut);
        for (int i = lane; i < p.C * kOut * kOut; i += 32) o[i] = 0.f;
      }
    }
  }

  // channel blocks per RoI inside a unit (compile-time for the two production shapes: the
  // idx / nblk split below is a 20-instruction integer division otherwise)
  const int nblk = NBLK_CT ? NBLK_CT : p.cg / (4 * CPL);
  const int cq = lane >> 3, pw = lane & 7;
  int g_base = 0;  // running pass-group counter (uniform): deals groups round-robin to warps
  float* my_stage = out_stage + warp * (4 * CPL * kOut * kOut);
  for (int it = 0;; ++it) {
    const int stage = it % p.stages;
    mbar_wait(&full[stage], (uint32_t)(it / p.stages) & 1u);
    const int4 d = s_desc[stage];
    if (d.x < 0) break;
    const int gidx = d.y, j0 = d.z, nroi = d.w;
#ifdef NAFAE_TRACE
    if (tid == 0) CtaTrace::count(cta_trace.idx, 1, nroi);  // items / RoI passes of this CTA
#endif
    const unsigned char* slab = reinterpret_cast<const unsigned char*>(slabs + stage_floats * stage);
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
#pragma unroll
      for (int ph = 0; ph < kS; ++ph) {
        const int dh = ph ? hoff[ph] - hoff[ph - 1] : -1;
        if (dh != 0) {
          const unsigned char* t = lane_base + hoff[ph];
          if (dh == W * 4) {
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
      // pooled block -> per-warp staging (channel-major like the output) -> one bulk store
      if (lane == 0) bulk_wait_read<0>();  // the previous pass's store has drained the buffer
      __syncwarp();
      float* o = my_stage + cq * (kOut * kOut) + pw;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        float* ok = o + k * 4 * (kOut * kOut);
        if (POOL == NAFAE_POOL_AVG) {
          float hs_prev = s[k][0] + __shfl_down_sync(0xffffffffu, s[k][0], 1);
#pragma unroll
          for (int i = 0; i < kOut; ++i) {
            const float hs = s[k][i + 1] + __shfl_down_sync(0xffffffffu, s[k][i + 1], 1);
            if (pw < kOut) ok[i * kOut] = hs_prev + hs;
            hs_prev = hs;
          }
        } else {
          float right_prev = __shfl_down_sync(0xffffffffu, s[k][0], 1);
#pragma unroll
          for (int i = 0; i < kOut; ++i) {
            const float right_next = __shfl_down_sync(0xffffffffu, s[k][i + 1], 1);
            const float v = pool4(POOL, s[k][i], right_prev, s[k][i + 1], right_next);
            if (pw < kOut) ok[i * kOut] = v;
            right_prev = right_next;
          }
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
      __syncwarp();
      if (lane == 0) {
        const int c_first = gidx * p.cg + blk * (4 * CPL);
        bulk_s2g(p.top + ((size_t)roi_id[j] * p.C + c_first) * (kOut * kOut), my_stage,
                 (uint32_t)(4 * CPL * kOut * kOut * sizeof(float)));
        bulk_commit();
      }
    }
    g_base = (g_base + ng) % kConsWarps;  // warp that takes the next pass-group
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

// Persistent grid for `units` work units: with dynamic claiming every available SM takes part;
// with the static split the kernel ends when the CTAs with ceil(units / grid) units end, so use the
// SMALLEST grid with that same maximum and leave the other SMs to whatever runs concurrently.
int slab_grid(int units, bool dynamic) {
  int grid = persistent_grid();
  if (grid > units) grid = units;
  if (grid < 1) return 1;
  if (dynamic) return grid;
  const int per_cta = (units + grid - 1) / grid;
  return (units + per_cta - 1) / per_cta;
}

template <int W_CT, int HWP_CT, int CPL, int NBLK_CT>
int launch_slab(const SlabParams& p, int pool, size_t smem, cudaStream_t stream) {
  auto kern = pool == NAFAE_POOL_AVG ? align_pool_fwd_slab<NAFAE_POOL_AVG, W_CT, HWP_CT, CPL, NBLK_CT>
                                     : align_pool_fwd_slab<NAFAE_POOL_MAX, W_CT, HWP_CT, CPL, NBLK_CT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("roi_align: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  int grid = slab_grid(p.units, p.sched != nullptr);
  if (grid > kMaxSlabCtas) grid = kMaxSlabCtas;
  kern<<<grid, kSlabThreads, smem, stream>>>(p);
  return launch_status("align_pool_fwd_slab");
}

// Returns 1 if the slab kernel was launched, 0 if the shape is not eligible (caller falls back),
// <0 on a launch error.
int try_launch_slab(const float* bottom, float scale, int B, int R, int H, int W, int C, int pool,
                    const float* rois, float* top, int* gate, int* sched, cudaStream_t stream) {
  const int hw = H * W;
  if (hw % 4 != 0 || C % 8 != 0 || H < 2 || W < 2) return 0;
  if ((reinterpret_cast<uintptr_t>(bottom) & 15) != 0) return 0;
  if ((reinterpret_cast<uintptr_t>(top) & 15) != 0) return 0;  // bulk stores of the pooled blocks
  if ((long long)R * 5 >= (1ll << 31)) return 0;
  int hwp = hw;
  while (hwp % 32 != 8) hwp += 4;
  const bool small_map = W == 14 && hwp == 200 && C % 32 == 0;  // 14x14: 16-channel passes (CPL 4)
  const size_t staging = (size_t)kConsWarps * 4 * (small_map ? 4 : 2) * kOut * kOut * sizeof(float);
  const size_t fixed = sizeof(RoiEntry) * kTabCap + sizeof(int) * 2 * kTabCap +
                       sizeof(uint64_t) * 2 * kStagesMax + staging + 128;
  const size_t budget = (size_t)smem_optin_limit() - 1024;  // static smem + slack
  int cg = 0, stages = 0;
  // prefer >= 3 stages with a slab of <= 64 KB
  for (int cand : {32, 16, 8}) {
    if (C % cand) continue;
    const size_t stage_bytes = (size_t)cand * hwp * 4;
    if (stage_bytes > 64 * 1024 && cand > 8) continue;
    if (budget < fixed + stage_bytes) continue;
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
  p.sched = sched;
  const size_t smem = (size_t)cg * hwp * 4 * stages + fixed;
  if (W == 50 && hwp == 1928 && cg == 8) return launch_slab<50, 1928, 2, 1>(p, pool, smem, stream);  // 38x50
  if (small_map && cg == 32) return launch_slab<14, 200, 4, 2>(p, pool, smem, stream);                // 14x14
  return launch_slab<0, 0, 2, 0>(p, pool, smem, stream);
}

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

NAFAE_API int nafae_roi_align_persistent_ctas(int num_units) { return slab_grid(num_units, true); }

NAFAE_API size_t nafae_roi_align_workspace_bytes(int batch_size, int num_rois) {
  (void)batch_size;
  (void)num_rois;
  // the residency gate + one tail counter per persistent CTA (the RoI tables live in shared memory)
  return (size_t)NAFAE_ROI_ALIGN_WS_BYTES + sizeof(int) * kMaxSlabCtas;
}

NAFAE_API int nafae_roi_align_forward(const float* bottom_data, float spatial_scale, int batch_size,
                                      int num_rois, int height, int width, int channels,
                                      int out_height, int out_width, int pool_mode,
                                      const float* bottom_rois, float* top_data, unsigned flags,
                                      void* workspace, size_t workspace_bytes,
                                      cudaStream_t stream) {
  NAFAE_REQUIRE(workspace == nullptr || workspace_bytes == 0 ||
                    workspace_bytes >= NAFAE_ROI_ALIGN_WS_BYTES,
                "roi_align: workspace must be NULL or >= %d bytes", NAFAE_ROI_ALIGN_WS_BYTES);
  int* ws = workspace != nullptr && workspace_bytes > 0 ? static_cast<int*>(workspace) : nullptr;
  int* gate = (flags & NAFAE_FLAG_NO_GATE) ? nullptr : ws;
  // dynamic unit scheduling needs the per-frame claim counters behind the gate words
  int* sched = ws != nullptr && workspace_bytes >= nafae_roi_align_workspace_bytes(batch_size, num_rois) &&
                       (reinterpret_cast<uintptr_t>(workspace) & 3) == 0
                   ? ws + kWsSched
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
  if (!exact && pool_mode != NAFAE_POOL_NONE && out_height == kOut && out_width == kOut &&
      batch_size > 0) {
    const int st = try_launch_slab(bottom_data, spatial_scale, batch_size, num_rois, height, width,
                                   channels, pool_mode, bottom_rois, top_data, gate, sched, stream);
    if (st != 0) return st;
  }
  if (gate) gate_open(gate, stream);  // no persistent kernel on this path: nothing to wait for
  const int grid = grid_for(total);
  if (exact)
    align_fwd_generic<true><<<grid, 256, 0, stream>>>(bottom_data, spatial_scale, batch_size, total,
                                                      height, width, channels, out_height,
                                                      out_width, pool_mode, bottom_rois, top_data);
  else
    align_fwd_generic<false><<<grid, 256, 0, stream>>>(bottom_data, spatial_scale, batch_size,
                                                       total, height, width, channels, out_height,
                                                       out_width, pool_mode, bottom_rois, top_data);
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
