// RoIAlignAvg backward without global atomics (sm_100a).
//
// Reference: ROIAlignBackward (lib/model/roi_align/src/roi_align_kernel.cu:94-143) scatters every
// sample's gradient to four cells with atomicAdd into a pre-zeroed (B, C, H, W) tensor, preceded by
// avg_pool2d's backward (autograd); at cfg2 that is 105 M float atomics to L2 and a 156 MB memset.
// Both kernels here give a CTA exclusive ownership of (frame, 8-channel) slabs of bottom_diff, so HBM
// sees every output byte once, written with plain / bulk stores:
//   align_avg_bwd_scatter  (default)  the slab lives in shared memory, RoIs are scattered into it with
//                          shared-memory atomics, the finished slab leaves by bulk async copy;
//   align_avg_bwd_gather   (NAFAE_FLAG_DETERMINISTIC)  threads own cells and gather their contributions
//                          from a per-frame index; fixed summation order, bitwise reproducible.
// The pool's backward is folded in: sample gradient gs[ph][pw] = sum of the <= 4 output gradients
// whose 2x2 window contains the sample, the 1/4 sits in the column weights.
//
// 7x7 outputs of an 8x8 sample grid with average pooling only (RoIAlignAvg(7, 7, s), the one
// configuration the reference instantiates); everything else takes the generic atomic kernel.
#include "roi_geom.cuh"

namespace nafae {
namespace {

constexpr int kS = 8, kOut = 7;
constexpr int kBwCg = 8;          // channels per pass
constexpr int kBwThreads = 512;
constexpr int kBwWarps = kBwThreads / 32;
constexpr int kBwCells = 4;       // cells per thread  => H*W <= 2048
constexpr int kBwMaxDim = 128;    // H, W <= 128

struct BwAxis {  // per RoI
  int hcell[kS];   // cell row of sample row ph (hstart), -1000 when the sample row is outside
  float h0[kS], h1[kS];
  int wcell[kS];
  float w0[kS], w1[kS];  // * 1/4 (average pool)
};

struct BwParams {
  const float* top_diff;   // (R, C, 7, 7)
  const float* rois;       // (R, 5)
  float* bottom_diff;      // (B, C, H, W), fully overwritten
  float scale;
  int B, R, H, W, C;
  int splits;              // CTAs per frame; each takes C / splits channels in groups of kBwCg
};

// The first `cap` RoIs of frame f at or after r_next, in index order, into s_ids (all kBwThreads threads
// call it).  Afterwards *s_n = how many, *s_next = the first RoI of the frame that did not fit (R when
// the frame is exhausted).  Ends with a CTA barrier.
__device__ __forceinline__ void frame_rois(const float* __restrict__ rois, int R, int f, int r_next, int cap,
                                           int* s_ids, int* s_wcnt, int* s_n, int* s_next) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    *s_n = 0;
    *s_next = R;
  }
  __syncthreads();
  for (int base = r_next; base < R; base += kBwThreads) {
    const int r = base + tid;
    const bool hit = r < R && (int)__ldg(rois + (size_t)r * 5) == f;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    const int have = *s_n;
    int before = have, tot = 0;
    for (int w = 0; w < kBwWarps; ++w) {
      const int cw = s_wcnt[w];
      if (w < warp) before += cw;
      tot += cw;
    }
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (hit && pos < cap) s_ids[pos] = r;
    if (hit && pos == cap) *s_next = r;  // first RoI that did not fit (unique thread)
    __syncthreads();
    if (have + tot > cap) {
      if (tid == 0) *s_n = cap;
      break;
    }
    if (tid == 0) {
      *s_n = have + tot;
      if (have + tot == cap) *s_next = min(base + kBwThreads, R);
    }
    if (have + tot == cap) break;
  }
  __syncthreads();
}

// Axis tables of one RoI (rows when !is_w, columns when is_w): cell and the two weights of every sample
// row / column; the 1/4 of the average pool sits in the column weights.
__device__ __forceinline__ void fill_axis(BwAxis* a, const float* __restrict__ roi, bool is_w, float scale,
                                          int H, int W) {
  const float lo = __ldg(roi + (is_w ? 1 : 2)), hi = __ldg(roi + (is_w ? 3 : 4));
  const float start = __fmul_rn(lo, scale);
  const float ext = fmaxf(__fadd_rn(__fmaf_rn(hi, scale, -start), 1.f), 0.f);
  const float bin = __double2float_rn(__ddiv_rn((double)ext, __dsub_rn((double)kS, 1.)));
#pragma unroll
  for (int k = 0; k < kS; ++k) {
    int cell;
    float ratio;
    const bool ok = axis_sample(start, bin, k, is_w ? W : H, &cell, &ratio);
    if (is_w) {
      a->wcell[k] = ok ? cell : -1000;
      a->w0[k] = ok ? 0.25f * (1.f - ratio) : 0.f;
      a->w1[k] = ok ? 0.25f * ratio : 0.f;
    } else {
      a->hcell[k] = ok ? cell : -1000;
      a->h0[k] = ok ? 1.f - ratio : 0.f;
      a->h1[k] = ok ? ratio : 0.f;
    }
  }
}

// ------------------------------------------------------------------ indexed cell-gather (deterministic) ----
// Which (RoI, sample) pairs reach a cell depends on the frame's RoIs only, not on the channel: a CTA
// builds that index ONCE per frame -- per cell the list of (sample slot, weight) contributions in
// (RoI, sample row, sample column) order, a CSR over the H x W cells -- and reuses it for every
// 8-channel group it owns.  Building enumerates (cell, RoI) pairs through the two byte-tables; the
// per-group work is then exactly the contributions (20 RoIs x 256 at cfg2), no tests.  Threads own
// cells (coalesced, exclusive stores); cells with more than kCsrLong contributions (RoIs piled on one
// spot, e.g. the zero-padded proposal rows) are reduced by a whole warp with a fixed shuffle tree.
// The grid is persistent: CTA b takes the contiguous unit range [b, b + 1) * units / grid of the
// (frame, channel group) units, so a CTA sees at most two frames and rebuilds the index that often.
constexpr int kCsrChunk = 20;                   // RoIs indexed at a time
constexpr int kCsrMaxEnt = kCsrChunk * 256;     // 64 samples x 4 cells per RoI
constexpr int kCsrLong = 48;                    // contributions per cell above which a warp takes the cell
constexpr int kCsrMaxLong = kCsrMaxEnt / (kCsrLong + 1) + 1;

struct CsrEnt {
  int slot;   // (RoI in chunk) * 64 + sample
  float w;
};
struct CsrLong {
  int cell, start, cnt;
};

__global__ void __launch_bounds__(kBwThreads, 2) align_avg_bwd_gather(const BwParams p) {
  extern __shared__ __align__(16) unsigned char bw_smem[];
  float* gs = reinterpret_cast<float*>(bw_smem);                        // [chunk][64 samples][8 ch]
  CsrEnt* ent = reinterpret_cast<CsrEnt*>(gs + kCsrChunk * 64 * kBwCg); // [kCsrMaxEnt]
  BwAxis* ax = reinterpret_cast<BwAxis*>(ent + kCsrMaxEnt);             // [chunk]
  uchar4* rowmap = reinterpret_cast<uchar4*>(gs);                       // [chunk][H]: lo0, n0, lo1, n1 (aliases gs
  uchar4* colmap = rowmap + kCsrChunk * p.H;                            // [chunk][W]   while the index is built)
  __shared__ int s_ids[kCsrChunk];
  __shared__ int s_wcnt[kBwWarps];
  __shared__ int s_n, s_next, s_nlong;
  __shared__ CsrLong s_long[kCsrMaxLong];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups = p.C / kBwCg;
  const int hw = p.H * p.W;
  const long long units = (long long)p.B * groups;
  long long u = units * blockIdx.x / gridDim.x;
  const long long u_end = units * (blockIdx.x + 1) / gridDim.x;

  int cy[kBwCells], cx[kBwCells];
#pragma unroll
  for (int i = 0; i < kBwCells; ++i) {
    const int cell = tid + i * kBwThreads;
    cy[i] = cell < hw ? cell / p.W : -1;
    cx[i] = cell < hw ? cell - (cell / p.W) * p.W : -1;
  }

  while (u < u_end) {
    const int f = (int)(u / groups);
    const int g_begin = (int)(u - (long long)f * groups);
    const int g_end = (int)min((long long)groups, (long long)g_begin + (u_end - u));
    int r_next = 0;
    bool first_chunk = true;
    for (;;) {
      __syncthreads();  // the previous gather is done with gs / ent
      frame_rois(p.rois, p.R, f, r_next, kCsrChunk, s_ids, s_wcnt, &s_n, &s_next);
      const int n = s_n;
      r_next = s_next;
      if (n == 0 && !first_chunk) break;
      if (tid < 2 * n) fill_axis(&ax[tid >> 1], p.rois + (size_t)s_ids[tid >> 1] * 5, tid & 1, p.scale, p.H, p.W);
      if (tid == 0) s_nlong = 0;
      __syncthreads();
      // row / column maps: which sample rows reach cell row y with weight h0 (hstart == y) / h1 (hstart == y - 1)
      for (int i = tid; i < n * (p.H + p.W); i += kBwThreads) {
        const int j = i / (p.H + p.W), k = i - j * (p.H + p.W);
        const bool is_w = k >= p.H;
        const int v = is_w ? k - p.H : k;
        const int* cells = is_w ? ax[j].wcell : ax[j].hcell;
        int lo0 = 0, n0 = 0, lo1 = 0, n1 = 0;
#pragma unroll
        for (int q = 0; q < kS; ++q) {
          const int cc = cells[q];
          if (cc == v) {
            if (n0 == 0) lo0 = q;
            ++n0;
          }
          if (cc == v - 1) {
            if (n1 == 0) lo1 = q;
            ++n1;
          }
        }
        const uchar4 m = make_uchar4((unsigned char)lo0, (unsigned char)n0, (unsigned char)lo1, (unsigned char)n1);
        if (is_w) colmap[j * p.W + v] = m;
        else rowmap[j * p.H + v] = m;
      }
      __syncthreads();
      // pass 1: contributions per owned cell; exclusive scan in cell order -> start of every cell's list
      int cnt[kBwCells], start[kBwCells];
#pragma unroll
      for (int i = 0; i < kBwCells; ++i) {
        int t = 0;
        if (cy[i] >= 0)
          for (int j = 0; j < n; ++j) {
            const uchar4 rm = rowmap[j * p.H + cy[i]], cm = colmap[j * p.W + cx[i]];
            t += (rm.y + rm.w) * (cm.y + cm.w);
          }
        cnt[i] = t;
      }
      int carry = 0;
#pragma unroll
      for (int i = 0; i < kBwCells; ++i) {
        int inc = cnt[i];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += t;
        }
        if (lane == 31) s_wcnt[warp] = inc;
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < kBwWarps; ++w) {
          const int cw = s_wcnt[w];
          if (w < warp) before += cw;
          tot += cw;
        }
        start[i] = carry + before + inc - cnt[i];
        carry += tot;
        __syncthreads();
      }
      // pass 2: the lists, in (RoI, sample row, sample column) order
#pragma unroll
      for (int i = 0; i < kBwCells; ++i) {
        if (cy[i] < 0 || cnt[i] == 0) continue;
        int pos = start[i];
        for (int j = 0; j < n; ++j) {
          const uchar4 rm = rowmap[j * p.H + cy[i]];
          if (rm.y + rm.w == 0) continue;
          const uchar4 cm = colmap[j * p.W + cx[i]];
          if (cm.y + cm.w == 0) continue;
          const BwAxis& a = ax[j];
          for (int rr = 0; rr < rm.y + rm.w; ++rr) {
            const int ph = rr < rm.y ? rm.x + rr : rm.z + (rr - rm.y);
            const float wy = rr < rm.y ? a.h0[ph] : a.h1[ph];
            for (int qq = 0; qq < cm.y + cm.w; ++qq) {
              const int pw = qq < cm.y ? cm.x + qq : cm.z + (qq - cm.y);
              CsrEnt e;
              e.slot = j * 64 + ph * 8 + pw;
              e.w = wy * (qq < cm.y ? a.w0[pw] : a.w1[pw]);
              ent[pos++] = e;
            }
          }
        }
        if (cnt[i] > kCsrLong) {
          const int k = atomicAdd(&s_nlong, 1);  // at most kCsrMaxLong such cells exist; their order is irrelevant
          CsrLong l;
          l.cell = tid + i * kBwThreads;
          l.start = start[i];
          l.cnt = cnt[i];
          s_long[k] = l;
        }
      }

      // ---- the channel groups of this CTA's range in frame f, all with the index above
      for (int g = g_begin; g < g_end; ++g) {
        const int c0 = g * kBwCg;
        __syncthreads();  // index complete / previous group's gather done with gs
        // sample gradients = avg_pool2d(2, 1) backward of the output gradients, layout [RoI][sample][channel]
#pragma unroll 4
        for (int i = tid; i < n * 64 * kBwCg; i += kBwThreads) {
          const int sidx = i & 63, c = (i >> 6) & (kBwCg - 1), j = i >> 9;
          const int ph = sidx >> 3, pw = sidx & 7;
          const float* gg = p.top_diff + ((size_t)s_ids[j] * p.C + c0 + c) * (kOut * kOut);
          float v = 0.f;
          if (ph > 0 && pw > 0) v += __ldg(gg + (ph - 1) * kOut + pw - 1);
          if (ph > 0 && pw < kOut) v += __ldg(gg + (ph - 1) * kOut + pw);
          if (ph < kOut && pw > 0) v += __ldg(gg + ph * kOut + pw - 1);
          if (ph < kOut && pw < kOut) v += __ldg(gg + ph * kOut + pw);
          gs[((size_t)j * 64 + sidx) * kBwCg + c] = v;
        }
        __syncthreads();
        float* out = p.bottom_diff + ((size_t)f * p.C + c0) * hw;
        // short lists: the owning thread
#pragma unroll
        for (int i = 0; i < kBwCells; ++i) {
          const int cell = tid + i * kBwThreads;
          if (cell >= hw || cnt[i] > kCsrLong) continue;
          float acc[kBwCg];
#pragma unroll
          for (int c = 0; c < kBwCg; ++c) acc[c] = first_chunk ? 0.f : out[(size_t)c * hw + cell];
          for (int e = start[i]; e < start[i] + cnt[i]; ++e) {
            const CsrEnt en = ent[e];
            const float4* g4 = reinterpret_cast<const float4*>(gs + (size_t)en.slot * kBwCg);
            const float4 ga = g4[0], gb = g4[1];
            acc[0] = fmaf(en.w, ga.x, acc[0]);
            acc[1] = fmaf(en.w, ga.y, acc[1]);
            acc[2] = fmaf(en.w, ga.z, acc[2]);
            acc[3] = fmaf(en.w, ga.w, acc[3]);
            acc[4] = fmaf(en.w, gb.x, acc[4]);
            acc[5] = fmaf(en.w, gb.y, acc[5]);
            acc[6] = fmaf(en.w, gb.z, acc[6]);
            acc[7] = fmaf(en.w, gb.w, acc[7]);
          }
#pragma unroll
          for (int c = 0; c < kBwCg; ++c) out[(size_t)c * hw + cell] = acc[c];
        }
        // long lists: one warp per cell, lanes stride the list, fixed shuffle tree
        for (int k = warp; k < s_nlong; k += kBwWarps) {
          const CsrLong l = s_long[k];
          float acc[kBwCg];
#pragma unroll
          for (int c = 0; c < kBwCg; ++c) acc[c] = 0.f;
          for (int e = l.start + lane; e < l.start + l.cnt; e += 32) {
            const CsrEnt en = ent[e];
            const float4* g4 = reinterpret_cast<const float4*>(gs + (size_t)en.slot * kBwCg);
            const float4 ga = g4[0], gb = g4[1];
            acc[0] = fmaf(en.w, ga.x, acc[0]);
            acc[1] = fmaf(en.w, ga.y, acc[1]);
            acc[2] = fmaf(en.w, ga.z, acc[2]);
            acc[3] = fmaf(en.w, ga.w, acc[3]);
            acc[4] = fmaf(en.w, gb.x, acc[4]);
            acc[5] = fmaf(en.w, gb.y, acc[5]);
            acc[6] = fmaf(en.w, gb.z, acc[6]);
            acc[7] = fmaf(en.w, gb.w, acc[7]);
          }
#pragma unroll
          for (int c = 0; c < kBwCg; ++c) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], d);
          }
          if (lane < kBwCg) {
            float v = acc[0];
#pragma unroll
            for (int c = 1; c < kBwCg; ++c) v = lane == c ? acc[c] : v;
            float* dst = out + (size_t)lane * hw + l.cell;
            *dst = first_chunk ? v : *dst + v;
          }
        }
      }
      first_chunk = false;
      if (r_next >= p.R) break;
    }
    u += g_end - g_begin;
  }
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory scatter (the default): a CTA owns one (frame, 8-channel) slab of bottom_diff IN SHARED
// MEMORY (8 x H x W floats, 61 KB at 38 x 50), scatters the frame's RoIs into it with shared-memory
// atomics and stores the finished slab -- which is one contiguous block of (B, C, H, W) -- with bulk
// async copies (cp.async.bulk shared -> global; cp.reduce...add.f32 when the caller accumulates into an
// existing tensor).  HBM sees every output byte once and no atomics; the 105 M scatter adds of cfg2
// stay on chip.  A warp handles one sample point per pass, lane = (channel, corner): the 32 addresses
// of a warp instruction are distinct (and bank-conflict free for even W), so the CAS loop behind a
// shared float atomicAdd spins only when two WARPS collide.  (Measured alternative: lane = (sample
// column, corner) walking the 8 channels halves the instruction count but lets neighbouring samples
// collide inside the warp -- 301 vs 256 us at cfg2, 336 vs 248 us on 14x14 maps.)  The sample gradients
// (the pool's backward) are formed once per (RoI, channel) in shared memory; the RoIs' output gradients
// are staged by 16-byte cp.async, double buffered in chunks of kScChunk RoIs.  The order in which warps
// reach a cell is not fixed: like the reference's atomicAdd the sum is not bitwise reproducible
// (NAFAE_FLAG_DETERMINISTIC selects the gather kernel above instead).
constexpr int kScChunk = 8;     // RoIs staged at a time (x2 buffers)
constexpr int kScIds = 128;     // RoI indices of the frame held at a time
constexpr int kScBlock = kBwCg * kOut * kOut;   // floats of one RoI's 8-channel gradient block (1568 B)
constexpr uint32_t kScPiece = 32768;            // bytes per bulk store
constexpr int kGsStride = 65;                   // floats per (RoI, channel) block of sample gradients: odd, so the
                                                // 8 channels of one sample sit in 8 different banks

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g_add_f32(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}

template <bool ACC>
__global__ void __launch_bounds__(kBwThreads, 2) align_avg_bwd_scatter(const BwParams p) {
  extern __shared__ __align__(128) unsigned char sc_smem[];
  const int hw = p.H * p.W;
  float* slab = reinterpret_cast<float*>(sc_smem);                         // [8][H*W]
  float* stage = slab + (size_t)kBwCg * hw;                                // [2][chunk][8][49]
  float* gs = stage + 2 * kScChunk * kScBlock;                             // [chunk][8][kGsStride] sample gradients
  BwAxis* ax = reinterpret_cast<BwAxis*>(gs + kScChunk * kBwCg * kGsStride);      // [2][chunk]
  __shared__ int s_ids[kScIds];
  __shared__ int s_wcnt[kBwWarps];
  __shared__ int s_n, s_next;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int groups = p.C / kBwCg;
  const int f = blockIdx.x / groups, c0 = (blockIdx.x - f * groups) * kBwCg;

  {
    float4* z = reinterpret_cast<float4*>(slab);
    for (int i = tid; i < kBwCg * hw / 4; i += kBwThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }

  // lane roles: channel, corner of the sample's bilinear footprint -- the 32 addresses of one warp
  // instruction are distinct (and bank-conflict free for even W)
  const int c = lane >> 2, ky = (lane >> 1) & 1, kx = lane & 1;
  float* my_plane = slab + (size_t)c * hw + ky * p.W + kx;
  bool any = false;

  int r_next = 0;
  for (;;) {
    frame_rois(p.rois, p.R, f, r_next, kScIds, s_ids, s_wcnt, &s_n, &s_next);
    const int n_ids = s_n;
    r_next = s_next;
    any |= n_ids > 0;
    const int nchunks = (n_ids + kScChunk - 1) / kScChunk;

    auto prefetch = [&](int k) {
      const int n = min(kScChunk, n_ids - k * kScChunk), buf = k & 1;
      float* dst = stage + (size_t)buf * kScChunk * kScBlock;
      constexpr int kPieces = kScBlock / 4;  // 16-byte pieces per RoI block
      for (int i = tid; i < n * kPieces; i += kBwThreads) {
        const int j = i / kPieces, q = i - j * kPieces;
        cp_async16(dst + j * kScBlock + q * 4,
                   p.top_diff + ((size_t)s_ids[k * kScChunk + j] * p.C + c0) * (kOut * kOut) + q * 4);
      }
      cp_async_commit();
      if (tid < 2 * n)
        fill_axis(&ax[buf * kScChunk + (tid >> 1)], p.rois + (size_t)s_ids[k * kScChunk + (tid >> 1)] * 5, tid & 1,
                  p.scale, p.H, p.W);
    };

    if (nchunks > 0) prefetch(0);
    for (int k = 0; k < nchunks; ++k) {
      cp_async_wait_all();
      __syncthreads();  // chunk k staged, its tables written; everyone is done with chunk k - 1
      if (k + 1 < nchunks) prefetch(k + 1);
      const int n = min(kScChunk, n_ids - k * kScChunk), buf = k & 1;
      const float* st = stage + (size_t)buf * kScChunk * kScBlock;
      // sample gradients = avg_pool2d(2, 1) backward: the <= 4 outputs whose window holds the sample
      for (int i = tid; i < n * kBwCg * 64; i += kBwThreads) {
        const int sidx = i & 63, ph = sidx >> 3, spw = sidx & 7;
        const float* gg = st + (i >> 6) * (kOut * kOut);  // (RoI, channel) block
        float v = 0.f;
        if (ph > 0 && spw > 0) v += gg[(ph - 1) * kOut + spw - 1];
        if (ph > 0 && spw < kOut) v += gg[(ph - 1) * kOut + spw];
        if (ph < kOut && spw > 0) v += gg[ph * kOut + spw - 1];
        if (ph < kOut && spw < kOut) v += gg[ph * kOut + spw];
        gs[(i >> 6) * kGsStride + sidx] = v;
      }
      __syncthreads();
      // a warp scatters one sample point per pass: lane = (channel, corner)
      const float* my_gs = gs + c * kGsStride;
      for (int it = warp; it < n * 64; it += kBwWarps) {
        const int j = it >> 6, sidx = it & 63, ph = sidx >> 3, pw = sidx & 7;
        const BwAxis& a = ax[buf * kScChunk + j];
        const int hc = a.hcell[ph], wc = a.wcell[pw];
        if (hc < 0 || wc < 0) continue;  // sample outside the map (warp-uniform)
        const float wgt = (ky ? a.h1[ph] : a.h0[ph]) * (kx ? a.w1[pw] : a.w0[pw]);
        atomicAdd(my_plane + hc * p.W + wc, wgt * my_gs[j * (kBwCg * kGsStride) + sidx]);
      }
    }
    __syncthreads();  // s_ids and the staging buffers are free again; the slab is complete when this was the last pass
    if (r_next >= p.R) break;
  }

  if (ACC && !any) return;  // nothing to add
  // slab -> its contiguous block of bottom_diff, in pieces, one issuing thread each
  fence_proxy_async_smem();
  __syncthreads();
  const uint32_t total = (uint32_t)kBwCg * hw * 4;
  const uint32_t off = (uint32_t)tid * kScPiece;
  if (off < total) {
    const uint32_t bytes = min(kScPiece, total - off);
    char* dst = reinterpret_cast<char*>(p.bottom_diff + ((size_t)f * p.C + c0) * hw) + off;
    const char* src = reinterpret_cast<const char*>(slab) + off;
    if (ACC) bulk_s2g_add_f32(dst, src, bytes);
    else bulk_s2g(dst, src, bytes);
    bulk_commit();
    bulk_wait_read<0>();  // shared memory must outlive the copy
  }
}

}  // namespace

// 1 launched, 0 not eligible (caller uses the generic kernel), < 0 launch error
int try_launch_avg_bwd_gather(const float* top_diff, float scale, int B, int R, int H, int W, int C,
                              const float* rois, float* bottom_diff, cudaStream_t stream) {
  if (C % kBwCg != 0 || H < 2 || W < 2 || H > kBwMaxDim || W > kBwMaxDim || H * W > kBwCells * kBwThreads) return 0;
  const size_t smem = (size_t)kCsrChunk * 64 * kBwCg * 4 + sizeof(CsrEnt) * kCsrMaxEnt + sizeof(BwAxis) * kCsrChunk;
  static_assert((size_t)kCsrChunk * 2 * kBwMaxDim * sizeof(uchar4) <= (size_t)kCsrChunk * 64 * kBwCg * 4,
                "the row / column maps alias the sample-gradient buffer");
  cudaError_t e = cudaFuncSetAttribute(align_avg_bwd_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("roi_align backward: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  BwParams p;
  p.top_diff = top_diff;
  p.rois = rois;
  p.bottom_diff = bottom_diff;
  p.scale = scale;
  p.B = B;
  p.R = R;
  p.H = H;
  p.W = W;
  p.C = C;
  p.splits = 0;
  const long long units = (long long)B * (C / kBwCg);
  const long long cap = (long long)sm_count() * 2;  // two CTAs per SM, persistent
  const int grid = (int)(units < cap ? units : cap);
  align_avg_bwd_gather<<<grid, kBwThreads, smem, stream>>>(p);
  return launch_status("align_avg_bwd_gather");
}

// 1 launched, 0 not eligible, < 0 launch error.  accumulate: add into bottom_diff instead of overwriting it.
int try_launch_avg_bwd_scatter(const float* top_diff, float scale, int B, int R, int H, int W, int C,
                               const float* rois, float* bottom_diff, bool accumulate, cudaStream_t stream) {
  if (C % kBwCg != 0 || H < 2 || W < 2) return 0;
  if ((long long)B * (C / kBwCg) > 0x7fffffffLL) return 0;
  if ((reinterpret_cast<uintptr_t>(top_diff) & 15) != 0 || (reinterpret_cast<uintptr_t>(bottom_diff) & 15) != 0) return 0;
  const size_t smem = (size_t)kBwCg * H * W * 4 + (size_t)2 * kScChunk * kScBlock * 4 +
                      (size_t)kScChunk * kBwCg * kGsStride * 4 + sizeof(BwAxis) * 2 * kScChunk;
  if (smem > 200 * 1024) return 0;  // slab does not fit: the generic kernel takes over
  auto* kern = accumulate ? align_avg_bwd_scatter<true> : align_avg_bwd_scatter<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("roi_align backward: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  BwParams p;
  p.top_diff = top_diff;
  p.rois = rois;
  p.bottom_diff = bottom_diff;
  p.scale = scale;
  p.B = B;
  p.R = R;
  p.H = H;
  p.W = W;
  p.C = C;
  p.splits = 0;
  kern<<<B * (C / kBwCg), kBwThreads, smem, stream>>>(p);
  return launch_status("align_avg_bwd_scatter");
}

}  // namespace nafae
