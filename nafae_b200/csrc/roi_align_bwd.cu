// RoIAlignAvg backward WITHOUT atomics (sm_100a): cell-gather with exclusive ownership.
//
// Reference: ROIAlignBackward (lib/model/roi_align/src/roi_align_kernel.cu:94-143) scatters every
// sample's gradient to four cells with atomicAdd into a pre-zeroed (B, C, H, W) tensor, preceded by
// avg_pool2d's backward (autograd); at cfg2 that is 26 M float atomics and a 156 MB memset.
// Here a CTA owns one (frame, 8-channel) slab of bottom_diff outright: every thread owns a few
// CELLS of that slab for all 8 channels and GATHERS what the frame's RoIs send there, so the output
// is written exactly once, coalesced, with plain stores -- no atomics, no zero-fill, a fixed summation
// order (deterministic).  Which samples of a RoI reach a cell is separable: sample row ph reaches cell
// row y with weight (1 - h_ratio) when hstart[ph] == y and h_ratio when hstart[ph] == y - 1 (ranges
// of ph, monotone), likewise for columns; two byte-tables per RoI (rows, columns) make the test two
// shared-memory reads, and most (RoI, cell) pairs are rejected by them.
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
constexpr int kBwChunk = 24;      // RoIs whose tables / gradients are resident at a time
constexpr int kBwCells = 4;       // cells per thread  => H*W <= 2048
constexpr int kBwMaxDim = 128;    // H, W <= 128
constexpr int kBwGroups = 4;      // channel groups (of kBwCg) a CTA works through with one set of tables

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

// Two CTAs per SM (<= 64 registers, ~85 KB of shared memory each): one CTA's global round trips
// (RoI scan, gradient staging) hide behind the other's gather.
__global__ void __launch_bounds__(kBwThreads, 2) align_avg_bwd_gather(const BwParams p) {
  extern __shared__ __align__(16) unsigned char bw_smem[];
  float* gs = reinterpret_cast<float*>(bw_smem);                       // [chunk][64 samples][8 ch]
  float* gst = gs + kBwChunk * 64 * kBwCg;                             // [chunk][8 ch][49]
  BwAxis* ax = reinterpret_cast<BwAxis*>(gst + kBwChunk * kBwCg * 49); // [chunk]
  uchar4* rowmap = reinterpret_cast<uchar4*>(ax + kBwChunk);           // [chunk][H]: lo0, n0, lo1, n1
  uchar4* colmap = rowmap + kBwChunk * p.H;                            // [chunk][W]
  __shared__ int s_ids[kBwChunk];
  __shared__ int s_wcnt[kBwWarps];
  __shared__ int s_n, s_next;

  const int tid = threadIdx.x;
  const int f = blockIdx.x / p.splits, split = blockIdx.x % p.splits;
  const int ch_per_split = p.C / p.splits;
  const int ch_begin = split * ch_per_split;
  const int ngroups = ch_per_split / kBwCg;
  const int hw = p.H * p.W;

  int cy[kBwCells], cx[kBwCells];
#pragma unroll
  for (int i = 0; i < kBwCells; ++i) {
    const int cell = tid + i * kBwThreads;
    cy[i] = cell < hw ? cell / p.W : -1;
    cx[i] = cell < hw ? cell - (cell / p.W) * p.W : -1;
  }

  int r_next = 0;
  bool first_chunk = true;
  for (;;) {
    // ---- next chunk: the first kBwChunk RoIs of frame f at or after r_next, in index order
    frame_rois(p.rois, p.R, f, r_next, kBwChunk, s_ids, s_wcnt, &s_n, &s_next);
    const int n = s_n;
    r_next = s_next;
    if (n == 0 && !first_chunk) break;

    // ---- per-RoI axis tables: one thread per (RoI, axis)
    if (tid < 2 * n) {
      fill_axis(&ax[tid >> 1], p.rois + (size_t)s_ids[tid >> 1] * 5, tid & 1, p.scale, p.H, p.W);
    }
    __syncthreads();
    // ---- row / column maps: which sample rows reach cell row y with weight h0 (hstart == y) / h1
    for (int i = tid; i < n * (p.H + p.W); i += kBwThreads) {
      const int j = i / (p.H + p.W), k = i - j * (p.H + p.W);
      const bool is_w = k >= p.H;
      const int v = is_w ? k - p.H : k;
      const int* cells = is_w ? ax[j].wcell : ax[j].hcell;
      int lo0 = 0, n0 = 0, lo1 = 0, n1 = 0;
#pragma unroll
      for (int s = 0; s < kS; ++s) {
        const int cc = cells[s];
        if (cc == v) {
          if (n0 == 0) lo0 = s;
          ++n0;
        }
        if (cc == v - 1) {
          if (n1 == 0) lo1 = s;
          ++n1;
        }
      }
      const uchar4 m = make_uchar4((unsigned char)lo0, (unsigned char)n0, (unsigned char)lo1, (unsigned char)n1);
      if (is_w) colmap[j * p.W + v] = m;
      else rowmap[j * p.H + v] = m;
    }

    // ---- the channel groups of this CTA, all with the tables above
    for (int g = 0; g < ngroups; ++g) {
      const int c0 = ch_begin + g * kBwCg;
      __syncthreads();  // previous group's gather is done with gs / gst (and the maps are complete)
      // output gradients of the chunk -> shared (coalesced: 8 channels x 49 are contiguous per RoI)
      for (int i = tid; i < n * kBwCg * 49; i += kBwThreads) {
        const int j = i / (kBwCg * 49), rem = i - j * (kBwCg * 49);
        gst[i] = __ldg(p.top_diff + ((size_t)s_ids[j] * p.C + c0) * 49 + rem);
      }
      // accumulators: zero for the first chunk of the frame, else what the earlier chunks left
      float acc[kBwCells][kBwCg];
#pragma unroll
      for (int i = 0; i < kBwCells; ++i) {
        const int cell = tid + i * kBwThreads;
#pragma unroll
        for (int c = 0; c < kBwCg; ++c)
          acc[i][c] = (first_chunk || cell >= hw) ? 0.f : p.bottom_diff[((size_t)f * p.C + c0 + c) * hw + cell];
      }
      __syncthreads();
      // sample gradients: avg_pool2d(2, 1) backward, layout [RoI][sample][channel]
      for (int i = tid; i < n * 64 * kBwCg; i += kBwThreads) {
        const int c = i % kBwCg, s = (i / kBwCg) % 64, j = i / (kBwCg * 64);
        const int ph = s >> 3, pw = s & 7;
        const float* gg = gst + (j * kBwCg + c) * 49;
        float v = 0.f;
        if (ph > 0 && pw > 0) v += gg[(ph - 1) * kOut + pw - 1];
        if (ph > 0 && pw < kOut) v += gg[(ph - 1) * kOut + pw];
        if (ph < kOut && pw > 0) v += gg[ph * kOut + pw - 1];
        if (ph < kOut && pw < kOut) v += gg[ph * kOut + pw];
        gs[i] = v;
      }
      __syncthreads();
      // gather: every owned cell collects from every RoI of the chunk
#pragma unroll
      for (int i = 0; i < kBwCells; ++i) {
        if (cy[i] < 0) continue;
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
              const float wgt = wy * (qq < cm.y ? a.w0[pw] : a.w1[pw]);
              const float4* g4 = reinterpret_cast<const float4*>(gs + ((size_t)j * 64 + ph * 8 + pw) * kBwCg);
              const float4 ga = g4[0], gb = g4[1];
              acc[i][0] = fmaf(wgt, ga.x, acc[i][0]);
              acc[i][1] = fmaf(wgt, ga.y, acc[i][1]);
              acc[i][2] = fmaf(wgt, ga.z, acc[i][2]);
              acc[i][3] = fmaf(wgt, ga.w, acc[i][3]);
              acc[i][4] = fmaf(wgt, gb.x, acc[i][4]);
              acc[i][5] = fmaf(wgt, gb.y, acc[i][5]);
              acc[i][6] = fmaf(wgt, gb.z, acc[i][6]);
              acc[i][7] = fmaf(wgt, gb.w, acc[i][7]);
            }
          }
        }
      }
      // every cell of the slab is written by its owner only (frames without RoIs: zeros)
#pragma unroll
      for (int i = 0; i < kBwCells; ++i) {
        const int cell = tid + i * kBwThreads;
        if (cell < hw) {
#pragma unroll
          for (int c = 0; c < kBwCg; ++c)
            p.bottom_diff[((size_t)f * p.C + c0 + c) * hw + cell] = acc[i][c];
        }
      }
    }
    first_chunk = false;
    if (r_next >= p.R) break;
    __syncthreads();  // the next chunk rewrites the shared tables
  }
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory scatter (the default): a CTA owns one (frame, 8-channel) slab of bottom_diff IN SHARED
// MEMORY (8 x H x W floats, 61 KB at 38 x 50), scatters the frame's RoIs into it with shared-memory
// atomics and stores the finished slab -- which is one contiguous block of (B, C, H, W) -- with bulk
// async copies (cp.async.bulk shared -> global; cp.reduce...add.f32 when the caller accumulates into an
// existing tensor).  HBM sees every output byte once and no atomics; the 105 M scatter adds of cfg2
// stay on chip.  A warp handles one sample point per instruction, lane = (channel, corner): the 32
// addresses of a warp instruction are distinct (and bank-conflict free for even W), so the CAS loop
// behind a shared float atomicAdd spins only when two WARPS collide.  The RoIs' output gradients are
// staged by 16-byte cp.async, double buffered in chunks of kScChunk RoIs.  The order in which warps
// reach a cell is not fixed: like the reference's atomicAdd the sum is not bitwise reproducible
// (NAFAE_FLAG_DETERMINISTIC selects the gather kernel above instead).
constexpr int kScChunk = 8;     // RoIs staged at a time (x2 buffers)
constexpr int kScIds = 128;     // RoI indices of the frame held at a time
constexpr int kScBlock = kBwCg * kOut * kOut;   // floats of one RoI's 8-channel gradient block (1568 B)
constexpr uint32_t kScPiece = 32768;            // bytes per bulk store

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
  BwAxis* ax = reinterpret_cast<BwAxis*>(stage + 2 * kScChunk * kScBlock); // [2][chunk]
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

  // lane roles: channel, corner
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
      for (int it = warp; it < n * 64; it += kBwWarps) {
        const int j = it >> 6, ph = (it >> 3) & 7, pw = it & 7;
        const BwAxis& a = ax[buf * kScChunk + j];
        const int hc = a.hcell[ph], wc = a.wcell[pw];
        if (hc < 0 || wc < 0) continue;  // sample outside the map (warp-uniform)
        const float* gg = st + (j * kBwCg + c) * (kOut * kOut);
        float v = 0.f;  // avg_pool2d(2, 1) backward: the <= 4 outputs whose window holds the sample
        if (ph > 0 && pw > 0) v += gg[(ph - 1) * kOut + pw - 1];
        if (ph > 0 && pw < kOut) v += gg[(ph - 1) * kOut + pw];
        if (ph < kOut && pw > 0) v += gg[ph * kOut + pw - 1];
        if (ph < kOut && pw < kOut) v += gg[ph * kOut + pw];
        const float wgt = (ky ? a.h1[ph] : a.h0[ph]) * (kx ? a.w1[pw] : a.w0[pw]);
        atomicAdd(my_plane + hc * p.W + wc, wgt * v);
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
  if ((long long)B * (C / kBwCg) > 0x7fffffffLL) return 0;
  const size_t smem = (size_t)kBwChunk * 64 * kBwCg * 4 + (size_t)kBwChunk * kBwCg * 49 * 4 +
                      sizeof(BwAxis) * kBwChunk + sizeof(uchar4) * kBwChunk * (size_t)(H + W);
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
  // CTAs per frame: each works through kBwGroups channel groups with one set of tables (fewer when C is small)
  int per_cta = kBwGroups * kBwCg;
  while (per_cta > kBwCg && C % per_cta != 0) per_cta -= kBwCg;
  p.splits = C / per_cta;
  align_avg_bwd_gather<<<B * p.splits, kBwThreads, smem, stream>>>(p);
  return launch_status("align_avg_bwd_gather");
}

// 1 launched, 0 not eligible, < 0 launch error.  accumulate: add into bottom_diff instead of overwriting it.
int try_launch_avg_bwd_scatter(const float* top_diff, float scale, int B, int R, int H, int W, int C,
                               const float* rois, float* bottom_diff, bool accumulate, cudaStream_t stream) {
  if (C % kBwCg != 0 || H < 2 || W < 2) return 0;
  if ((long long)B * (C / kBwCg) > 0x7fffffffLL) return 0;
  if ((reinterpret_cast<uintptr_t>(top_diff) & 15) != 0 || (reinterpret_cast<uintptr_t>(bottom_diff) & 15) != 0) return 0;
  const size_t smem = (size_t)kBwCg * H * W * 4 + (size_t)2 * kScChunk * kScBlock * 4 + sizeof(BwAxis) * 2 * kScChunk;
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
