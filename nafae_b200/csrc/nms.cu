// Batched greedy NMS for RPN proposals on sm_100a.
//
// Replaces (reference tree paths):
//   lib/model/nms/src/nms_cuda_kernel.cu:31-39   devIoU
//   lib/model/nms/src/nms_cuda_kernel.cu:41-85   nms_kernel (64x64 IoU bitmask tiles)
//   lib/model/nms/src/nms_cuda_kernel.cu:87-161  nms_cuda_compute (malloc, D2H mask, HOST sweep)
//   lib/model/rpn/proposal_layer.py:127-163      per-frame Python loop around nms()
//
// Design (B200-first, not a translation):
//   * nms_mask_kernel   warp-ballot IoU bitmask: a CTA owns one 64x64 tile of one frame, row and
//                       column boxes are staged in shared memory, every lane owns two column
//                       boxes and a __ballot_sync over the comparison IS the mask word.  Only
//                       the upper triangle is computed (the reference computes and ignores the
//                       lower one, nms_cuda_kernel.cu:46,140).  Frames are batched in gridDim.z.
//   * nms_sweep_kernel  the greedy sweep runs ON THE DEVICE (one CTA per frame): no 696 KB D2H
//                       mask copy, no host loop, no second sync for num_out.
//   * proposal_tail_kernel  the hot-path form: NMS + "first post_nms_topN keeps" + zero padding
//                       for all frames in one launch.  Greedy NMS's first N keeps depend only on
//                       a prefix of the sorted boxes, so each frame stops scanning as soon as N
//                       boxes are kept: O(scanned * N) IoUs instead of O(n^2), no mask in HBM.
//
// Bit-exactness: the IoU is evaluated with the reference's fp32 operation order *as nvcc
// compiles it* (column-box area fused into the union by an FMA, IEEE division), pinned with
// __f*_rn intrinsics so this file's own compilation cannot re-associate it.  A cheap two-sided
// filter decides almost every pair without the division; only pairs within 2^-20 of the
// threshold take the exact division, so the decision is always the reference's.
#include <mutex>

#include "common.cuh"

namespace nafae {
namespace {

constexpr int kTile = 64;  // boxes per mask word, nms_cuda_kernel.cu:29

struct Thresh {
  float t;     // nms_overlap_thresh
  float t_hi;  // t * (1 + 2^-20)
  float t_lo;  // t * (1 - 2^-20)
  int fast;    // filter usable (0 < t < 2^20)
};

__host__ inline Thresh make_thresh(float t) {
  Thresh r;
  r.t = t;
  r.fast = (t > 0.f && t < 1048576.f) ? 1 : 0;
  r.t_hi = t * (1.f + 9.5367431640625e-7f);
  r.t_lo = t * (1.f - 9.5367431640625e-7f);
  return r;
}

struct RowBox {  // "a" of devIoU: the earlier (higher score) box
  float x1, y1, x2, y2;
  float area;  // RN((x2-x1+1)*(y2-y1+1))
};
struct ColBox {  // "b" of devIoU
  float x1, y1, x2, y2;
  float w, h;  // (x2-x1)+1, (y2-y1)+1 : multiplied inside the FMA
};

__device__ __forceinline__ RowBox make_row(float x1, float y1, float x2, float y2) {
  RowBox r{x1, y1, x2, y2, 0.f};
  r.area = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
  return r;
}
__device__ __forceinline__ ColBox make_col(float x1, float y1, float x2, float y2) {
  ColBox c{x1, y1, x2, y2, 0.f, 0.f};
  c.w = __fadd_rn(__fsub_rn(x2, x1), 1.f);
  c.h = __fadd_rn(__fsub_rn(y2, y1), 1.f);
  return c;
}

// devIoU(a, b) > thresh, decision identical to the compiled reference.
__device__ __forceinline__ bool iou_exceeds(const RowBox& a, const ColBox& b, const Thresh& th) {
  float left = fmaxf(a.x1, b.x1), right = fminf(a.x2, b.x2);
  float top = fmaxf(a.y1, b.y1), bottom = fminf(a.y2, b.y2);
  float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
  float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
  float inter = __fmul_rn(width, height);
  float uni = __fsub_rn(__fmaf_rn(b.w, b.h, a.area), inter);
  if (th.fast && uni > 0.f) {
    // inter/uni is within 2^-24 of the true quotient; the two products below are within 2^-24
    // of thresh*(1 +- 2^-20)*uni, so outside the band the rounded quotient compares the same way.
    if (inter > __fmul_rn(uni, th.t_hi)) return true;
    if (inter < __fmul_rn(uni, th.t_lo)) return false;
  }
  return __fdiv_rn(inter, uni) > th.t;
}

// ------------------------------------------------------------------ full keep list ----

// grid (col_blocks, col_blocks, F), block 128.  mask layout (F, n, col_blocks) u64, identical to
// the reference's dev_mask; words with column block < row block are never written nor read.
__global__ void __launch_bounds__(128)
nms_mask_kernel(const float* __restrict__ boxes, int n, int dim, Thresh th,
                unsigned long long* __restrict__ mask) {
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (cb < rb) return;
  const int col_blocks = gridDim.x;
  const float* fb = boxes + (size_t)blockIdx.z * n * dim;
  unsigned long long* fm = mask + (size_t)blockIdx.z * n * col_blocks;

  __shared__ RowBox rows[kTile];
  __shared__ ColBox cols[kTile];
  const int row_size = min(n - rb * kTile, kTile);
  const int col_size = min(n - cb * kTile, kTile);
  const int tid = threadIdx.x;
  if (tid < kTile) {
    if (tid < row_size) {
      const float* p = fb + (size_t)(rb * kTile + tid) * dim;
      rows[tid] = make_row(p[0], p[1], p[2], p[3]);
    }
  } else {
    const int j = tid - kTile;
    if (j < col_size) {
      const float* p = fb + (size_t)(cb * kTile + j) * dim;
      cols[j] = make_col(p[0], p[1], p[2], p[3]);
    }
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const bool v0 = lane < col_size, v1 = lane + 32 < col_size;
  ColBox c0 = cols[v0 ? lane : 0], c1 = cols[v1 ? lane + 32 : 0];
  unsigned long long mine = 0;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int r = warp * 16 + i;
    if (r >= row_size) break;  // warp-uniform
    const RowBox a = rows[r];
    bool p0 = v0 && iou_exceeds(a, c0, th);
    bool p1 = v1 && iou_exceeds(a, c1, th);
    if (rb == cb) {  // diagonal tile: only later boxes (nms_cuda_kernel.cu:74-76)
      p0 = p0 && (lane > r);
      p1 = p1 && (lane + 32 > r);
    }
    const unsigned lo = __ballot_sync(0xffffffffu, p0);
    const unsigned hi = __ballot_sync(0xffffffffu, p1);
    if (lane == i) mine = ((unsigned long long)hi << 32) | lo;
  }
  const int r = warp * 16 + lane;
  if (lane < 16 && r < row_size) fm[(size_t)(rb * kTile + r) * col_blocks + cb] = mine;
}

// grid F, block 256.  Device-side restatement of the host sweep (nms_cuda_kernel.cu:132-144):
// walk the 64-box blocks in order; inside a block the greedy choice is a 64-step scan over the
// diagonal tile's words, then the kept rows' words are OR-ed into the later blocks' removal
// words by all threads (independent loads, shared-memory atomics).
constexpr int kSweepMaxBlocks = 4096;  // n <= 262144

__global__ void __launch_bounds__(256)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, int n, int col_blocks,
                 int* __restrict__ keep_out, int* __restrict__ num_out) {
  extern __shared__ unsigned long long remv[];  // col_blocks words
  __shared__ unsigned long long diag[kTile];
  __shared__ int kept_row[kTile];
  __shared__ unsigned long long kept_bits;
  __shared__ int s_count;

  const unsigned long long* fm = mask + (size_t)blockIdx.x * n * col_blocks;
  int* keep = keep_out + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  for (int j = tid; j < col_blocks; j += blockDim.x) remv[j] = 0;
  if (tid == 0) s_count = 0;
  __syncthreads();

  for (int blk = 0; blk < col_blocks; ++blk) {
    const int size = min(n - blk * kTile, kTile);
    if (tid < kTile) diag[tid] = tid < size ? fm[(size_t)(blk * kTile + tid) * col_blocks + blk] : 0;
    __syncthreads();
    if (tid == 0) {
      unsigned long long w = remv[blk];
      if (size < kTile) w |= ~0ULL << size;  // boxes past n do not exist
      unsigned long long kb = 0;
      if (w != ~0ULL) {
#pragma unroll 8
        for (int k = 0; k < kTile; ++k) {
          const unsigned long long d = diag[k];
          if (!((w >> k) & 1ULL)) {
            kb |= 1ULL << k;
            w |= d;
          }
        }
      }
      kept_bits = kb;
    }
    __syncthreads();
    const unsigned long long kb = kept_bits;
    const int base = s_count;
    const int nk = __popcll(kb);
    if (tid < kTile && ((kb >> tid) & 1ULL)) {
      const int pos = __popcll(kb & ((1ULL << tid) - 1ULL));
      keep[base + pos] = blk * kTile + tid;
      kept_row[pos] = blk * kTile + tid;
    }
    __syncthreads();
    const int later = col_blocks - blk - 1;
    for (int p = tid; p < nk * later; p += blockDim.x) {
      const int k = p / later, j = blk + 1 + p % later;
      const unsigned long long wv = fm[(size_t)kept_row[k] * col_blocks + j];
      if (wv) atomicOr(&remv[j], wv);
    }
    if (tid == 0) s_count = base + nk;
    __syncthreads();
  }
  if (tid == 0) num_out[blockIdx.x] = s_count;
}

// ------------------------------------------------------- fused proposal-layer tail ----

// grid F, block 256 (4 threads per candidate), dynamic smem = post_topn * sizeof(RowBox).
// One chunk = 64 consecutive candidates (score order).  Per chunk:
//   1. every candidate is tested against the kept list (a = kept box, b = candidate); the four
//      threads of a candidate split the list and OR their verdicts
//   2. survivors compute their row of the chunk's diagonal tile (a = this box, b = later box),
//      again split four ways
//   3. lane 0 resolves the chunk greedily (64-step bit scan), stopping at post_topn
//   4. the newly kept boxes are appended to the kept list and written out
constexpr int kTailThreads = 4 * kTile;

__global__ void __launch_bounds__(kTailThreads)
proposal_tail_kernel(const float* __restrict__ proposals, const float* __restrict__ scores, int n,
                     int m, int post_topn, Thresh th, float* __restrict__ rois,
                     float* __restrict__ roi_scores, int* __restrict__ num_kept) {
  extern __shared__ RowBox kept[];  // post_topn entries
  __shared__ ColBox chunk[kTile];
  __shared__ unsigned long long diag[kTile];
  __shared__ int s_flag[kTile];
  __shared__ unsigned long long s_new;
  __shared__ int s_count;

  NAFAE_CTA_TRACE(cta_trace, 2);
  const int f = blockIdx.x, tid = threadIdx.x;
  const int t = tid >> 2, q = tid & 3;  // candidate slot, quarter
  const float* fp = proposals + (size_t)f * n * 4;
  const float* fs = scores + (size_t)f * n;
  float* out = rois + (size_t)f * post_topn * 5;
  float* out_s = roi_scores + (size_t)f * post_topn;
  if (tid == 0) s_count = 0;
  __syncthreads();

  for (int c0 = 0; c0 < m; c0 += kTile) {
    const int count = s_count;
    if (count >= post_topn) break;  // uniform: early exit, the rest of the frame is never read
    const int size = min(m - c0, kTile);
    const bool valid = t < size;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) bx = __ldg(reinterpret_cast<const float4*>(fp + (size_t)(c0 + t) * 4));
    const ColBox cb = make_col(bx.x, bx.y, bx.z, bx.w);
    if (q == 0) chunk[t] = cb;
    int sup = valid ? 0 : 1;
    if (valid) {
#pragma unroll 2
      for (int k = q; k < count; k += 4) sup |= iou_exceeds(kept[k], cb, th) ? 1 : 0;
    }
    sup |= __shfl_xor_sync(0xffffffffu, sup, 1);
    sup |= __shfl_xor_sync(0xffffffffu, sup, 2);
    if (q == 0) s_flag[t] = sup;
    __syncthreads();
    unsigned dlo = 0, dhi = 0;
    if (!sup) {
      const RowBox a = make_row(bx.x, bx.y, bx.z, bx.w);
#pragma unroll 2
      for (int j = t + 1 + q; j < size; j += 4)
        if (iou_exceeds(a, chunk[j], th)) {
          if (j < 32) dlo |= 1u << j;
          else dhi |= 1u << (j - 32);
        }
    }
    dlo |= __shfl_xor_sync(0xffffffffu, dlo, 1);
    dhi |= __shfl_xor_sync(0xffffffffu, dhi, 1);
    dlo |= __shfl_xor_sync(0xffffffffu, dlo, 2);
    dhi |= __shfl_xor_sync(0xffffffffu, dhi, 2);
    if (q == 0) diag[t] = ((unsigned long long)dhi << 32) | dlo;
    __syncthreads();
    if (tid < 32) {
      const unsigned lo = __ballot_sync(0xffffffffu, s_flag[tid] != 0);
      const unsigned hi = __ballot_sync(0xffffffffu, s_flag[tid + 32] != 0);
      if (tid == 0) {
        unsigned long long w = ((unsigned long long)hi << 32) | lo;
        unsigned long long kb = 0;
        int cnt = count;
#pragma unroll 8
        for (int k = 0; k < kTile; ++k) {
          const unsigned long long dk = diag[k];
          if (!((w >> k) & 1ULL) && cnt < post_topn) {
            kb |= 1ULL << k;
            w |= dk;
            ++cnt;
          }
        }
        s_new = kb;
        s_count = cnt;
      }
    }
    __syncthreads();
    const unsigned long long kb = s_new;
    if (q == 0 && ((kb >> t) & 1ULL)) {
      const int pos = count + __popcll(kb & ((1ULL << t) - 1ULL));
      kept[pos] = make_row(bx.x, bx.y, bx.z, bx.w);
      float* o = out + (size_t)pos * 5;
      o[0] = (float)f;
      o[1] = bx.x;
      o[2] = bx.y;
      o[3] = bx.z;
      o[4] = bx.w;
      out_s[pos] = fs[c0 + t];
    }
    __syncthreads();
  }
  // zero padding, frame index in column 0 of every row (proposal_layer.py:127,160)
  const int count = s_count;
  for (int k = count + tid; k < post_topn; k += kTailThreads) {
    float* o = out + (size_t)k * 5;
    o[0] = (float)f;
    o[1] = o[2] = o[3] = o[4] = 0.f;
    out_s[k] = 0.f;
  }
  if (tid == 0 && num_kept) num_kept[f] = count;
}

struct ScratchCache {
  void* ptr = nullptr;
  size_t bytes = 0;
  int device = -1;
};
ScratchCache g_nms_scratch;
std::mutex g_nms_scratch_mu;  // nms_cuda_compute may be called from several host threads

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_CTA_TRACE_READER(nafae_debug_cta_trace_nms)

NAFAE_API size_t nafae_nms_workspace_bytes(int num_frames, int boxes_num) {
  if (num_frames <= 0 || boxes_num <= 0) return 0;
  const size_t col_blocks = (size_t)ceil_div(boxes_num, kTile);
  return align_up((size_t)num_frames * boxes_num * col_blocks * sizeof(unsigned long long), 256);
}

NAFAE_API int nafae_nms_batched(int* keep_out, int* num_out, const float* boxes, int num_frames,
                                int boxes_num, int boxes_dim, float nms_overlap_thresh,
                                void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  NAFAE_REQUIRE(num_frames >= 0 && boxes_num >= 0, "nms: negative sizes");
  if (num_frames == 0) return 1;
  NAFAE_REQUIRE(num_out != nullptr, "nms: num_out is NULL");
  if (boxes_num == 0) {
    cudaMemsetAsync(num_out, 0, sizeof(int) * num_frames, stream);
    return launch_status("nms memset");
  }
  NAFAE_REQUIRE(boxes_dim >= 4, "nms: boxes_dim must be >= 4, got %d", boxes_dim);
  NAFAE_REQUIRE(keep_out && boxes, "nms: NULL buffer");
  const int col_blocks = ceil_div(boxes_num, kTile);
  NAFAE_REQUIRE(col_blocks <= kSweepMaxBlocks, "nms: boxes_num %d exceeds %d", boxes_num,
                kSweepMaxBlocks * kTile);
  NAFAE_REQUIRE(num_frames <= 65535, "nms: more than 65535 frames per call");
  const size_t need = nafae_nms_workspace_bytes(num_frames, boxes_num);
  NAFAE_REQUIRE(workspace && workspace_bytes >= need, "nms: workspace too small (%zu < %zu)",
                workspace_bytes, need);
  const Thresh th = make_thresh(nms_overlap_thresh);
  auto* mask = static_cast<unsigned long long*>(workspace);
  dim3 grid(col_blocks, col_blocks, num_frames);
  nms_mask_kernel<<<grid, 128, 0, stream>>>(boxes, boxes_num, boxes_dim, th, mask);
  int st = launch_status("nms_mask_kernel");
  if (st != 1) return st;
  nms_sweep_kernel<<<num_frames, 256, col_blocks * sizeof(unsigned long long), stream>>>(
      mask, boxes_num, col_blocks, keep_out, num_out);
  return launch_status("nms_sweep_kernel");
}

// Reference-named entry point (lib/model/nms/src/nms_cuda_kernel.h:5-6): legacy default stream,
// internal grow-only scratch so that the signature stays exactly the reference's.
NAFAE_API void nms_cuda_compute(int* keep_out, int* num_out, float* boxes_host, int boxes_num,
                                int boxes_dim, float nms_overlap_thresh) {
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t need = nafae_nms_workspace_bytes(1, boxes_num);
  std::lock_guard<std::mutex> lock(g_nms_scratch_mu);  // held across the launch: it uses the scratch
  ScratchCache& sc = g_nms_scratch;
  if (need > 0 && (sc.device != dev || sc.bytes < need)) {
    if (sc.ptr) {
      cudaStreamSynchronize(0);
      cudaFree(sc.ptr);
    }
    sc.ptr = nullptr;
    sc.bytes = 0;
    if (cudaMalloc(&sc.ptr, need) != cudaSuccess) {
      set_error("nms_cuda_compute: cudaMalloc(%zu) failed", need);
      return;
    }
    sc.bytes = need;
    sc.device = dev;
  }
  nafae_nms_batched(keep_out, num_out, boxes_host, 1, boxes_num, boxes_dim, nms_overlap_thresh,
                    sc.ptr, sc.bytes, 0);
}

NAFAE_API int nafae_proposal_tail(const float* proposals, const float* scores, int num_frames,
                                  int boxes_num, int pre_nms_topn, int post_nms_topn,
                                  float nms_thresh, float* rois, float* roi_scores, int* num_kept,
                                  cudaStream_t stream) {
  NAFAE_REQUIRE(num_frames >= 0 && boxes_num >= 0, "proposal_tail: negative sizes");
  NAFAE_REQUIRE(post_nms_topn > 0, "proposal_tail: post_nms_topn must be > 0");
  if (num_frames == 0) return 1;
  NAFAE_REQUIRE(rois && roi_scores, "proposal_tail: NULL output");
  NAFAE_REQUIRE(boxes_num == 0 || (proposals && scores), "proposal_tail: NULL input");
  NAFAE_REQUIRE((reinterpret_cast<uintptr_t>(proposals) & 15) == 0,
                "proposal_tail: proposals must be 16-byte aligned");
  // proposal_layer.py:139 compares pre_nms_topN with scores_keep.numel() (= F*n)
  int m = boxes_num;
  if (pre_nms_topn > 0 && (long long)pre_nms_topn < (long long)num_frames * boxes_num &&
      pre_nms_topn < boxes_num)
    m = pre_nms_topn;
  const size_t smem = (size_t)post_nms_topn * sizeof(RowBox);
  NAFAE_REQUIRE(smem <= 200 * 1024, "proposal_tail: post_nms_topn %d too large", post_nms_topn);
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(proposal_tail_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("proposal_tail: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return -(int)e;
    }
  }
  const Thresh th = make_thresh(nms_thresh);
  proposal_tail_kernel<<<num_frames, kTailThreads, smem, stream>>>(proposals, scores, boxes_num, m,
                                                            post_nms_topn, th, rois, roi_scores,
                                                            num_kept);
  return launch_status("proposal_tail_kernel");
}
