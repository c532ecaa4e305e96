// Proposal front end for sm_100a: everything _ProposalLayer.forward does BEFORE its per-frame NMS loop
// (reference lib/model/rpn/proposal_layer.py:66-125), as two launches for the whole batch:
//
//   proposal_decode_kernel   anchors (generate_anchors.py base windows + feature-stride shifts,
//                            proposal_layer.py:80-93), bbox_transform_inv (bbox_transform.py:77-103),
//                            clip_boxes (:125-133) and the (H, W, A) re-ordering of the NCHW RPN outputs
//                            (:98-103) -- one thread per feature cell, coalesced NCHW reads
//   proposal_sort_kernel     torch.sort(scores, 1, True) (:125) per frame: one CTA per frame, a STABLE
//                            least-significant-digit radix sort (4 x 8 bits) that lives entirely in
//                            shared memory (keys + two 16-bit index buffers), then the gather of the
//                            sorted proposals / scores that nafae_proposal_tail consumes
//
// The reference leaves the order of equal scores to torch.sort (unstable); here ties keep ascending
// anchor index (what a stable descending sort gives), which is one of the orders the reference allows.
// Arithmetic: the reference's op order with one rounding per op (no FMA contraction), expf as torch's
// CUDA exp -- decoded boxes are bit-identical to the reference run on the same GPU.
#include "common.cuh"

namespace nafae {
namespace {

constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortMaxN = 25600;  // 8 bytes of shared memory per box (key + two 16-bit indices)

__device__ __forceinline__ float clampf(float x, float lo, float hi) {
  if (x != x) return x;  // clamp_ propagates NaN
  return fminf(fmaxf(x, lo), hi);
}

// order-preserving map float -> uint32 (ascending), then complemented: ascending key = descending score
__device__ __forceinline__ uint32_t desc_key(float s) {
  if (s == 0.f) s = 0.f;  // -0 and +0 compare equal in torch.sort
  uint32_t u = __float_as_uint(s);
  u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
  return ~u;
}

struct DecodeParams {
  const float* cls_prob;   // (B, 2A, H, W): channels [A, 2A) are the foreground probabilities
  const float* deltas;     // (B, 4A, H, W)
  const float* im_info;    // (B, 3): height, width, scale
  const float* anchors;    // (A, 4) base windows
  float* proposals;        // (B, H*W*A, 4)
  float* scores;           // (B, H*W*A)
  int B, A, H, W;
  float stride;
};

__global__ void __launch_bounds__(256) proposal_decode_kernel(const DecodeParams p) {
  const int hw = p.H * p.W;
  const long long total = (long long)p.B * hw;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / hw), k = (int)(t % hw);
    const int y = k / p.W, x = k - y * p.W;
    const float sx = (float)x * p.stride, sy = (float)y * p.stride;  // shifts (:80-84), exact integers
    const float im_h = __ldg(p.im_info + b * 3), im_w = __ldg(p.im_info + b * 3 + 1);
    const float max_x = __fsub_rn(im_w, 1.f), max_y = __fsub_rn(im_h, 1.f);
    const float* dl = p.deltas + (size_t)b * 4 * p.A * hw + k;
    const float* sc = p.cls_prob + ((size_t)b * 2 * p.A + p.A) * hw + k;
    float* out = p.proposals + ((size_t)b * hw + k) * p.A * 4;
    float* so = p.scores + ((size_t)b * hw + k) * p.A;
    for (int a = 0; a < p.A; ++a) {
      const float ax1 = __fadd_rn(__ldg(p.anchors + a * 4 + 0), sx), ay1 = __fadd_rn(__ldg(p.anchors + a * 4 + 1), sy);
      const float ax2 = __fadd_rn(__ldg(p.anchors + a * 4 + 2), sx), ay2 = __fadd_rn(__ldg(p.anchors + a * 4 + 3), sy);
      const float dx = __ldg(dl + (size_t)(a * 4 + 0) * hw), dy = __ldg(dl + (size_t)(a * 4 + 1) * hw);
      const float dw = __ldg(dl + (size_t)(a * 4 + 2) * hw), dh = __ldg(dl + (size_t)(a * 4 + 3) * hw);
      // bbox_transform.py:78-94, one rounding per torch op
      const float w = __fadd_rn(__fsub_rn(ax2, ax1), 1.f), h = __fadd_rn(__fsub_rn(ay2, ay1), 1.f);
      const float cx = __fadd_rn(ax1, __fmul_rn(0.5f, w)), cy = __fadd_rn(ay1, __fmul_rn(0.5f, h));
      const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
      const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
      float4 box;
      box.x = clampf(__fsub_rn(pcx, __fmul_rn(0.5f, pw)), 0.f, max_x);
      box.y = clampf(__fsub_rn(pcy, __fmul_rn(0.5f, ph)), 0.f, max_y);
      box.z = clampf(__fadd_rn(pcx, __fmul_rn(0.5f, pw)), 0.f, max_x);
      box.w = clampf(__fadd_rn(pcy, __fmul_rn(0.5f, ph)), 0.f, max_y);
      *reinterpret_cast<float4*>(out + a * 4) = box;
      so[a] = __ldg(sc + (size_t)a * hw);
    }
  }
}

struct SortParams {
  const float* proposals;  // (B, n, 4) in anchor order
  const float* scores;     // (B, n)
  float* out_proposals;    // (B, m, 4) score-descending
  float* out_scores;       // (B, m)
  int* out_order;          // (B, m) anchor index of every sorted position, or NULL
  int n, m;
};

__global__ void __launch_bounds__(kSortThreads, 1) proposal_sort_kernel(const SortParams p) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  uint32_t* keys = reinterpret_cast<uint32_t*>(sort_smem);              // [n]
  uint16_t* idx0 = reinterpret_cast<uint16_t*>(keys + p.n);             // [n]
  uint16_t* idx1 = idx0 + ((p.n + 1) & ~1);                             // [n]
  __shared__ uint32_t s_base[256];
  __shared__ uint16_t s_tile[kSortWarps][256];
  __shared__ int s_single;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x, n = p.n;
  const float* sc = p.scores + (size_t)b * n;
  for (int i = tid; i < n; i += kSortThreads) {
    keys[i] = desc_key(__ldg(sc + i));
    idx0[i] = (uint16_t)i;
  }
  __syncthreads();
  uint16_t* in = idx0;
  uint16_t* out = idx1;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = pass * 8;
    if (tid < 256) s_base[tid] = 0;
    if (tid == 0) s_single = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kSortThreads) atomicAdd(&s_base[(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (tid < 256 && s_base[tid] == (uint32_t)n) s_single = 1;  // every key has the same digit: nothing moves
    __syncthreads();
    if (s_single) continue;
    if (tid < 32) {  // exclusive scan of the 256 digit counts by one warp
      uint32_t v[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = s_base[lane * 8 + j];
        sum += v[j];
      }
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      uint32_t run = incl - sum;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s_base[lane * 8 + j] = run;
        run += v[j];
      }
    }
    __syncthreads();
    for (int tile0 = 0; tile0 < n; tile0 += kSortThreads) {
      // clear the per-warp digit counts of this tile
      uint32_t* tz = reinterpret_cast<uint32_t*>(&s_tile[0][0]);
      for (int i = tid; i < kSortWarps * 256 / 2; i += kSortThreads) tz[i] = 0u;
      __syncthreads();
      const int pos = tile0 + tid;
      const bool valid = pos < n;
      const uint16_t e = valid ? in[pos] : (uint16_t)0;
      const uint32_t d = valid ? ((keys[e] >> shift) & 255u) : (256u + (uint32_t)lane);  // invalid: unique groups
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int rank_w = __popc(peers & ((1u << lane) - 1u));
      if (valid && rank_w == 0) s_tile[warp][d] = (uint16_t)__popc(peers);
      __syncthreads();
      if (tid < 256) {  // stable: earlier warps of the tile first, after everything of earlier tiles
        uint32_t run = s_base[tid];
#pragma unroll 8
        for (int w = 0; w < kSortWarps; ++w) {
          const uint32_t c = s_tile[w][tid];
          s_tile[w][tid] = (uint16_t)run;   // fits: run < n <= 25600
          run += c;
        }
        s_base[tid] = run;
      }
      __syncthreads();
      if (valid) out[(uint32_t)s_tile[warp][d] + (uint32_t)rank_w] = e;
      __syncthreads();
    }
    uint16_t* t = in;
    in = out;
    out = t;
  }
  // gather: sorted position -> anchor index -> box / score
  const float4* src = reinterpret_cast<const float4*>(p.proposals + (size_t)b * n * 4);
  float4* dst = reinterpret_cast<float4*>(p.out_proposals + (size_t)b * p.m * 4);
  for (int i = tid; i < p.m; i += kSortThreads) {
    const int e = in[i];
    dst[i] = __ldg(src + e);
    p.out_scores[(size_t)b * p.m + i] = __ldg(sc + e);
    if (p.out_order) p.out_order[(size_t)b * p.m + i] = e;
  }
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_API size_t nafae_proposal_front_workspace_bytes(int batch_size, int num_anchors, int height, int width) {
  if (batch_size <= 0 || num_anchors <= 0 || height <= 0 || width <= 0) return 0;
  const size_t n = (size_t)height * width * num_anchors;
  return align_up((size_t)batch_size * n * 5 * sizeof(float), 256);  // decoded boxes + scores in anchor order
}

NAFAE_API int nafae_proposal_front(const float* rpn_cls_prob, const float* rpn_bbox_pred, const float* im_info,
                                   const float* anchors, int batch_size, int num_anchors, int height, int width,
                                   float feat_stride, int pre_nms_topn, float* proposals_sorted,
                                   float* scores_sorted, int* order, void* workspace, size_t workspace_bytes,
                                   cudaStream_t stream) {
  NAFAE_REQUIRE(batch_size > 0 && num_anchors > 0 && height > 0 && width > 0, "proposal_front: sizes must be positive");
  NAFAE_REQUIRE(rpn_cls_prob && rpn_bbox_pred && im_info && anchors && proposals_sorted && scores_sorted,
                "proposal_front: NULL buffer");
  NAFAE_REQUIRE(batch_size <= 65535, "proposal_front: more than 65535 frames per call");
  const long long n = (long long)height * width * num_anchors;
  NAFAE_REQUIRE(n <= kSortMaxN, "proposal_front: %lld anchors per frame exceed the in-shared-memory sort (%d)", n,
                kSortMaxN);
  const size_t need = nafae_proposal_front_workspace_bytes(batch_size, num_anchors, height, width);
  NAFAE_REQUIRE(workspace && workspace_bytes >= need, "proposal_front: workspace too small (%zu < %zu)",
                workspace_bytes, need);
  NAFAE_REQUIRE(((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(proposals_sorted)) & 15) == 0,
                "proposal_front: workspace / proposals must be 16-byte aligned");
  // proposal_layer.py:139-140 compares pre_nms_topN with the element count of the WHOLE batch
  int m = (int)n;
  if (pre_nms_topn > 0 && (long long)pre_nms_topn < (long long)batch_size * n && pre_nms_topn < m) m = pre_nms_topn;
  DecodeParams d;
  d.cls_prob = rpn_cls_prob;
  d.deltas = rpn_bbox_pred;
  d.im_info = im_info;
  d.anchors = anchors;
  d.proposals = static_cast<float*>(workspace);
  d.scores = d.proposals + (size_t)batch_size * n * 4;
  d.B = batch_size;
  d.A = num_anchors;
  d.H = height;
  d.W = width;
  d.stride = feat_stride;
  const long long cells = (long long)batch_size * height * width;
  int grid = (int)((cells + 255) / 256);
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  proposal_decode_kernel<<<grid, 256, 0, stream>>>(d);
  int st = launch_status("proposal_decode_kernel");
  if (st != 1) return st;
  SortParams s;
  s.proposals = d.proposals;
  s.scores = d.scores;
  s.out_proposals = proposals_sorted;
  s.out_scores = scores_sorted;
  s.out_order = order;
  s.n = (int)n;
  s.m = m;
  const size_t smem = (size_t)n * 4 + 2 * (((size_t)n + 1) & ~(size_t)1) * 2 + 16;
  cudaError_t e = cudaFuncSetAttribute(proposal_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("proposal_front: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  proposal_sort_kernel<<<batch_size, kSortThreads, smem, stream>>>(s);
  return launch_status("proposal_sort_kernel");
}
