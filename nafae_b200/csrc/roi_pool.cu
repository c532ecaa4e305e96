// Max RoIPool forward / backward for sm_100a.
//
// Replaces (reference tree paths):
//   lib/model/roi_pooling/src/roi_pooling_kernel.cu:24-93    ROIPoolForward
//   lib/model/roi_pooling/src/roi_pooling_kernel.cu:128-203  ROIPoolBackward
//   lib/model/roi_pooling/src/roi_pooling_kernel.cu:95-125,205-234  launchers (same symbols)
//
// Integer / compare work only, bit-exact with the reference: RoI corners round(coord*scale)
// (half away from zero), bins [floor(ph*bh), ceil((ph+1)*bh)) shifted and clipped to the map, empty
// bin -> 0 with argmax -1, first maximum wins (strict '>' scanning h then w), argmax is the flat
// index into the whole (B,C,H,W) batch.
//
// The reference backward lets every input cell scan ALL RoIs of the batch (O(cells * R)).  Here a
// CTA serves one frame: it first compacts that frame's RoIs (ascending index, geometry
// precomputed) into shared memory, then every cell walks only those -- the same gather order
// (RoI, ph, pw ascending), hence the same fp32 sum, without atomics.
#include <float.h>

#include "common.cuh"

namespace nafae {
namespace {

struct PoolRoi {
  int n;  // RoI index
  int start_w, start_h, end_w, end_h;
  float bin_h, bin_w;
};

// roi_pooling_kernel.cu:44-55
__device__ __forceinline__ PoolRoi pool_roi(const float* __restrict__ roi, int n, float scale,
                                            int ph_n, int pw_n) {
  PoolRoi r;
  r.n = n;
  r.start_w = (int)roundf(__fmul_rn(roi[1], scale));
  r.start_h = (int)roundf(__fmul_rn(roi[2], scale));
  r.end_w = (int)roundf(__fmul_rn(roi[3], scale));
  r.end_h = (int)roundf(__fmul_rn(roi[4], scale));
  const int roi_w = (int)fmaxf((float)(r.end_w - r.start_w + 1), 1.f);
  const int roi_h = (int)fmaxf((float)(r.end_h - r.start_h + 1), 1.f);
  r.bin_h = __fdiv_rn((float)roi_h, (float)ph_n);
  r.bin_w = __fdiv_rn((float)roi_w, (float)pw_n);
  return r;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) {
  return (int)fminf(fmaxf((float)v, (float)lo), (float)hi);
}

__global__ void __launch_bounds__(256)
roi_pool_fwd_kernel(const float* __restrict__ bottom, float scale, long long total, int H, int W,
                    int C, int ph_n, int pw_n, const float* __restrict__ rois,
                    float* __restrict__ top, int* __restrict__ argmax) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pw = (int)(idx % pw_n);
    const int ph = (int)((idx / pw_n) % ph_n);
    const int c = (int)((idx / pw_n / ph_n) % C);
    const int n = (int)(idx / pw_n / ph_n / C);
    const float* roi = rois + (size_t)n * 5;
    const PoolRoi r = pool_roi(roi, n, scale, ph_n, pw_n);
    const int batch = (int)roi[0];
    int hstart = (int)floorf(__fmul_rn((float)ph, r.bin_h));
    int wstart = (int)floorf(__fmul_rn((float)pw, r.bin_w));
    int hend = (int)ceilf(__fmul_rn((float)(ph + 1), r.bin_h));
    int wend = (int)ceilf(__fmul_rn((float)(pw + 1), r.bin_w));
    hstart = clampi(hstart + r.start_h, 0, H);
    hend = clampi(hend + r.start_h, 0, H);
    wstart = clampi(wstart + r.start_w, 0, W);
    wend = clampi(wend + r.start_w, 0, W);
    const bool is_empty = (hend <= hstart) || (wend <= wstart);
    float maxval = is_empty ? 0.f : -FLT_MAX;
    int maxidx = -1;
    const int off = (batch * C + c) * H * W;  // int like the reference (:73-74)
    for (int h = hstart; h < hend; ++h)
      for (int w = wstart; w < wend; ++w) {
        const float v = __ldg(bottom + off + h * W + w);
        if (v > maxval) {
          maxval = v;
          maxidx = off + h * W + w;
        }
      }
    top[idx] = maxval;
    if (argmax) argmax[idx] = maxidx;
  }
}

constexpr int kBwdThreads = 256;
constexpr int kBwdList = 256;  // RoIs of the frame resident per chunk

// grid (ceil(C*H*W / 256), B)
__global__ void __launch_bounds__(kBwdThreads)
roi_pool_bwd_kernel(const float* __restrict__ top_diff, const int* __restrict__ argmax, int R,
                    float scale, int H, int W, int C, int ph_n, int pw_n,
                    float* __restrict__ bottom_diff, const float* __restrict__ rois) {
  __shared__ PoolRoi list[kBwdList];
  __shared__ int s_cnt, s_next, s_wcnt[kBwdThreads / 32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y;
  const int cell = blockIdx.x * kBwdThreads + tid;  // (c, h, w) inside frame b
  const bool active = cell < C * H * W;
  const int w = cell % W, h = (cell / W) % H, c = cell / (W * H);
  const int index = (b * C + c) * H * W + h * W + w;  // flat index the argmax stores
  float gradient = 0.f;

  int r_start = 0;
  while (r_start < R) {
    __syncthreads();
    if (tid == 0) {
      s_cnt = 0;
      s_next = R;
    }
    __syncthreads();
    for (int base = r_start; base < R; base += kBwdThreads) {
      const int r = base + tid;
      const bool hit = r < R && (int)rois[(size_t)r * 5] == b;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) s_wcnt[warp] = __popc(bal);
      __syncthreads();
      const int have = s_cnt;
      int before = have, tot = 0;
      for (int k = 0; k < kBwdThreads / 32; ++k) {
        const int cw = s_wcnt[k];
        if (k < warp) before += cw;
        tot += cw;
      }
      const int pos = before + __popc(bal & ((1u << lane) - 1u));
      if (hit && pos < kBwdList) list[pos] = pool_roi(rois + (size_t)r * 5, r, scale, ph_n, pw_n);
      if (hit && pos == kBwdList) s_next = r;
      __syncthreads();
      if (have + tot >= kBwdList) {
        if (tid == 0) {
          s_cnt = kBwdList;
          if (have + tot == kBwdList) s_next = min(base + kBwdThreads, R);
        }
        break;
      }
      if (tid == 0) s_cnt = have + tot;
    }
    __syncthreads();
    const int cnt = s_cnt;
    r_start = s_next;
    if (active) {
      for (int k = 0; k < cnt; ++k) {
        const PoolRoi& r = list[k];
        // roi_pooling_kernel.cu:162-166
        if (!(w >= r.start_w && w <= r.end_w && h >= r.start_h && h <= r.end_h)) continue;
        const size_t offset = (size_t)r.n * ph_n * pw_n * C;
        int phstart = (int)floorf(__fdiv_rn((float)(h - r.start_h), r.bin_h));
        int phend = (int)ceilf(__fdiv_rn((float)(h - r.start_h + 1), r.bin_h));
        int pwstart = (int)floorf(__fdiv_rn((float)(w - r.start_w), r.bin_w));
        int pwend = (int)ceilf(__fdiv_rn((float)(w - r.start_w + 1), r.bin_w));
        phstart = clampi(phstart, 0, ph_n);
        phend = clampi(phend, 0, ph_n);
        pwstart = clampi(pwstart, 0, pw_n);
        pwend = clampi(pwend, 0, pw_n);
        for (int ph = phstart; ph < phend; ++ph)
          for (int pw = pwstart; pw < pwend; ++pw) {
            const size_t o = offset + ((size_t)c * ph_n + ph) * pw_n + pw;
            if (__ldg(argmax + o) == index) gradient = __fadd_rn(gradient, __ldg(top_diff + o));
          }
      }
    }
  }
  if (active) bottom_diff[index] = gradient;
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_API int ROIPoolForwardLaucher(const float* bottom_data, const float spatial_scale,
                                    const int num_rois, const int height, const int width,
                                    const int channels, const int pooled_height,
                                    const int pooled_width, const float* bottom_rois,
                                    float* top_data, int* argmax_data, cudaStream_t stream) {
  NAFAE_REQUIRE(num_rois >= 0 && channels >= 0 && pooled_height >= 1 && pooled_width >= 1,
                "roi_pool: bad sizes");
  const long long total = (long long)num_rois * channels * pooled_height * pooled_width;
  if (total == 0) return 1;
  NAFAE_REQUIRE(bottom_data && bottom_rois && top_data, "roi_pool: NULL buffer");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  roi_pool_fwd_kernel<<<(int)blocks, 256, 0, stream>>>(bottom_data, spatial_scale, total, height,
                                                      width, channels, pooled_height,
                                                      pooled_width, bottom_rois, top_data,
                                                      argmax_data);
  return launch_status("roi_pool_fwd_kernel");
}

NAFAE_API int ROIPoolBackwardLaucher(const float* top_diff, const float spatial_scale,
                                     const int batch_size, const int num_rois, const int height,
                                     const int width, const int channels, const int pooled_height,
                                     const int pooled_width, const float* bottom_rois,
                                     float* bottom_diff, const int* argmax_data,
                                     cudaStream_t stream) {
  NAFAE_REQUIRE(batch_size >= 0 && num_rois >= 0 && channels >= 0 && pooled_height >= 1 &&
                    pooled_width >= 1,
                "roi_pool: bad sizes");
  const long long per_frame = (long long)channels * height * width;
  if (per_frame == 0 || batch_size == 0) return 1;
  NAFAE_REQUIRE(per_frame * batch_size < (1ll << 31), "roi_pool: batch too large for int argmax");
  NAFAE_REQUIRE(batch_size <= 65535, "roi_pool: more than 65535 frames per call");
  NAFAE_REQUIRE(bottom_diff && (num_rois == 0 || (top_diff && bottom_rois && argmax_data)),
                "roi_pool: NULL buffer");
  dim3 grid((unsigned)((per_frame + kBwdThreads - 1) / kBwdThreads), batch_size);
  roi_pool_bwd_kernel<<<grid, kBwdThreads, 0, stream>>>(top_diff, argmax_data, num_rois,
                                                        spatial_scale, height, width, channels,
                                                        pooled_height, pooled_width, bottom_diff,
                                                        bottom_rois);
  return launch_status("roi_pool_bwd_kernel");
}
