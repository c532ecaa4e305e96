// Step wrapper tail of the training path (reference model.py:773-774):
//     torch.nn.utils.clip_grad_norm_(ground_model.parameters(), args.clip)     # clip = 100
//     optimizer.step()                      # Adam(lr 1e-3, weight_decay 1e-5), model.py:1030-1036
// over the FLAT fp32 bucket the data-parallel all-reduce averages (vis_ebd.fc1, word_ebd.fc1,
// word_ebd.bn: 2.20 M floats), as two launches with no host synchronisation (graph-capturable):
//   grad_sumsq_kernel   deterministic sum of squares (per-CTA partials, the last CTA adds them in
//                       index order) -> total norm, clip coefficient, Adam step count and bias
//                       corrections in the workspace
//   clip_adam_kernel    g <- g * clip_coef (written back, like clip_grad_norm_ does in place),
//                       g += wd * p, m/v update, p <- p - step_size * m / (sqrt(v)/sqrt(bc2) + eps)
// In data-parallel training it runs AFTER the gradient all-reduce, so every replica clips the same
// averaged gradient and stays bit-identical.
#include <math.h>

#include "common.cuh"

namespace nafae {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptMaxCtas = 1024;

// workspace (bytes): [0] int ticket, [4] int step, [8] float total_norm, [12] float clip_coef,
// [16] float step_size, [20] float inv_sqrt_bc2, [64 ...] double partial[kOptMaxCtas]
struct OptWs {
  int* ticket;
  int* step;
  float* scal;  // [0] total_norm [1] clip_coef [2] step_size [3] 1/sqrt(bias_correction2)
  double* partial;
};
__host__ __device__ inline OptWs opt_carve(void* ws) {
  OptWs w;
  char* b = static_cast<char*>(ws);
  w.ticket = reinterpret_cast<int*>(b);
  w.step = reinterpret_cast<int*>(b + 4);
  w.scal = reinterpret_cast<float*>(b + 8);
  w.partial = reinterpret_cast<double*>(b + 64);
  return w;
}
constexpr size_t kOptWsBytes = 64 + sizeof(double) * kOptMaxCtas;

__global__ void __launch_bounds__(kOptThreads)
grad_sumsq_kernel(const float* __restrict__ grad, size_t n, float max_norm, float lr, float beta1,
                  float beta2, void* ws) {
  const OptWs w = opt_carve(ws);
  __shared__ double s_part[kOptThreads / 32];
  __shared__ int s_ticket;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // contiguous chunk per CTA, fixed traversal order => the same bits on every replica / every run
  const size_t n4 = n >> 2;
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t b = min(n4, (size_t)blockIdx.x * per), e = min(n4, b + per);
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  for (size_t i = b + tid; i < e; i += kOptThreads) {
    const float4 v = __ldg(g4 + i);
    acc0 = fmaf(v.x, v.x, acc0);
    acc1 = fmaf(v.y, v.y, acc1);
    acc2 = fmaf(v.z, v.z, acc2);
    acc3 = fmaf(v.w, v.w, acc3);
  }
  double acc = ((double)acc0 + (double)acc1) + ((double)acc2 + (double)acc3);
  if (blockIdx.x == gridDim.x - 1)  // tail elements (n % 4)
    for (size_t i = (n4 << 2) + tid; i < n; i += kOptThreads) acc += (double)grad[i] * (double)grad[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_part[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.;
    for (int k = 0; k < kOptThreads / 32; ++k) t += s_part[k];
    w.partial[blockIdx.x] = t;
  }
  __syncthreads();
  if (tid == 0) s_ticket = ticket_acq_rel(w.ticket);
  __syncthreads();
  if (s_ticket != (int)gridDim.x - 1) return;
  if (tid == 0) {
    double t = 0.;
    for (unsigned k = 0; k < gridDim.x; ++k) t += __ldcg(w.partial + k);
    const float total = (float)sqrt(t);
    // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    float coef = max_norm / (total + 1e-6f);
    if (!(max_norm > 0.f)) coef = 1.f;  // max_norm <= 0: clipping disabled
    coef = fminf(coef, 1.f);
    const int step = *w.step + 1;
    *w.step = step;
    const double bc1 = 1. - pow((double)beta1, (double)step);
    const double bc2 = 1. - pow((double)beta2, (double)step);
    w.scal[0] = total;
    w.scal[1] = coef;
    w.scal[2] = (float)((double)lr / bc1);
    w.scal[3] = (float)(1. / sqrt(bc2));
    *w.ticket = 0;
  }
}

__global__ void __launch_bounds__(kOptThreads)
clip_adam_kernel(float* __restrict__ param, float* __restrict__ grad, float* __restrict__ exp_avg,
                 float* __restrict__ exp_avg_sq, size_t n, float beta1, float beta2, float eps,
                 float weight_decay, const void* ws) {
  const OptWs w = opt_carve(const_cast<void*>(ws));
  const float coef = w.scal[1], step_size = w.scal[2], inv_bc2s = w.scal[3];
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  auto upd = [&](float& p, float& g, float& m, float& v) {
    g = g * coef;                          // clip_grad_norm_ scales .grad in place
    const float gd = fmaf(weight_decay, p, g);  // Adam's L2 term: grad + wd * param
    m = fmaf(gd - m, omb1, m);             // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(omb2 * gd, gd, v * beta2);    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * inv_bc2s + eps;
    p = p - step_size * (m / denom);
  };
  const size_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(param);
  float4* g4 = reinterpret_cast<float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(exp_avg);
  float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
  for (size_t i = (size_t)blockIdx.x * kOptThreads + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * kOptThreads) {
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    upd(p.x, g.x, m.x, v.x);
    upd(p.y, g.y, m.y, v.y);
    upd(p.z, g.z, m.z, v.z);
    upd(p.w, g.w, m.w, v.w);
    p4[i] = p;
    g4[i] = g;
    m4[i] = m;
    v4[i] = v;
  }
  if (blockIdx.x == 0)
    for (size_t i = (n4 << 2) + threadIdx.x; i < n; i += kOptThreads)
      upd(param[i], grad[i], exp_avg[i], exp_avg_sq[i]);
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_API size_t nafae_clip_adam_workspace_bytes(void) { return kOptWsBytes; }

NAFAE_API int nafae_clip_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq,
                                   size_t n, float lr, float beta1, float beta2, float eps,
                                   float weight_decay, float max_norm, void* workspace,
                                   size_t workspace_bytes, cudaStream_t stream) {
  NAFAE_REQUIRE(param && grad && exp_avg && exp_avg_sq, "clip_adam: NULL buffer");
  NAFAE_REQUIRE(workspace && workspace_bytes >= kOptWsBytes, "clip_adam: workspace too small");
  NAFAE_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                  reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0,
                "clip_adam: buffers must be 16-byte aligned");
  NAFAE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "clip_adam: workspace must be 8-byte aligned");
  if (n == 0) return 1;
  size_t want = (n / 4 + (size_t)kOptThreads * 8 - 1) / ((size_t)kOptThreads * 8);
  int grid = (int)(want < 1 ? 1 : (want > (size_t)kOptMaxCtas ? kOptMaxCtas : want));
  const int cap = sm_count() * 4;
  if (grid > cap) grid = cap;
  grad_sumsq_kernel<<<grid, kOptThreads, 0, stream>>>(grad, n, max_norm, lr, beta1, beta2, workspace);
  int st = launch_status("grad_sumsq_kernel");
  if (st != 1) return st;
  clip_adam_kernel<<<grid, kOptThreads, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2,
                                                     eps, weight_decay, workspace);
  return launch_status("clip_adam_kernel");
}
