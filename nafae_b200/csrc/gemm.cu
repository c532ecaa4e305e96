// Bridge GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// First slice of SURVEY.md section 8(f) rank 2: the frozen, inference-only fully connected layers
// between RoIAlign and the scoring head -- reference `RCNN_top` = VGG16 fc6 / fc7
// (lib/model/faster_rcnn/vgg16_rpn.py:35,56-61: Linear + ReLU (+ Dropout, a no-op in eval mode)):
//     C[M, N] = act(A[M, K] . B[N, K]^T + bias[N])       A = pooled RoI features (R x 25088, bf16),
//                                                        B = the layer's weight as PyTorch stores it
// Both operands are K-major, which is exactly the (R, C*7*7) row-major layout RoIAlign writes and
// the (out_features, in_features) layout of nn.Linear.weight: no transposes anywhere.
//
// One CTA per 128 x bn output tile (bn chosen per problem so that the tiles fill the SMs), 192 threads,
// warp-specialised:
//   warp 0   TMA producer: cp.async.bulk.tensor (128-byte swizzle) into a kStages-deep ring
//   warp 1   TMEM allocation + MMA issue: one elected thread issues tcgen05.mma (M 128, N BN, K 16),
//            tcgen05.commit frees each ring slot and finally signals the epilogue
//   warps 2-5 epilogue: tcgen05.ld the fp32 accumulator (warp w owns TMEM lanes 32*(w%4)..), add the
//            bias, ReLU, convert, store
// bf16 inputs, fp32 accumulation: a looser bound than the fp32 path (tests state it).
#include "tc05.cuh"

namespace nafae {
namespace tc05 {

namespace {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}
}  // namespace

bool make_tensor_map_2d(CUtensorMap* out, const void* base, int elem_bytes, bool is_bf16, long long rows,
                        long long cols, int box_rows) {
  EncodeTiledFn fn = encode_tiled();
  if (fn == nullptr) {
    set_error("tensor map: cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return false;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * (cuuint64_t)elem_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)(128 / elem_bytes), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("tensor map: cuTensorMapEncodeTiled(rows %lld, cols %lld, box %d x %d) failed (CUresult %d)", rows,
              cols, box_rows, 128 / elem_bytes, (int)r);
    return false;
  }
  return true;
}

}  // namespace tc05

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;  // bf16 elements per ring slot = one 128-byte swizzle row
constexpr int kGemmThreads = 192;
constexpr int kABytes = kBM * 128;
constexpr int kMaxStages = 8;

struct GemmParams {
  const float* bias;
  void* C;
  int M, N, K;
  int relu, out_bf16;
  int bn;      // tile width (multiple of 16, <= 256): chosen on the host so that the tiles fill the SMs
  int stages;  // ring depth that fits shared memory at this width
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const GemmParams p) {
  const int S = p.stages, BN = p.bn;
  const int b_bytes = BN * 128;
  extern __shared__ unsigned char gemm_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(
      (reinterpret_cast<uintptr_t>(gemm_smem_raw) + 1023) & ~(uintptr_t)1023);  // swizzle atoms: 1024-byte aligned
  unsigned char* sa = smem;
  unsigned char* sb = smem + (size_t)S * kABytes;   // (b_bytes is a multiple of 2048: stays 1024-aligned)
  uint64_t* full = reinterpret_cast<uint64_t*>(sb + (size_t)S * b_bytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* acc_ready = empty + kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int nkb = (p.K + kBK - 1) / kBK;
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < BN) tmem_cols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_ready, 1);
    fence_mbar_init();
    tc05::tma_prefetch_desc(&map_a);
    tc05::tma_prefetch_desc(&map_b);
  }
  if (warp == 1) tc05::tmem_alloc(tmem_slot, tmem_cols);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        mbar_wait(&empty[s], ((uint32_t)(kb / S) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[s], (uint32_t)(kABytes + b_bytes));
        tc05::tma_load_2d(sa + (size_t)s * kABytes, &map_a, &full[s], kb * kBK, m0);
        tc05::tma_load_2d(sb + (size_t)s * b_bytes, &map_b, &full[s], kb * kBK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = tc05::instr_desc(tc05::kFmtBF16, kBM, BN);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % S;
        mbar_wait(&full[s], (uint32_t)(kb / S) & 1u);
        tc05::fence_after_sync();
        const uint64_t da = tc05::smem_desc_sw128(smem_u32(sa + (size_t)s * kABytes));
        const uint64_t db = tc05::smem_desc_sw128(smem_u32(sb + (size_t)s * b_bytes));
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)  // 32 bytes of K per instruction: +2 in the 16-byte address field
          tc05::mma_bf16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
        tc05::commit(&empty[s]);  // slot reusable once these MMAs have read it
      }
      tc05::commit(acc_ready);    // accumulator complete
    }
  } else {
    // ---- epilogue: warps 2..5 -> TMEM lane quarters 2, 3, 0, 1
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(acc_ready, 0);
    tc05::fence_after_sync();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tc05::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (row < p.M) {
        const int col = n0 + c0;
        const int ncol = min(min(32, BN - c0), p.N - col);  // columns of this chunk that belong to the tile
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]);
          if (p.bias != nullptr && j < ncol) x += __ldg(p.bias + col + j);
          f[j] = p.relu ? fmaxf(x, 0.f) : x;
        }
        if (p.out_bf16) {
          __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.C) + (size_t)row * p.N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (j + 8 <= ncol && (p.N & 7) == 0) {
              uint4 pk;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j], f[j + 1]), h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
              pk.x = *reinterpret_cast<uint32_t*>(&h0);
              pk.y = *reinterpret_cast<uint32_t*>(&h1);
              pk.z = *reinterpret_cast<uint32_t*>(&h2);
              pk.w = *reinterpret_cast<uint32_t*>(&h3);
              *reinterpret_cast<uint4*>(dst + j) = pk;
            } else {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                if (j + t < ncol) dst[j + t] = __float2bfloat16_rn(f[j + t]);
            }
          }
        } else {
          float* dst = static_cast<float*>(p.C) + (size_t)row * p.N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j + 4 <= ncol && (p.N & 3) == 0) {
              *reinterpret_cast<float4*>(dst + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (j + t < ncol) dst[j + t] = f[j + t];
            }
          }
        }
      }
    }
  }
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc05::tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace
}  // namespace nafae

using namespace nafae;

NAFAE_API int nafae_gemm_bf16_tn(const void* A, const void* B, const float* bias, void* C, int M, int N, int K,
                                 unsigned flags, cudaStream_t stream) {
  NAFAE_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: sizes must be positive");
  NAFAE_REQUIRE(A && B && C, "gemm: NULL buffer");
  NAFAE_REQUIRE(K % 8 == 0, "gemm: K must be a multiple of 8 (16-byte row pitch), got %d", K);
  NAFAE_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
                "gemm: buffers must be 16-byte aligned");
  GemmParams p;
  p.bias = bias;
  p.C = C;
  p.M = M;
  p.N = N;
  p.K = K;
  p.relu = (flags & NAFAE_GEMM_RELU) ? 1 : 0;
  p.out_bf16 = (flags & NAFAE_GEMM_OUT_BF16) ? 1 : 0;
  // Tile width: one CTA per 128 x bn tile, so the kernel takes ceil(tiles / SMs) rounds of a time
  // proportional to bn -- pick the width with the cheapest rounds * bn (fc6 at R = 800: 208 columns =
  // 140 tiles on 148 SMs instead of 112 tiles of 256).
  const int m_tiles = (M + kBM - 1) / kBM;
  const int sms = sm_count();
  int bn = 256;
  long long best = -1;
  for (int cand = 256; cand >= 32; cand -= 16) {
    if (cand > 32 && cand >= 2 * ((N + 15) / 16 * 16)) continue;  // far wider than the problem
    const long long tiles = (long long)m_tiles * ((N + cand - 1) / cand);
    const long long cost = ((tiles + sms - 1) / sms) * (long long)(cand + 16);  // +16: per-tile fixed cost
    if (best < 0 || cost < best) {
      best = cost;
      bn = cand;
    }
  }
  p.bn = bn;
  const size_t stage = (size_t)kABytes + (size_t)bn * 128;
  int stages = (int)(((size_t)220 * 1024) / stage);
  if (stages > kMaxStages) stages = kMaxStages;
  NAFAE_REQUIRE(stages >= 2, "gemm: tile does not fit shared memory");
  p.stages = stages;
  CUtensorMap ma, mb;
  if (!tc05::make_tensor_map_2d(&ma, A, 2, true, M, K, kBM)) return 0;
  if (!tc05::make_tensor_map_2d(&mb, B, 2, true, N, K, bn)) return 0;
  const size_t smem = 1024 + (size_t)stages * stage + 256;
  cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("gemm: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    return -(int)e;
  }
  dim3 grid((N + bn - 1) / bn, m_tiles);
  gemm_bf16_tn_kernel<<<grid, kGemmThreads, smem, stream>>>(ma, mb, p);
  return launch_status("gemm_bf16_tn_kernel");
}
