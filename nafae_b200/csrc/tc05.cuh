// tcgen05 / TMEM / TMA-tensor-map building blocks shared by the tensor-core kernels (gemm.cu, the
// tensor-core form of the similarity contraction in ground.cu).  sm_100a only.
//
// Conventions (checked against the bit layouts in CUTLASS' cute/arch/mma_sm100_desc.hpp):
//   * operands are K-major tiles in shared memory whose rows are 128 bytes (64 bf16 / 32 tf32
//     elements) in the 128-byte swizzle TMA writes (CU_TENSOR_MAP_SWIZZLE_128B), 8-row groups 1024
//     bytes apart; one tcgen05.mma consumes 32 bytes of K, so a 128-byte row holds four K steps
//     and the descriptor's start address advances by 32 bytes per step;
//   * accumulators live in TMEM: M = 128 uses all 128 lanes, one 32-bit column per fp32 element;
//     warp w may read lanes 32*(w % 4) .. +31 (tcgen05.ld 32x32b).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace nafae {
namespace tc05 {

// ---- shared-memory matrix descriptor: K-major, SWIZZLE_128B --------------------------------
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);  // start address, 16-byte units        bits [0,14)
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused here)   bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B   bits [32,46)
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)       bits [46,48)
  d |= (uint64_t)2 << 61;                       // layout type SWIZZLE_128B             bits [61,64)
  return d;
}

// ---- instruction descriptor: dense, fp32 accumulate, both operands K-major -----------------
constexpr uint32_t kFmtBF16 = 1, kFmtTF32 = 2;
__host__ __device__ constexpr uint32_t instr_desc(uint32_t fmt, int M, int N) {
  return (1u << 4)                    // c_format = F32       bits [4,6)
         | (fmt << 7)                 // a_format             bits [7,10)
         | (fmt << 10)                // b_format             bits [10,13)
         | ((uint32_t)(N >> 3) << 17) // n_dim                bits [17,23)
         | ((uint32_t)(M >> 4) << 24);// m_dim                bits [24,29)
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the whole CTA
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (count 1) on an mbarrier once every previously issued tcgen05.mma of this thread is done
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------
// whole warp; ncols a power of two >= 32; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
// 32 lanes (this warp's quarter) x 32 consecutive columns -> 32 registers per thread (thread = lane/row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- TMA (tensor-map) loads ------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner,
                                            int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Host: row-major (rows x cols) matrix of 2- or 4-byte elements -> tensor map whose box is
// (box_rows x 128 bytes) with the 128-byte swizzle.  Returns false (and sets the library error)
// when the driver entry point is unavailable or rejects the arguments.
bool make_tensor_map_2d(CUtensorMap* out, const void* base, int elem_bytes, bool is_bf16, long long rows,
                        long long cols, int box_rows);

}  // namespace tc05
}  // namespace nafae
