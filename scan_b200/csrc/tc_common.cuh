// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace scan {

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
// one lane of a converged warp (CUTLASS' elect_one_sync): tcgen05.mma / commit issued under this predicate from
// warp-uniform code compile to a single predicated UTCxMMA instead of the compiler's generic one-thread-at-a-time loop
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred;
}
// warp index the compiler can prove uniform
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// Drain 16 accumulator columns of this thread's TMEM lane into an fp32 row in global memory:
//   dst[0..15] (+)= mul * sum_{p < PARTS} tmem[taddr + 64 p + 0..15]        (plain store when first, else reduction)
// The tensor core adds into its accumulator with TRUNCATION (measured: a 1100-step chain is off by a coherent 5e-5), so long
// reductions are cut into <= 64-step pieces that are combined here with round-to-nearest adds.  The add is a fire-and-
// forget vector reduction performed by the L2 (red.global.add.v4.f32): no load latency on the SM, and deterministic because
// each address belongs to exactly one thread, which issues its pieces in order.  Warp-collective.
template <int PARTS>
__device__ __forceinline__ void tmem_drain16(uint32_t taddr, float* dst, float mul, bool first, bool valid) {
  uint32_t r[PARTS][16];
#pragma unroll
  for (int p = 0; p < PARTS; ++p)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[p][0]), "=r"(r[p][1]), "=r"(r[p][2]), "=r"(r[p][3]), "=r"(r[p][4]), "=r"(r[p][5]), "=r"(r[p][6]), "=r"(r[p][7]),
          "=r"(r[p][8]), "=r"(r[p][9]), "=r"(r[p][10]), "=r"(r[p][11]), "=r"(r[p][12]), "=r"(r[p][13]), "=r"(r[p][14]), "=r"(r[p][15])
        : "r"(taddr + 64 * p)
        : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (valid) {
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float sum = __uint_as_float(r[0][4 * e4 + i]);
#pragma unroll
        for (int p = 1; p < PARTS; ++p) sum += __uint_as_float(r[p][4 * e4 + i]);
        v[i] = sum * mul;
      }
      if (first)
        reinterpret_cast<float4*>(dst)[e4] = make_float4(v[0], v[1], v[2], v[3]);
      else
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * e4), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3])
                     : "memory");
    }
  }
}

// K-major operand tile with 128-byte rows and the 128B swizzle: 8-row groups are 1024 B apart (SBO),
// LBO unused for swizzled K-major layouts, descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major:
// D=f32 (bits 4-5 = 1), A=B=tf32 (2 at bits 7-9 / 10-12), N>>3 at bit 17, M>>4 at bit 24
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// driver entry point for tensor-map encoding (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int get_tensormap_encoder(EncodeTiledFn* fn);
// 2-D fp32 row-major [n_rows, n_cols] matrix, box = [box_rows x 32 columns] (128 bytes), 128B swizzle, zero OOB fill
int make_rowmajor_map(CUtensorMap* m, const float* base, uint64_t n_rows, uint64_t n_cols, uint32_t box_rows);

}  // namespace scan
