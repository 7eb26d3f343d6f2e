// K3a backward on tcgen05 (see attention_t5.cu for the forward and the numerics).  Two kernels, both recompute the
// probabilities from the saved log-sum-exp and keep their accumulators in tensor memory across the whole loop:
//   attn_bwd_dq_t5_kernel   CTA = 128 queries (TMEM lanes), loops over 64-key tiles:
//        S = Q K^T, dP = dO V^T -> dS = P (dP*keep - D) -> hi/lo to TMEM -> dQ += dS K          (B = K^T planes)
//   attn_bwd_dkv_t5_kernel  CTA = 128 keys (TMEM lanes), loops over 32-query tiles:
//        S^T = K Q^T, dP^T = V dO^T -> P~^T, dS^T -> hi/lo to TMEM -> dV += P~^T dO, dK += dS^T Q (B = dO^T / Q^T planes)
// Every product is 3xTF32 with the "hi x (hi | lo)" pair fused into one wide-N MMA.  No atomics: dQ, dK, dV are each
// written once.  Operand planes come from attn_bwd_prep_kernel.
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace scan {

constexpr int B5_THREADS = 640;           // warps 0-3 control, 4..19 elementwise (lane quarter = w % 4, column block = w / 4)
constexpr int B5_EW = 512;
constexpr int B5_BOX128 = 128 * 32 * 4;   // 16 KB  [128 rows x 32 cols]
constexpr int B5_BOX64 = 64 * 32 * 4;     // 8 KB
constexpr int B5_BOX32 = 32 * 32 * 4;     // 4 KB
constexpr uint32_t B5_ID128 = umma_idesc_tf32(128, 128);
constexpr uint32_t B5_ID64 = umma_idesc_tf32(128, 64);
constexpr uint32_t B5_ID32 = umma_idesc_tf32(128, 32);

__device__ __forceinline__ void b5_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void b5_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void b5_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void b5_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void b5_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void b5_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void b5_commit(uint32_t bar) {
  if (elect_one_sync()) umma_commit(bar);
}
__device__ __forceinline__ void b5_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void b5_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void b5_split(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// ---------------------------------------------------------------------------- pre-pass
// rows:   dst_hl[r] = hi(src[r]) | lo(src[r])                       ([4M, 64] -> [4M, 128])
// planes: dst_t[chunk][d | 64 + d][key] = hi | lo of src[chunk*M + key][d], zero for key >= M   ([4][128][Mp])
struct PrepArgs {
  const float* rows_src[4];
  float* rows_dst[4];
  const float* t_src[3];
  float* t_dst[3];
  int n_rows, n_t;
  const float *lse, *delta;   // [4M]
  float *lse2p, *dlp;         // [4][Mp]: lse * log2(e) padded with +huge (probability 0), delta padded with 0
};

__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(PrepArgs a, int m, int mp) {
  __shared__ float th[64][33], tl[64][33];
  const long long n_rows = 4ll * m;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * 16; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i >> 4;
    const int c4 = (int)(i & 15);
    for (int which = 0; which < a.n_rows; ++which) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(a.rows_src[which] + row * 64) + c4);
      float4 h, l;
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
      float* dst = a.rows_dst[which] + row * 128;
      reinterpret_cast<float4*>(dst)[c4] = h;
      reinterpret_cast<float4*>(dst + 64)[c4] = l;
    }
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 4ll * mp; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = (int)(i / mp), r = (int)(i % mp);
    a.lse2p[i] = r < m ? __ldg(a.lse + (long long)chunk * m + r) * 1.4426950408889634f : 1e30f;
    a.dlp[i] = r < m ? __ldg(a.delta + (long long)chunk * m + r) : 0.f;
  }
  const int tiles_per_chunk = mp / 32;
  for (int tile = blockIdx.x; tile < a.n_t * 4 * tiles_per_chunk; tile += gridDim.x) {
    const int which = tile / (4 * tiles_per_chunk);
    const int rem = tile % (4 * tiles_per_chunk);
    const int chunk = rem / tiles_per_chunk, j0 = (rem % tiles_per_chunk) * 32;
    const float* src = a.t_src[which];
    float* dst = a.t_dst[which];
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
      const int key = i >> 6, d = i & 63;
      float x = 0.f;
      if (j0 + key < m) x = __ldg(src + ((long long)chunk * m + j0 + key) * 64 + d);
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
      th[d][key] = __uint_as_float(u);
      tl[d][key] = x - __uint_as_float(u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
      const int d = i >> 5, key = i & 31;
      dst[((long long)chunk * 128 + d) * mp + j0 + key] = th[d][key];
      dst[((long long)chunk * 128 + 64 + d) * mp + j0 + key] = tl[d][key];
    }
  }
}

// ---------------------------------------------------------------------------- dQ
// smem: Q hi/lo (4 x 16 KB) | dO hi/lo (4 x 16 KB) | K stage 32 KB | V stage 32 KB | K^T stage 32 KB
// TMEM: S [0,192) (wide: hi.hi | hi.lo | lo.hi)   dP [192,256) (three MMAs into one accumulator)
//       dS operand slot hi [256,320) lo [320,384)   dQ accumulator [384,512)
// Pipeline: the elementwise warps pull S/dP of tile t into registers and release the TMEM region at once, so the tensor
// pipe computes S/dP of tile t+1 while they work; dQ MMAs of tile t follow as soon as dS(t) is in the operand slot.
constexpr int DQ_SMEM = 1024 + 8 * B5_BOX128 + 3 * 4 * B5_BOX64 + 1024;
constexpr float B5_LOG2E = 1.4426950408889634f;
// The tensor core adds into its fp32 accumulator with truncation: the gradient accumulators are drained into the fp32
// output rows every 64 accumulation steps (see T5_FLUSH in attention_t5.cu).
constexpr int DQ_FLUSH = 8;     // key tiles of 64 (8 steps each)
constexpr int DKV_FLUSH = 16;   // query tiles of 32 (4 steps each)

__device__ __forceinline__ float b5_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool DROP>
__global__ void __launch_bounds__(B5_THREADS, 1)
    attn_bwd_dq_t5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                          const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                          const __grid_constant__ CUtensorMap map_kt, const float* __restrict__ lse, const float* __restrict__ delta,
                          int m, float scale, float drop_p, uint64_t seed, float* __restrict__ dq) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;
  uint8_t* do_s = q_s + 4 * B5_BOX128;
  uint8_t* k_s = do_s + 4 * B5_BOX128;
  uint8_t* v_s = k_s + 4 * B5_BOX64;
  uint8_t* kt_s = v_s + 4 * B5_BOX64;
  uint64_t* bars = (uint64_t*)(kt_s + 4 * B5_BOX64);
  uint64_t* r_full = bars;        // resident Q, dO
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* kt_full = bars + 5;
  uint64_t* kt_empty = bars + 6;
  uint64_t* sp_full = bars + 7;   // S and dP of a tile are in TMEM
  uint64_t* sp_free = bars + 8;   // ... and have been pulled into registers (512 arrivals)
  uint64_t* ds_full = bars + 9;   // dS operand written (512 arrivals)
  uint64_t* dq_done = bars + 10;  // dQ MMAs of the tile retired: the dS slot may be overwritten
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int i0 = blockIdx.x * 128;
  // gridDim.z > 1: the key tiles are split over z (fills the machine when the (query tile, chunk) grid leaves SMs idle in
  // the last wave); every split then ADDS its partial dQ into the zero-initialised output with the same vector reductions
  // the drains use
  const int n_tiles_all = (m + 63) / 64;
  const int t0 = (int)((long long)n_tiles_all * blockIdx.z / gridDim.z);
  const int n_tiles = (int)((long long)n_tiles_all * (blockIdx.z + 1) / gridDim.z) - t0;
  const bool accum = gridDim.z > 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) mbar_init(smem_u32(bars + i), (i == 8 || i == 9) ? B5_EW : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(r_full), 8 * B5_BOX128);
      for (int part = 0; part < 2; ++part)
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(smem_u32(q_s + (part * 2 + kb) * B5_BOX128), &map_q, smem_u32(r_full), part * 64 + kb * 32, (int)(base + i0));
          tma_load_2d(smem_u32(do_s + (part * 2 + kb) * B5_BOX128), &map_do, smem_u32(r_full), part * 64 + kb * 32, (int)(base + i0));
        }
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = (t0 + t) * 64;
        const uint32_t ph = (uint32_t)(t & 1);
        mbar_wait(smem_u32(k_empty), ph ^ 1);
        mbar_expect_tx(smem_u32(k_full), 4 * B5_BOX64);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part)
            tma_load_2d(smem_u32(k_s + (kb * 2 + part) * B5_BOX64), &map_k, smem_u32(k_full), part * 64 + kb * 32, (int)(base + j0));
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = (t0 + t) * 64;
        mbar_wait(smem_u32(v_empty), (uint32_t)(t & 1) ^ 1);
        mbar_expect_tx(smem_u32(v_full), 4 * B5_BOX64);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part)
            tma_load_2d(smem_u32(v_s + (kb * 2 + part) * B5_BOX64), &map_v, smem_u32(v_full), part * 64 + kb * 32, (int)(base + j0));
      }
    }
  } else if (warp == 3) {
    // K^T planes have their own producer: their slot frees only when the dQ MMAs retire, which must not hold back the
    // K/V prefetch for the next S/dP
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = (t0 + t) * 64;
        mbar_wait(smem_u32(kt_empty), (uint32_t)(t & 1) ^ 1);
        mbar_expect_tx(smem_u32(kt_full), 4 * B5_BOX64);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part)
            tma_load_2d(smem_u32(kt_s + (kb * 2 + part) * B5_BOX64), &map_kt, smem_u32(kt_full), j0 + kb * 32, chunk * 128 + part * 64);
      }
    }
  } else if (warp == 1) {
    {   // all 32 lanes run the issue loop; each tcgen05 instruction is issued by one elected lane
      mbar_wait(smem_u32(r_full), 0);
      tcgen05_fence_after();
      const uint32_t qa = smem_u32(q_s), da = smem_u32(do_s), ka = smem_u32(k_s), va = smem_u32(v_s), kta = smem_u32(kt_s);
      // S first, then dP: the K stage frees a whole dP-phase earlier, so the (single-stage) K and V loads of the next tile
      // hide behind the dP MMAs and the dQ MMAs respectively
      auto issue_sdp = [&](int t) {
        const uint32_t ph = (uint32_t)(t & 1);
        mbar_wait(smem_u32(k_full), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (kb | k) != 0;
            const uint64_t bk = umma_desc_sw128(ka + kb * 2 * B5_BOX64 + k * 32);
            b5_mma_ss(tmem_base + 0, umma_desc_sw128(qa + kb * B5_BOX128 + k * 32), bk, B5_ID128, acc);         // Q_hi . [K_hi ; K_lo]
            b5_mma_ss(tmem_base + 128, umma_desc_sw128(qa + (2 + kb) * B5_BOX128 + k * 32), bk, B5_ID64, acc);   // Q_lo . K_hi
          }
        b5_commit(smem_u32(k_empty));
        mbar_wait(smem_u32(v_full), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (kb | k) != 0;
            const uint64_t bvh = umma_desc_sw128(va + kb * 2 * B5_BOX64 + k * 32);
            const uint64_t bvl = umma_desc_sw128(va + (kb * 2 + 1) * B5_BOX64 + k * 32);
            const uint64_t adh = umma_desc_sw128(da + kb * B5_BOX128 + k * 32);
            const uint64_t adl = umma_desc_sw128(da + (2 + kb) * B5_BOX128 + k * 32);
            b5_mma_ss(tmem_base + 192, adh, bvh, B5_ID64, acc);   // dO_hi . V_hi
            b5_mma_ss(tmem_base + 192, adh, bvl, B5_ID64, 1);     // dO_hi . V_lo
            b5_mma_ss(tmem_base + 192, adl, bvh, B5_ID64, 1);     // dO_lo . V_hi
          }
        b5_commit(smem_u32(v_empty));
        b5_commit(smem_u32(sp_full));
      };
      for (int tt = 0; tt <= n_tiles; ++tt) {
        if (tt < n_tiles) {
          if (tt > 0) mbar_wait(smem_u32(sp_free), (uint32_t)((tt - 1) & 1));   // S/dP of tile tt-1 now live in registers
          issue_sdp(tt);
        }
        if (tt == 0) continue;
        const int t = tt - 1;
        const uint32_t ph = (uint32_t)(t & 1);
        mbar_wait(smem_u32(ds_full), ph);
        mbar_wait(smem_u32(kt_full), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = ((t % DQ_FLUSH) | kb | k) != 0;   // restart after every drain
            const uint64_t b = umma_desc_sw128(kta + kb * 2 * B5_BOX64 + k * 32);
            b5_mma_ts(tmem_base + 384, tmem_base + 256 + kb * 32 + k * 8, b, B5_ID128, acc);   // dS_hi . [Kt_hi ; Kt_lo]
            b5_mma_ts(tmem_base + 384, tmem_base + 320 + kb * 32 + k * 8, b, B5_ID64, 1);      // dS_lo . Kt_hi
          }
        b5_commit(smem_u32(kt_empty));
        b5_commit(smem_u32(dq_done));
      }
      b5_commit(smem_u32(acc_full));
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int w = warp - 4;
    const int qd = w & 3, cq = w >> 2;
    const int row = qd * 32 + lane, grow = i0 + row;
    const uint32_t lb = (uint32_t)(qd * 32) << 16;
    // rows past the chunk: lse = +huge makes every probability (and dS) exactly zero
    const float lse2 = grow < m ? __ldg(lse + base + grow) * B5_LOG2E : 1e30f;
    const float dl_r = grow < m ? __ldg(delta + base + grow) : 0.f;
    const float sl2 = scale * B5_LOG2E;
    const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t drop_thr = DROP ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
    const uint32_t hrow = attn_drop_pre(seed, chunk) ^ ((uint32_t)grow * ATTN_DROP_CI);
    const uint32_t tb = tmem_base + lb;
    float a[16], b[16], c[16], x[16];
    auto drain_dq = [&](bool first) {
      tmem_drain16<2>(tb + 384 + cq * 16, dq + (base + grow) * 64 + cq * 16, scale, first, grow < m);
    };
    // LAST: only the final key tile can hold keys past the chunk
    auto ew_tile = [&](int t, auto last_c) {
      constexpr bool LAST = decltype(last_c)::value;
      // Every DQ_FLUSH tiles: tiles [t - DQ_FLUSH, t) are complete in the accumulator once the dQ MMAs of tile t-1 have
      // retired -> drain them now, while few registers are live.  The dQ MMAs of tile t (which restart the accumulator) are
      // issued only after every elementwise thread has arrived on ds_full at the end of this tile, i.e. after this read.
      const bool drain = t > 0 && t % DQ_FLUSH == 0;
      if (drain) {
        mbar_wait(smem_u32(dq_done), (uint32_t)((t - 1) & 1));
        tcgen05_fence_after();
        drain_dq(!accum && t == DQ_FLUSH);
        tcgen05_fence_before();
      }
      mbar_wait(smem_u32(sp_full), (uint32_t)(t & 1));
      tcgen05_fence_after();
      b5_ld16(tb + cq * 16, a);
      b5_ld16(tb + 64 + cq * 16, b);
      b5_ld16(tb + 128 + cq * 16, c);
      b5_ld16(tb + 192 + cq * 16, x);
      b5_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(sp_free));
      const int j0 = (t0 + t) * 64 + cq * 16;
      const int lim = m - j0;
      [[maybe_unused]] const uint32_t hcol = (uint32_t)j0 * ATTN_DROP_CJ;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float p = b5_ex2(fmaf(a[e] + b[e] + c[e], sl2, -lse2));
        float dp = x[e];
        if constexpr (DROP) dp = (attn_drop_mix(hrow ^ (hcol + (uint32_t)e * ATTN_DROP_CJ)) >= drop_thr) ? dp * inv_keep : 0.f;
        float ds = p * (dp - dl_r);
        if (LAST) ds = e < lim ? ds : 0.f;
        b5_split(ds, hi[e], lo[e]);
      }
      if (t > 0 && !drain) mbar_wait(smem_u32(dq_done), (uint32_t)((t - 1) & 1));
      tcgen05_fence_after();
      b5_st16(tb + 256 + cq * 16, hi);
      b5_st16(tb + 320 + cq * 16, lo);
      b5_st_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(ds_full));
    };
    for (int t = 0; t < n_tiles; ++t) {
      if (t0 + t == n_tiles_all - 1) ew_tile(t, std::true_type{});
      else ew_tile(t, std::false_type{});
    }
    mbar_wait(smem_u32(acc_full), 0);
    tcgen05_fence_after();
    drain_dq(!accum && n_tiles <= DQ_FLUSH);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------- dK, dV
// smem: K hi/lo (4 x 16 KB) | V hi/lo (4 x 16 KB) | 2 stages of (Q 16 KB + dO 16 KB) | dO^T stage 16 KB | Q^T stage 16 KB
// TMEM: S^T [0,96) (wide)   dP^T [96,128) (three MMAs)   operand slots P~^T hi [128,160) lo [160,192), dS^T hi [192,224)
//       lo [224,256)   dV accumulator [256,384)   dK accumulator [384,512)
// Same register-buffered pipeline as the dQ kernel.
constexpr int DKV_QSTAGE = 8 * B5_BOX32;   // Q tile (16 KB) + dO tile (16 KB); two stages
constexpr int DKV_SMEM = 1024 + 8 * B5_BOX128 + 2 * DKV_QSTAGE + 8 * B5_BOX32 + 1024;

template <bool DROP>
__global__ void __launch_bounds__(B5_THREADS, 1)
    attn_bwd_dkv_t5_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                           const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                           const __grid_constant__ CUtensorMap map_dot, const __grid_constant__ CUtensorMap map_qt,
                           const float* __restrict__ lse2p, const float* __restrict__ dlp, int m, int mp, float scale,
                           float drop_p, uint64_t seed, float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* k_s = smem;
  uint8_t* v_s = k_s + 4 * B5_BOX128;
  uint8_t* q_s = v_s + 4 * B5_BOX128;      // stage s at q_s + s*DKV_QSTAGE: Q [kb][hi 4 KB | lo 4 KB], then dO likewise
  uint8_t* dot_s = q_s + 2 * DKV_QSTAGE;   // [hi d 0..63 (8 KB) | lo (8 KB)] x 32 queries
  uint8_t* qt_s = dot_s + 4 * B5_BOX32;
  uint64_t* bars = (uint64_t*)(qt_s + 4 * B5_BOX32);
  uint64_t* r_full = bars;
  uint64_t* q_full = bars + 1;     // [2]  Q and dO row tiles (S-type operands)
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* t_full = bars + 5;
  uint64_t* t_empty = bars + 6;    // dO^T and Q^T planes (accumulation operands)
  uint64_t* sp_full = bars + 7;
  uint64_t* sp_free = bars + 8;    // 512 arrivals
  uint64_t* op_full = bars + 9;    // 512 arrivals
  uint64_t* acc_done = bars + 10;
  uint64_t* acc_full = bars + 11;
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int j0 = blockIdx.x * 128;            // first key of this CTA
  const int n_tiles_all = (m + 31) / 32;       // gridDim.z > 1: query tiles split over z, partial dK / dV added (see the dQ kernel)
  const int t0 = (int)((long long)n_tiles_all * blockIdx.z / gridDim.z);
  const int n_tiles = (int)((long long)n_tiles_all * (blockIdx.z + 1) / gridDim.z) - t0;
  const bool accum = gridDim.z > 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) mbar_init(smem_u32(bars + i), (i == 8 || i == 9) ? B5_EW : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(smem_u32(r_full), 8 * B5_BOX128);
      for (int part = 0; part < 2; ++part)
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(smem_u32(k_s + (part * 2 + kb) * B5_BOX128), &map_k, smem_u32(r_full), part * 64 + kb * 32, (int)(base + j0));
          tma_load_2d(smem_u32(v_s + (part * 2 + kb) * B5_BOX128), &map_v, smem_u32(r_full), part * 64 + kb * 32, (int)(base + j0));
        }
      for (int t = 0; t < n_tiles; ++t) {
        const int i0 = (t0 + t) * 32, st = t & 1;
        uint8_t* qs = q_s + st * DKV_QSTAGE;
        mbar_wait(smem_u32(q_empty + st), (uint32_t)((t >> 1) & 1) ^ 1);
        mbar_expect_tx(smem_u32(q_full + st), 8 * B5_BOX32);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part) {
            tma_load_2d(smem_u32(qs + (kb * 2 + part) * B5_BOX32), &map_q, smem_u32(q_full + st), part * 64 + kb * 32, (int)(base + i0));
            tma_load_2d(smem_u32(qs + (4 + kb * 2 + part) * B5_BOX32), &map_do, smem_u32(q_full + st), part * 64 + kb * 32,
                        (int)(base + i0));
          }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      for (int t = 0; t < n_tiles; ++t) {
        const int i0 = (t0 + t) * 32;
        mbar_wait(smem_u32(t_empty), (uint32_t)(t & 1) ^ 1);
        mbar_expect_tx(smem_u32(t_full), 4 * B5_BOX64);
        for (int part = 0; part < 2; ++part) {
          tma_load_2d(smem_u32(dot_s + part * B5_BOX64), &map_dot, smem_u32(t_full), i0, chunk * 128 + part * 64);
          tma_load_2d(smem_u32(qt_s + part * B5_BOX64), &map_qt, smem_u32(t_full), i0, chunk * 128 + part * 64);
        }
      }
    }
  } else if (warp == 1) {
    {   // all 32 lanes run the issue loop; each tcgen05 instruction is issued by one elected lane
      mbar_wait(smem_u32(r_full), 0);
      tcgen05_fence_after();
      const uint32_t ka = smem_u32(k_s), va = smem_u32(v_s);
      const uint32_t dota = smem_u32(dot_s), qta = smem_u32(qt_s);
      auto issue_sdp = [&](int t) {
        const int st = t & 1;
        const uint32_t qa = smem_u32(q_s + st * DKV_QSTAGE), da = qa + 4 * B5_BOX32;
        mbar_wait(smem_u32(q_full + st), (uint32_t)((t >> 1) & 1));
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (kb | k) != 0;
            const uint64_t bq = umma_desc_sw128(qa + kb * 2 * B5_BOX32 + k * 32);
            const uint64_t bdh = umma_desc_sw128(da + kb * 2 * B5_BOX32 + k * 32);
            const uint64_t bdl = umma_desc_sw128(da + (kb * 2 + 1) * B5_BOX32 + k * 32);
            const uint64_t akh = umma_desc_sw128(ka + kb * B5_BOX128 + k * 32);
            const uint64_t akl = umma_desc_sw128(ka + (2 + kb) * B5_BOX128 + k * 32);
            const uint64_t avh = umma_desc_sw128(va + kb * B5_BOX128 + k * 32);
            const uint64_t avl = umma_desc_sw128(va + (2 + kb) * B5_BOX128 + k * 32);
            b5_mma_ss(tmem_base + 0, akh, bq, B5_ID64, acc);      // S^T: K_hi . [Q_hi ; Q_lo]
            b5_mma_ss(tmem_base + 96, avh, bdh, B5_ID32, acc);    // dP^T: V_hi . dO_hi
            b5_mma_ss(tmem_base + 64, akl, bq, B5_ID32, acc);     // S^T: K_lo . Q_hi
            b5_mma_ss(tmem_base + 96, avh, bdl, B5_ID32, 1);      // dP^T: V_hi . dO_lo
            b5_mma_ss(tmem_base + 96, avl, bdh, B5_ID32, 1);      // dP^T: V_lo . dO_hi
          }
        b5_commit(smem_u32(q_empty + st));
        b5_commit(smem_u32(sp_full));
      };
      for (int tt = 0; tt <= n_tiles; ++tt) {
        if (tt < n_tiles) {
          if (tt > 0) mbar_wait(smem_u32(sp_free), (uint32_t)((tt - 1) & 1));
          issue_sdp(tt);
        }
        if (tt == 0) continue;
        const int t = tt - 1;
        const uint32_t ph = (uint32_t)(t & 1);
        mbar_wait(smem_u32(op_full), ph);
        mbar_wait(smem_u32(t_full), ph);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t acc = ((t % DKV_FLUSH) | k) != 0;   // restart after every drain
          const uint64_t bdo = umma_desc_sw128(dota + k * 32);
          const uint64_t bq = umma_desc_sw128(qta + k * 32);
          b5_mma_ts(tmem_base + 256, tmem_base + 128 + k * 8, bdo, B5_ID128, acc);   // dV += P~^T_hi . [dO^T_hi ; dO^T_lo]
          b5_mma_ts(tmem_base + 384, tmem_base + 192 + k * 8, bq, B5_ID128, acc);    // dK += dS^T_hi . [Q^T_hi ; Q^T_lo]
          b5_mma_ts(tmem_base + 256, tmem_base + 160 + k * 8, bdo, B5_ID64, 1);      // dV += P~^T_lo . dO^T_hi
          b5_mma_ts(tmem_base + 384, tmem_base + 224 + k * 8, bq, B5_ID64, 1);       // dK += dS^T_lo . Q^T_hi
        }
        b5_commit(smem_u32(t_empty));
        b5_commit(smem_u32(acc_done));
      }
      b5_commit(smem_u32(acc_full));
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int w = warp - 4;
    const int qd = w & 3, cq = w >> 2;             // 8 query columns per warp
    const int row = qd * 32 + lane, gkey = j0 + row;
    const uint32_t lb = (uint32_t)(qd * 32) << 16;
    const float sl2 = scale * B5_LOG2E;
    const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t drop_thr = DROP ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
    const uint32_t hkey = attn_drop_pre(seed, chunk) ^ ((uint32_t)gkey * ATTN_DROP_CJ);
    const uint32_t tb = tmem_base + lb;
    const float4* lse4 = reinterpret_cast<const float4*>(lse2p + (long long)chunk * mp) + cq * 2;
    const float4* dl4 = reinterpret_cast<const float4*>(dlp + (long long)chunk * mp) + cq * 2;
    float a[8], b[8], c[8], x[8];
    auto drain_one = [&](uint32_t col, float* out, float mul, bool first) {
      tmem_drain16<2>(tb + col + cq * 16, out + (base + gkey) * 64 + cq * 16, mul, first, gkey < m);
    };
    for (int t = 0; t < n_tiles; ++t) {
      const bool drain = t > 0 && t % DKV_FLUSH == 0;   // see the dQ kernel
      if (drain) {
        mbar_wait(smem_u32(acc_done), (uint32_t)((t - 1) & 1));
        tcgen05_fence_after();
        drain_one(256, dv, 1.f, !accum && t == DKV_FLUSH);
        drain_one(384, dk, scale, !accum && t == DKV_FLUSH);
        tcgen05_fence_before();
      }
      // per-query statistics of the 8 columns of this warp (padded rows: lse = +huge -> probability exactly 0)
      const float4 l0 = __ldg(lse4 + (t0 + t) * 8), l1 = __ldg(lse4 + (t0 + t) * 8 + 1);
      const float4 d0 = __ldg(dl4 + (t0 + t) * 8), d1 = __ldg(dl4 + (t0 + t) * 8 + 1);
      const float lse_c[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      const float dl_c[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      mbar_wait(smem_u32(sp_full), (uint32_t)(t & 1));
      tcgen05_fence_after();
      b5_ld8(tb + cq * 8, a);
      b5_ld8(tb + 32 + cq * 8, b);
      b5_ld8(tb + 64 + cq * 8, c);
      b5_ld8(tb + 96 + cq * 8, x);
      b5_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(sp_free));
      const uint32_t hq = (uint32_t)((t0 + t) * 32 + cq * 8) * ATTN_DROP_CI;
      uint32_t ph_[8], pl_[8], sh_[8], sl_[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float p = b5_ex2(fmaf(a[e] + b[e] + c[e], sl2, -lse_c[e]));
        float pt = p, dp = x[e];
        if constexpr (DROP) {
          const bool keep = attn_drop_mix(hkey ^ (hq + (uint32_t)e * ATTN_DROP_CI)) >= drop_thr;
          pt = keep ? p * inv_keep : 0.f;
          dp = keep ? dp * inv_keep : 0.f;
        }
        b5_split(pt, ph_[e], pl_[e]);
        b5_split(p * (dp - dl_c[e]), sh_[e], sl_[e]);
      }
      if (t > 0 && !drain) mbar_wait(smem_u32(acc_done), (uint32_t)((t - 1) & 1));
      tcgen05_fence_after();
      b5_st8(tb + 128 + cq * 8, ph_);
      b5_st8(tb + 160 + cq * 8, pl_);
      b5_st8(tb + 192 + cq * 8, sh_);
      b5_st8(tb + 224 + cq * 8, sl_);
      b5_st_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(op_full));
    }
    mbar_wait(smem_u32(acc_full), 0);
    tcgen05_fence_after();
    drain_one(256, dv, 1.f, !accum && n_tiles <= DKV_FLUSH);
    drain_one(384, dk, scale, !accum && n_tiles <= DKV_FLUSH);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------- host side
static unsigned long long g_b5_attr = 0;

int64_t attn_t5_bwd_workspace_bytes(int m) {
  const long long mp = ((long long)m + 63) / 64 * 64;
  return (4ll * (4ll * m * 128) + 3ll * (4ll * 128 * mp) + 2ll * 4 * mp) * 4 + 1024;
}

int launch_attn_bwd_t5(const float* q, const float* k, const float* v, const float* lse, const float* delta, const float* d_ctx, int m,
                       float scale, float drop_p, uint64_t seed, float* dq, float* dk, float* dv, void* workspace, cudaStream_t st) {
  const int mp = (m + 63) / 64 * 64;
  const long long hl = 4ll * m * 128, pl = 4ll * 128 * mp;
  float* w0 = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float *q_hl = w0, *k_hl = w0 + hl, *v_hl = w0 + 2 * hl, *do_hl = w0 + 3 * hl;
  float *kt = w0 + 4 * hl, *dot = kt + pl, *qt = dot + pl;
  float *lse2p = qt + pl, *dlp = lse2p + 4ll * mp;
  PrepArgs pa;
  pa.n_rows = 4;
  pa.rows_src[0] = q; pa.rows_src[1] = k; pa.rows_src[2] = v; pa.rows_src[3] = d_ctx;
  pa.rows_dst[0] = q_hl; pa.rows_dst[1] = k_hl; pa.rows_dst[2] = v_hl; pa.rows_dst[3] = do_hl;
  pa.n_t = 3;
  pa.t_src[0] = k; pa.t_src[1] = d_ctx; pa.t_src[2] = q;
  pa.t_dst[0] = kt; pa.t_dst[1] = dot; pa.t_dst[2] = qt;
  pa.lse = lse; pa.delta = delta; pa.lse2p = lse2p; pa.dlp = dlp;
  attn_bwd_prep_kernel<<<4 * sm_count(), 256, 0, st>>>(pa, m, mp);
  SCAN_LAUNCH_CHECK("attn_bwd_prep_kernel");
  CUtensorMap mq128, mdo128, mk128, mv128, mk64, mv64, mq32, mdo32, mkt, mdot, mqt;
  int rc = 0;
  rc |= make_rowmajor_map(&mq128, q_hl, 4ull * m, 128, 128);
  rc |= make_rowmajor_map(&mdo128, do_hl, 4ull * m, 128, 128);
  rc |= make_rowmajor_map(&mk128, k_hl, 4ull * m, 128, 128);
  rc |= make_rowmajor_map(&mv128, v_hl, 4ull * m, 128, 128);
  rc |= make_rowmajor_map(&mk64, k_hl, 4ull * m, 128, 64);
  rc |= make_rowmajor_map(&mv64, v_hl, 4ull * m, 128, 64);
  rc |= make_rowmajor_map(&mq32, q_hl, 4ull * m, 128, 32);
  rc |= make_rowmajor_map(&mdo32, do_hl, 4ull * m, 128, 32);
  rc |= make_rowmajor_map(&mkt, kt, 4ull * 128, (uint64_t)mp, 64);
  rc |= make_rowmajor_map(&mdot, dot, 4ull * 128, (uint64_t)mp, 64);
  rc |= make_rowmajor_map(&mqt, qt, 4ull * 128, (uint64_t)mp, 64);
  if (rc) return SCAN_ECUDA;
  if (first_use_on_device(&g_b5_attr)) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_t5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dq_t5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_t5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_dkv_t5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
  }
  // split the inner loop over gridDim.z when the (tile, chunk) grid fills the last wave badly: pick the split count that
  // minimises waves / split (ties -> fewer splits), and only if it buys at least 8 %
  const int ctas = ((m + 127) / 128) * 4, sms = sm_count();
  int split = 1;
  double best = (double)((ctas + sms - 1) / sms);
  const double base_cost = best;
  for (int sp = 2; sp <= 6 && sp * 8 <= (m + 63) / 64; ++sp) {
    const double cost = (double)((ctas * sp + sms - 1) / sms) / sp + 0.02 * sp;   // + per-split prologue
    if (cost < best - 1e-9) { best = cost; split = sp; }
  }
  if (best > 0.92 * base_cost) split = 1;
  static const int forced = getenv("SCAN_B200_ATTN_SPLIT") ? atoi(getenv("SCAN_B200_ATTN_SPLIT")) : 0;
  if (forced > 0) split = forced;
  if (split > 1) {   // partial results are added: outputs start from zero (nondeterministic fp32 add order across splits)
    SCAN_CUDA_CHECK(cudaMemsetAsync(dq, 0, sizeof(float) * 4ull * m * 64, st));
    SCAN_CUDA_CHECK(cudaMemsetAsync(dk, 0, sizeof(float) * 4ull * m * 64, st));
    SCAN_CUDA_CHECK(cudaMemsetAsync(dv, 0, sizeof(float) * 4ull * m * 64, st));
  }
  dim3 grid((m + 127) / 128, 4, split);
  if (drop_p > 0.f) {
    attn_bwd_dq_t5_kernel<true><<<grid, B5_THREADS, DQ_SMEM, st>>>(mq128, mdo128, mk64, mv64, mkt, lse, delta, m, scale, drop_p, seed, dq);
    SCAN_LAUNCH_CHECK("attn_bwd_dq_t5_kernel");
    attn_bwd_dkv_t5_kernel<true><<<grid, B5_THREADS, DKV_SMEM, st>>>(mk128, mv128, mq32, mdo32, mdot, mqt, lse2p, dlp, m, mp, scale,
                                                                    drop_p, seed, dk, dv);
  } else {
    attn_bwd_dq_t5_kernel<false><<<grid, B5_THREADS, DQ_SMEM, st>>>(mq128, mdo128, mk64, mv64, mkt, lse, delta, m, scale, drop_p, seed, dq);
    SCAN_LAUNCH_CHECK("attn_bwd_dq_t5_kernel");
    attn_bwd_dkv_t5_kernel<false><<<grid, B5_THREADS, DKV_SMEM, st>>>(mk128, mv128, mq32, mdo32, mdot, mqt, lse2p, dlp, m, mp, scale,
                                                                     drop_p, seed, dk, dv);
  }
  SCAN_LAUNCH_CHECK("attn_bwd_dkv_t5_kernel");
  return SCAN_OK;
}

}  // namespace scan
