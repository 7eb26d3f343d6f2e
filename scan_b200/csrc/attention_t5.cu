// K3a on tcgen05: streaming attention over the 4 chunks of M 64-d sub-tokens (layers/transformer.py:5-34 under the
// .view of :66-68) with the 128 x 64 score tiles and the P.V products on the 5th-gen tensor cores.
//
// Numerics: 3xTF32 (x = hi + lo): the scores feed an exponential, plain tf32 would break the 1e-3 contract.  The two
// "hi x (hi | lo)" products are fused into ONE wide MMA by stacking [hi ; lo] of the B operand along N, so a tile needs
// two short, independent accumulation chains instead of three dependent ones (dependent tcgen05.mma chains are bound by
// the tensor-pipe latency, see condconv_ts.inl).
//
// Forward = two passes over the key tiles (no online rescaling of a TMEM-resident O):
//   pass 1  S = Q K^T -> per-row running max / sum -> lse           (S double-buffered in TMEM)
//   pass 2  S again, P = exp(S*scale - lse) (already normalised), dropout, split into hi / lo, stored to TENSOR MEMORY
//           as the A operand of O += P V (tcgen05.mma with A from TMEM); O accumulates across all key tiles.
// Operands come from a small pre-pass (attn_prep_kernel): q_hl / k_hl [4M, 128] = (hi | lo) per sub-token and the
// transposed V^T planes [chunk][hi d 0..63 | lo d 0..63][Mp] (keys contiguous), all loaded by TMA (128-byte swizzle).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4..19 softmax (lane quarter = w % 4, 16-key block = w / 4):
// four softmax warps per scheduler hide the TMEM / MUFU / ALU latencies (one or two per scheduler run at IPC ~0.1).
#include <stdlib.h>

#include <type_traits>

#include "tc_common.cuh"

namespace scan {

constexpr int T5_BQ = 128;   // query rows per CTA (TMEM lanes)
constexpr int T5_BK = 64;    // keys per tile
constexpr int T5_D = 64;
constexpr int T5_QBOX = T5_BQ * 32 * 4;   // 16 KB: [128 rows x 32 cols]
constexpr int T5_KBOX = T5_BK * 32 * 4;   // 8 KB:  [64 rows x 32 cols]
constexpr int T5_Q_BYTES = 4 * T5_QBOX;   // hi kb0, hi kb1, lo kb0, lo kb1
constexpr int T5_STAGE = 4 * T5_KBOX;     // [kb0: hi | lo][kb1: hi | lo] = 32 KB
constexpr int T5_SMEM = 1024 + T5_Q_BYTES + 2 * T5_STAGE + 2 * T5_STAGE + 6144;   // + barriers and the 4 KB row statistics
constexpr int T5_SM_WARPS = 16;           // softmax warps: lane quarter = w % 4, 16-column block = w / 4
constexpr int T5_THREADS = 128 + 32 * T5_SM_WARPS;
constexpr int T5_S_COLS = 192;            // Sa (hi.hi | hi.lo) 128 + Sb (lo.hi) 64
constexpr int T5_P_COL0 = 192;            // P operand slot: hi 64 | lo 64
constexpr int T5_O_COL0 = 320;            // Oa 128 + Ob 64
// The tensor core adds into its fp32 accumulator with truncation, not round-to-nearest: over the ~1100 accumulation steps
// of a full-size O row that is a coherent 5e-5 relative error (measured, tools/diag_attn_precision.py).  O is therefore
// drained every T5_FLUSH key tiles (64 steps) into the fp32 output row with ordinary round-to-nearest adds.
constexpr int T5_FLUSH = 8;
constexpr uint32_t T5_IDESC_N128 = umma_idesc_tf32(T5_BQ, 128);
constexpr uint32_t T5_IDESC_N64 = umma_idesc_tf32(T5_BQ, 64);


__device__ __forceinline__ void t5_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void t5_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void t5_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void t5_ld16(uint32_t taddr, float (&v)[16]) {
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
__device__ __forceinline__ void t5_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void t5_commit(uint32_t bar) {
  if (elect_one_sync()) umma_commit(bar);
}
__device__ __forceinline__ void t5_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float t5_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void t5_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

// ---------------------------------------------------------------------------- pre-pass
// q_hl, k_hl: [4M, 128] = hi(64) | lo(64);  vt: [4][128][Mp], rows 0..63 hi of dim d, rows 64..127 lo, zero for keys >= M
__global__ void __launch_bounds__(256) attn_prep_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                        int m, int mp, float* __restrict__ q_hl, float* __restrict__ k_hl,
                                                        float* __restrict__ vt) {
  __shared__ float th[64][33], tl[64][33];
  const long long n_rows = 4ll * m;
  // part 1: row-wise hi/lo of q and k (grid-stride over float4)
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows * 16; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i >> 4;
    const int c4 = (int)(i & 15);
    for (int which = 0; which < 2; ++which) {
      const float4 x = __ldg(reinterpret_cast<const float4*>((which ? k : q) + row * 64) + c4);
      float4 h, l;
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
      float* dst = (which ? k_hl : q_hl) + row * 128;
      reinterpret_cast<float4*>(dst)[c4] = h;
      reinterpret_cast<float4*>(dst + 64)[c4] = l;
    }
  }
  // part 2: V^T planes, 32-key x 64-dim tiles through shared memory
  const int tiles_per_chunk = mp / 32;
  for (int tile = blockIdx.x; tile < 4 * tiles_per_chunk; tile += gridDim.x) {
    const int chunk = tile / tiles_per_chunk, j0 = (tile % tiles_per_chunk) * 32;
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
      const int key = i >> 6, d = i & 63;
      float x = 0.f;
      if (j0 + key < m) x = __ldg(v + ((long long)chunk * m + j0 + key) * 64 + d);
      uint32_t u;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
      th[d][key] = __uint_as_float(u);
      tl[d][key] = x - __uint_as_float(u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
      const int d = i >> 5, key = i & 31;
      vt[((long long)chunk * 128 + d) * mp + j0 + key] = th[d][key];
      vt[((long long)chunk * 128 + 64 + d) * mp + j0 + key] = tl[d][key];
    }
  }
}

// ---------------------------------------------------------------------------- forward
template <bool DROP>
__global__ void __launch_bounds__(T5_THREADS, 1)
    attn_fwd_t5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                       const __grid_constant__ CUtensorMap map_v, int m, float scale, float drop_p, uint64_t seed,
                       float* __restrict__ ctx, float* __restrict__ lse, const int* __restrict__ run_flag) {
  // fallback of the single-pass kernel below: runs only when that kernel flagged a row whose probabilities underflowed
  if (run_flag && *run_flag == 0) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + T5_Q_BYTES;
  uint8_t* v_s = k_s + 2 * T5_STAGE;
  uint64_t* bars = (uint64_t*)(v_s + 2 * T5_STAGE);
  uint64_t* k_full = bars;          // [2]
  uint64_t* k_empty = bars + 2;     // [2]
  uint64_t* v_full = bars + 4;      // [2]
  uint64_t* v_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [2]
  uint64_t* s_empty = bars + 10;    // [2] (256 arrivals)
  uint64_t* p_full = bars + 12;     // (256 arrivals)
  uint64_t* p_empty = bars + 13;
  uint64_t* o_full = bars + 14;
  uint64_t* q_full = bars + 15;
  uint32_t* tmem_slot = (uint32_t*)(bars + 16);
  float* stat_m = (float*)(bars + 18);      // [4][128] per-column-block running max
  float* stat_l = stat_m + 512;             // [4][128] per-column-block running sum

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;   // first sub-token row of the chunk
  const int i0 = blockIdx.x * T5_BQ;
  const int n_tiles = (m + T5_BK - 1) / T5_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(k_full + i), 1);
      mbar_init(smem_u32(k_empty + i), 1);
      mbar_init(smem_u32(v_full + i), 1);
      mbar_init(smem_u32(v_empty + i), 1);
      mbar_init(smem_u32(s_full + i), 1);
      mbar_init(smem_u32(s_empty + i), 32 * T5_SM_WARPS);
    }
    mbar_init(smem_u32(p_full), 32 * T5_SM_WARPS);
    mbar_init(smem_u32(p_empty), 1);
    mbar_init(smem_u32(o_full), 1);
    mbar_init(smem_u32(q_full), 1);
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
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), T5_Q_BYTES);
      for (int part = 0; part < 2; ++part)      // hi, lo
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(smem_u32(q_s + (part * 2 + kb) * T5_QBOX), &map_q, smem_u32(q_full), part * 64 + kb * 32, (int)(base + i0));
      int ks = 0;
      uint32_t kph = 0;
      for (int pass = 0; pass < 2; ++pass) {
        for (int t = 0; t < n_tiles; ++t) {
          const int j0 = t * T5_BK;
          mbar_wait(smem_u32(k_empty + ks), kph ^ 1);
          mbar_expect_tx(smem_u32(k_full + ks), T5_STAGE);
          for (int kb = 0; kb < 2; ++kb)
            for (int part = 0; part < 2; ++part)
              tma_load_2d(smem_u32(k_s + ks * T5_STAGE + (kb * 2 + part) * T5_KBOX), &map_k, smem_u32(k_full + ks), part * 64 + kb * 32,
                          (int)(base + j0));
          if (++ks == 2) { ks = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ===== second TMA producer: V^T planes (pass 2 only), independent of the K ring =====
    if (lane == 0) {
      int vs = 0;
      uint32_t vph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = t * T5_BK;
        mbar_wait(smem_u32(v_empty + vs), vph ^ 1);
        mbar_expect_tx(smem_u32(v_full + vs), T5_STAGE);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part)
            tma_load_2d(smem_u32(v_s + vs * T5_STAGE + (kb * 2 + part) * T5_KBOX), &map_v, smem_u32(v_full + vs), j0 + kb * 32,
                        chunk * 128 + part * 64);
        if (++vs == 2) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: all 32 lanes run the loop, each tcgen05 instruction is issued by one elected lane =====
    {
      mbar_wait(smem_u32(q_full), 0);
      tcgen05_fence_after();
      const uint32_t q_addr = smem_u32(q_s);
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      // S tile: Sa[128 x 128] = Qhi . [Khi ; Klo]^T, Sb[128 x 64] = Qlo . Khi^T
      auto issue_s = [&](uint32_t s_col) {
        mbar_wait(smem_u32(k_full + ks), kph);
        tcgen05_fence_after();
        const uint32_t kaddr = smem_u32(k_s + ks * T5_STAGE);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t b = umma_desc_sw128(kaddr + kb * 2 * T5_KBOX + k * 32);
            t5_mma_ss(tmem_base + s_col, umma_desc_sw128(q_addr + kb * T5_QBOX + k * 32), b, T5_IDESC_N128, (kb | k) != 0);
            t5_mma_ss(tmem_base + s_col + 128, umma_desc_sw128(q_addr + (2 + kb) * T5_QBOX + k * 32), b, T5_IDESC_N64, (kb | k) != 0);
          }
        t5_commit(smem_u32(k_empty + ks));
        if (++ks == 2) { ks = 0; kph ^= 1; }
      };
      // ---- pass 1: S double-buffered at columns 0 / 192
      {
        uint32_t se_ph[2] = {0, 0};
        for (int t = 0; t < n_tiles; ++t) {
          const int buf = t & 1;
          mbar_wait(smem_u32(s_empty + buf), se_ph[buf] ^ 1);
          se_ph[buf] ^= 1;
          tcgen05_fence_after();
          issue_s(buf * T5_S_COLS);
          t5_commit(smem_u32(s_full + buf));
        }
      }
      // ---- pass 2: single S buffer (columns 0..191), P slot, O accumulators
      {
        // s_empty[0] phase bookkeeping continues from pass 1: count the completed phases so far
        uint32_t se0 = (uint32_t)(((n_tiles + 1) / 2) & 1);   // parity of completions of s_empty[0] consumed in pass 1
        uint32_t pf = 0;
        // wait until the softmax warps have drained the last pass-1 tiles from BOTH buffers (buffer 1 overlaps P / O)
        // -> handled by the s_empty waits below plus an explicit wait on buffer 1
        uint32_t se1 = (uint32_t)((n_tiles / 2) & 1);
        if (n_tiles >= 2) mbar_wait(smem_u32(s_empty + 1), se1 ^ 1);
        tcgen05_fence_after();
        for (int t = 0; t <= n_tiles; ++t) {
          if (t < n_tiles) {
            mbar_wait(smem_u32(s_empty + 0), se0 ^ 1);
            se0 ^= 1;
            tcgen05_fence_after();
            issue_s(0);
            t5_commit(smem_u32(s_full + 0));
          }
          if (t > 0) {  // O += P(t-1) . V(t-1)
            mbar_wait(smem_u32(p_full), pf);
            pf ^= 1;
            mbar_wait(smem_u32(v_full + vs), vph);
            tcgen05_fence_after();
            const uint32_t vaddr = smem_u32(v_s + vs * T5_STAGE);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t b = umma_desc_sw128(vaddr + kb * 2 * T5_KBOX + k * 32);
                const uint32_t acc = (((t - 1) % T5_FLUSH) | kb | k) != 0;   // restart after every drain
                t5_mma_ts(tmem_base + T5_O_COL0, tmem_base + T5_P_COL0 + kb * 32 + k * 8, b, T5_IDESC_N128, acc);
                t5_mma_ts(tmem_base + T5_O_COL0 + 128, tmem_base + T5_P_COL0 + 64 + kb * 32 + k * 8, b, T5_IDESC_N64, acc);
              }
            t5_commit(smem_u32(v_empty + vs));
            t5_commit(smem_u32(p_empty));
            if (++vs == 2) { vs = 0; vph ^= 1; }
          }
        }
        t5_commit(smem_u32(o_full));
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== softmax / epilogue warps: thread = query row (TMEM lane), 16-column block cq of the 64 key columns =====
    const int w = warp - 4;
    const int qd = w & 3, cq = w >> 2;
    const int row = qd * 32 + lane;                 // row inside the tile
    const int grow = i0 + row;                      // row inside the chunk
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    // everything in the log2 domain: s2 = S * scale * log2(e), p = 2^(s2 - lse2)
    const float sl2 = scale * 1.4426950408889634f;
    float mrun = -INFINITY, lrun = 0.f;
    uint32_t sf_ph[2] = {0, 0};
    float a[16], b[16], c[16];
    // ---- pass 1: statistics.  LAST: only the final key tile can hold keys past the chunk.
    auto stat_tile = [&](int t, auto last_c) {
      constexpr bool LAST = decltype(last_c)::value;
      const int buf = t & 1;
      mbar_wait(smem_u32(s_full + buf), sf_ph[buf]);
      sf_ph[buf] ^= 1;
      tcgen05_fence_after();
      const uint32_t sc = tmem_base + lane_base + buf * T5_S_COLS;
      t5_ld16(sc + cq * 16, a);
      t5_ld16(sc + 64 + cq * 16, b);
      t5_ld16(sc + 128 + cq * 16, c);
      t5_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(s_empty + buf));
      const int lim = m - (t * T5_BK + cq * 16);
      float mx = -INFINITY;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        a[e] = (a[e] + b[e] + c[e]) * sl2;
        if (LAST) a[e] = e < lim ? a[e] : -INFINITY;
        mx = fmaxf(mx, a[e]);
      }
      const float mnew = fmaxf(mrun, mx);
      if (mnew != -INFINITY) {   // ex2(-inf) = 0 takes care of masked entries and of the first tile (mrun = -inf)
        float sum = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) sum += t5_ex2(a[e] - mnew);
        lrun = lrun * t5_ex2(mrun - mnew) + sum;
        mrun = mnew;
      }
    };
    for (int t = 0; t + 1 < n_tiles; ++t) stat_tile(t, std::false_type{});
    stat_tile(n_tiles - 1, std::true_type{});
    // combine the four column blocks -> log-sum-exp per row
    stat_m[cq * 128 + row] = mrun;
    stat_l[cq * 128 + row] = lrun;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * T5_SM_WARPS) : "memory");
    float lse2_row;
    {
      float mm = -INFINITY;
#pragma unroll
      for (int x = 0; x < 4; ++x) mm = fmaxf(mm, stat_m[x * 128 + row]);
      float ll = 0.f;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const float mx_ = stat_m[x * 128 + row];
        if (mx_ != -INFINITY) ll += stat_l[x * 128 + row] * exp2f(mx_ - mm);
      }
      lse2_row = mm + log2f(ll);
    }
    // The backward recomputes P from the STORED natural-log value: use exactly that value (same two roundings) here as well,
    // otherwise every probability of a row differs from the backward's by a common factor ~1e-6, which the dP - D
    // cancellation of the softmax gradient amplifies a thousandfold.
    const float lse_nat = lse2_row * 0.6931471805599453f;
    lse2_row = lse_nat * 1.4426950408889634f;
    if (cq == 0 && grow < m) lse[base + grow] = lse_nat;
    // ---- pass 2: P = 2^(s2 - lse2), dropout, hi/lo -> TMEM operand slot
    const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t drop_thr = DROP ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
    const uint32_t hrow = attn_drop_pre(seed, chunk) ^ ((uint32_t)grow * ATTN_DROP_CI);
    uint32_t sf0 = sf_ph[0], pe = 0;
    // drain the O accumulator (Oa[0..63] + Oa[64..127] + Ob, 16 of the 64 d columns per warp) into the output row
    auto drain_o = [&](bool first) {
      tmem_drain16<3>(tmem_base + lane_base + T5_O_COL0 + cq * 16, ctx + (base + grow) * T5_D + cq * 16, 1.f, first, grow < m);
    };
    auto prob_tile = [&](int t, auto last_c) {
      constexpr bool LAST = decltype(last_c)::value;
      // Every T5_FLUSH tiles: tiles [t - T5_FLUSH, t) are complete in O once the P.V MMAs of tile t-1 have retired -> drain
      // them now, while few registers are live.  The P.V MMAs of tile t restart the accumulator and are issued only after
      // every softmax thread has arrived on p_full at the end of this tile, i.e. after this read.
      const bool drain = t > 0 && t % T5_FLUSH == 0;
      if (drain) {
        mbar_wait(smem_u32(p_empty), pe);
        pe ^= 1;
        tcgen05_fence_after();
        drain_o(t == T5_FLUSH);
        tcgen05_fence_before();
      }
      mbar_wait(smem_u32(s_full + 0), sf0);
      sf0 ^= 1;
      tcgen05_fence_after();
      const uint32_t sc = tmem_base + lane_base;
      t5_ld16(sc + cq * 16, a);
      t5_ld16(sc + 64 + cq * 16, b);
      t5_ld16(sc + 128 + cq * 16, c);
      t5_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(s_empty + 0));
      const int j0 = t * T5_BK + cq * 16;
      const int lim = m - j0;
      [[maybe_unused]] const uint32_t hcol = (uint32_t)j0 * ATTN_DROP_CJ;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float p = t5_ex2(fmaf(a[e] + b[e] + c[e], sl2, -lse2_row));
        if constexpr (DROP) p = (attn_drop_mix(hrow ^ (hcol + (uint32_t)e * ATTN_DROP_CJ)) >= drop_thr) ? p * inv_keep : 0.f;
        if (LAST) p = e < lim ? p : 0.f;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(p));
        hi[e] = u;
        lo[e] = __float_as_uint(p - __uint_as_float(u));
      }
      if (t > 0 && !drain) {   // the previous P must have been consumed by its P.V MMAs
        mbar_wait(smem_u32(p_empty), pe);
        pe ^= 1;
      }
      tcgen05_fence_after();
      t5_st16(tmem_base + lane_base + T5_P_COL0 + cq * 16, hi);
      t5_st16(tmem_base + lane_base + T5_P_COL0 + 64 + cq * 16, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(smem_u32(p_full));
    };
    for (int t = 0; t + 1 < n_tiles; ++t) prob_tile(t, std::false_type{});
    prob_tile(n_tiles - 1, std::true_type{});
    // ---- epilogue: the tiles since the last drain
    mbar_wait(smem_u32(o_full), 0);
    tcgen05_fence_after();
    drain_o(n_tiles <= T5_FLUSH);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------------------- forward, single pass (product)
// The two-pass kernel above computes every score tile twice and its 16 elementwise warps execute ~35 instructions per score
// (ncu, round 1: 47 % issue-slot utilisation next to a 48 % busy tensor pipe -- instruction issue, not the MMAs, bounds it).
// This kernel visits every key tile ONCE:
//   * softmax reference = an upper bound instead of the row maximum: s_ij <= |q_i| max_j |k_j| (Cauchy-Schwarz; the norms come
//     from attn_norm_kernel), so p = 2^(s - bound) <= 1 never overflows and no statistics pass is needed; O accumulates the
//     UN-normalised P~ V, the row sums l accumulate in registers, attn_finish_kernel divides and writes the log-sum-exp.
//     A bound far above the true maximum only costs exponent range (fp32 and tf32 share the 8-bit exponent, the hi/lo split
//     is relative); if a row sum drops below 2^-100 attn_finish_kernel raises a flag and the two-pass kernel re-runs.
//   * the three 3xTF32 score products go into ONE 64-column accumulator (three N = 64 MMAs instead of the wide-N pair): the
//     elementwise warps load 16 instead of 48 TMEM values per tile and skip two adds per score, and S shrinks from 192 to 64
//     columns, which buys a THREE-deep S ring (the tensor pipe runs two tiles ahead of the softmax);
//   * P.V keeps the wide-N form (its accumulator is only read at the drains).
// TMEM: S ring [0,192) = 3 x 64   P operand slot [192,320) hi | lo   O [320,512) = Oa 128 + Ob 64.
constexpr int T1_SBUF = 3;

// |q_g| per sub-token row and max_j |k_j| per chunk (bits of a non-negative float order like unsigned integers)
__global__ void __launch_bounds__(256) attn_norm_kernel(const float* __restrict__ q, const float* __restrict__ k, int m, float* __restrict__ qn,
                                                        unsigned int* __restrict__ kmax_bits) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= 4ll * m) return;
  const float4* q4 = reinterpret_cast<const float4*>(q + r * 64);
  const float4* k4 = reinterpret_cast<const float4*>(k + r * 64);
  float sq = 0.f, sk = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 a = __ldg(q4 + i), b = __ldg(k4 + i);
    sq += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    sk += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  }
  qn[r] = sqrtf(sq);
  atomicMax(kmax_bits + (int)(r / m), __float_as_uint(sqrtf(sk)));
}

// ctx[g, :] /= l[g];  lse[g] = ln 2 * (bound2[g] + log2 l[g]);  flag |= (l[g] underflowed)
__global__ void __launch_bounds__(256) attn_finish_kernel(float* __restrict__ ctx, const float* __restrict__ lsum, const float* __restrict__ qn,
                                                          const unsigned int* __restrict__ kmax_bits, int m, float sl2,
                                                          float* __restrict__ lse, int* __restrict__ flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long g = i >> 4;
  if (g >= 4ll * m) return;
  const float l = __ldg(lsum + g);
  const float inv = 1.f / l;
  float4* c4 = reinterpret_cast<float4*>(ctx + g * 64) + (i & 15);
  float4 v = *c4;
  v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
  *c4 = v;
  if ((i & 15) == 0) {
    const float b2 = sl2 * __ldg(qn + g) * __uint_as_float(__ldg(kmax_bits + (int)(g / m))) * 1.000002f;
    lse[g] = (b2 + log2f(l)) * 0.6931471805599453f;
    if (!(l >= 7.9e-31f) || !(l <= 3.0e38f)) atomicOr(flag, 1);   // 2^-100; also catches NaN / inf
  }
}

template <bool DROP>
__global__ void __launch_bounds__(T5_THREADS, 1)
    attn_fwd1_t5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                        const __grid_constant__ CUtensorMap map_v, int m, float scale, float drop_p, uint64_t seed,
                        const float* __restrict__ qn, const unsigned int* __restrict__ kmax_bits, float* __restrict__ ctx,
                        float* __restrict__ lsum) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* q_s = smem;
  uint8_t* k_s = q_s + T5_Q_BYTES;
  uint8_t* v_s = k_s + 2 * T5_STAGE;
  uint64_t* bars = (uint64_t*)(v_s + 2 * T5_STAGE);
  uint64_t* k_full = bars;          // [2]
  uint64_t* k_empty = bars + 2;     // [2]
  uint64_t* v_full = bars + 4;      // [2]
  uint64_t* v_empty = bars + 6;     // [2]
  uint64_t* s_full = bars + 8;      // [3]
  uint64_t* s_empty = bars + 11;    // [3] (512 arrivals)
  uint64_t* p_full = bars + 14;     // (512 arrivals)
  uint64_t* p_empty = bars + 15;
  uint64_t* o_full = bars + 16;
  uint64_t* q_full = bars + 17;
  uint32_t* tmem_slot = (uint32_t*)(bars + 18);
  float* stat_l = (float*)(bars + 20);      // [4][128] per-column-block row sums

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int i0 = blockIdx.x * T5_BQ;
  const int n_tiles = (m + T5_BK - 1) / T5_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(k_full + i), 1);
      mbar_init(smem_u32(k_empty + i), 1);
      mbar_init(smem_u32(v_full + i), 1);
      mbar_init(smem_u32(v_empty + i), 1);
    }
    for (int i = 0; i < T1_SBUF; ++i) {
      mbar_init(smem_u32(s_full + i), 1);
      mbar_init(smem_u32(s_empty + i), 32 * T5_SM_WARPS);
    }
    mbar_init(smem_u32(p_full), 32 * T5_SM_WARPS);
    mbar_init(smem_u32(p_empty), 1);
    mbar_init(smem_u32(o_full), 1);
    mbar_init(smem_u32(q_full), 1);
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
    // ===== TMA producer: Q once, then the K ring =====
    if (lane == 0) {
      mbar_expect_tx(smem_u32(q_full), T5_Q_BYTES);
      for (int part = 0; part < 2; ++part)      // hi, lo
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(smem_u32(q_s + (part * 2 + kb) * T5_QBOX), &map_q, smem_u32(q_full), part * 64 + kb * 32, (int)(base + i0));
      int ks = 0;
      uint32_t kph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = t * T5_BK;
        mbar_wait(smem_u32(k_empty + ks), kph ^ 1);
        mbar_expect_tx(smem_u32(k_full + ks), T5_STAGE);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part)
            tma_load_2d(smem_u32(k_s + ks * T5_STAGE + (kb * 2 + part) * T5_KBOX), &map_k, smem_u32(k_full + ks), part * 64 + kb * 32,
                        (int)(base + j0));
        if (++ks == 2) { ks = 0; kph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ===== second TMA producer: the V^T ring =====
    if (lane == 0) {
      int vs = 0;
      uint32_t vph = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = t * T5_BK;
        mbar_wait(smem_u32(v_empty + vs), vph ^ 1);
        mbar_expect_tx(smem_u32(v_full + vs), T5_STAGE);
        for (int kb = 0; kb < 2; ++kb)
          for (int part = 0; part < 2; ++part)
            tma_load_2d(smem_u32(v_s + vs * T5_STAGE + (kb * 2 + part) * T5_KBOX), &map_v, smem_u32(v_full + vs), j0 + kb * 32,
                        chunk * 128 + part * 64);
        if (++vs == 2) { vs = 0; vph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, elected lane per instruction) =====
    mbar_wait(smem_u32(q_full), 0);
    tcgen05_fence_after();
    const uint32_t q_addr = smem_u32(q_s);
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0, pf = 0;
    // S(t) -> ring slot t % 3: Q_hi.K_hi + Q_hi.K_lo + Q_lo.K_hi, three N = 64 MMAs per k-step into the same 64 columns
    auto issue_s = [&](int t) {
      const int buf = t % T1_SBUF;
      mbar_wait(smem_u32(s_empty + buf), (uint32_t)(((t / T1_SBUF) & 1) ^ 1));
      mbar_wait(smem_u32(k_full + ks), kph);
      tcgen05_fence_after();
      const uint32_t kaddr = smem_u32(k_s + ks * T5_STAGE);
      const uint32_t d = tmem_base + buf * 64;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t bh = umma_desc_sw128(kaddr + kb * 2 * T5_KBOX + k * 32);
          const uint64_t bl = umma_desc_sw128(kaddr + (kb * 2 + 1) * T5_KBOX + k * 32);
          const uint64_t ah = umma_desc_sw128(q_addr + kb * T5_QBOX + k * 32);
          const uint64_t al = umma_desc_sw128(q_addr + (2 + kb) * T5_QBOX + k * 32);
          t5_mma_ss(d, ah, bh, T5_IDESC_N64, (kb | k) != 0);
          t5_mma_ss(d, ah, bl, T5_IDESC_N64, 1);
          t5_mma_ss(d, al, bh, T5_IDESC_N64, 1);
        }
      t5_commit(smem_u32(k_empty + ks));
      t5_commit(smem_u32(s_full + buf));
      if (++ks == 2) { ks = 0; kph ^= 1; }
    };
    issue_s(0);
    if (n_tiles > 1) issue_s(1);
    for (int t = 0; t < n_tiles; ++t) {
      if (t + 2 < n_tiles) issue_s(t + 2);
      // O += P(t) . V(t)
      mbar_wait(smem_u32(p_full), pf);
      pf ^= 1;
      mbar_wait(smem_u32(v_full + vs), vph);
      tcgen05_fence_after();
      const uint32_t vaddr = smem_u32(v_s + vs * T5_STAGE);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t b = umma_desc_sw128(vaddr + kb * 2 * T5_KBOX + k * 32);
          const uint32_t acc = ((t % T5_FLUSH) | kb | k) != 0;   // restart after every drain
          t5_mma_ts(tmem_base + T5_O_COL0, tmem_base + T5_P_COL0 + kb * 32 + k * 8, b, T5_IDESC_N128, acc);
          t5_mma_ts(tmem_base + T5_O_COL0 + 128, tmem_base + T5_P_COL0 + 64 + kb * 32 + k * 8, b, T5_IDESC_N64, acc);
        }
      t5_commit(smem_u32(v_empty + vs));
      t5_commit(smem_u32(p_empty));
      if (++vs == 2) { vs = 0; vph ^= 1; }
    }
    t5_commit(smem_u32(o_full));
    __syncwarp();
  } else if (warp >= 4) {
    // ===== softmax / epilogue warps: thread = query row (TMEM lane), 16-column block cq of the 64 key columns =====
    const int w = warp - 4;
    const int qd = w & 3, cq = w >> 2;
    const int row = qd * 32 + lane;
    const int grow = i0 + row;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const float sl2 = scale * 1.4426950408889634f;
    // bound of the row's scores in the log2 domain (the SAME expression attn_finish_kernel uses for the log-sum-exp)
    const float b2 = grow < m ? sl2 * __ldg(qn + base + grow) * __uint_as_float(__ldg(kmax_bits + chunk)) * 1.000002f : 0.f;
    const float inv_keep = DROP ? 1.f / (1.f - drop_p) : 1.f;
    const uint32_t drop_thr = DROP ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
    const uint32_t hrow = attn_drop_pre(seed, chunk) ^ ((uint32_t)grow * ATTN_DROP_CI);
    float lrun = 0.f;
    uint32_t pe = 0;
    float s[16];
    auto drain_o = [&](bool first) {
      tmem_drain16<3>(tmem_base + lane_base + T5_O_COL0 + cq * 16, ctx + (base + grow) * T5_D + cq * 16, 1.f, first, grow < m);
    };
    auto prob_tile = [&](int t, auto last_c) {
      constexpr bool LAST = decltype(last_c)::value;
      const bool drain = t > 0 && t % T5_FLUSH == 0;
      if (drain) {   // tiles [t - T5_FLUSH, t) are complete in O once the P.V MMAs of tile t-1 have retired
        mbar_wait(smem_u32(p_empty), pe);
        pe ^= 1;
        tcgen05_fence_after();
        drain_o(t == T5_FLUSH);
        tcgen05_fence_before();
      }
      const int buf = t % T1_SBUF;
      mbar_wait(smem_u32(s_full + buf), (uint32_t)((t / T1_SBUF) & 1));
      tcgen05_fence_after();
      t5_ld16(tmem_base + lane_base + buf * 64 + cq * 16, s);
      t5_ld_wait();
      tcgen05_fence_before();
      mbar_arrive(smem_u32(s_empty + buf));
      const int j0 = t * T5_BK + cq * 16;
      const int lim = m - j0;
      [[maybe_unused]] const uint32_t hcol = (uint32_t)j0 * ATTN_DROP_CJ;
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float p = t5_ex2(fmaf(s[e], sl2, -b2));
        if (LAST) p = e < lim ? p : 0.f;
        lrun += p;                                   // the normaliser sums the UN-dropped probabilities
        if constexpr (DROP) p = (attn_drop_mix(hrow ^ (hcol + (uint32_t)e * ATTN_DROP_CJ)) >= drop_thr) ? p * inv_keep : 0.f;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(p));
        hi[e] = u;
        lo[e] = __float_as_uint(p - __uint_as_float(u));
      }
      if (t > 0 && !drain) {   // the previous P must have been consumed by its P.V MMAs
        mbar_wait(smem_u32(p_empty), pe);
        pe ^= 1;
      }
      tcgen05_fence_after();
      t5_st16(tmem_base + lane_base + T5_P_COL0 + cq * 16, hi);
      t5_st16(tmem_base + lane_base + T5_P_COL0 + 64 + cq * 16, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tcgen05_fence_before();
      mbar_arrive(smem_u32(p_full));
    };
    for (int t = 0; t + 1 < n_tiles; ++t) prob_tile(t, std::false_type{});
    prob_tile(n_tiles - 1, std::true_type{});
    // row sums: combine the four column blocks
    stat_l[cq * 128 + row] = lrun;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * T5_SM_WARPS) : "memory");
    if (cq == 0 && grow < m) lsum[base + grow] = (stat_l[row] + stat_l[128 + row]) + (stat_l[256 + row] + stat_l[384 + row]);
    // epilogue: the tiles since the last drain
    mbar_wait(smem_u32(o_full), 0);
    tcgen05_fence_after();
    drain_o(n_tiles <= T5_FLUSH);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

static unsigned long long g_t5_attr = 0;

int64_t attn_t5_workspace_bytes(int m) {
  const long long mp = ((long long)m + 63) / 64 * 64;
  // q_hl, k_hl [4M,128] | V^T planes [4][128][Mp] | qn [4M] | lsum [4M] | kmax[4] + flag
  return (4ll * m * 128 * 2 + 4ll * 128 * mp + 8ll * m + 64) * 4 + 1024;
}

int launch_attn_fwd_t5(const float* q, const float* k, const float* v, int m, float scale, float drop_p, uint64_t seed, float* ctx,
                       float* lse, void* workspace, cudaStream_t st) {
  const int mp = (m + 63) / 64 * 64;
  float* q_hl = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float* k_hl = q_hl + 4ll * m * 128;
  float* vt = k_hl + 4ll * m * 128;
  float* qn = vt + 4ll * 128 * mp;
  float* lsum = qn + 4ll * m;
  unsigned int* kmax_bits = (unsigned int*)(lsum + 4ll * m);   // [4] + flag at [8]
  int* flag = (int*)(kmax_bits + 8);
  static const int two_pass = getenv("SCAN_B200_ATTN_2PASS") ? atoi(getenv("SCAN_B200_ATTN_2PASS")) : 0;   // tests: force the fallback kernel
  SCAN_CUDA_CHECK(cudaMemsetAsync(kmax_bits, 0, 16 * sizeof(int), st));
  attn_prep_kernel<<<2 * sm_count(), 256, 0, st>>>(q, k, v, m, mp, q_hl, k_hl, vt);
  SCAN_LAUNCH_CHECK("attn_prep_kernel");
  attn_norm_kernel<<<(unsigned)ceil_div(4ll * m, 256), 256, 0, st>>>(q, k, m, qn, kmax_bits);
  SCAN_LAUNCH_CHECK("attn_norm_kernel");
  CUtensorMap mq, mk, mv;
  int rc = make_rowmajor_map(&mq, q_hl, 4ull * m, 128, T5_BQ);
  if (rc) return rc;
  rc = make_rowmajor_map(&mk, k_hl, 4ull * m, 128, T5_BK);
  if (rc) return rc;
  rc = make_rowmajor_map(&mv, vt, 4ull * 128, (uint64_t)mp, 64);
  if (rc) return rc;
  if (first_use_on_device(&g_t5_attr)) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_t5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_t5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd1_t5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd1_t5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM));
  }
  dim3 grid((m + T5_BQ - 1) / T5_BQ, 4);
  if (!two_pass) {
    if (drop_p > 0.f)
      attn_fwd1_t5_kernel<true><<<grid, T5_THREADS, T5_SMEM, st>>>(mq, mk, mv, m, scale, drop_p, seed, qn, kmax_bits, ctx, lsum);
    else
      attn_fwd1_t5_kernel<false><<<grid, T5_THREADS, T5_SMEM, st>>>(mq, mk, mv, m, scale, drop_p, seed, qn, kmax_bits, ctx, lsum);
    SCAN_LAUNCH_CHECK("attn_fwd1_t5_kernel");
    attn_finish_kernel<<<(unsigned)ceil_div(4ll * m * 16, 256), 256, 0, st>>>(ctx, lsum, qn, kmax_bits, m, scale * 1.4426950408889634f, lse, flag);
    SCAN_LAUNCH_CHECK("attn_finish_kernel");
  }
  // two-pass kernel: exits at once unless the single-pass kernel flagged an underflowed row (or the test hook forces it)
  const int* run_flag = two_pass ? nullptr : flag;
  if (drop_p > 0.f)
    attn_fwd_t5_kernel<true><<<grid, T5_THREADS, T5_SMEM, st>>>(mq, mk, mv, m, scale, drop_p, seed, ctx, lse, run_flag);
  else
    attn_fwd_t5_kernel<false><<<grid, T5_THREADS, T5_SMEM, st>>>(mq, mk, mv, m, scale, drop_p, seed, ctx, lse, run_flag);
  SCAN_LAUNCH_CHECK("attn_fwd_t5_kernel");
  return SCAN_OK;
}

}  // namespace scan
