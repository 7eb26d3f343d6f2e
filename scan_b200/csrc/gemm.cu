// Dense layers of the graph-aggregation stage on the 5th-gen tensor cores: the projections linear_q / linear_k / linear_v /
// linear_final of MultiHeadAttention (layers/transformer.py:43-49, 61-88), the residual LayerNorm (:49, :88) and the node
// classifier proto_cls(relu(proto_cls_hidden(.))) with its cross-entropy (modeling/rpn/fcos/condgraph.py:186-188, 400-402),
// forward and backward.  torch routes these through cuBLAS fp32 SIMT kernels (allow_tf32 is off for matmul by default);
// here they are ONE templated tcgen05 kernel:
//
//     C[m, n] = sum_k A[m, k] * B[n, k]            A [M, K] and B [N, K] both K-major (row-major, K contiguous)
//
//   * 128 x BN output tile per CTA (BN = 16 / 128 / 256), fp32 accumulator in tensor memory, K streamed in 32-column
//     (128-byte) TMA boxes with the 128-byte swizzle through an mbarrier ring;
//   * 3xTF32 error compensation: the tensor core reads an fp32 word as tf32 by IGNORING the low 13 mantissa bits, so the
//     raw tile IS the `hi` operand; converter warps compute lo = rna_tf32(x - trunc_tf32(x)) (exact difference, 11 more
//     bits) into a second buffer, element-wise and therefore swizzle-agnostic, and the MMA warp issues
//     hi.hi + hi.lo + lo.hi into the same accumulator (a.b error ~ 2^-21 |a||b|: fp32-level);
//   * optional split over K (gridDim.z) with per-split partial tiles that a second kernel sums in a fixed order
//     (weight gradients reduce over the M nodes: each split stays below the ~64-k-step accumulation length at which the
//     truncating tensor-core accumulator becomes visible);
//   * A may be a stack of `a_part_rows`-row parts concatenated along K ([dq | dk | dv] without materialising the concat),
//     and the output may be scattered to column parts (q, k, v as three contiguous [M, 256] matrices);
//   * epilogues (one TMEM lane = one output row): bias (+ReLU) store; residual + dropout + LayerNorm over the full 256-wide
//     row; softmax cross-entropy over <= 16 logits with the logit gradient as a by-product.
// Weight-gradient and data-gradient GEMMs use the same kernel on transposed copies (transpose_pad_kernel); bias gradients
// are deterministic two-level column sums.
#include "tc_common.cuh"

namespace scan {

constexpr int GM_BM = 128;
constexpr int GM_BK = 32;                       // fp32 columns per TMA box row = 128 bytes
constexpr int GM_A_BYTES = GM_BM * GM_BK * 4;   // 16 KB
constexpr int GM_CONV_WARPS = 8;
constexpr int GM_THREADS = 256 + 32 * GM_CONV_WARPS;   // warps 0-3: TMA / MMA / TMEM alloc / idle, 4-7: epilogue, 8-15: converters

enum { GM_EPI_STORE = 0, GM_EPI_LN = 1, GM_EPI_CE = 2 };

template <int BN>
struct GemmCfg {
  static constexpr int B_BYTES = BN * GM_BK * 4;
  static constexpr int RAW_BYTES = GM_A_BYTES + B_BYTES;          // A | B raw (= hi)
  static constexpr int STAGE_BYTES = 2 * RAW_BYTES;               // raw | lo
  static constexpr int STAGES = BN >= 256 ? 2 : (BN >= 128 ? 3 : 4);
  static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + 256;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

struct GemmArgs {
  int M, N;              // valid output rows / columns
  int k_blocks;          // 32-column k-blocks per split
  int kb_per_part;       // k-blocks per A part (>= total k-blocks when A is a single matrix)
  int a_part_rows;       // row offset between consecutive A parts inside the A tensor map
  // GM_EPI_STORE
  float* c;              // output, row pitch ldc
  int ldc;
  int part_cols;         // output columns per part (>= N: a single part)
  long long c_part_stride;   // elements between output parts
  long long c_split_stride;  // elements between split-K partial outputs
  const float* bias;     // [N] or null
  int relu;
  int accumulate;        // out += result (the owning thread reads its own previous value)
  // GM_EPI_LN: y = LayerNorm(resid + dropout(acc + bias)) * gamma + beta; saves xhat and rstd
  const float* resid;
  const float* gamma;
  const float* beta;
  float eps;
  float drop_p;
  unsigned long long seed;
  float* xhat;
  float* rstd;
  // GM_EPI_CE: logits = acc + bias; loss_partials[cta] = sum_rows -log softmax(logits)[label - shift]; dlogits = softmax - onehot
  const long long* labels;
  int label_shift;
  float* dlogits;        // [M, 16]
  double* loss_partials;
};

__device__ __forceinline__ void gm_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void gm_commit(uint32_t bar) {
  if (elect_one_sync()) umma_commit(bar);
}
__device__ __forceinline__ void gm_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void gm_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void gm_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
      "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]),
      "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]),
      "f"(v[31])
      : "memory");
}

// Bernoulli source of the output dropout after linear_final (transformer.py:86): same counter hash as the attention mask,
// keyed on (seed, row, column); forward and backward recompute it.
__device__ __forceinline__ bool gm_keep(unsigned long long seed, uint32_t row, uint32_t col, uint32_t thr) {
  return attn_drop_hash(seed, 0x5EEDu, row, col) >= thr;
}
__host__ __device__ inline uint32_t gm_drop_threshold(float p) {
  const float t = p * 4294967296.f;
  return p <= 0.f ? 0u : (t >= 4294967295.f ? 4294967295u : (uint32_t)t);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GM_THREADS, 1)
    gemm3x_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmArgs g) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  uint64_t* bars = (uint64_t*)(stages + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                          // TMA -> converters
  uint64_t* ready_bar = bars + Cfg::STAGES;           // converters -> MMA
  uint64_t* empty_bar = bars + 2 * Cfg::STAGES;       // MMA -> TMA
  uint64_t* acc_full = bars + 3 * Cfg::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  __shared__ double red[4];

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * GM_BM, n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * g.k_blocks;
  constexpr uint32_t IDESC = umma_idesc_tf32(GM_BM, BN);

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(ready_bar + i), 32 * GM_CONV_WARPS);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    mbar_init(smem_u32(acc_full), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < g.k_blocks; ++i) {
        const int kb = kb0 + i;
        const int part = kb / g.kb_per_part;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        mbar_expect_tx(smem_u32(full_bar + stage), Cfg::RAW_BYTES);
        uint8_t* dst = stages + stage * Cfg::STAGE_BYTES;
        tma_load_2d(smem_u32(dst), &map_a, smem_u32(full_bar + stage), (kb - part * g.kb_per_part) * GM_BK, part * g.a_part_rows + m0);
        tma_load_2d(smem_u32(dst + GM_A_BYTES), &map_b, smem_u32(full_bar + stage), kb * GM_BK, n0);
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: converged warp, one elected lane per instruction =====
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < g.k_blocks; ++i) {
      mbar_wait(smem_u32(ready_bar + stage), phase);
      tcgen05_fence_after();
      const uint32_t a_hi = smem_u32(stages + stage * Cfg::STAGE_BYTES);
      const uint32_t b_hi = a_hi + GM_A_BYTES;
      const uint32_t a_lo = a_hi + Cfg::RAW_BYTES;
      const uint32_t b_lo = a_lo + GM_A_BYTES;
#pragma unroll
      for (int k = 0; k < GM_BK / 8; ++k) {
        const uint64_t da = umma_desc_sw128(a_hi + k * 32), db = umma_desc_sw128(b_hi + k * 32);
        gm_mma(tmem_base, da, db, IDESC, (i | k) != 0);
        gm_mma(tmem_base, da, umma_desc_sw128(b_lo + k * 32), IDESC, 1);
        gm_mma(tmem_base, umma_desc_sw128(a_lo + k * 32), db, IDESC, 1);
      }
      gm_commit(smem_u32(empty_bar + stage));
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    gm_commit(smem_u32(acc_full));
    __syncwarp();
  } else if (warp >= 8) {
    // ===== converters: lo = rna_tf32(x - trunc_tf32(x)) of the landed A and B boxes =====
    const int t = threadIdx.x - 256;
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < g.k_blocks; ++i) {
      mbar_wait(smem_u32(full_bar + stage), phase);
      const float4* raw = reinterpret_cast<const float4*>(stages + stage * Cfg::STAGE_BYTES);
      float4* lo = reinterpret_cast<float4*>(stages + stage * Cfg::STAGE_BYTES + Cfg::RAW_BYTES);
#pragma unroll 4
      for (int j = t; j < Cfg::RAW_BYTES / 16; j += 32 * GM_CONV_WARPS) {
        const float4 v = raw[j];
        float4 l;
        uint32_t u;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
        l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
        l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(l.x)); l.x = __uint_as_float(u);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(l.y)); l.y = __uint_as_float(u);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(l.z)); l.z = __uint_as_float(u);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(l.w)); l.w = __uint_as_float(u);
        lo[j] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core (async proxy) reads
      mbar_arrive(smem_u32(ready_bar + stage));
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp (4 + q) owns TMEM lanes [32 q, 32 q + 32): thread = output row =====
    const int q = warp - 4;
    const int row = m0 + q * 32 + lane;
    const bool valid = row < g.M;
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(smem_u32(acc_full), 0);
    tcgen05_fence_after();
    if constexpr (EPI == GM_EPI_STORE) {
      const int part = n0 / g.part_cols;
      float* out = g.c + (long long)blockIdx.z * g.c_split_stride + (long long)part * g.c_part_stride + (long long)row * g.ldc +
                   (n0 - part * g.part_cols);
      if constexpr (BN >= 32) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          gm_ld32(tl + c * 32, v);
          if (!valid) continue;
          const int col = n0 + c * 32;
          if (col + 32 <= g.N) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              float4 o = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
              if (g.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + e));
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
              if (g.accumulate) {
                const float4 p = *reinterpret_cast<const float4*>(out + c * 32 + e);
                o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
              }
              *reinterpret_cast<float4*>(out + c * 32 + e) = o;
            }
          } else {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (col + e < g.N) {
                float o = v[e] + (g.bias ? __ldg(g.bias + col + e) : 0.f);
                o = g.relu ? fmaxf(o, 0.f) : o;
                out[c * 32 + e] = g.accumulate ? out[c * 32 + e] + o : o;
              }
          }
        }
      } else {
        float v[16];
        gm_ld16(tl, v);
        if (valid) {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (n0 + e < g.N) {
              float o = v[e] + (g.bias ? __ldg(g.bias + n0 + e) : 0.f);
              out[e] = g.relu ? fmaxf(o, 0.f) : o;
            }
        }
      }
    } else if constexpr (EPI == GM_EPI_LN) {
      // full 256-wide row in this lane's TMEM columns.  Pass 1 forms v = resid + dropout(acc + bias) ONCE (residual row read with
      // 128-bit loads, 1 KB per thread), accumulates the mean and writes v back into the accumulator columns (tcgen05.st);
      // passes 2 and 3 (variance, normalise) read TMEM only.
      static_assert(EPI != GM_EPI_LN || BN == 256, "LayerNorm epilogue needs the whole row in one tile");
      const float4* xr4 = reinterpret_cast<const float4*>(g.resid + (long long)(valid ? row : 0) * 256);
      const uint32_t thr = gm_drop_threshold(g.drop_p);
      const float inv_keep = g.drop_p > 0.f ? 1.f / (1.f - g.drop_p) : 1.f;
      float sum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[32];
        gm_ld32(tl + c * 32, v);
#pragma unroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + c * 32) + e4);
          const float4 x = __ldg(xr4 + c * 8 + e4);
          const float bb[4] = {b.x, b.y, b.z, b.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int e = e4 * 4 + j;
            float o = v[e] + bb[j];
            if (g.drop_p > 0.f) o = gm_keep(g.seed, (uint32_t)row, (uint32_t)(c * 32 + e), thr) ? o * inv_keep : 0.f;
            v[e] = o + xx[j];
            sum += v[e];
          }
        }
        gm_st32(tl + c * 32, v);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      const float mean = sum * (1.f / 256.f);
      float sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[32];
        gm_ld32(tl + c * 32, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float d = v[e] - mean;
          sq = fmaf(d, d, sq);
        }
      }
      const float rs = rsqrtf(sq * (1.f / 256.f) + g.eps);
      if (valid) g.rstd[row] = rs;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[32];
        gm_ld32(tl + c * 32, v);
        if (!valid) continue;
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 xh, y;
          xh.x = (v[e] - mean) * rs;
          xh.y = (v[e + 1] - mean) * rs;
          xh.z = (v[e + 2] - mean) * rs;
          xh.w = (v[e + 3] - mean) * rs;
          const float4 ga = __ldg(reinterpret_cast<const float4*>(g.gamma + c * 32 + e));
          const float4 be = __ldg(reinterpret_cast<const float4*>(g.beta + c * 32 + e));
          y.x = fmaf(xh.x, ga.x, be.x); y.y = fmaf(xh.y, ga.y, be.y); y.z = fmaf(xh.z, ga.z, be.z); y.w = fmaf(xh.w, ga.w, be.w);
          *reinterpret_cast<float4*>(g.xhat + (long long)row * 256 + c * 32 + e) = xh;
          *reinterpret_cast<float4*>(g.c + (long long)row * 256 + c * 32 + e) = y;
        }
      }
    } else {  // GM_EPI_CE
      float v[16];
      gm_ld16(tl, v);
      double loss = 0.0;
      if (valid) {
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          v[e] = e < g.N ? v[e] + __ldg(g.bias + e) : -INFINITY;
          mx = fmaxf(mx, v[e]);
        }
        float s = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          v[e] = e < g.N ? expf(v[e] - mx) : 0.f;
          s += v[e];
        }
        const int lab = (int)(g.labels[row] - g.label_shift);
        const float inv = 1.f / s;
        float pl = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p = v[e] * inv;
          if (e == lab) pl = p;
          v[e] = p - (e == lab ? 1.f : 0.f);
        }
        loss = -(double)logf(pl);
        float4* d = reinterpret_cast<float4*>(g.dlogits + (long long)row * 16);
        d[0] = make_float4(v[0], v[1], v[2], v[3]);
        d[1] = make_float4(v[4], v[5], v[6], v[7]);
        d[2] = make_float4(v[8], v[9], v[10], v[11]);
        d[3] = make_float4(v[12], v[13], v[14], v[15]);
      }
      loss = warp_sum_d(loss);
      if (lane == 0) red[q] = loss;
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (q == 0 && lane == 0) g.loss_partials[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
  }
}

// dst[c, r] = src[r, c] for r < n_rows, zero for n_rows <= r < ld_dst (the pad keeps the TMA row pitch a multiple of 16 bytes
// and lets a split-K tail read zeros); 32 x 32 tiles through padded shared memory, both sides coalesced.
__global__ void __launch_bounds__(256) transpose_pad_kernel(const float* __restrict__ src, int n_rows, int n_cols, int ld_src,
                                                            float* __restrict__ dst, int ld_dst) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < n_rows && c < n_cols) ? __ldg(src + (long long)r * ld_src + c) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (c < n_cols && r < ld_dst) dst[(long long)c * ld_dst + r] = tile[tx][i];
  }
}

// out[i] = sum_s partial[s][i] (fixed order), i < n
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, long long n, float* __restrict__ out) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float4 acc = *reinterpret_cast<const float4*>(partial + i4);
  for (int s = 1; s < splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (long long)s * n + i4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(out + i4) = acc;
}

// out[r] = sum_{c < n_valid} xt[r, c]: bias gradients as row sums of the transposed gradient copies; one warp per row, fixed
// lane pattern + shuffle tree (deterministic)
__global__ void __launch_bounds__(256) rowsum_kernel(const float* __restrict__ xt, int n_rows, int n_valid, int ld, float* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  float s = 0.f;
  for (int c = lane; c < n_valid; c += 32) s += __ldg(xt + (long long)row * ld + c);
  s = warp_sum(s);
  if (lane == 0) out[row] = s;
}
// out[c] = sum_i partial[i][c] (fixed order)
__global__ void __launch_bounds__(256) partial_sum_kernel(const float* __restrict__ partial, int parts, int n_cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  float s = 0.f;
  for (int i = 0; i < parts; ++i) s += partial[(long long)i * n_cols + c];
  out[c] = s;
}

// ---------------------------------------------------------------------------- LayerNorm backward (+ output dropout, residual)
// y = xhat * gamma + beta, v = resid + drop(lin): d_v = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = dy * gamma;
// d_resid = d_v, d_lin = drop'(d_v).  One warp per row; per-CTA partial sums of d_gamma / d_beta (fixed-order reduce afterwards).
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xhat, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, int m, float drop_p, unsigned long long seed,
                                                     float* __restrict__ d_resid, float* __restrict__ d_lin, float* __restrict__ part_gb) {
  __shared__ float sg[8][256], sb[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t thr = gm_drop_threshold(drop_p);
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float ag[8], ab[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ag[e] = ab[e] = 0.f;
  float gam[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) gam[e] = __ldg(gamma + lane * 8 + e);
  for (int row = blockIdx.x * 8 + warp; row < m; row += gridDim.x * 8) {
    const float4* dy4 = reinterpret_cast<const float4*>(dy + (long long)row * 256 + lane * 8);
    const float4* xh4 = reinterpret_cast<const float4*>(xhat + (long long)row * 256 + lane * 8);
    const float4 a0 = __ldg(dy4), a1 = __ldg(dy4 + 1), b0 = __ldg(xh4), b1 = __ldg(xh4 + 1);
    const float d[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float x[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float gg[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gg[e] = d[e] * gam[e];
      s1 += gg[e];
      s2 = fmaf(gg[e], x[e], s2);
      ag[e] = fmaf(d[e], x[e], ag[e]);
      ab[e] += d[e];
    }
    s1 = warp_sum(s1) * (1.f / 256.f);
    s2 = warp_sum(s2) * (1.f / 256.f);
    const float rs = __ldg(rstd + row);
    float dv[8], dl[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      dv[e] = rs * (gg[e] - s1 - x[e] * s2);
      dl[e] = dv[e];
      if (drop_p > 0.f) dl[e] = gm_keep(seed, (uint32_t)row, (uint32_t)(lane * 8 + e), thr) ? dv[e] * inv_keep : 0.f;
    }
    float4* o1 = reinterpret_cast<float4*>(d_resid + (long long)row * 256 + lane * 8);
    float4* o2 = reinterpret_cast<float4*>(d_lin + (long long)row * 256 + lane * 8);
    o1[0] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    o1[1] = make_float4(dv[4], dv[5], dv[6], dv[7]);
    o2[0] = make_float4(dl[0], dl[1], dl[2], dl[3]);
    o2[1] = make_float4(dl[4], dl[5], dl[6], dl[7]);
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sg[warp][lane * 8 + e] = ag[e];
    sb[warp][lane * 8 + e] = ab[e];
  }
  __syncthreads();
  const int c = threadIdx.x;
  float tg = 0.f, tb = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    tg += sg[w][c];
    tb += sb[w][c];
  }
  part_gb[(long long)blockIdx.x * 512 + c] = tg;
  part_gb[(long long)blockIdx.x * 512 + 256 + c] = tb;
}

// ---------------------------------------------------------------------------- classifier backward (the K <= 16 side)
// g = *d_loss * loss_scale (d(total)/d(mean CE) * lambda / M):
//   d_hidden[m, j] = (hidden[m, j] > 0) * g * sum_k dlogits[m, k] W2[k, j]         (written in place of nothing: dense [M, H])
//   part_w2[cta][k][j] = sum_{m in cta} dlogits[m, k] hidden[m, j]                   (x g in the final reduce)
// Thread = hidden column j (H = 512: two columns per thread), 64-row tiles, dlogits tile broadcast from shared memory.
constexpr int CLS_ROWS = 64;
__global__ void __launch_bounds__(256) cls_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ hidden, const float* __restrict__ w2,
                                                      int m, int h, int k, float loss_scale, const float* __restrict__ d_loss,
                                                      float* __restrict__ d_hidden, float* __restrict__ part_w2) {
  __shared__ float4 dz[CLS_ROWS][4];
  const float gscale = loss_scale * __ldg(d_loss);
  for (int j = threadIdx.x; j < h; j += 256) {    // h <= 512: at most two columns per thread, handled one after the other
    float w[16], acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      w[e] = e < k ? __ldg(w2 + (long long)e * h + j) : 0.f;
      acc[e] = 0.f;
    }
    for (int r0 = blockIdx.x * CLS_ROWS; r0 < m; r0 += gridDim.x * CLS_ROWS) {
      __syncthreads();
      {
        const int r = r0 + (threadIdx.x >> 2), part = threadIdx.x & 3;
        dz[threadIdx.x >> 2][part] = r < m ? __ldg(reinterpret_cast<const float4*>(dlogits + (long long)r * 16) + part) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      const int rows = min(CLS_ROWS, m - r0);
      for (int i = 0; i < rows; ++i) {
        const float hv = __ldg(hidden + (long long)(r0 + i) * h + j);
        const float4 z0 = dz[i][0], z1 = dz[i][1], z2 = dz[i][2], z3 = dz[i][3];
        const float z[16] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w, z2.x, z2.y, z2.z, z2.w, z3.x, z3.y, z3.z, z3.w};
        float dh = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          dh = fmaf(z[e], w[e], dh);
          acc[e] = fmaf(z[e], hv, acc[e]);
        }
        d_hidden[(long long)(r0 + i) * h + j] = hv > 0.f ? dh * gscale : 0.f;
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) part_w2[((long long)blockIdx.x * 16 + e) * h + j] = acc[e];
  }
}
// d_w2[k, j] = g * sum_cta part_w2[cta][k][j];  d_b2[k] = g * sum_m dlogits[m, k]  (one block per class row k; fixed order)
__global__ void __launch_bounds__(256) cls_bwd_reduce_kernel(const float* __restrict__ part_w2, int parts, const float* __restrict__ dlogits, int m,
                                                             int h, float loss_scale, const float* __restrict__ d_loss,
                                                             float* __restrict__ d_w2, float* __restrict__ d_b2) {
  __shared__ float red[256];
  const int k = blockIdx.x;
  const float gscale = loss_scale * __ldg(d_loss);
  for (int j = threadIdx.x; j < h; j += 256) {
    float s = 0.f;
    for (int p = 0; p < parts; ++p) s += part_w2[((long long)p * 16 + k) * h + j];
    d_w2[(long long)k * h + j] = s * gscale;
  }
  float s = 0.f;
  for (int r = threadIdx.x; r < m; r += 256) s += dlogits[(long long)r * 16 + k];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) d_b2[k] = red[0] * gscale;
}

// sum of the per-CTA double partials -> fp32 scalar * scale (mean cross-entropy * lambda)
__global__ void __launch_bounds__(32) loss_finalize_kernel(const double* __restrict__ partials, int n, double scale, float* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += partials[i];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) *out = (float)(s * scale);
}

// d_nodes[m, :] (+)= d_mean[label[m] - shift, :] / max(count, 1): backward of the per-class means (condgraph.py:395-398)
__global__ void __launch_bounds__(256) class_mean_bwd_kernel(const float* __restrict__ d_mean, const float* __restrict__ packed, const long long* __restrict__ labels,
                                                             int m, int channels, int label_shift, float* __restrict__ d_nodes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = channels / 4;
  if (i >= (long long)m * c4) return;
  const int row = (int)(i / c4), c = (int)(i % c4) * 4;
  const int cls = (int)(labels[row] - label_shift);
  const float cnt = fmaxf(__ldg(packed + (long long)cls * (channels + 1) + channels), 1.f);
  const float4 d = __ldg(reinterpret_cast<const float4*>(d_mean + (long long)cls * channels + c));
  const float inv = 1.f / cnt;
  *reinterpret_cast<float4*>(d_nodes + (long long)row * channels + c) = make_float4(d.x * inv, d.y * inv, d.z * inv, d.w * inv);
}

// ---------------------------------------------------------------------------- generic helpers of the per-class GCN (a6)
// out[r, c] (+)= act(sum_s partial[s][r][c] + bias[c]) for c < n (fixed order over s)
__global__ void __launch_bounds__(256) splitk_reduce_act_kernel(const float* __restrict__ partial, int splits, int m, int n, long long ld_part,
                                                                const float* __restrict__ bias, int relu, int accumulate, float* __restrict__ out,
                                                                long long ldc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)m * n) return;
  const int r = (int)(i / n), c = (int)(i % n);
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += partial[((long long)s * m + r) * ld_part + c];
  if (bias) acc += __ldg(bias + c);
  if (relu) acc = fmaxf(acc, 0.f);
  float* o = out + (long long)r * ldc + c;
  *o = accumulate ? *o + acc : acc;
}

// in-place softmax over the first n_cols entries of every row (block per row); columns n_cols .. ld-1 are zeroed
__global__ void __launch_bounds__(256) rows_softmax_kernel(float* __restrict__ x, int n_cols, long long ld) {
  __shared__ float red[8];
  __shared__ float bc;
  float* row = x + (long long)blockIdx.x * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < n_cols; c += 256) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]);
    bc = v;
  }
  __syncthreads();
  mx = bc;
  float sum = 0.f;
  for (int c = threadIdx.x; c < n_cols; c += 256) {
    const float e = expf(row[c] - mx);
    row[c] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i];
    bc = 1.f / v;
  }
  __syncthreads();
  const float inv = bc;
  for (int c = threadIdx.x; c < (int)ld; c += 256) row[c] = c < n_cols ? row[c] * inv : 0.f;
}

// y[r, :] = x[r, :] / max(|x[r, :]|, eps) (sim_matrix, condgraph.py:35-43); 256 channels, one warp per row
__global__ void __launch_bounds__(256) rows_l2norm_kernel(const float* __restrict__ x, int m, float eps, float* __restrict__ y) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  const float4* x4 = reinterpret_cast<const float4*>(x + (long long)row * 256 + lane * 8);
  const float4 a = __ldg(x4), b = __ldg(x4 + 1);
  float q = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
  q = warp_sum(q);
  const float inv = 1.f / fmaxf(sqrtf(q), eps);
  float4* y4 = reinterpret_cast<float4*>(y + (long long)row * 256 + lane * 8);
  y4[0] = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  y4[1] = make_float4(b.x * inv, b.y * inv, b.z * inv, b.w * inv);
}

// GCN output activation (condgraph.py:274-281) over rows of 256 channels, one warp per row.
// mode 0 NO, 1 relu, 2 sigmoid, 3 tanh, 4 softmax(dim=-1).  a = act(z); y = a (+ shortcut row).
__global__ void __launch_bounds__(256) gcn_act_fwd_kernel(const float* __restrict__ z, const float* __restrict__ shortcut, int m, int mode,
                                                          float* __restrict__ a_out, float* __restrict__ y) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  float v[8];
  {
    const float4* z4 = reinterpret_cast<const float4*>(z + (long long)row * 256 + lane * 8);
    const float4 p = __ldg(z4), q = __ldg(z4 + 1);
    v[0] = p.x; v[1] = p.y; v[2] = p.z; v[3] = p.w; v[4] = q.x; v[5] = q.y; v[6] = q.z; v[7] = q.w;
  }
  if (mode == 4) {
    float mx = v[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) mx = fmaxf(mx, v[e]);
    mx = warp_max(mx);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { v[e] = expf(v[e] - mx); s += v[e]; }
    s = 1.f / warp_sum(s);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= s;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      v[e] = mode == 1 ? fmaxf(v[e], 0.f) : mode == 2 ? 1.f / (1.f + expf(-v[e])) : mode == 3 ? tanhf(v[e]) : v[e];
  }
  float4* a4 = reinterpret_cast<float4*>(a_out + (long long)row * 256 + lane * 8);
  a4[0] = make_float4(v[0], v[1], v[2], v[3]);
  a4[1] = make_float4(v[4], v[5], v[6], v[7]);
  if (shortcut) {
    const float4* s4 = reinterpret_cast<const float4*>(shortcut + (long long)row * 256 + lane * 8);
    const float4 p = __ldg(s4), q = __ldg(s4 + 1);
    v[0] += p.x; v[1] += p.y; v[2] += p.z; v[3] += p.w; v[4] += q.x; v[5] += q.y; v[6] += q.z; v[7] += q.w;
  }
  float4* y4 = reinterpret_cast<float4*>(y + (long long)row * 256 + lane * 8);
  y4[0] = make_float4(v[0], v[1], v[2], v[3]);
  y4[1] = make_float4(v[4], v[5], v[6], v[7]);
}
// dz = dy * act'(z) expressed through a = act(z)
__global__ void __launch_bounds__(256) gcn_act_bwd_kernel(const float* __restrict__ a, const float* __restrict__ dy, int m, int mode,
                                                          float* __restrict__ dz) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= m) return;
  float av[8], g[8];
  {
    const float4* a4 = reinterpret_cast<const float4*>(a + (long long)row * 256 + lane * 8);
    const float4* g4 = reinterpret_cast<const float4*>(dy + (long long)row * 256 + lane * 8);
    const float4 p = __ldg(a4), q = __ldg(a4 + 1), r = __ldg(g4), t = __ldg(g4 + 1);
    av[0] = p.x; av[1] = p.y; av[2] = p.z; av[3] = p.w; av[4] = q.x; av[5] = q.y; av[6] = q.z; av[7] = q.w;
    g[0] = r.x; g[1] = r.y; g[2] = r.z; g[3] = r.w; g[4] = t.x; g[5] = t.y; g[6] = t.z; g[7] = t.w;
  }
  if (mode == 4) {
    float dot = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) dot = fmaf(g[e], av[e], dot);
    dot = warp_sum(dot);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] = av[e] * (g[e] - dot);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      g[e] = mode == 1 ? (av[e] > 0.f ? g[e] : 0.f) : mode == 2 ? g[e] * av[e] * (1.f - av[e]) : mode == 3 ? g[e] * (1.f - av[e] * av[e]) : g[e];
  }
  float4* o = reinterpret_cast<float4*>(dz + (long long)row * 256 + lane * 8);
  o[0] = make_float4(g[0], g[1], g[2], g[3]);
  o[1] = make_float4(g[4], g[5], g[6], g[7]);
}

// ---------------------------------------------------------------------------- host side
static int make_map_pitch(CUtensorMap* m, const float* base, uint64_t n_rows, uint64_t n_cols, uint64_t pitch_floats, uint32_t box_rows) {
  EncodeTiledFn enc;
  int rc = get_tensormap_encoder(&enc);
  if (rc) return rc;
  if (((uintptr_t)base & 15) || (pitch_floats % 4)) return SCAN_EINVAL;
  cuuint64_t dims[2] = {n_cols, n_rows};
  cuuint64_t strides[1] = {pitch_floats * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed (gemm)");
    return SCAN_ECUDA;
  }
  return SCAN_OK;
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmArgs& g, int splits, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr))
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(gemm3x_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  dim3 grid((unsigned)ceil_div(g.M, GM_BM), (unsigned)ceil_div(g.N, BN), (unsigned)splits);
  gemm3x_kernel<BN, EPI><<<grid, GM_THREADS, Cfg::SMEM, st>>>(ma, mb, g);
  SCAN_LAUNCH_CHECK("gemm3x_kernel");
  return SCAN_OK;
}

static GemmArgs base_args(int M, int N, int K) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.M = M;
  g.N = N;
  g.k_blocks = K / GM_BK;
  g.kb_per_part = 1 << 28;
  g.part_cols = 1 << 30;
  return g;
}

static int transpose_pad(const float* src, int n_rows, int n_cols, int ld_src, float* dst, int ld_dst, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(ld_dst, 32), (unsigned)ceil_div(n_cols, 32));
  transpose_pad_kernel<<<grid, 256, 0, st>>>(src, n_rows, n_cols, ld_src, dst, ld_dst);
  SCAN_LAUNCH_CHECK("transpose_pad_kernel");
  return SCAN_OK;
}

constexpr int GM_SPLIT_K = 512;   // reduction rows per split of a weight-gradient GEMM: 64 k-steps

static inline int pad32(int m) { return (m + 31) / 32 * 32; }
static inline char* align256(char* p) { return (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255); }

// d_w [n_out, n_in] = dy^T [n_out, m] . x [m, n_in], both operands given transposed-and-padded ([*, mp] row-major)
static int wgrad(const float* dyt, const float* xt, int n_out, int n_in, int m, int mp, float* partial, float* d_w, cudaStream_t st) {
  const int splits = (int)ceil_div(m, GM_SPLIT_K);
  CUtensorMap ma, mb;
  int rc = make_map_pitch(&ma, dyt, (uint64_t)n_out, (uint64_t)mp, (uint64_t)mp, GM_BM);
  if (rc) return rc;
  rc = make_map_pitch(&mb, xt, (uint64_t)n_in, (uint64_t)mp, (uint64_t)mp, 128);
  if (rc) return rc;
  GemmArgs g = base_args(n_out, n_in, GM_SPLIT_K);
  g.c = splits > 1 ? partial : d_w;
  g.ldc = n_in;
  g.c_split_stride = (long long)n_out * n_in;
  rc = launch_gemm<128, GM_EPI_STORE>(ma, mb, g, splits, st);
  if (rc) return rc;
  if (splits > 1) {
    const long long n = (long long)n_out * n_in;
    splitk_reduce_kernel<<<(unsigned)ceil_div(n / 4, 256), 256, 0, st>>>(partial, splits, n, d_w);
    SCAN_LAUNCH_CHECK("splitk_reduce_kernel");
  }
  return SCAN_OK;
}

static int rowsum(const float* xt, int n_rows, int n_valid, int ld, float* out, cudaStream_t st) {
  rowsum_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, st>>>(xt, n_rows, n_valid, ld, out);
  SCAN_LAUNCH_CHECK("rowsum_kernel");
  return SCAN_OK;
}

// bump allocator over the caller's workspace
struct Bump {
  char* p;
  char* end;
  Bump(void* base, int64_t bytes) : p(align256((char*)base)), end((char*)base + bytes) {}
  float* take(long long floats) {
    float* r = (float*)p;
    p = align256(p + floats * 4);
    return r;
  }
  bool ok() const { return p <= end; }
};

}  // namespace scan

using namespace scan;

// ---------------------------------------------------------------------------- C ABI
extern "C" int64_t scan_graph_workspace_bytes(int32_t m) {
  // upper bound over scan_attn_out_ln_bwd / scan_qkv_bwd / scan_node_cls_fwd / scan_node_cls_bwd
  const long long mp = pad32(m > 0 ? m : 1);
  const long long splits = ceil_div(mp, GM_SPLIT_K);
  const long long transposed = (long long)(768 + 256 + 512 + 256) * mp;        // gradient^T and input^T copies
  const long long dense = (long long)mp * 512;                                   // d_hidden / d_lin
  const long long weights_t = 768ll * 256;                                       // transposed weight copy
  const long long partials = splits * 768 * 256 + 2ll * sm_count() * 512 + 2ll * sm_count() * 16 * 512 + 4096;
  return (transposed + dense + weights_t + partials) * 4 + 64 * 256;
}

// q|k|v [3, M, 256] = x [M, 256] . w_qkv [768, 256]^T + b_qkv     (linear_q / linear_k / linear_v, transformer.py:61-63)
extern "C" int scan_qkv_fwd(const float* x, const float* w_qkv, const float* b_qkv, int32_t m, float* qkv, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!x || !w_qkv || !b_qkv || !qkv || m < 0) return SCAN_EINVAL;
  CUtensorMap ma, mb;
  int rc = make_map_pitch(&ma, x, (uint64_t)m, 256, 256, GM_BM);
  if (rc) return rc;
  rc = make_map_pitch(&mb, w_qkv, 768, 256, 256, 128);
  if (rc) return rc;
  GemmArgs g = base_args(m, 768, 256);
  g.c = qkv;
  g.ldc = 256;
  g.part_cols = 256;
  g.c_part_stride = (long long)m * 256;
  g.bias = b_qkv;
  return launch_gemm<128, GM_EPI_STORE>(ma, mb, g, 1, (cudaStream_t)stream);
}

// d_x [M,256] += d_qkv . w_qkv ; d_w_qkv [768,256] = d_qkv^T . x ; d_b_qkv [768] = column sums of d_qkv
extern "C" int scan_qkv_bwd(const float* d_qkv, const float* x, const float* w_qkv, int32_t m, int32_t accumulate_dx, float* d_x,
                            float* d_w_qkv, float* d_b_qkv, void* workspace, int64_t workspace_bytes, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!d_qkv || !x || !w_qkv || !d_x || !d_w_qkv || !d_b_qkv || !workspace || m < 0) return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int mp = pad32(m);
  Bump ws(workspace, workspace_bytes);
  float* dyt = ws.take(768ll * mp);       // [768, mp]: (d_q | d_k | d_v)^T
  float* xt = ws.take(256ll * mp);        // x^T
  float* wt = ws.take(768ll * 256);       // w_qkv^T [256, 768]
  float* part = ws.take(ceil_div(m, GM_SPLIT_K) * 768 * 256);
  if (!ws.ok()) return SCAN_ECAPACITY;
  int rc;
  for (int p = 0; p < 3; ++p)
    if ((rc = transpose_pad(d_qkv + (long long)p * m * 256, m, 256, 256, dyt + (long long)p * 256 * mp, mp, st))) return rc;
  if ((rc = transpose_pad(x, m, 256, 256, xt, mp, st))) return rc;
  if ((rc = transpose_pad(w_qkv, 768, 256, 256, wt, 768, st))) return rc;
  if ((rc = rowsum(dyt, 768, m, mp, d_b_qkv, st))) return rc;
  if ((rc = wgrad(dyt, xt, 768, 256, m, mp, part, d_w_qkv, st))) return rc;
  CUtensorMap ma, mb;
  if ((rc = make_map_pitch(&ma, d_qkv, 3ull * m, 256, 256, GM_BM))) return rc;
  if ((rc = make_map_pitch(&mb, wt, 256, 768, 768, 128))) return rc;
  GemmArgs g = base_args(m, 256, 768);
  g.kb_per_part = 8;
  g.a_part_rows = m;
  g.c = d_x;
  g.ldc = 256;
  g.accumulate = accumulate_dx;
  return launch_gemm<128, GM_EPI_STORE>(ma, mb, g, 1, st);
}

// y = LayerNorm(x + dropout(ctx . w_f^T + b_f)) * gamma + beta   (transformer.py:84-88); saves xhat [M,256], rstd [M]
extern "C" int scan_attn_out_ln_fwd(const float* ctx, const float* w_f, const float* b_f, const float* x, const float* gamma,
                                    const float* beta, int32_t m, float eps, float drop_p, uint64_t seed, float* y, float* xhat,
                                    float* rstd, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!ctx || !w_f || !b_f || !x || !gamma || !beta || !y || !xhat || !rstd || m < 0 || drop_p < 0.f || drop_p >= 1.f) return SCAN_EINVAL;
  CUtensorMap ma, mb;
  int rc = make_map_pitch(&ma, ctx, (uint64_t)m, 256, 256, GM_BM);
  if (rc) return rc;
  rc = make_map_pitch(&mb, w_f, 256, 256, 256, 256);
  if (rc) return rc;
  GemmArgs g = base_args(m, 256, 256);
  g.c = y;
  g.bias = b_f;
  g.resid = x;
  g.gamma = gamma;
  g.beta = beta;
  g.eps = eps;
  g.drop_p = drop_p;
  g.seed = seed;
  g.xhat = xhat;
  g.rstd = rstd;
  return launch_gemm<256, GM_EPI_LN>(ma, mb, g, 1, (cudaStream_t)stream);
}

// d_y -> d_x (the residual branch; scan_qkv_bwd adds the projection branch), d_ctx, d_w_f, d_b_f, d_gamma_beta [512]
extern "C" int scan_attn_out_ln_bwd(const float* d_y, const float* xhat, const float* rstd, const float* gamma, const float* ctx,
                                    const float* w_f, int32_t m, float drop_p, uint64_t seed, float* d_x, float* d_ctx, float* d_w_f,
                                    float* d_b_f, float* d_gamma_beta, void* workspace, int64_t workspace_bytes, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!d_y || !xhat || !rstd || !gamma || !ctx || !w_f || !d_x || !d_ctx || !d_w_f || !d_b_f || !d_gamma_beta || !workspace || m < 0)
    return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int mp = pad32(m);
  const int ln_blocks = std::min((int)ceil_div(m, 8), 2 * sm_count());
  Bump ws(workspace, workspace_bytes);
  float* d_lin = ws.take((long long)m * 256);
  float* dlt = ws.take(256ll * mp);
  float* ctxt = ws.take(256ll * mp);
  float* wft = ws.take(256 * 256);
  float* part = ws.take(ceil_div(m, GM_SPLIT_K) * 256 * 256);
  float* part_gb = ws.take((long long)ln_blocks * 512);
  if (!ws.ok()) return SCAN_ECAPACITY;
  ln_bwd_kernel<<<ln_blocks, 256, 0, st>>>(d_y, xhat, rstd, gamma, m, drop_p, seed, d_x, d_lin, part_gb);
  SCAN_LAUNCH_CHECK("ln_bwd_kernel");
  partial_sum_kernel<<<2, 256, 0, st>>>(part_gb, ln_blocks, 512, d_gamma_beta);
  SCAN_LAUNCH_CHECK("partial_sum_kernel");
  int rc;
  if ((rc = transpose_pad(d_lin, m, 256, 256, dlt, mp, st))) return rc;
  if ((rc = transpose_pad(ctx, m, 256, 256, ctxt, mp, st))) return rc;
  if ((rc = transpose_pad(w_f, 256, 256, 256, wft, 256, st))) return rc;
  if ((rc = rowsum(dlt, 256, m, mp, d_b_f, st))) return rc;
  if ((rc = wgrad(dlt, ctxt, 256, 256, m, mp, part, d_w_f, st))) return rc;
  CUtensorMap ma, mb;
  if ((rc = make_map_pitch(&ma, d_lin, (uint64_t)m, 256, 256, GM_BM))) return rc;
  if ((rc = make_map_pitch(&mb, wft, 256, 256, 256, 128))) return rc;
  GemmArgs g = base_args(m, 256, 256);
  g.c = d_ctx;
  g.ldc = 256;
  return launch_gemm<128, GM_EPI_STORE>(ma, mb, g, 1, st);
}

// node classifier (condgraph.py:400-402): hidden = relu(nodes . w1^T + b1) [M,H]; logits = hidden . w2^T + b2 [M,K];
// loss = loss_weight * mean_m CE(logits, labels - shift); dlogits [M,16] = softmax - onehot (saved for the backward)
extern "C" int scan_node_cls_fwd(const float* nodes, const float* w1, const float* b1, const float* w2, const float* b2,
                                 const int64_t* labels, int32_t m, int32_t hidden_dim, int32_t num_classes, int32_t label_shift,
                                 float loss_weight, float* hidden, float* dlogits, float* loss, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  if (!nodes || !w1 || !b1 || !w2 || !b2 || !labels || !hidden || !dlogits || !loss || !workspace || m < 1) return SCAN_EINVAL;
  if (hidden_dim != 512 || num_classes < 1 || num_classes > 16) return SCAN_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (int)ceil_div(m, GM_BM);
  if (workspace_bytes < (int64_t)tiles * 8 + 256) return SCAN_ECAPACITY;
  double* partials = (double*)align256((char*)workspace);
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map_pitch(&ma, nodes, (uint64_t)m, 256, 256, GM_BM))) return rc;
  if ((rc = make_map_pitch(&mb, w1, 512, 256, 256, 128))) return rc;
  GemmArgs g = base_args(m, 512, 256);
  g.c = hidden;
  g.ldc = 512;
  g.bias = b1;
  g.relu = 1;
  if ((rc = launch_gemm<128, GM_EPI_STORE>(ma, mb, g, 1, st))) return rc;
  if ((rc = make_map_pitch(&ma, hidden, (uint64_t)m, 512, 512, GM_BM))) return rc;
  if ((rc = make_map_pitch(&mb, w2, (uint64_t)num_classes, 512, 512, 16))) return rc;
  GemmArgs c = base_args(m, num_classes, 512);
  c.bias = b2;
  c.labels = (const long long*)labels;
  c.label_shift = label_shift;
  c.dlogits = dlogits;
  c.loss_partials = partials;
  if ((rc = launch_gemm<16, GM_EPI_CE>(ma, mb, c, 1, st))) return rc;
  loss_finalize_kernel<<<1, 32, 0, st>>>(partials, tiles, (double)loss_weight / (double)m, loss);
  SCAN_LAUNCH_CHECK("loss_finalize_kernel");
  return SCAN_OK;
}

// backward of scan_node_cls_fwd: d_loss = device scalar d(total)/d(loss)
extern "C" int scan_node_cls_bwd(const float* dlogits, const float* hidden, const float* nodes, const float* w1, const float* w2,
                                 int32_t m, int32_t hidden_dim, int32_t num_classes, float loss_weight, const float* d_loss,
                                 float* d_nodes, float* d_w1, float* d_b1, float* d_w2, float* d_b2, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  if (!dlogits || !hidden || !nodes || !w1 || !w2 || !d_loss || !d_nodes || !d_w1 || !d_b1 || !d_w2 || !d_b2 || !workspace || m < 1)
    return SCAN_EINVAL;
  if (hidden_dim != 512 || num_classes < 1 || num_classes > 16) return SCAN_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  const int mp = pad32(m);
  const int parts = std::min((int)ceil_div(m, CLS_ROWS), 2 * sm_count());
  Bump ws(workspace, workspace_bytes);
  float* d_hidden = ws.take((long long)m * 512);
  float* dht = ws.take(512ll * mp);
  float* nt = ws.take(256ll * mp);
  float* w1t = ws.take(256 * 512);
  float* part = ws.take(ceil_div(m, GM_SPLIT_K) * 512 * 256);
  float* part_w2 = ws.take((long long)parts * 16 * 512);
  if (!ws.ok()) return SCAN_ECAPACITY;
  const float scale = loss_weight / (float)m;
  cls_bwd_kernel<<<parts, 256, 0, st>>>(dlogits, hidden, w2, m, 512, num_classes, scale, d_loss, d_hidden, part_w2);
  SCAN_LAUNCH_CHECK("cls_bwd_kernel");
  cls_bwd_reduce_kernel<<<num_classes, 256, 0, st>>>(part_w2, parts, dlogits, m, 512, scale, d_loss, d_w2, d_b2);
  SCAN_LAUNCH_CHECK("cls_bwd_reduce_kernel");
  int rc;
  if ((rc = transpose_pad(d_hidden, m, 512, 512, dht, mp, st))) return rc;
  if ((rc = transpose_pad(nodes, m, 256, 256, nt, mp, st))) return rc;
  if ((rc = transpose_pad(w1, 512, 256, 256, w1t, 512, st))) return rc;
  if ((rc = rowsum(dht, 512, m, mp, d_b1, st))) return rc;
  if ((rc = wgrad(dht, nt, 512, 256, m, mp, part, d_w1, st))) return rc;
  CUtensorMap ma, mb;
  if ((rc = make_map_pitch(&ma, d_hidden, (uint64_t)m, 512, 512, GM_BM))) return rc;
  if ((rc = make_map_pitch(&mb, w1t, 256, 512, 512, 128))) return rc;
  GemmArgs g = base_args(m, 256, 512);
  g.c = d_nodes;
  g.ldc = 256;
  return launch_gemm<128, GM_EPI_STORE>(ma, mb, g, 1, st);
}

// d_nodes[m, :] = d_mean[label[m] - shift, :] / max(count, 1)   (backward of the per-class node means, condgraph.py:395-398)
extern "C" int scan_class_mean_bwd(const float* d_mean, const float* packed, const int64_t* labels, int32_t m, int32_t channels,
                                   int32_t label_shift, float* d_nodes, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!d_mean || !packed || !labels || !d_nodes || m < 0 || channels % 4) return SCAN_EINVAL;
  const long long n = (long long)m * (channels / 4);
  class_mean_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_mean, packed, (const long long*)labels, m, channels,
                                                                                      label_shift, d_nodes);
  SCAN_LAUNCH_CHECK("class_mean_bwd_kernel");
  return SCAN_OK;
}

// ---------------------------------------------------------------------------- generic entry points (per-class GCN, a6)
static inline long long pad4ll(long long v) { return (v + 3) / 4 * 4; }

extern "C" int64_t scan_gemm_nt_workspace_bytes(int32_t m, int32_t n, int32_t k) {
  const long long splits = ceil_div(k > 0 ? k : 1, GM_SPLIT_K);
  return splits > 1 ? splits * (long long)m * pad4ll(n) * 4 + 1024 : 1024;
}

// c [m, n] (pitch ldc) (+)= act(a [m, k] (pitch lda) . b [n, k]^T (pitch ldb) + bias); reductions longer than 512 are split and
// summed in a fixed order (the tensor-core accumulator truncates)
extern "C" int scan_gemm_nt(const float* a, int64_t lda, const float* b, int64_t ldb, int32_t m, int32_t n, int32_t k, const float* bias,
                            int32_t relu, int32_t accumulate, float* c, int64_t ldc, void* workspace, int64_t workspace_bytes, void* stream) {
  if (m == 0 || n == 0) return SCAN_OK;
  if (!a || !b || !c || m < 0 || n < 0 || k < 1 || lda < k || ldb < k || ldc < n) return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map_pitch(&ma, a, (uint64_t)m, (uint64_t)k, (uint64_t)lda, GM_BM))) return rc;
  if ((rc = make_map_pitch(&mb, b, (uint64_t)n, (uint64_t)k, (uint64_t)ldb, 128))) return rc;
  const int splits = (int)ceil_div(k, GM_SPLIT_K);
  if (splits == 1) {
    GemmArgs g = base_args(m, n, (k + GM_BK - 1) / GM_BK * GM_BK);
    g.c = c;
    g.ldc = (int)ldc;
    g.bias = bias;
    g.relu = relu;
    g.accumulate = accumulate;
    if ((ldc % 4) || ((uintptr_t)c & 15)) return SCAN_EINVAL;
    return launch_gemm<128, GM_EPI_STORE>(ma, mb, g, 1, st);
  }
  if (!workspace || workspace_bytes < scan_gemm_nt_workspace_bytes(m, n, k)) return SCAN_ECAPACITY;
  float* part = (float*)align256((char*)workspace);
  const long long ldp = pad4ll(n);
  GemmArgs g = base_args(m, n, GM_SPLIT_K);
  g.c = part;
  g.ldc = (int)ldp;
  g.c_split_stride = (long long)m * ldp;
  if ((rc = launch_gemm<128, GM_EPI_STORE>(ma, mb, g, splits, st))) return rc;
  const long long total = (long long)m * n;
  splitk_reduce_act_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(part, splits, m, n, ldp, bias, relu, accumulate, c, ldc);
  SCAN_LAUNCH_CHECK("splitk_reduce_act_kernel");
  return SCAN_OK;
}

// dst [n_cols, ld_dst] = src [n_rows, n_cols]^T, zero in columns n_rows .. ld_dst-1
extern "C" int scan_transpose(const float* src, int32_t n_rows, int32_t n_cols, int64_t ld_src, float* dst, int64_t ld_dst, void* stream) {
  if (n_rows == 0 || n_cols == 0) return SCAN_OK;
  if (!src || !dst || n_rows < 0 || n_cols < 0 || ld_src < n_cols || ld_dst < n_rows) return SCAN_EINVAL;
  return transpose_pad(src, n_rows, n_cols, (int)ld_src, dst, (int)ld_dst, (cudaStream_t)stream);
}

extern "C" int64_t scan_linear_wgrad_workspace_bytes(int32_t m, int32_t n_out, int32_t n_in) {
  const long long mp = pad32(m > 0 ? m : 1);
  return ((long long)(n_out + n_in) * mp + ceil_div(mp, GM_SPLIT_K) * (long long)n_out * n_in + (long long)n_out * n_in + n_out) * 4 + 4096;
}

// d_w [n_out, n_in] (+)= dz [m, n_out]^T . x [m, n_in];  d_b [n_out] (+)= column sums of dz   (n_in % 4 == 0)
extern "C" int scan_linear_wgrad(const float* dz, const float* x, int32_t m, int32_t n_out, int32_t n_in, int32_t accumulate, float* d_w,
                                 float* d_b, void* workspace, int64_t workspace_bytes, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!dz || !x || !d_w || !workspace || m < 0 || n_out < 1 || n_in < 4 || n_in % 4 || n_out % 4) return SCAN_EINVAL;
  if (workspace_bytes < scan_linear_wgrad_workspace_bytes(m, n_out, n_in)) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  const int mp = pad32(m);
  Bump ws(workspace, workspace_bytes);
  float* dzt = ws.take((long long)n_out * mp);
  float* xt = ws.take((long long)n_in * mp);
  float* part = ws.take(ceil_div(m, GM_SPLIT_K) * (long long)n_out * n_in);
  float* tmp_w = ws.take((long long)n_out * n_in);
  float* tmp_b = ws.take(n_out);
  if (!ws.ok()) return SCAN_ECAPACITY;
  int rc;
  if ((rc = transpose_pad(dz, m, n_out, n_out, dzt, mp, st))) return rc;
  if ((rc = transpose_pad(x, m, n_in, n_in, xt, mp, st))) return rc;
  if (!accumulate) {
    if (d_b && (rc = rowsum(dzt, n_out, m, mp, d_b, st))) return rc;
    return wgrad(dzt, xt, n_out, n_in, m, mp, part, d_w, st);
  }
  if ((rc = wgrad(dzt, xt, n_out, n_in, m, mp, part, tmp_w, st))) return rc;
  const long long n = (long long)n_out * n_in;
  splitk_reduce_act_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(tmp_w, 1, 1, (int)n, n, nullptr, 0, 1, d_w, n);
  SCAN_LAUNCH_CHECK("splitk_reduce_act_kernel");
  if (d_b) {
    if ((rc = rowsum(dzt, n_out, m, mp, tmp_b, st))) return rc;
    splitk_reduce_act_kernel<<<(unsigned)ceil_div(n_out, 256), 256, 0, st>>>(tmp_b, 1, 1, n_out, n_out, nullptr, 0, 1, d_b, n_out);
    SCAN_LAUNCH_CHECK("splitk_reduce_act_kernel");
  }
  return SCAN_OK;
}

extern "C" int scan_rows_softmax(float* x, int32_t n_rows, int32_t n_cols, int64_t ld, void* stream) {
  if (n_rows == 0) return SCAN_OK;
  if (!x || n_rows < 0 || n_cols < 1 || ld < n_cols) return SCAN_EINVAL;
  rows_softmax_kernel<<<n_rows, 256, 0, (cudaStream_t)stream>>>(x, n_cols, ld);
  SCAN_LAUNCH_CHECK("rows_softmax_kernel");
  return SCAN_OK;
}

extern "C" int scan_rows_l2normalize(const float* x, int32_t m, float eps, float* y, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!x || !y || m < 0) return SCAN_EINVAL;
  rows_l2norm_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(x, m, eps, y);
  SCAN_LAUNCH_CHECK("rows_l2norm_kernel");
  return SCAN_OK;
}

extern "C" int scan_gcn_act_fwd(const float* z, const float* shortcut, int32_t m, int32_t mode, float* act_out, float* y, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!z || !act_out || !y || m < 0 || mode < 0 || mode > 4) return SCAN_EINVAL;
  gcn_act_fwd_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(z, shortcut, m, mode, act_out, y);
  SCAN_LAUNCH_CHECK("gcn_act_fwd_kernel");
  return SCAN_OK;
}

extern "C" int scan_gcn_act_bwd(const float* act_out, const float* dy, int32_t m, int32_t mode, float* dz, void* stream) {
  if (m == 0) return SCAN_OK;
  if (!act_out || !dy || !dz || m < 0 || mode < 0 || mode > 4) return SCAN_EINVAL;
  gcn_act_bwd_kernel<<<(unsigned)ceil_div(m, 8), 256, 0, (cudaStream_t)stream>>>(act_out, dy, m, mode, dz);
  SCAN_LAUNCH_CHECK("gcn_act_bwd_kernel");
  return SCAN_OK;
}
