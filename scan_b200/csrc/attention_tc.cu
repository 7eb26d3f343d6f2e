// K3a on the tensor cores: the same streaming attention as attention.cu (4 chunks of M 64-d sub-tokens, online
// softmax, M x M never materialised, backward recomputes P from the saved log-sum-exp), with every 64x64x64 tile product
// on mma.sync.m16n8k8 tf32 and 3xTF32 error compensation (x = hi + lo; lo*hi + hi*lo + hi*hi), which keeps fp32-level
// accuracy: the scores feed an exponential, so plain tf32 (1e-3 relative) is not acceptable for the north-star's 1e-3
// tolerance on the aggregated nodes.
// Reference: layers/transformer.py:5-34 under the .view of :66-68.
//
// Data movement (v2; v1 split every fragment in registers and was ALU/LDS-issue bound, no faster than the FFMA kernels):
//   * every operand tile is stored in shared memory in "row = output index, columns = reduction index" form (B operands
//     transposed while loading where needed), with the reduction index PERMUTED inside each group of 8
//     (k -> 2*(k%4) + k/4), so that the two values an mma fragment needs (k = t and k = t+4) are adjacent:
//     one 64-bit shared load per fragment half; leading dimension 72 floats makes those loads conflict-free;
//   * B operands are split into tf32 hi / lo ONCE per tile when they are written to shared memory (each element is
//     consumed by 4 warps x many k-steps); A fragments (4 values per 24 MMAs) are split in registers.
// Block = 4 warps; warp w owns rows [16w, 16w+16) of the 64-row output tile.
#include "common.cuh"

namespace scan {

constexpr int TA_D = 64;
constexpr int TA_T = 64;
constexpr int LDT = 72;                 // leading dimension of every tile (floats)
constexpr int TILE_F = TA_T * LDT;      // floats per tile plane


__device__ __forceinline__ int perm8(int k) { return (k & ~7) | (((k & 3) << 1) | ((k >> 2) & 1)); }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[nt] (16 x 8*NT warp tile) += A(16 x 64) . B(8*NT x 64)^T with
//   A[m][k] at A[m*LDT + perm8(k)]      (fp32, split in registers)            -- A points at the warp's first row
//   B[n][k] at Bh/Bl[n*LDT + perm8(k)]  (tf32 hi / lo planes, split at store time) -- B points at the warp's first column
// A_TRANS: A[m][k] is read from a [k][perm8(m)] tile instead (scalar loads; A then points at the tile origin and m0 is
//          the warp's first row).  B_SPLIT: B is a single fp32 plane (Bl unused) and is split in registers.
template <int NT, bool A_TRANS, bool B_SPLIT>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ A, const float* __restrict__ Bh, const float* __restrict__ Bl,
                                          float (&acc)[NT][4], int g, int t, int m0 = 0) {
#pragma unroll 2
  for (int ks = 0; ks < 8; ++ks) {
    const int k0 = ks * 8 + 2 * t;  // permuted position of (k = t, k = t + 4)
    float a0, a1, a2, a3;
    if (!A_TRANS) {
      const float2 a02 = *reinterpret_cast<const float2*>(A + g * LDT + k0);
      const float2 a13 = *reinterpret_cast<const float2*>(A + (g + 8) * LDT + k0);
      a0 = a02.x; a2 = a02.y; a1 = a13.x; a3 = a13.y;
    } else {
      const int pm0 = perm8(m0 + g), pm1 = perm8(m0 + g + 8);
      a0 = A[(ks * 8 + t) * LDT + pm0];
      a1 = A[(ks * 8 + t) * LDT + pm1];
      a2 = A[(ks * 8 + t + 4) * LDT + pm0];
      a3 = A[(ks * 8 + t + 4) * LDT + pm1];
    }
    uint32_t ah[4], al[4];
    split_tf32(a0, ah[0], al[0]);
    split_tf32(a1, ah[1], al[1]);
    split_tf32(a2, ah[2], al[2]);
    split_tf32(a3, ah[3], al[3]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      uint32_t bh0, bh1, bl0, bl1;
      const float2 bh = *reinterpret_cast<const float2*>(Bh + (nt * 8 + g) * LDT + k0);
      if (B_SPLIT) {
        split_tf32(bh.x, bh0, bl0);
        split_tf32(bh.y, bh1, bl1);
      } else {
        const float2 bl = *reinterpret_cast<const float2*>(Bl + (nt * 8 + g) * LDT + k0);
        bh0 = __float_as_uint(bh.x); bh1 = __float_as_uint(bh.y);
        bl0 = __float_as_uint(bl.x); bl1 = __float_as_uint(bl.y);
      }
      mma_tf32(acc[nt], al, bh0, bh1);
      mma_tf32(acc[nt], ah, bl0, bl1);
      mma_tf32(acc[nt], ah, bh0, bh1);
    }
  }
}

// global [n_rows, 64] rows r0.. -> smem tile, row-major, reduction index (the 64 columns) permuted.
// kSplit: write tf32 hi / lo planes (B operand); otherwise a single fp32 plane (A operand).
template <bool kSplit>
__device__ __forceinline__ void load_rowmajor(float* hi, float* lo, const float* __restrict__ src, long long r0, long long n_rows) {
  for (int i = threadIdx.x; i < TA_T * (TA_D / 4); i += blockDim.x) {
    const int r = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(src + (r0 + r) * TA_D) + c4);
    const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int p = r * LDT + perm8(c4 * 4 + e);
      if (kSplit) {
        uint32_t h, l;
        split_tf32(x[e], h, l);
        hi[p] = __uint_as_float(h);
        lo[p] = __uint_as_float(l);
      } else {
        hi[p] = x[e];
      }
    }
  }
}

// global [n_rows, 64] rows r0.. -> smem TRANSPOSED tile T[c][perm8(r)] (the reduction index is the global row)
template <bool kSplit>
__device__ __forceinline__ void load_transposed(float* hi, float* lo, const float* __restrict__ src, long long r0, long long n_rows) {
  for (int i = threadIdx.x; i < TA_T * (TA_D / 4); i += blockDim.x) {
    const int r = i & 63, c4 = i >> 6;  // consecutive threads take consecutive rows: conflict-free transposed stores
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(src + (r0 + r) * TA_D) + c4);
    const float x[4] = {v.x, v.y, v.z, v.w};
    const int pr = perm8(r);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int p = (c4 * 4 + e) * LDT + pr;
      if (kSplit) {
        uint32_t h, l;
        split_tf32(x[e], h, l);
        hi[p] = __uint_as_float(h);
        lo[p] = __uint_as_float(l);
      } else {
        hi[p] = x[e];
      }
    }
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---------------------------------------------------------------------------- forward
// planes: Q (A), P (A, per-warp rows), K hi/lo (B, [key][d]), V^T hi/lo (B, [d][key])
constexpr int TAF_SMEM = 6 * TILE_F * 4;

__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                          int m, float scale, float drop_p, uint64_t seed, float* __restrict__ ctx,
                                                          float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Ps = Qs + TILE_F;
  float* Kh = Ps + TILE_F;
  float* Kl = Kh + TILE_F;
  float* Vh = Kl + TILE_F;
  float* Vl = Vh + TILE_F;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int i0 = blockIdx.x * TA_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* qc = q + base * TA_D;
  const float* kc = k + base * TA_D;
  const float* vc = v + base * TA_D;
  load_rowmajor<false>(Qs, nullptr, qc, i0, m);
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nt][i] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t drop_thr = drop_p > 0.f ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
  const int row_a = i0 + warp * 16 + g, row_b = row_a + 8;
  float* Pw = Ps + warp * 16 * LDT;
  for (int j0 = 0; j0 < m; j0 += TA_T) {
    __syncthreads();
    load_rowmajor<true>(Kh, Kl, kc, j0, m);
    load_transposed<true>(Vh, Vl, vc, j0, m);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
    warp_gemm<8, false, false>(Qs + warp * 16 * LDT, Kh, Kl, s, g, t);  // S[i][j] = q_i . k_j
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = j0 + nt * 8 + 2 * t + (i & 1);
        s[nt][i] = (col < m) ? s[nt][i] * scale : -INFINITY;
        mx[i >> 1] = fmaxf(mx[i >> 1], s[nt][i]);
      }
    float alpha[2], sum[2] = {0.f, 0.f}, mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = quad_max(mx[r]);
      mnew[r] = fmaxf(mrow[r], mx[r]);
      alpha[r] = (mrow[r] == -INFINITY) ? 0.f : expf(mrow[r] - mnew[r]);
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i >> 1;
        float pv = (s[nt][i] == -INFINITY) ? 0.f : expf(s[nt][i] - mnew[r]);
        sum[r] += pv;
        if (drop_p > 0.f) {
          const uint32_t h = attn_drop_hash(seed, chunk, r ? row_b : row_a, j0 + nt * 8 + 2 * t + (i & 1));
          pv = (h >= drop_thr) ? pv * inv_keep : 0.f;
        }
        // P[row][key] is the A operand of P.V: key index permuted
        Pw[(g + 8 * r) * LDT + perm8(nt * 8 + 2 * t + (i & 1))] = pv;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      sum[r] = quad_sum(sum[r]);
      lrow[r] = lrow[r] * alpha[r] + sum[r];
      mrow[r] = mnew[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= alpha[0];
      o[nt][1] *= alpha[0];
      o[nt][2] *= alpha[1];
      o[nt][3] *= alpha[1];
    }
    __syncwarp();
    warp_gemm<8, false, false>(Pw, Vh, Vl, o, g, t);  // O[i][c] += P[i][j] V[j][c]   (B = V^T[c][j])
  }
  const float inv[2] = {1.f / lrow[0], 1.f / lrow[1]};
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (row_a < m) *reinterpret_cast<float2*>(ctx + (base + row_a) * TA_D + nt * 8 + 2 * t) = make_float2(o[nt][0] * inv[0], o[nt][1] * inv[0]);
    if (row_b < m) *reinterpret_cast<float2*>(ctx + (base + row_b) * TA_D + nt * 8 + 2 * t) = make_float2(o[nt][2] * inv[1], o[nt][3] * inv[1]);
  }
  if (t == 0) {
    if (row_a < m) lse[base + row_a] = mrow[0] + logf(lrow[0]);
    if (row_b < m) lse[base + row_b] = mrow[1] + logf(lrow[1]);
  }
}

// ---------------------------------------------------------------------------- backward
// One block (8 warps) per (key tile, chunk); warp = (row block wr = warp % 4, column half wc = warp / 4).  Per query tile:
//   S = Q K^T, dP = dO V^T              (A = Q / dO rows, B = K / V rows: [key][d], hi/lo planes, resident)
//   dV += P~^T dO, dK += dS^T Q         (A = P~^T / dS^T: [key][query], B = dO^T / Q^T: [d][query], fp32, split in registers)
//   dQ += dS K                          (A = dS read transposed from the [key][query] tile, B = K^T: [d][key], hi/lo, resident)
// 12 planes of 18 KB: Kh Kl Vh Vl KTh KTl | Qa Oa | QT OT | PT ST
constexpr int TAB_PLANES = 12;
constexpr int TAB_SMEM = TAB_PLANES * TILE_F * 4;

__global__ void __launch_bounds__(256) attn_bwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          const float* __restrict__ d_ctx, int m, float scale, float drop_p, uint64_t seed,
                                                          float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  float* Kh = sm;
  float* Kl = Kh + TILE_F;
  float* Vh = Kl + TILE_F;
  float* Vl = Vh + TILE_F;
  float* KTh = Vl + TILE_F;
  float* KTl = KTh + TILE_F;
  float* Qa = KTl + TILE_F;
  float* Oa = Qa + TILE_F;
  float* QT = Oa + TILE_F;
  float* OT = QT + TILE_F;
  float* PT = OT + TILE_F;
  float* ST = PT + TILE_F;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int j0 = blockIdx.x * TA_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = warp & 3, wc = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const float* qc = q + base * TA_D;
  const float* kc = k + base * TA_D;
  const float* vc = v + base * TA_D;
  const float* doc = d_ctx + base * TA_D;
  load_rowmajor<true>(Kh, Kl, kc, j0, m);
  load_rowmajor<true>(Vh, Vl, vc, j0, m);
  load_transposed<true>(KTh, KTl, kc, j0, m);
  float dkacc[4][4], dvacc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) dkacc[nt][i] = dvacc[nt][i] = 0.f;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t drop_thr = drop_p > 0.f ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
  const int cb = wc * 32;  // first column (of the 64) this warp produces in every GEMM
  for (int i0 = 0; i0 < m; i0 += TA_T) {
    __syncthreads();
    load_rowmajor<false>(Qa, nullptr, qc, i0, m);
    load_rowmajor<false>(Oa, nullptr, doc, i0, m);
    load_transposed<false>(QT, nullptr, qc, i0, m);
    load_transposed<false>(OT, nullptr, doc, i0, m);
    __syncthreads();
    {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = dp[nt][i] = 0.f;
      warp_gemm<4, false, false>(Qa + wr * 16 * LDT, Kh + cb * LDT, Kl + cb * LDT, s, g, t);   // S[i][j], keys cb..cb+31
      warp_gemm<4, false, false>(Oa + wr * 16 * LDT, Vh + cb * LDT, Vl + cb * LDT, dp, g, t);  // dP[i][j]
      const int row_a = i0 + wr * 16 + g, row_b = row_a + 8;
      const float l[2] = {row_a < m ? __ldg(lse + base + row_a) : 0.f, row_b < m ? __ldg(lse + base + row_b) : 0.f};
      const float dl[2] = {row_a < m ? __ldg(delta + base + row_a) : 0.f, row_b < m ? __ldg(delta + base + row_b) : 0.f};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i >> 1;
          const int row = r ? row_b : row_a;
          const int lq = wr * 16 + g + 8 * r;             // query index inside the tile
          const int lk = cb + nt * 8 + 2 * t + (i & 1);   // key index inside the tile
          const int col = j0 + lk;
          float p = 0.f, keep = 1.f;
          if (row < m && col < m) {
            p = expf(s[nt][i] * scale - l[r]);
            if (drop_p > 0.f) keep = (attn_drop_hash(seed, chunk, row, col) >= drop_thr) ? inv_keep : 0.f;
          }
          PT[lk * LDT + perm8(lq)] = p * keep;                        // P~^T  [key][query]
          ST[lk * LDT + perm8(lq)] = p * (dp[nt][i] * keep - dl[r]);  // dS^T  [key][query]
        }
      }
    }
    __syncthreads();
    // rows of these results are keys wr*16.., columns are d = cb..cb+31
    warp_gemm<4, false, true>(PT + wr * 16 * LDT, OT + cb * LDT, nullptr, dvacc, g, t);  // dV[j][c] += P~[i][j] dO[i][c]
    warp_gemm<4, false, true>(ST + wr * 16 * LDT, QT + cb * LDT, nullptr, dkacc, g, t);  // dK[j][c] += dS[i][j] Q[i][c]
    float dqt[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dqt[nt][i] = 0.f;
    warp_gemm<4, true, false>(ST, KTh + cb * LDT, KTl + cb * LDT, dqt, g, t, wr * 16);    // dQ[i][c] += dS[i][j] K[j][c]
    const int row_a = i0 + wr * 16 + g, row_b = row_a + 8;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int c = cb + nt * 8 + 2 * t;
      if (row_a < m) {
        atomicAdd(dq + (base + row_a) * TA_D + c, dqt[nt][0] * scale);
        atomicAdd(dq + (base + row_a) * TA_D + c + 1, dqt[nt][1] * scale);
      }
      if (row_b < m) {
        atomicAdd(dq + (base + row_b) * TA_D + c, dqt[nt][2] * scale);
        atomicAdd(dq + (base + row_b) * TA_D + c + 1, dqt[nt][3] * scale);
      }
    }
  }
  const int key_a = j0 + wr * 16 + g, key_b = key_a + 8;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = cb + nt * 8 + 2 * t;
    if (key_a < m) {
      *reinterpret_cast<float2*>(dk + (base + key_a) * TA_D + c) = make_float2(dkacc[nt][0] * scale, dkacc[nt][1] * scale);
      *reinterpret_cast<float2*>(dv + (base + key_a) * TA_D + c) = make_float2(dvacc[nt][0], dvacc[nt][1]);
    }
    if (key_b < m) {
      *reinterpret_cast<float2*>(dk + (base + key_b) * TA_D + c) = make_float2(dkacc[nt][2] * scale, dkacc[nt][3] * scale);
      *reinterpret_cast<float2*>(dv + (base + key_b) * TA_D + c) = make_float2(dvacc[nt][2], dvacc[nt][3]);
    }
  }
}

static int g_ta_attr = 0;
static int ta_attrs() {
  if (!g_ta_attr) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAF_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAB_SMEM));
    g_ta_attr = 1;
  }
  return SCAN_OK;
}

int launch_attn_fwd_tc(const float* q, const float* k, const float* v, int m, float scale, float drop_p, uint64_t seed, float* ctx,
                       float* lse, cudaStream_t st) {
  int rc = ta_attrs();
  if (rc) return rc;
  dim3 grid((m + TA_T - 1) / TA_T, 4);
  attn_fwd_tc_kernel<<<grid, 128, TAF_SMEM, st>>>(q, k, v, m, scale, drop_p, seed, ctx, lse);
  SCAN_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return SCAN_OK;
}

int launch_attn_bwd_tc(const float* q, const float* k, const float* v, const float* lse, const float* delta, const float* d_ctx, int m,
                       float scale, float drop_p, uint64_t seed, float* dq, float* dk, float* dv, cudaStream_t st) {
  int rc = ta_attrs();
  if (rc) return rc;
  dim3 grid((m + TA_T - 1) / TA_T, 4);
  attn_bwd_tc_kernel<<<grid, 256, TAB_SMEM, st>>>(q, k, v, lse, delta, d_ctx, m, scale, drop_p, seed, dq, dk, dv);
  SCAN_LAUNCH_CHECK("attn_bwd_tc_kernel");
  return SCAN_OK;
}

}  // namespace scan
