// K3a on the tensor cores: the same streaming attention as attention.cu (4 chunks of M 64-d sub-tokens, online
// softmax, M x M never materialised, backward recomputes P from the saved log-sum-exp), but every 64x64x64 tile product
// runs on mma.sync.m16n8k8 tf32 with 3xTF32 error compensation (x = hi + lo, hi*hi + lo*hi + hi*lo), which keeps
// fp32-level accuracy: the scores feed an exponential, so plain tf32 (1e-3 relative) is not acceptable for the
// north-star's 1e-3 tolerance on the aggregated nodes.
// Reference: layers/transformer.py:5-34 under the .view of :66-68.
//
// Block = 4 warps; warp w owns rows [16w, 16w+16) of the 64-row output tile.  Operands live in shared memory with
// padded leading dimensions (68 / 72 floats) chosen so that the mma fragment loads are bank-conflict free in the form
// each tile is used most.  A tcgen05 version is a round-2 item.
#include "common.cuh"

namespace scan {

constexpr int TA_D = 64;
constexpr int TA_T = 64;
constexpr int LD_A = 68;  // tiles read as row-major A / "n-major" B fragments: bank = 4g + t
constexpr int LD_B = 72;  // tiles read as k-major B / transposed A fragments: bank = 8t + g

__device__ __forceinline__ uint32_t ta_drop_hash(uint64_t seed, uint32_t chunk, uint32_t i, uint32_t j) {
  uint64_t x = seed ^ ((uint64_t)chunk << 60) ^ ((uint64_t)i << 30) ^ (uint64_t)j;
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  x ^= x >> 31;
  return (uint32_t)(x >> 32);
}

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

// acc[nt][.] (16 x 64 warp tile, nt = 8-column block) += A(16 x 64) * B(64 x 64)
//   A element (m, k): A_TRANS ? A[k*LDA + m] : A[m*LDA + k]        (A already offset to the warp's first row / column)
//   B element (k, n): B_NMAJOR ? B[n*LDB + k] : B[k*LDB + n]
template <int LDA, bool A_TRANS, int LDB, bool B_NMAJOR>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ A, const float* __restrict__ B, float (&acc)[8][4], int g, int t) {
#pragma unroll 2
  for (int ks = 0; ks < 8; ++ks) {
    const int k0 = ks * 8;
    float av[4];
    if (!A_TRANS) {
      av[0] = A[g * LDA + k0 + t];
      av[1] = A[(g + 8) * LDA + k0 + t];
      av[2] = A[g * LDA + k0 + t + 4];
      av[3] = A[(g + 8) * LDA + k0 + t + 4];
    } else {
      av[0] = A[(k0 + t) * LDA + g];
      av[1] = A[(k0 + t) * LDA + g + 8];
      av[2] = A[(k0 + t + 4) * LDA + g];
      av[3] = A[(k0 + t + 4) * LDA + g + 8];
    }
    uint32_t ah[4], al[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(av[i], ah[i], al[i]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float b0, b1;
      if (B_NMAJOR) {
        b0 = B[(nt * 8 + g) * LDB + k0 + t];
        b1 = B[(nt * 8 + g) * LDB + k0 + t + 4];
      } else {
        b0 = B[(k0 + t) * LDB + nt * 8 + g];
        b1 = B[(k0 + t + 4) * LDB + nt * 8 + g];
      }
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(b0, bh0, bl0);
      split_tf32(b1, bh1, bl1);
      mma_tf32(acc[nt], al, bh0, bh1);
      mma_tf32(acc[nt], ah, bl0, bl1);
      mma_tf32(acc[nt], ah, bh0, bh1);
    }
  }
}

template <int LD>
__device__ __forceinline__ void ta_load_tile(float* s, const float* __restrict__ gsrc, long long r0, long long n_rows) {
  for (int i = threadIdx.x; i < TA_T * (TA_D / 4); i += blockDim.x) {
    const int r = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(gsrc + (r0 + r) * TA_D) + c4);
    *reinterpret_cast<float4*>(s + r * LD + c4 * 4) = v;
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
constexpr int TAF_SMEM = (TA_T * LD_A * 3 + TA_T * LD_B) * 4;  // Q, K, P (LD_A) + V (LD_B)

__global__ void __launch_bounds__(128) attn_fwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                          int m, float scale, float drop_p, uint64_t seed, float* __restrict__ ctx,
                                                          float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Ks = Qs + TA_T * LD_A;
  float* Ps = Ks + TA_T * LD_A;
  float* Vs = Ps + TA_T * LD_A;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int i0 = blockIdx.x * TA_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* qc = q + base * TA_D;
  const float* kc = k + base * TA_D;
  const float* vc = v + base * TA_D;
  ta_load_tile<LD_A>(Qs, qc, i0, m);
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[nt][i] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t drop_thr = drop_p > 0.f ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
  const int row_a = i0 + warp * 16 + g, row_b = row_a + 8;
  float* Pw = Ps + warp * 16 * LD_A;
  for (int j0 = 0; j0 < m; j0 += TA_T) {
    __syncthreads();
    ta_load_tile<LD_A>(Ks, kc, j0, m);
    ta_load_tile<LD_B>(Vs, vc, j0, m);
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[nt][i] = 0.f;
    warp_gemm<LD_A, false, LD_A, true>(Qs + warp * 16 * LD_A, Ks, s, g, t);
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
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i >> 1;
        float pv = (s[nt][i] == -INFINITY) ? 0.f : expf(s[nt][i] - mnew[r]);
        sum[r] += pv;
        if (drop_p > 0.f) {
          const uint32_t h = ta_drop_hash(seed, chunk, r ? row_b : row_a, j0 + nt * 8 + 2 * t + (i & 1));
          pv = (h >= drop_thr) ? pv * inv_keep : 0.f;
        }
        p[i] = pv;
      }
      *reinterpret_cast<float2*>(Pw + g * LD_A + nt * 8 + 2 * t) = make_float2(p[0], p[1]);
      *reinterpret_cast<float2*>(Pw + (g + 8) * LD_A + nt * 8 + 2 * t) = make_float2(p[2], p[3]);
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
    warp_gemm<LD_A, false, LD_B, false>(Pw, Vs, o, g, t);
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
// smem: Qs, Os (dO), Ks, Vs with LD_A; Ps (P~), Ss (dS) with LD_B
constexpr int TAB_SMEM = (TA_T * LD_A * 4 + TA_T * LD_B * 2) * 4;

__global__ void __launch_bounds__(128) attn_bwd_tc_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                          const float* __restrict__ lse, const float* __restrict__ delta,
                                                          const float* __restrict__ d_ctx, int m, float scale, float drop_p, uint64_t seed,
                                                          float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Os = Qs + TA_T * LD_A;
  float* Ks = Os + TA_T * LD_A;
  float* Vs = Ks + TA_T * LD_A;
  float* Ps = Vs + TA_T * LD_A;
  float* Ss = Ps + TA_T * LD_B;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int j0 = blockIdx.x * TA_T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float* qc = q + base * TA_D;
  const float* kc = k + base * TA_D;
  const float* vc = v + base * TA_D;
  const float* doc = d_ctx + base * TA_D;
  ta_load_tile<LD_A>(Ks, kc, j0, m);
  ta_load_tile<LD_A>(Vs, vc, j0, m);
  float dkacc[8][4], dvacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int i = 0; i < 4; ++i) dkacc[nt][i] = dvacc[nt][i] = 0.f;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t drop_thr = drop_p > 0.f ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
  for (int i0 = 0; i0 < m; i0 += TA_T) {
    __syncthreads();
    ta_load_tile<LD_A>(Qs, qc, i0, m);
    ta_load_tile<LD_A>(Os, doc, i0, m);
    __syncthreads();
    {
      float s[8][4], dp[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) s[nt][i] = dp[nt][i] = 0.f;
      warp_gemm<LD_A, false, LD_A, true>(Qs + warp * 16 * LD_A, Ks, s, g, t);   // S[i][j] = q_i . k_j
      warp_gemm<LD_A, false, LD_A, true>(Os + warp * 16 * LD_A, Vs, dp, g, t);  // dP[i][j] = dO_i . v_j
      const int row_a = i0 + warp * 16 + g, row_b = row_a + 8;
      const float l[2] = {row_a < m ? __ldg(lse + base + row_a) : 0.f, row_b < m ? __ldg(lse + base + row_b) : 0.f};
      const float dl[2] = {row_a < m ? __ldg(delta + base + row_a) : 0.f, row_b < m ? __ldg(delta + base + row_b) : 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float pt[4], ds[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i >> 1;
          const int row = r ? row_b : row_a;
          const int col = j0 + nt * 8 + 2 * t + (i & 1);
          float p = 0.f, keep = 1.f;
          if (row < m && col < m) {
            p = expf(s[nt][i] * scale - l[r]);
            if (drop_p > 0.f) keep = (ta_drop_hash(seed, chunk, row, col) >= drop_thr) ? inv_keep : 0.f;
          }
          pt[i] = p * keep;
          ds[i] = p * (dp[nt][i] * keep - dl[r]);
        }
        const int lr = warp * 16 + g, lc = nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(Ps + lr * LD_B + lc) = make_float2(pt[0], pt[1]);
        *reinterpret_cast<float2*>(Ps + (lr + 8) * LD_B + lc) = make_float2(pt[2], pt[3]);
        *reinterpret_cast<float2*>(Ss + lr * LD_B + lc) = make_float2(ds[0], ds[1]);
        *reinterpret_cast<float2*>(Ss + (lr + 8) * LD_B + lc) = make_float2(ds[2], ds[3]);
      }
    }
    __syncthreads();
    // dV[j][c] += P~[i][j] dO[i][c] ; dK[j][c] += dS[i][j] Q[i][c]   (A transposed: rows of the result are keys)
    warp_gemm<LD_B, true, LD_A, false>(Ps + warp * 16, Os, dvacc, g, t);
    warp_gemm<LD_B, true, LD_A, false>(Ss + warp * 16, Qs, dkacc, g, t);
    // dQ[i][c] += dS[i][j] K[j][c]
    float dqt[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) dqt[nt][i] = 0.f;
    warp_gemm<LD_B, false, LD_A, false>(Ss + warp * 16 * LD_B, Ks, dqt, g, t);
    const int row_a = i0 + warp * 16 + g, row_b = row_a + 8;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int c = nt * 8 + 2 * t;
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
  const int key_a = j0 + warp * 16 + g, key_b = key_a + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * t;
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

int launch_attn_fwd_tc(const float* q, const float* k, const float* v, int m, float scale, float drop_p, uint64_t seed, float* ctx,
                       float* lse, cudaStream_t st) {
  if (!g_ta_attr) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAF_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAB_SMEM));
    g_ta_attr = 1;
  }
  dim3 grid((m + TA_T - 1) / TA_T, 4);
  attn_fwd_tc_kernel<<<grid, 128, TAF_SMEM, st>>>(q, k, v, m, scale, drop_p, seed, ctx, lse);
  SCAN_LAUNCH_CHECK("attn_fwd_tc_kernel");
  return SCAN_OK;
}

int launch_attn_bwd_tc(const float* q, const float* k, const float* v, const float* lse, const float* delta, const float* d_ctx, int m,
                       float scale, float drop_p, uint64_t seed, float* dq, float* dk, float* dv, cudaStream_t st) {
  if (!g_ta_attr) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAF_SMEM));
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TAB_SMEM));
    g_ta_attr = 1;
  }
  dim3 grid((m + TA_T - 1) / TA_T, 4);
  attn_bwd_tc_kernel<<<grid, 128, TAB_SMEM, st>>>(q, k, v, lse, delta, d_ctx, m, scale, drop_p, seed, dq, dk, dv);
  SCAN_LAUNCH_CHECK("attn_bwd_tc_kernel");
  return SCAN_OK;
}

}  // namespace scan
