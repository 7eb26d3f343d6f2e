// K3a: cross-image graph aggregation, global variant: node affinity (q k^T * scale), row softmax and
// aggregation (P v) in one streaming kernel -- the M x M affinity matrix is never materialised.
// Reference: layers/transformer.py:5-34 (dot_attention) inside MultiHeadAttention (:36-90), called at
// modeling/rpn/fcos/condgraph.py:390-393.  The reference's `.view(4, -1, 64)` on a [1, M, 256] projection
// makes 4 independent chunks of M 64-d sub-tokens; chunk b = rows [bM, (b+1)M) of the
// row-major [4M, 64] reinterpretation.  fp32 FFMA arithmetic, 64x64 tiles staged in shared memory,
// 4x4 register blocking, online softmax; backward recomputes P from the saved log-sum-exp.
#include <stdlib.h>

#include "common.cuh"

namespace scan {

// tcgen05 version (attention_t5.cu): used by scan_attn_fwd when the caller provides a workspace
int64_t attn_t5_workspace_bytes(int m);
int launch_attn_fwd_t5(const float* q, const float* k, const float* v, int m, float scale, float drop_p, uint64_t seed, float* ctx,
                       float* lse, void* workspace, cudaStream_t st);
int64_t attn_t5_bwd_workspace_bytes(int m);
int launch_attn_bwd_t5(const float* q, const float* k, const float* v, const float* lse, const float* delta, const float* d_ctx, int m,
                       float scale, float drop_p, uint64_t seed, float* dq, float* dk, float* dv, void* workspace, cudaStream_t st);
constexpr int AT_D = 64;   // sub-token width
constexpr int AT_T = 64;   // tile edge
constexpr int AT_LD = 68;  // padded leading dimension (floats): 272-byte rows, 16-byte aligned
constexpr int AT_TILE = AT_T * AT_LD;


// tile loader: rows [r0, r0+64) of a [n_rows, 64] matrix -> smem [64][AT_LD], zero beyond n_rows
__device__ __forceinline__ void load_tile(float* s, const float* __restrict__ g, long long r0, long long n_rows) {
  for (int i = threadIdx.x; i < AT_T * (AT_D / 4); i += blockDim.x) {
    const int r = i >> 4, c4 = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < n_rows) v = __ldg(reinterpret_cast<const float4*>(g + (r0 + r) * AT_D) + c4);
    *reinterpret_cast<float4*>(s + r * AT_LD + c4 * 4) = v;
  }
}

// C[ty*4+a][tx+16c] = sum_d A[ty*4+a][d] * B[tx+16c][d]
__device__ __forceinline__ void gemm_nt(const float* A, const float* B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 4
  for (int d = 0; d < AT_D; d += 4) {
    float4 av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(A + (ty * 4 + a) * AT_LD + d);
#pragma unroll
    for (int c = 0; c < 4; ++c) bv[c] = *reinterpret_cast<const float4*>(B + (tx + 16 * c) * AT_LD + d);
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[a][c] = fmaf(av[a].x, bv[c].x, acc[a][c]);
        acc[a][c] = fmaf(av[a].y, bv[c].y, acc[a][c]);
        acc[a][c] = fmaf(av[a].z, bv[c].z, acc[a][c]);
        acc[a][c] = fmaf(av[a].w, bv[c].w, acc[a][c]);
      }
  }
}

// acc[a][cc] += sum_j A[ty*4+a][j] * B[j][tx*4+cc]
__device__ __forceinline__ void gemm_nn_acc(const float* A, const float* B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 4
  for (int j = 0; j < AT_T; j += 4) {
    float4 av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(A + (ty * 4 + a) * AT_LD + j);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) bv[jj] = *reinterpret_cast<const float4*>(B + (j + jj) * AT_LD + tx * 4);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float pa[4] = {av[a].x, av[a].y, av[a].z, av[a].w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        acc[a][0] = fmaf(pa[jj], bv[jj].x, acc[a][0]);
        acc[a][1] = fmaf(pa[jj], bv[jj].y, acc[a][1]);
        acc[a][2] = fmaf(pa[jj], bv[jj].z, acc[a][2]);
        acc[a][3] = fmaf(pa[jj], bv[jj].w, acc[a][3]);
      }
    }
  }
}

// acc[a][cc] += sum_i A[i][ty*4+a] * B[i][tx*4+cc]
__device__ __forceinline__ void gemm_tn_acc(const float* A, const float* B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 8
  for (int i = 0; i < AT_T; ++i) {
    const float4 av = *reinterpret_cast<const float4*>(A + i * AT_LD + ty * 4);
    const float4 bv = *reinterpret_cast<const float4*>(B + i * AT_LD + tx * 4);
    const float pa[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      acc[a][0] = fmaf(pa[a], bv.x, acc[a][0]);
      acc[a][1] = fmaf(pa[a], bv.y, acc[a][1]);
      acc[a][2] = fmaf(pa[a], bv.z, acc[a][2]);
      acc[a][3] = fmaf(pa[a], bv.w, acc[a][3]);
    }
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(256) attn_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                       int m, float scale, float drop_p, uint64_t seed, float* __restrict__ ctx,
                                                       float* __restrict__ lse) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Ks = sm + AT_TILE;
  float* Vs = sm + 2 * AT_TILE;
  float* Ps = sm + 3 * AT_TILE;
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;  // first sub-token row of the chunk
  const int i0 = blockIdx.x * AT_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qc = q + base * AT_D;
  const float* kc = k + base * AT_D;
  const float* vc = v + base * AT_D;
  load_tile(Qs, qc, i0, m);
  float o[4][4];
  float mrow[4], lrow[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    mrow[a] = -INFINITY;
    lrow[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) o[a][c] = 0.f;
  }
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t drop_thr = drop_p > 0.f ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
  for (int j0 = 0; j0 < m; j0 += AT_T) {
    __syncthreads();  // previous iteration done with Ks / Vs / Ps (and Qs visible on the first one)
    load_tile(Ks, kc, j0, m);
    load_tile(Vs, vc, j0, m);
    __syncthreads();
    float s[4][4];
    gemm_nt(Qs, Ks, ty, tx, s);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = j0 + tx + 16 * c;
        s[a][c] = (col < m) ? s[a][c] * scale : -INFINITY;
        mx = fmaxf(mx, s[a][c]);
      }
      mx = group16_max(mx);
      const float mnew = fmaxf(mrow[a], mx);
      const float alpha = (mrow[a] == -INFINITY) ? 0.f : expf(mrow[a] - mnew);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float p = (s[a][c] == -INFINITY) ? 0.f : expf(s[a][c] - mnew);
        sum += p;
        if (drop_p > 0.f) {
          const uint32_t h = attn_drop_hash(seed, chunk, i0 + ty * 4 + a, j0 + tx + 16 * c);
          p = (h >= drop_thr) ? p * inv_keep : 0.f;
        }
        Ps[(ty * 4 + a) * AT_LD + tx + 16 * c] = p;
      }
      sum = group16_sum(sum);
      lrow[a] = lrow[a] * alpha + sum;
      mrow[a] = mnew;
#pragma unroll
      for (int c = 0; c < 4; ++c) o[a][c] *= alpha;
    }
    __syncthreads();
    gemm_nn_acc(Ps, Vs, ty, tx, o);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = i0 + ty * 4 + a;
    if (row < m) {
      const float inv = 1.f / lrow[a];
      float4 r = make_float4(o[a][0] * inv, o[a][1] * inv, o[a][2] * inv, o[a][3] * inv);
      *reinterpret_cast<float4*>(ctx + (base + row) * AT_D + tx * 4) = r;
      if (tx == 0) lse[base + row] = mrow[a] + logf(lrow[a]);
    }
  }
}

// ---------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(256) attn_delta_kernel(const float* __restrict__ ctx, const float* __restrict__ d_ctx, long long n_rows,
                                                         float* __restrict__ delta) {
  // one 16-lane group per sub-token row (64 floats = 16 float4)
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int l16 = threadIdx.x & 15;
  float s = 0.f;
  if (row < n_rows) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(ctx + row * AT_D) + l16);
    const float4 b = __ldg(reinterpret_cast<const float4*>(d_ctx + row * AT_D) + l16);
    s = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  s = group16_sum(s);
  if (row < n_rows && l16 == 0) delta[row] = s;
}

// one block per (key tile, chunk): dK, dV in registers, dQ by atomicAdd
__global__ void __launch_bounds__(256) attn_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                       const float* __restrict__ lse, const float* __restrict__ delta,
                                                       const float* __restrict__ d_ctx, int m, float scale, float drop_p, uint64_t seed,
                                                       float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
  extern __shared__ __align__(16) float sm[];
  float* Qs = sm;
  float* Os = sm + AT_TILE;      // dO tile
  float* Ks = sm + 2 * AT_TILE;
  float* Vs = sm + 3 * AT_TILE;
  float* Ps = sm + 4 * AT_TILE;  // dropped probabilities P~
  float* Ss = sm + 5 * AT_TILE;  // dS
  const int chunk = blockIdx.y;
  const long long base = (long long)chunk * m;
  const int j0 = blockIdx.x * AT_T;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const float* qc = q + base * AT_D;
  const float* kc = k + base * AT_D;
  const float* vc = v + base * AT_D;
  const float* doc = d_ctx + base * AT_D;
  load_tile(Ks, kc, j0, m);
  load_tile(Vs, vc, j0, m);
  float dkacc[4][4], dvacc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) dkacc[a][c] = dvacc[a][c] = 0.f;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t drop_thr = drop_p > 0.f ? (uint32_t)fminf(drop_p * 4294967296.f, 4294967295.f) : 0u;
  for (int i0 = 0; i0 < m; i0 += AT_T) {
    __syncthreads();
    load_tile(Qs, qc, i0, m);
    load_tile(Os, doc, i0, m);
    __syncthreads();
    float s[4][4], dp[4][4];
    gemm_nt(Qs, Ks, ty, tx, s);
    gemm_nt(Os, Vs, ty, tx, dp);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = i0 + ty * 4 + a;
      const float l = row < m ? __ldg(lse + base + row) : 0.f;
      const float dl = row < m ? __ldg(delta + base + row) : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int col = j0 + tx + 16 * c;
        float p = 0.f, keep = 1.f;
        if (row < m && col < m) {
          p = expf(s[a][c] * scale - l);
          if (drop_p > 0.f) keep = (attn_drop_hash(seed, chunk, row, col) >= drop_thr) ? inv_keep : 0.f;
        }
        const float pt = p * keep;
        Ps[(ty * 4 + a) * AT_LD + tx + 16 * c] = pt;
        Ss[(ty * 4 + a) * AT_LD + tx + 16 * c] = p * (dp[a][c] * keep - dl);
      }
    }
    __syncthreads();
    gemm_tn_acc(Ps, Os, ty, tx, dvacc);  // dV[j][c] += P~[i][j] dO[i][c]
    gemm_tn_acc(Ss, Qs, ty, tx, dkacc);  // dK[j][c] += dS[i][j] Q[i][c]
    float dqt[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) dqt[a][c] = 0.f;
    gemm_nn_acc(Ss, Ks, ty, tx, dqt);  // dQ[i][c] += dS[i][j] K[j][c]
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int row = i0 + ty * 4 + a;
      if (row < m) {
        float* dst = dq + (base + row) * AT_D + tx * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(dst + c, dqt[a][c] * scale);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int row = j0 + ty * 4 + a;
    if (row < m) {
      *reinterpret_cast<float4*>(dk + (base + row) * AT_D + tx * 4) =
          make_float4(dkacc[a][0] * scale, dkacc[a][1] * scale, dkacc[a][2] * scale, dkacc[a][3] * scale);
      *reinterpret_cast<float4*>(dv + (base + row) * AT_D + tx * 4) = make_float4(dvacc[a][0], dvacc[a][1], dvacc[a][2], dvacc[a][3]);
    }
  }
}

static unsigned long long g_attn_attr = 0;
static int set_attn_attrs() {
  if (!first_use_on_device(&g_attn_attr)) return SCAN_OK;
  SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * AT_TILE * 4));
  SCAN_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * AT_TILE * 4));
  return SCAN_OK;
}

}  // namespace scan

extern "C" int64_t scan_attn_workspace_bytes(int32_t m) { return m > 0 ? scan::attn_t5_workspace_bytes(m) : 0; }

extern "C" int64_t scan_attn_bwd_workspace_bytes(int32_t m) { return m > 0 ? scan::attn_t5_bwd_workspace_bytes(m) : 0; }

extern "C" int scan_attn_fwd(const float* q, const float* k, const float* v, int32_t m, float scale, float dropout_p, uint64_t seed,
                             float* ctx, float* lse, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace scan;
  if (m == 0) return SCAN_OK;
  if (!q || !k || !v || !ctx || !lse || m < 0 || dropout_p < 0.f || dropout_p >= 1.f) return SCAN_EINVAL;
  if (workspace) {  // tcgen05 path
    if (workspace_bytes < attn_t5_workspace_bytes(m)) return SCAN_ECAPACITY;
    return launch_attn_fwd_t5(q, k, v, m, scale, dropout_p, seed, ctx, lse, workspace, (cudaStream_t)stream);
  }
  int rc = set_attn_attrs();
  if (rc) return rc;
  dim3 grid((m + AT_T - 1) / AT_T, 4);
  attn_fwd_kernel<<<grid, 256, 4 * AT_TILE * 4, (cudaStream_t)stream>>>(q, k, v, m, scale, dropout_p, seed, ctx, lse);
  SCAN_LAUNCH_CHECK("attn_fwd_kernel");
  return SCAN_OK;
}

extern "C" int scan_attn_bwd(const float* q, const float* k, const float* v, const float* ctx, const float* lse, const float* d_ctx,
                             int32_t m, float scale, float dropout_p, uint64_t seed, float* dq, float* dk, float* dv,
                             float* delta_ws, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace scan;
  if (m == 0) return SCAN_OK;
  if (!q || !k || !v || !ctx || !lse || !d_ctx || !dq || !dk || !dv || !delta_ws || m < 0) return SCAN_EINVAL;
  int rc = set_attn_attrs();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n_rows = 4ll * m;
  attn_delta_kernel<<<(unsigned)ceil_div(n_rows * 16, 256), 256, 0, st>>>(ctx, d_ctx, n_rows, delta_ws);
  SCAN_LAUNCH_CHECK("attn_delta_kernel");
  if (workspace) {  // tcgen05 path: every gradient row is written exactly once, no zero-fill needed
    if (workspace_bytes < attn_t5_bwd_workspace_bytes(m)) return SCAN_ECAPACITY;
    return launch_attn_bwd_t5(q, k, v, lse, delta_ws, d_ctx, m, scale, dropout_p, seed, dq, dk, dv, workspace, st);
  }
  SCAN_CUDA_CHECK(cudaMemsetAsync(dq, 0, sizeof(float) * n_rows * AT_D, st));
  dim3 grid((m + AT_T - 1) / AT_T, 4);
  attn_bwd_kernel<<<grid, 256, 6 * AT_TILE * 4, st>>>(q, k, v, lse, delta_ws, d_ctx, m, scale, dropout_p, seed, dq, dk, dv);
  SCAN_LAUNCH_CHECK("attn_bwd_kernel");
  return SCAN_OK;
}
