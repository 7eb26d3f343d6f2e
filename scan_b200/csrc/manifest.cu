// K4a: paradigm manifestation, RNN variant (condgraph.py:313-336 get_conded_weight with MODEL.MIDDLE_HEAD.USE_RNN):
//   h = nn.RNN(I -> H, 2 layers, tanh)(prototype.permute(2, 0, 1))       # P steps, "batch" = the K class paradigms
//   kernel[k, o] = sum_{p, c} cond_nx1.weight[o, c, p, 0] * h[p, k, c] + cond_nx1.bias[o]
// ~60 Mflop on 5.3 MB of weights, independent of the image count: pure latency / weight streaming.  torch runs it as a
// few dozen cuDNN / cuBLAS launches per direction; here the forward is P + 2 launches (input transpose, the wavefront of
// the two RNN layers, the (P x 1) convolution) and the backward P + 3 (output layer, reverse wavefront, all weight
// gradients in one launch).  Row-times-matrix products with K <= 16 rows: one warp per output neuron streams its weight
// row once (coalesced float4) against all K rows; transposed products: one thread per output column, the weight matrix
// read row by row (coalesced across threads), 4-way split of the reduction inside the block.  Everything is fp32 FFMA with
// fixed summation order: deterministic.
#include "common.cuh"

namespace scan {

constexpr int MF_KMAX = 16;

struct MfDims {
  int K, P, I, H, O;
};

// xs[t][k][c] = proto[k][c][t]
__global__ void mf_prep_kernel(MfDims d, const float* __restrict__ proto, float* __restrict__ xs) {
  const int n = d.P * d.K * d.I;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i % d.I, k = (i / d.I) % d.K, t = i / (d.I * d.K);
    xs[i] = __ldg(proto + ((long long)k * d.I + c) * d.P + t);
  }
}

// acc[k] += sum_c w_row[c] * X[k][c]   (this lane's share; C % 4 == 0)
__device__ __forceinline__ void mf_row_dot(const float* __restrict__ w_row, const float* __restrict__ X, int C, int K, int lane,
                                           float (&acc)[MF_KMAX]) {
  for (int c = lane * 4; c < C; c += 128) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(w_row + c));
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k)
      if (k < K) {
        const float4 x = *reinterpret_cast<const float4*>(X + (long long)k * C + c);
        acc[k] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[k]))));
      }
  }
}

// wavefront step s: warps [0, H) compute h1[s] (if s < P), warps [H, 2H) compute h2[s-1] (if s >= 1)
__global__ void __launch_bounds__(256) mf_stage_kernel(MfDims d, int s, const float* __restrict__ xs, const float* __restrict__ w_ih0,
                                                       const float* __restrict__ w_hh0, const float* __restrict__ b_ih0,
                                                       const float* __restrict__ b_hh0, const float* __restrict__ w_ih1,
                                                       const float* __restrict__ w_hh1, const float* __restrict__ b_ih1,
                                                       const float* __restrict__ b_hh1, float* __restrict__ h1, float* __restrict__ h2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * d.H) return;
  const int layer = warp / d.H, j = warp % d.H;
  float acc[MF_KMAX];
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
  float* out;
  float bias;
  if (layer == 0) {
    if (s >= d.P) return;
    mf_row_dot(w_ih0 + (long long)j * d.I, xs + (long long)s * d.K * d.I, d.I, d.K, lane, acc);
    if (s > 0) mf_row_dot(w_hh0 + (long long)j * d.H, h1 + (long long)(s - 1) * d.K * d.H, d.H, d.K, lane, acc);
    bias = __ldg(b_ih0 + j) + __ldg(b_hh0 + j);
    out = h1 + (long long)s * d.K * d.H;
  } else {
    const int t = s - 1;
    if (t < 0) return;
    mf_row_dot(w_ih1 + (long long)j * d.H, h1 + (long long)t * d.K * d.H, d.H, d.K, lane, acc);
    if (t > 0) mf_row_dot(w_hh1 + (long long)j * d.H, h2 + (long long)(t - 1) * d.K * d.H, d.H, d.K, lane, acc);
    bias = __ldg(b_ih1 + j) + __ldg(b_hh1 + j);
    out = h2 + (long long)t * d.K * d.H;
  }
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k)
    if (k < d.K) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) out[(long long)k * d.H + j] = tanhf(v + bias);
    }
}

// kernel[k][o] = bc[o] + sum_q wc[o][q] * h2[p][k][c],  q = c * P + p   (one warp per o)
__global__ void __launch_bounds__(256) mf_out_kernel(MfDims d, const float* __restrict__ wc, const float* __restrict__ bc,
                                                     const float* __restrict__ h2, float* __restrict__ out) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= d.O) return;
  float acc[MF_KMAX];
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
  const int Q = d.H * d.P;
  for (int q = lane; q < Q; q += 32) {
    const float w = __ldg(wc + (long long)o * Q + q);
    const int c = q / d.P, p = q - c * d.P;
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k)
      if (k < d.K) acc[k] = fmaf(w, h2[((long long)p * d.K + k) * d.H + c], acc[k]);
  }
  const float b = __ldg(bc + o);
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k)
    if (k < d.K) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) out[(long long)k * d.O + o] = v + b;
    }
}

// ---------------------------------------------------------------------------- backward
// blocks [0, nb_q): dh2[p][k][c] = sum_o d_out[k][o] wc[o][q]      (thread per q)
// blocks [nb_q, nb_q + nb_w): d_wc[o][q] = sum_k d_out[k][o] h2[p][k][c]   (thread per (o, q)); d_bc[o] = sum_k d_out[k][o]
__global__ void __launch_bounds__(256) mf_bwd_out_kernel(MfDims d, int nb_q, const float* __restrict__ d_out, const float* __restrict__ wc,
                                                         const float* __restrict__ h2, float* __restrict__ dh2, float* __restrict__ d_wc,
                                                         float* __restrict__ d_bc) {
  const int Q = d.H * d.P;
  if ((int)blockIdx.x < nb_q) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    float acc[MF_KMAX];
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
    for (int o = 0; o < d.O; ++o) {
      const float w = __ldg(wc + (long long)o * Q + q);
#pragma unroll
      for (int k = 0; k < MF_KMAX; ++k)
        if (k < d.K) acc[k] = fmaf(w, __ldg(d_out + (long long)k * d.O + o), acc[k]);
    }
    const int c = q / d.P, p = q - c * d.P;
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k)
      if (k < d.K) dh2[((long long)p * d.K + k) * d.H + c] = acc[k];
  } else {
    const long long i = (long long)(blockIdx.x - nb_q) * blockDim.x + threadIdx.x;
    if (i >= (long long)d.O * Q) return;
    const int o = (int)(i / Q), q = (int)(i - (long long)o * Q);
    const int c = q / d.P, p = q - c * d.P;
    float s = 0.f, sb = 0.f;
    for (int k = 0; k < d.K; ++k) {
      const float g = __ldg(d_out + (long long)k * d.O + o);
      s = fmaf(g, __ldg(h2 + ((long long)p * d.K + k) * d.H + c), s);
      sb += g;
    }
    d_wc[i] = s;
    if (q == 0) d_bc[o] = sb;
  }
}

// y[k][j] (this thread's j) = sum_i W[i][j] * a[k][i]; the i range is split 4 ways inside the block (blockDim = 4 x 128)
__device__ __forceinline__ void mf_col_dot(const float* __restrict__ W, int ld, int j, const float* __restrict__ a_s, int n_i, int K,
                                           int part, float (&acc)[MF_KMAX]) {
  const int i0 = part * (n_i / 4), i1 = (part == 3) ? n_i : i0 + n_i / 4;
  for (int i = i0; i < i1; ++i) {
    const float w = __ldg(W + (long long)i * ld + j);
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k)
      if (k < K) acc[k] = fmaf(w, a_s[k * n_i + i], acc[k]);
  }
}

// reverse wavefront step r: blocks [0, nb) compute da2[t] for t = P-1-r (if t >= 0); blocks [nb, 2 nb) compute da1[t] for
// t = P-r (if r >= 1 and t <= P-1).  nb = ceil(H / 128); block = 512 threads = 4 reduction parts x 128 columns.
//   da2[t] = (dh2[t] + W_hh1^T da2[t+1]) * (1 - h2[t]^2)
//   da1[t] = (W_ih1^T da2[t] + W_hh0^T da1[t+1]) * (1 - h1[t]^2)
__global__ void __launch_bounds__(512) mf_bwd_stage_kernel(MfDims d, int r, int nb, const float* __restrict__ w_hh0,
                                                           const float* __restrict__ w_ih1, const float* __restrict__ w_hh1,
                                                           const float* __restrict__ h1, const float* __restrict__ h2,
                                                           const float* __restrict__ dh2, float* __restrict__ da1, float* __restrict__ da2) {
  extern __shared__ float sm[];
  float* a_s = sm;                               // [K][H] activations-gradient of the later step / upper layer
  float* red = sm + d.K * d.H;                   // [4][128][K]
  const int which = blockIdx.x / nb, jb = blockIdx.x % nb;
  const int part = threadIdx.x >> 7, jj = threadIdx.x & 127, j = jb * 128 + jj;
  const int t = which == 0 ? d.P - 1 - r : d.P - r;
  if (t < 0 || t > d.P - 1 || (which == 1 && r < 1)) return;
  const long long step = (long long)d.K * d.H;
  float acc[MF_KMAX];
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
  if (which == 0) {
    if (t < d.P - 1) {
      for (int i = threadIdx.x; i < d.K * d.H; i += blockDim.x) a_s[i] = da2[(t + 1) * step + i];
      __syncthreads();
      if (j < d.H) mf_col_dot(w_hh1, d.H, j, a_s, d.H, d.K, part, acc);
    }
  } else {
    for (int i = threadIdx.x; i < d.K * d.H; i += blockDim.x) a_s[i] = da2[t * step + i];
    __syncthreads();
    if (j < d.H) mf_col_dot(w_ih1, d.H, j, a_s, d.H, d.K, part, acc);
    if (t < d.P - 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < d.K * d.H; i += blockDim.x) a_s[i] = da1[(t + 1) * step + i];
      __syncthreads();
      if (j < d.H) mf_col_dot(w_hh0, d.H, j, a_s, d.H, d.K, part, acc);
    }
  }
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k)
    if (k < d.K) red[(part * 128 + jj) * d.K + k] = acc[k];
  __syncthreads();
  if (part == 0 && j < d.H) {
    for (int k = 0; k < d.K; ++k) {
      float v = (red[(0 * 128 + jj) * d.K + k] + red[(1 * 128 + jj) * d.K + k]) + (red[(2 * 128 + jj) * d.K + k] + red[(3 * 128 + jj) * d.K + k]);
      const long long idx = t * step + (long long)k * d.H + j;
      if (which == 0) {
        const float h = h2[idx];
        da2[idx] = (dh2[idx] + v) * (1.f - h * h);
      } else {
        const float h = h1[idx];
        da1[idx] = v * (1.f - h * h);
      }
    }
  }
}

// all RNN weight / bias gradients: thread per matrix element, c fastest
//   m = 0: d_w_ih0[j][c] = sum_{t,k} da1[t][k][j] xs[t][k][c]         (H x I)
//   m = 1: d_w_hh0[j][c] = sum_{t>=1,k} da1[t][k][j] h1[t-1][k][c]    (H x H)
//   m = 2: d_w_ih1[j][c] = sum_{t,k} da2[t][k][j] h1[t][k][c]
//   m = 3: d_w_hh1[j][c] = sum_{t>=1,k} da2[t][k][j] h2[t-1][k][c]
//   biases (c == 0 threads of m = 0 / m = 2): d_b*0[j] = sum_{t,k} da1[t][k][j], d_b*1[j] = sum_{t,k} da2[t][k][j]
__global__ void __launch_bounds__(256) mf_bwd_weights_kernel(MfDims d, const float* __restrict__ xs, const float* __restrict__ h1,
                                                             const float* __restrict__ h2, const float* __restrict__ da1,
                                                             const float* __restrict__ da2, float* __restrict__ d_w_ih0,
                                                             float* __restrict__ d_w_hh0, float* __restrict__ d_w_ih1,
                                                             float* __restrict__ d_w_hh1, float* __restrict__ d_b_ih0,
                                                             float* __restrict__ d_b_hh0, float* __restrict__ d_b_ih1,
                                                             float* __restrict__ d_b_hh1) {
  const long long n0 = (long long)d.H * d.I, n1 = (long long)d.H * d.H;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int m;
  if (i < n0) m = 0;
  else if ((i -= n0) < n1) m = 1;
  else if ((i -= n1) < n1) m = 2;
  else if ((i -= n1) < n1) m = 3;
  else return;
  const int C = m == 0 ? d.I : d.H;
  const int j = (int)(i / C), c = (int)(i - (long long)j * C);
  const float* a = (m < 2) ? da1 : da2;
  const float* x = m == 0 ? xs : (m == 3 ? h2 : h1);
  const int shift = (m == 1 || m == 3) ? 1 : 0;      // recurrent matrices pair step t with the state of step t-1
  float s = 0.f, sb = 0.f;
  for (int t = shift; t < d.P; ++t)
    for (int k = 0; k < d.K; ++k) {
      const float g = __ldg(a + ((long long)t * d.K + k) * d.H + j);
      s = fmaf(g, __ldg(x + ((long long)(t - shift) * d.K + k) * C + c), s);
      sb += g;
    }
  float* out = m == 0 ? d_w_ih0 : (m == 1 ? d_w_hh0 : (m == 2 ? d_w_ih1 : d_w_hh1));
  out[i] = s;
  if (c == 0 && m == 0) { d_b_ih0[j] = sb; d_b_hh0[j] = sb; }
  if (c == 0 && m == 2) { d_b_ih1[j] = sb; d_b_hh1[j] = sb; }
}

static int mf_check(const MfDims& d) {
  if (d.K < 1 || d.K > MF_KMAX || d.P < 1 || d.P > 16 || d.I < 4 || d.H < 4 || d.O < 1 || (d.I & 3) || (d.H & 3)) return SCAN_EINVAL;
  return SCAN_OK;
}

}  // namespace scan

// saved activations: xs [P,K,I] | h1 [P,K,H] | h2 [P,K,H]
extern "C" int64_t scan_manifest_rnn_saved_floats(int32_t K, int32_t P, int32_t I, int32_t H) {
  return (int64_t)P * K * (I + 2 * (int64_t)H);
}
// backward scratch: dh2 | da1 | da2, each [P,K,H]
extern "C" int64_t scan_manifest_rnn_workspace_bytes(int32_t K, int32_t P, int32_t H) { return 3ll * P * K * H * 4 + 256; }

extern "C" int scan_manifest_rnn_fwd(const float* proto, int32_t K, int32_t P, int32_t I, int32_t H, int32_t O, const float* w_ih0,
                                     const float* w_hh0, const float* b_ih0, const float* b_hh0, const float* w_ih1, const float* w_hh1,
                                     const float* b_ih1, const float* b_hh1, const float* wc, const float* bc, float* kernel_out,
                                     float* saved, void* stream) {
  using namespace scan;
  const MfDims d{K, P, I, H, O};
  int rc = mf_check(d);
  if (rc) return rc;
  if (!proto || !w_ih0 || !w_hh0 || !b_ih0 || !b_hh0 || !w_ih1 || !w_hh1 || !b_ih1 || !b_hh1 || !wc || !bc || !kernel_out || !saved)
    return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  float* xs = saved;
  float* h1 = xs + (long long)P * K * I;
  float* h2 = h1 + (long long)P * K * H;
  mf_prep_kernel<<<(P * K * I + 255) / 256, 256, 0, st>>>(d, proto, xs);
  SCAN_LAUNCH_CHECK("mf_prep_kernel");
  for (int s = 0; s <= P; ++s) {
    mf_stage_kernel<<<(2 * H * 32 + 255) / 256, 256, 0, st>>>(d, s, xs, w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, h1, h2);
    SCAN_LAUNCH_CHECK("mf_stage_kernel");
  }
  mf_out_kernel<<<(O * 32 + 255) / 256, 256, 0, st>>>(d, wc, bc, h2, kernel_out);
  SCAN_LAUNCH_CHECK("mf_out_kernel");
  return SCAN_OK;
}

extern "C" int scan_manifest_rnn_bwd(const float* d_kernel, int32_t K, int32_t P, int32_t I, int32_t H, int32_t O, const float* w_hh0,
                                     const float* w_ih1, const float* w_hh1, const float* wc, const float* saved, float* d_w_ih0,
                                     float* d_w_hh0, float* d_b_ih0, float* d_b_hh0, float* d_w_ih1, float* d_w_hh1, float* d_b_ih1,
                                     float* d_b_hh1, float* d_wc, float* d_bc, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace scan;
  const MfDims d{K, P, I, H, O};
  int rc = mf_check(d);
  if (rc) return rc;
  if (!d_kernel || !w_hh0 || !w_ih1 || !w_hh1 || !wc || !saved || !d_w_ih0 || !d_w_hh0 || !d_b_ih0 || !d_b_hh0 || !d_w_ih1 || !d_w_hh1 ||
      !d_b_ih1 || !d_b_hh1 || !d_wc || !d_bc || !workspace)
    return SCAN_EINVAL;
  if (workspace_bytes < scan_manifest_rnn_workspace_bytes(K, P, H)) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  const float* xs = saved;
  const float* h1 = xs + (long long)P * K * I;
  const float* h2 = h1 + (long long)P * K * H;
  float* dh2 = (float*)workspace;
  float* da1 = dh2 + (long long)P * K * H;
  float* da2 = da1 + (long long)P * K * H;
  const int Q = H * P;
  const int nb_q = (Q + 255) / 256;
  const int nb_w = (int)(((long long)O * Q + 255) / 256);
  mf_bwd_out_kernel<<<nb_q + nb_w, 256, 0, st>>>(d, nb_q, d_kernel, wc, h2, dh2, d_wc, d_bc);
  SCAN_LAUNCH_CHECK("mf_bwd_out_kernel");
  const int nb = (H + 127) / 128;
  const size_t smem = ((size_t)K * H + 4 * 128 * K) * sizeof(float);
  static int attr_set = 0;
  if (!attr_set) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(mf_bwd_stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = 1;
  }
  if (smem > 96 * 1024) return SCAN_ENOTSUP;
  for (int r = 0; r <= P; ++r) {
    mf_bwd_stage_kernel<<<2 * nb, 512, smem, st>>>(d, r, nb, w_hh0, w_ih1, w_hh1, h1, h2, dh2, da1, da2);
    SCAN_LAUNCH_CHECK("mf_bwd_stage_kernel");
  }
  const long long n_w = (long long)H * I + 3ll * H * H;
  mf_bwd_weights_kernel<<<(unsigned)((n_w + 255) / 256), 256, 0, st>>>(d, xs, h1, h2, da1, da2, d_w_ih0, d_w_hh0, d_w_ih1, d_w_hh1, d_b_ih0,
                                                                      d_b_hh0, d_b_ih1, d_b_hh1);
  SCAN_LAUNCH_CHECK("mf_bwd_weights_kernel");
  return SCAN_OK;
}
