// a9, the two manifestation variants without the RNN (modeling/rpn/fcos/condgraph.py:320-334):
//   PROTO_ITER > 1:  cond_2( relu( GroupNorm32( cond_nx1(prototype) ) ) )      cond_nx1 = Conv2d(256, hid, (P,1)) == Linear(256 P, hid)
//   PROTO_ITER == 1: cond_2( relu( cond_1(prototype) ) )
// Inputs are the K <= 16 paradigm rows, so these are "tall weight, tiny batch" layers: latency / weight-streaming bound (the
// weights, 0.5 - 1.6 MB, are read exactly once per call).  Building blocks, forward and backward:
//   rows_linear   y[k, o] = act(sum_i x[k, i] W[o, i] + b[o]): one warp per output neuron streams its weight row once (float4,
//                 coalesced) against all K input rows held in shared memory; backward: d_W as K-term outer products (one thread
//                 per 4 weights), d_b, and d_x with one thread per input column (weights read coalesced, transposed use).
//   rows_gn_relu  GroupNorm(32 groups) over the channels of each row + ReLU (F.group_norm on a [K, C] tensor), backward incl.
//                 d_gamma / d_beta.
// Fixed summation orders everywhere (deterministic).
#include "common.cuh"

namespace scan {

constexpr int RM_MAXK = SCAN_MAX_CLASSES;

// grid: ceil(O / 8) blocks of 8 warps; dynamic smem: K * I floats
__global__ void __launch_bounds__(256) rows_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                              int k, int in_dim, int out_dim, int relu, float* __restrict__ y) {
  extern __shared__ float xs[];   // [k][in_dim]
  for (int i = threadIdx.x; i < k * in_dim; i += blockDim.x) xs[i] = __ldg(x + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= out_dim) return;
  float acc[RM_MAXK];
#pragma unroll
  for (int r = 0; r < RM_MAXK; ++r) acc[r] = 0.f;
  const float4* w4 = reinterpret_cast<const float4*>(w + (long long)o * in_dim);
  for (int i4 = lane; i4 < in_dim / 4; i4 += 32) {
    const float4 wv = __ldg(w4 + i4);
#pragma unroll
    for (int r = 0; r < RM_MAXK; ++r)
      if (r < k) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + r * in_dim + i4 * 4);
        acc[r] = fmaf(wv.x, xv.x, fmaf(wv.y, xv.y, fmaf(wv.z, xv.z, fmaf(wv.w, xv.w, acc[r]))));
      }
  }
  const float bias = b ? __ldg(b + o) : 0.f;
#pragma unroll
  for (int r = 0; r < RM_MAXK; ++r)
    if (r < k) {
      float v = warp_sum(acc[r]) + bias;
      if (relu) v = fmaxf(v, 0.f);
      if (lane == 0) y[(long long)r * out_dim + o] = v;
    }
}

// d_y_eff[k, o] = dy[k, o] * (relu ? y[k, o] > 0 : 1);  d_w[o, i] = sum_k d_y_eff[k, o] x[k, i];  d_b[o] = sum_k d_y_eff[k, o]
// grid: (ceil(I / 4 / 256), O); one thread per 4 consecutive weights of output o
__global__ void __launch_bounds__(256) rows_linear_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
                                                                int k, int in_dim, int out_dim, int relu, float* __restrict__ d_w,
                                                                float* __restrict__ d_b) {
  __shared__ float g[RM_MAXK];
  const int o = blockIdx.y;
  if (threadIdx.x < RM_MAXK) {
    float v = 0.f;
    if (threadIdx.x < k) {
      v = __ldg(dy + (long long)threadIdx.x * out_dim + o);
      if (relu && !(__ldg(y + (long long)threadIdx.x * out_dim + o) > 0.f)) v = 0.f;
    }
    g[threadIdx.x] = v;
  }
  __syncthreads();
  const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 < in_dim / 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < k; ++r) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)r * in_dim) + i4);
      acc.x = fmaf(g[r], xv.x, acc.x); acc.y = fmaf(g[r], xv.y, acc.y); acc.z = fmaf(g[r], xv.z, acc.z); acc.w = fmaf(g[r], xv.w, acc.w);
    }
    reinterpret_cast<float4*>(d_w + (long long)o * in_dim)[i4] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && d_b) {
    float s = 0.f;
    for (int r = 0; r < k; ++r) s += g[r];
    d_b[o] = s;
  }
}

// d_x[k, i] = sum_o d_y_eff[k, o] W[o, i]: one thread per input column i (W read coalesced along i), dy staged in smem [k][O]
__global__ void __launch_bounds__(256) rows_linear_dgrad_kernel(const float* __restrict__ w, const float* __restrict__ dy, const float* __restrict__ y,
                                                                int k, int in_dim, int out_dim, int relu, float* __restrict__ d_x) {
  extern __shared__ float gs[];   // [k][out_dim]
  for (int i = threadIdx.x; i < k * out_dim; i += blockDim.x) {
    float v = __ldg(dy + i);
    if (relu && !(__ldg(y + i) > 0.f)) v = 0.f;
    gs[i] = v;
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= in_dim) return;
  float acc[RM_MAXK];
#pragma unroll
  for (int r = 0; r < RM_MAXK; ++r) acc[r] = 0.f;
  for (int o = 0; o < out_dim; ++o) {
    const float wv = __ldg(w + (long long)o * in_dim + i);
#pragma unroll
    for (int r = 0; r < RM_MAXK; ++r)
      if (r < k) acc[r] = fmaf(gs[r * out_dim + o], wv, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < RM_MAXK; ++r)
    if (r < k) d_x[(long long)r * in_dim + i] = acc[r];
}

// y = relu(GroupNorm(x)) per row: grid = K blocks, thread = channel (C <= 1024), groups of C / G consecutive channels
__global__ void __launch_bounds__(1024) rows_gn_relu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                int c_dim, int groups, float eps, float* __restrict__ y, float* __restrict__ stats) {
  __shared__ float xs[1024];
  const int r = blockIdx.x, c = threadIdx.x;
  const int gsz = c_dim / groups;
  if (c < c_dim) xs[c] = __ldg(x + (long long)r * c_dim + c);
  __syncthreads();
  if (c >= c_dim) return;
  const int g0 = (c / gsz) * gsz;
  float mean = 0.f;
  for (int i = 0; i < gsz; ++i) mean += xs[g0 + i];
  mean /= (float)gsz;
  float var = 0.f;
  for (int i = 0; i < gsz; ++i) {
    const float d = xs[g0 + i] - mean;
    var = fmaf(d, d, var);
  }
  const float rstd = rsqrtf(var / (float)gsz + eps);
  const float v = (xs[c] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
  y[(long long)r * c_dim + c] = fmaxf(v, 0.f);
  if (c == g0) {
    stats[((long long)r * groups + c / gsz) * 2] = mean;
    stats[((long long)r * groups + c / gsz) * 2 + 1] = rstd;
  }
}

// backward of the above.  grid = K blocks for d_x; d_gamma / d_beta by a second kernel (thread per channel, loop over rows)
__global__ void __launch_bounds__(1024) rows_gn_relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                                                const float* __restrict__ gamma, const float* __restrict__ stats, int c_dim,
                                                                int groups, float* __restrict__ d_x) {
  __shared__ float gx[1024], xh[1024];
  const int r = blockIdx.x, c = threadIdx.x;
  const int gsz = c_dim / groups;
  if (c < c_dim) {
    const float mean = stats[((long long)r * groups + c / gsz) * 2], rstd = stats[((long long)r * groups + c / gsz) * 2 + 1];
    const float d = __ldg(y + (long long)r * c_dim + c) > 0.f ? __ldg(dy + (long long)r * c_dim + c) : 0.f;
    gx[c] = d * __ldg(gamma + c);
    xh[c] = (__ldg(x + (long long)r * c_dim + c) - mean) * rstd;
  }
  __syncthreads();
  if (c >= c_dim) return;
  const int g0 = (c / gsz) * gsz;
  float s1 = 0.f, s2 = 0.f;
  for (int i = 0; i < gsz; ++i) {
    s1 += gx[g0 + i];
    s2 = fmaf(gx[g0 + i], xh[g0 + i], s2);
  }
  const float rstd = stats[((long long)r * groups + c / gsz) * 2 + 1];
  d_x[(long long)r * c_dim + c] = rstd * (gx[c] - s1 / (float)gsz - xh[c] * s2 / (float)gsz);
}
__global__ void __launch_bounds__(256) rows_gn_relu_bwd_affine_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                                                                      const float* __restrict__ stats, int k, int c_dim, int groups,
                                                                      float* __restrict__ d_gamma, float* __restrict__ d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_dim) return;
  const int gsz = c_dim / groups;
  float sg = 0.f, sb = 0.f;
  for (int r = 0; r < k; ++r) {
    const float d = __ldg(y + (long long)r * c_dim + c) > 0.f ? __ldg(dy + (long long)r * c_dim + c) : 0.f;
    const float mean = stats[((long long)r * groups + c / gsz) * 2], rstd = stats[((long long)r * groups + c / gsz) * 2 + 1];
    sg = fmaf(d, (__ldg(x + (long long)r * c_dim + c) - mean) * rstd, sg);
    sb += d;
  }
  d_gamma[c] = sg;
  d_beta[c] = sb;
}

}  // namespace scan

using namespace scan;

extern "C" int scan_rows_linear_fwd(const float* x, const float* w, const float* b, int32_t k, int32_t in_dim, int32_t out_dim, int32_t relu,
                                    float* y, void* stream) {
  if (!x || !w || !y || k < 1 || k > RM_MAXK || in_dim < 4 || in_dim % 4 || out_dim < 1) return SCAN_EINVAL;
  const size_t smem = (size_t)k * in_dim * sizeof(float);
  if (smem > 200 * 1024) return SCAN_ENOTSUP;
  if (smem > 48 * 1024) SCAN_CUDA_CHECK(cudaFuncSetAttribute(rows_linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rows_linear_fwd_kernel<<<(unsigned)ceil_div(out_dim, 8), 256, smem, (cudaStream_t)stream>>>(x, w, b, k, in_dim, out_dim, relu, y);
  SCAN_LAUNCH_CHECK("rows_linear_fwd_kernel");
  return SCAN_OK;
}

// y: the forward output (needed when relu != 0); d_x may be NULL (first layer: the paradigm buffer takes no gradient)
extern "C" int scan_rows_linear_bwd(const float* x, const float* w, const float* dy, const float* y, int32_t k, int32_t in_dim, int32_t out_dim,
                                    int32_t relu, float* d_w, float* d_b, float* d_x, void* stream) {
  if (!x || !w || !dy || !d_w || k < 1 || k > RM_MAXK || in_dim < 4 || in_dim % 4 || out_dim < 1 || (relu && !y)) return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  rows_linear_wgrad_kernel<<<dim3((unsigned)ceil_div(in_dim / 4, 256), (unsigned)out_dim), 256, 0, st>>>(x, dy, y, k, in_dim, out_dim, relu, d_w, d_b);
  SCAN_LAUNCH_CHECK("rows_linear_wgrad_kernel");
  if (d_x) {
    const size_t smem = (size_t)k * out_dim * sizeof(float);
    if (smem > 200 * 1024) return SCAN_ENOTSUP;
    if (smem > 48 * 1024) SCAN_CUDA_CHECK(cudaFuncSetAttribute(rows_linear_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rows_linear_dgrad_kernel<<<(unsigned)ceil_div(in_dim, 256), 256, smem, st>>>(w, dy, y, k, in_dim, out_dim, relu, d_x);
    SCAN_LAUNCH_CHECK("rows_linear_dgrad_kernel");
  }
  return SCAN_OK;
}

extern "C" int scan_rows_gn_relu_fwd(const float* x, const float* gamma, const float* beta, int32_t k, int32_t channels, int32_t groups, float eps,
                                     float* y, float* stats, void* stream) {
  if (!x || !gamma || !beta || !y || !stats || k < 1 || channels < 1 || channels > 1024 || groups < 1 || channels % groups) return SCAN_EINVAL;
  rows_gn_relu_fwd_kernel<<<k, 1024, 0, (cudaStream_t)stream>>>(x, gamma, beta, channels, groups, eps, y, stats);
  SCAN_LAUNCH_CHECK("rows_gn_relu_fwd_kernel");
  return SCAN_OK;
}

extern "C" int scan_rows_gn_relu_bwd(const float* x, const float* y, const float* dy, const float* gamma, const float* stats, int32_t k,
                                     int32_t channels, int32_t groups, float* d_x, float* d_gamma, float* d_beta, void* stream) {
  if (!x || !y || !dy || !gamma || !stats || !d_x || !d_gamma || !d_beta || k < 1 || channels < 1 || channels > 1024 || channels % groups) return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  rows_gn_relu_bwd_kernel<<<k, 1024, 0, st>>>(x, y, dy, gamma, stats, channels, groups, d_x);
  SCAN_LAUNCH_CHECK("rows_gn_relu_bwd_kernel");
  rows_gn_relu_bwd_affine_kernel<<<(unsigned)ceil_div(channels, 256), 256, 0, st>>>(x, y, dy, stats, k, channels, groups, d_gamma, d_beta);
  SCAN_LAUNCH_CHECK("rows_gn_relu_bwd_affine_kernel");
  return SCAN_OK;
}
