// K4a: paradigm manifestation, RNN variant (condgraph.py:313-336 get_conded_weight with MODEL.MIDDLE_HEAD.USE_RNN):
//   h = nn.RNN(I -> H, 2 layers, tanh)(prototype.permute(2, 0, 1))       # P steps, "batch" = the K class paradigms
//   kernel[k, o] = sum_{p, c} cond_nx1.weight[o, c, p, 0] * h[p, k, c] + cond_nx1.bias[o]
// ~60 Mflop on 5.3 MB of weights, independent of the image count: pure latency / weight streaming.  torch runs it as a
// few dozen cuDNN / cuBLAS launches per direction; here the forward is P + 2 launches (input transpose, the wavefront of
// the two RNN layers, the (P x 1) convolution) and the backward P + 3 (output layer, reverse wavefront, all weight
// gradients in one launch).  Every product is "K <= 16 rows times a matrix": one warp per output neuron streams its weight
// row once (coalesced float4, all loads of a row issued before the first FMA: the weights come cold from HBM, the kernels
// are latency-bound) against all K rows.  The backward needs the transposed products; it transposes the four matrices once
// (one launch, 4 MB) and then uses the same row-streaming kernels.  fp32 FFMA, fixed summation order: deterministic.
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

// acc[k] += sum_c w_row[c] * X[k][c]   (this lane's share; C % 4 == 0).  Four weight loads in flight per lane.
__device__ __forceinline__ void mf_row_dot(const float* __restrict__ w_row, const float* __restrict__ X, int C, int K, int lane,
                                           float (&acc)[MF_KMAX]) {
  for (int c0 = lane * 4; c0 < C; c0 += 512) {
    float4 w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      w[u] = (c0 + u * 128 < C) ? __ldg(reinterpret_cast<const float4*>(w_row + c0 + u * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u * 128;
      if (c < C) {
#pragma unroll
        for (int k = 0; k < MF_KMAX; ++k)
          if (k < K) {
            const float4 x = *reinterpret_cast<const float4*>(X + (long long)k * C + c);
            acc[k] = fmaf(w[u].x, x.x, fmaf(w[u].y, x.y, fmaf(w[u].z, x.z, fmaf(w[u].w, x.w, acc[k]))));
          }
      }
    }
  }
}

// wavefront step s: warps [0, H) compute h1[s] (if s < P), warps [H, 2H) compute h2[s-1] (if s >= 1)
__global__ void __launch_bounds__(256) mf_stage_kernel(MfDims d, int s, const float* __restrict__ xs, const float* __restrict__ w_ih0,
                                                       const float* __restrict__ w_hh0, const float* __restrict__ b_ih0,
                                                       const float* __restrict__ b_hh0, const float* __restrict__ w_ih1,
                                                       const float* __restrict__ w_hh1, const float* __restrict__ b_ih1,
                                                       const float* __restrict__ b_hh1, float* __restrict__ h1, float* __restrict__ h2,
                                                       float* __restrict__ h2q /* [K][H*P], q = c * P + p */) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * d.H) return;
  const int layer = warp / d.H, j = warp % d.H;
  float acc[MF_KMAX];
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
  float* out;
  float bias;
  int tq = -1;
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
    tq = t;
  }
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k)
    if (k < d.K) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) {
        const float hv = tanhf(v + bias);
        out[(long long)k * d.H + j] = hv;
        if (tq >= 0) h2q[(long long)k * d.H * d.P + (long long)j * d.P + tq] = hv;
      }
    }
}

// kernel[k][o] = bc[o] + sum_q wc[o][q] * h2q[k][q],  q = c * P + p   (one warp per o)
__global__ void __launch_bounds__(256) mf_out_kernel(MfDims d, const float* __restrict__ wc, const float* __restrict__ bc,
                                                     const float* __restrict__ h2q, float* __restrict__ out) {
  const int o = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (o >= d.O) return;
  float acc[MF_KMAX];
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
  const int Q = d.H * d.P;
  mf_row_dot(wc + (long long)o * Q, h2q, Q, d.K, lane, acc);
  const float b = __ldg(bc + o);
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k)
    if (k < d.K) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) out[(long long)k * d.O + o] = v + b;
    }
}

// ---------------------------------------------------------------------------- backward
// one launch transposes the four matrices the backward multiplies from the other side (32 x 32 tiles through shared memory)
struct MfTr {
  const float* src[4];
  float* dst[4];
  int rows[4], cols[4];
  int tile_off[5];
};
__global__ void __launch_bounds__(256) mf_transpose_kernel(MfTr a) {
  __shared__ float tile[32][33];
  int m = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if ((int)blockIdx.x >= a.tile_off[i]) m = i;
  const int t = blockIdx.x - a.tile_off[m];
  const int tcols = (a.cols[m] + 31) / 32;
  const int r0 = (t / tcols) * 32, c0 = (t % tcols) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    tile[r][tx] = (r0 + r < a.rows[m] && c0 + tx < a.cols[m]) ? __ldg(a.src[m] + (long long)(r0 + r) * a.cols[m] + c0 + tx) : 0.f;
  __syncthreads();
  for (int c = ty; c < 32; c += 8)
    if (c0 + c < a.cols[m] && r0 + tx < a.rows[m]) a.dst[m][(long long)(c0 + c) * a.rows[m] + r0 + tx] = tile[tx][c];
}

// warps [0, Q): dh2[p][k][c] = sum_o wcT[q][o] d_out[k][o]                          (one warp per q, q = c * P + p)
// then threads: d_wc[o][q] = sum_k d_out[k][o] h2q[k][q];  d_bc[o] = sum_k d_out[k][o]
__global__ void __launch_bounds__(256) mf_bwd_out_kernel(MfDims d, int nb_q, const float* __restrict__ d_out, const float* __restrict__ wcT,
                                                         const float* __restrict__ h2q, float* __restrict__ dh2, float* __restrict__ d_wc,
                                                         float* __restrict__ d_bc) {
  const int Q = d.H * d.P;
  if ((int)blockIdx.x < nb_q) {
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= Q) return;
    float acc[MF_KMAX];
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
    mf_row_dot(wcT + (long long)q * d.O, d_out, d.O, d.K, lane, acc);
    const int c = q / d.P, p = q - c * d.P;
#pragma unroll
    for (int k = 0; k < MF_KMAX; ++k)
      if (k < d.K) {
        const float v = warp_sum(acc[k]);
        if (lane == 0) dh2[((long long)p * d.K + k) * d.H + c] = v;
      }
  } else {
    const long long i = (long long)(blockIdx.x - nb_q) * blockDim.x + threadIdx.x;
    if (i >= (long long)d.O * Q) return;
    const int o = (int)(i / Q), q = (int)(i - (long long)o * Q);
    float s = 0.f, sb = 0.f;
#pragma unroll 4
    for (int k = 0; k < d.K; ++k) {
      const float g = __ldg(d_out + (long long)k * d.O + o);
      s = fmaf(g, __ldg(h2q + (long long)k * Q + q), s);
      sb += g;
    }
    d_wc[i] = s;
    if (q == 0) d_bc[o] = sb;
  }
}

// reverse wavefront step r: warps [0, H) compute da2[t] for t = P-1-r (if t >= 0); warps [H, 2H) compute da1[t] for
// t = P-r (if r >= 1 and t <= P-1), with the TRANSPOSED matrices (row j of W^T = column j of W):
//   da2[t] = (dh2[t] + da2[t+1] . W_hh1) * (1 - h2[t]^2)
//   da1[t] = (da2[t] . W_ih1 + da1[t+1] . W_hh0) * (1 - h1[t]^2)
__global__ void __launch_bounds__(256) mf_bwd_stage_kernel(MfDims d, int r, const float* __restrict__ w_hh0T, const float* __restrict__ w_ih1T,
                                                           const float* __restrict__ w_hh1T, const float* __restrict__ h1,
                                                           const float* __restrict__ h2, const float* __restrict__ dh2,
                                                           float* __restrict__ da1, float* __restrict__ da2) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= 2 * d.H) return;
  const int which = warp / d.H, j = warp % d.H;
  const int t = which == 0 ? d.P - 1 - r : d.P - r;
  if (t < 0 || t > d.P - 1 || (which == 1 && r < 1)) return;
  const long long step = (long long)d.K * d.H;
  float acc[MF_KMAX];
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k) acc[k] = 0.f;
  if (which == 0) {
    if (t < d.P - 1) mf_row_dot(w_hh1T + (long long)j * d.H, da2 + (t + 1) * step, d.H, d.K, lane, acc);
  } else {
    mf_row_dot(w_ih1T + (long long)j * d.H, da2 + t * step, d.H, d.K, lane, acc);
    if (t < d.P - 1) mf_row_dot(w_hh0T + (long long)j * d.H, da1 + (t + 1) * step, d.H, d.K, lane, acc);
  }
#pragma unroll
  for (int k = 0; k < MF_KMAX; ++k)
    if (k < d.K) {
      const float v = warp_sum(acc[k]);
      if (lane == 0) {
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

// all RNN weight / bias gradients.  Block = 16 output rows j x 256 columns c of one matrix; thread = column, 16 accumulators;
// the (t, k) terms of the 16 rows are staged in shared memory, each x value is loaded once for 16 FMAs.
//   m = 0: d_w_ih0[j][c] = sum_{t,k} da1[t][k][j] xs[t][k][c]         (H x I)
//   m = 1: d_w_hh0[j][c] = sum_{t>=1,k} da1[t][k][j] h1[t-1][k][c]    (H x H)
//   m = 2: d_w_ih1[j][c] = sum_{t,k} da2[t][k][j] h1[t][k][c]
//   m = 3: d_w_hh1[j][c] = sum_{t>=1,k} da2[t][k][j] h2[t-1][k][c]
//   biases (first column block of m = 0 / m = 2): d_b*0[j] = sum_{t,k} da1[t][k][j], d_b*1[j] = sum_{t,k} da2[t][k][j]
struct MfWg {
  int blk_off[5];     // first block of each matrix
  int cblocks[4];     // column blocks (of 256) per matrix
};
__global__ void __launch_bounds__(256) mf_bwd_weights_kernel(MfDims d, MfWg wg, const float* __restrict__ xs, const float* __restrict__ h1,
                                                             const float* __restrict__ h2, const float* __restrict__ da1,
                                                             const float* __restrict__ da2, float* __restrict__ d_w_ih0,
                                                             float* __restrict__ d_w_hh0, float* __restrict__ d_w_ih1,
                                                             float* __restrict__ d_w_hh1, float* __restrict__ d_b_ih0,
                                                             float* __restrict__ d_b_hh0, float* __restrict__ d_b_ih1,
                                                             float* __restrict__ d_b_hh1) {
  __shared__ float a_s[16 * MF_KMAX][16];   // [(t, k)][j in tile]   (P <= 16, K <= 16)
  int m = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if ((int)blockIdx.x >= wg.blk_off[i]) m = i;
  const int rel = blockIdx.x - wg.blk_off[m];
  const int jb = rel / wg.cblocks[m], cb = rel % wg.cblocks[m];
  const int C = m == 0 ? d.I : d.H;
  const int j0 = jb * 16, c = cb * 256 + threadIdx.x;
  const float* a = (m < 2) ? da1 : da2;
  const float* x = m == 0 ? xs : (m == 3 ? h2 : h1);
  const int shift = (m == 1 || m == 3) ? 1 : 0;      // recurrent matrices pair step t with the state of step t-1
  const int n_tk = d.P * d.K;
  for (int i = threadIdx.x; i < n_tk * 16; i += 256) {
    const int tk = i >> 4, jj = i & 15;
    a_s[tk][jj] = (j0 + jj < d.H) ? __ldg(a + (long long)tk * d.H + j0 + jj) : 0.f;
  }
  __syncthreads();
  float acc[16];
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) acc[jj] = 0.f;
  if (c < C) {
    for (int tk = shift * d.K; tk < n_tk; ++tk) {
      const float xv = __ldg(x + (long long)(tk - shift * d.K) * C + c);
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) acc[jj] = fmaf(a_s[tk][jj], xv, acc[jj]);
    }
    float* out = m == 0 ? d_w_ih0 : (m == 1 ? d_w_hh0 : (m == 2 ? d_w_ih1 : d_w_hh1));
#pragma unroll
    for (int jj = 0; jj < 16; ++jj)
      if (j0 + jj < d.H) out[(long long)(j0 + jj) * C + c] = acc[jj];
  }
  if (cb == 0 && (m == 0 || m == 2) && threadIdx.x < 16 && j0 + threadIdx.x < d.H) {
    float sb = 0.f;
    for (int tk = 0; tk < n_tk; ++tk) sb += a_s[tk][threadIdx.x];
    if (m == 0) { d_b_ih0[j0 + threadIdx.x] = sb; d_b_hh0[j0 + threadIdx.x] = sb; }
    else { d_b_ih1[j0 + threadIdx.x] = sb; d_b_hh1[j0 + threadIdx.x] = sb; }
  }
}

static int mf_check(const MfDims& d) {
  if (d.K < 1 || d.K > MF_KMAX || d.P < 1 || d.P > 16 || d.I < 4 || d.H < 4 || d.O < 1 || (d.I & 3) || (d.H & 3) || (d.O & 3)) return SCAN_EINVAL;
  return SCAN_OK;
}

}  // namespace scan

// saved activations: xs [P,K,I] | h1 [P,K,H] | h2 [P,K,H] | h2q [K,H*P]
extern "C" int64_t scan_manifest_rnn_saved_floats(int32_t K, int32_t P, int32_t I, int32_t H) {
  return (int64_t)P * K * (I + 3 * (int64_t)H);
}
// backward scratch: dh2 | da1 | da2, each [P,K,H]; transposed W_hh0, W_ih1, W_hh1 [H,H] and Wc^T [H*P, O]
extern "C" int64_t scan_manifest_rnn_workspace_bytes(int32_t K, int32_t P, int32_t H, int32_t O) {
  return (3ll * P * K * H + 3ll * H * H + (int64_t)H * P * O) * 4 + 256;
}

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
  float* h2q = h2 + (long long)P * K * H;
  mf_prep_kernel<<<(P * K * I + 255) / 256, 256, 0, st>>>(d, proto, xs);
  SCAN_LAUNCH_CHECK("mf_prep_kernel");
  for (int s = 0; s <= P; ++s) {
    mf_stage_kernel<<<(2 * H * 32 + 255) / 256, 256, 0, st>>>(d, s, xs, w_ih0, w_hh0, b_ih0, b_hh0, w_ih1, w_hh1, b_ih1, b_hh1, h1, h2, h2q);
    SCAN_LAUNCH_CHECK("mf_stage_kernel");
  }
  mf_out_kernel<<<(O * 32 + 255) / 256, 256, 0, st>>>(d, wc, bc, h2q, kernel_out);
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
  if (workspace_bytes < scan_manifest_rnn_workspace_bytes(K, P, H, O)) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  const float* xs = saved;
  const float* h1 = xs + (long long)P * K * I;
  const float* h2 = h1 + (long long)P * K * H;
  const float* h2q = h2 + (long long)P * K * H;
  float* dh2 = (float*)workspace;
  float* da1 = dh2 + (long long)P * K * H;
  float* da2 = da1 + (long long)P * K * H;
  float* w_hh0T = da2 + (long long)P * K * H;
  float* w_ih1T = w_hh0T + (long long)H * H;
  float* w_hh1T = w_ih1T + (long long)H * H;
  float* wcT = w_hh1T + (long long)H * H;
  const int Q = H * P;
  MfTr tr;
  const float* srcs[4] = {w_hh0, w_ih1, w_hh1, wc};
  float* dsts[4] = {w_hh0T, w_ih1T, w_hh1T, wcT};
  int off = 0;
  for (int i = 0; i < 4; ++i) {
    tr.src[i] = srcs[i];
    tr.dst[i] = dsts[i];
    tr.rows[i] = i < 3 ? H : O;
    tr.cols[i] = i < 3 ? H : Q;
    tr.tile_off[i] = off;
    off += ((tr.rows[i] + 31) / 32) * ((tr.cols[i] + 31) / 32);
  }
  tr.tile_off[4] = off;
  mf_transpose_kernel<<<off, 256, 0, st>>>(tr);
  SCAN_LAUNCH_CHECK("mf_transpose_kernel");
  const int nb_q = (Q * 32 + 255) / 256;
  const int nb_w = (int)(((long long)O * Q + 255) / 256);
  mf_bwd_out_kernel<<<nb_q + nb_w, 256, 0, st>>>(d, nb_q, d_kernel, wcT, h2q, dh2, d_wc, d_bc);
  SCAN_LAUNCH_CHECK("mf_bwd_out_kernel");
  for (int r = 0; r <= P; ++r) {
    mf_bwd_stage_kernel<<<(2 * H * 32 + 255) / 256, 256, 0, st>>>(d, r, w_hh0T, w_ih1T, w_hh1T, h1, h2, dh2, da1, da2);
    SCAN_LAUNCH_CHECK("mf_bwd_stage_kernel");
  }
  MfWg wg;
  off = 0;
  for (int i = 0; i < 4; ++i) {
    wg.blk_off[i] = off;
    wg.cblocks[i] = ((i == 0 ? I : H) + 255) / 256;
    off += ((H + 15) / 16) * wg.cblocks[i];
  }
  wg.blk_off[4] = off;
  mf_bwd_weights_kernel<<<off, 256, 0, st>>>(d, wg, xs, h1, h2, da1, da2, d_w_ih0, d_w_hh0, d_w_ih1, d_w_hh1, d_b_ih0, d_b_hh0, d_b_ih1,
                                             d_b_hh1);
  SCAN_LAUNCH_CHECK("mf_bwd_weights_kernel");
  return SCAN_OK;
}
