// a14: the transfer (graph-matching) losses of the target branch, modeling/rpn/fcos/condgraph.py:457-498 (get_transfer_loss)
// with sim_matrix of :35-43, nn.KLDivLoss() (reduction 'mean': divide by ALL elements) and nn.CosineEmbeddingLoss() with
// target +1 (loss = 1 - cos).  The reference runs ~25 tiny torch ops here, two of which (boolean-mask indexing by
// `tg_prototype.sum(-1).bool()`) force a device->host synchronisation; these kernels keep the class mask on the device.
//
//   NODES        KL( softmax(sr_proto[label_m]) || softmax(node_m) ) averaged over the M x 256 elements
//                -> one warp per node row (256 channels = 8 per lane), class targets softmax(sr_proto) staged once per block in
//                   shared memory; the forward saves diff = softmax(node) - target, which IS the gradient up to a scalar.
//   PROTOTYPE    the same KL between the rows of tg_proto and sr_proto for the classes present in the target batch
//   ADJ          1 - cos( vec(S S^T), vec(T T^T) ), S / T = row-normalised sr_proto / tg_proto restricted to the present classes
//   ADJ_COMPLETE the same over all K classes with absent target rows replaced by the source rows (no gradient through those)
//                -> K <= 16 rows of 256 channels: one CTA, everything in shared memory, analytic backward.
// sr_proto = prototype.mean(-1) for PROTO_ITER > 1 (condgraph.py:459-460), computed in-kernel from the [K, 256, P] buffer.
#include "common.cuh"

namespace scan {

constexpr int TR_C = 256;
constexpr int TR_MAXK = SCAN_MAX_CLASSES;

__device__ __forceinline__ float sr_value(const float* __restrict__ proto, int p_iter, int c, int j) {
  if (p_iter == 1) return __ldg(proto + (long long)c * TR_C + j);
  float s = 0.f;
  for (int p = 0; p < p_iter; ++p) s += __ldg(proto + ((long long)c * TR_C + j) * p_iter + p);
  return s / (float)p_iter;      // torch.mean: sum / P
}

// softmax of the 256 values a warp holds as 8 per lane (lane l: columns 8l .. 8l+7); returns log-sum-exp pieces
__device__ __forceinline__ void warp_softmax8(const float (&x)[8], float (&p)[8], float& mx, float& lse) {
  float m = x[0];
#pragma unroll
  for (int e = 1; e < 8; ++e) m = fmaxf(m, x[e]);
  m = warp_max(m);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    p[e] = expf(x[e] - m);
    s += p[e];
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int e = 0; e < 8; ++e) p[e] *= inv;
  mx = m;
  lse = logf(s);
}

// ---------------------------------------------------------------------------- NODES
__global__ void __launch_bounds__(256) transfer_nodes_fwd_kernel(const float* __restrict__ nodes, const long long* __restrict__ labels,
                                                                 const float* __restrict__ proto, int p_iter, int m, int k,
                                                                 float* __restrict__ diff, double* __restrict__ partials) {
  __shared__ float tgt[TR_MAXK][TR_C];       // softmax(sr_proto[c])
  __shared__ float tlog[TR_MAXK][TR_C];      // its logarithm (t log t term)
  __shared__ double red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < k; c += 8) {
    float x[8], p[8], mx, lse;
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = sr_value(proto, p_iter, c, lane * 8 + e);
    warp_softmax8(x, p, mx, lse);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      tgt[c][lane * 8 + e] = p[e];
      tlog[c][lane * 8 + e] = logf(p[e]);      // the reference takes softmax(...) then KLDivLoss computes t * (log t - input)
    }
  }
  __syncthreads();
  double acc = 0.0;
  for (int row = blockIdx.x * 8 + warp; row < m; row += gridDim.x * 8) {
    const float4* n4 = reinterpret_cast<const float4*>(nodes + (long long)row * TR_C + lane * 8);
    const float4 a = __ldg(n4), b = __ldg(n4 + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float p[8], mx, lse;
    warp_softmax8(x, p, mx, lse);
    const int c = (int)labels[row];
    float d[8], l = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = tgt[c][lane * 8 + e];
      const float logp = logf(p[e]);             // input = softmax(nodes).log() (condgraph.py:463)
      l += t * (tlog[c][lane * 8 + e] - logp);
      d[e] = p[e] - t;
    }
    acc += (double)warp_sum(l);
    float4* o = reinterpret_cast<float4*>(diff + (long long)row * TR_C + lane * 8);
    o[0] = make_float4(d[0], d[1], d[2], d[3]);
    o[1] = make_float4(d[4], d[5], d[6], d[7]);
  }
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    partials[blockIdx.x] = s;
  }
}

// d_nodes = diff * (g * scale), g = device scalar d(total)/d(loss)
__global__ void __launch_bounds__(256) transfer_nodes_bwd_kernel(const float* __restrict__ diff, const float* __restrict__ g, float scale,
                                                                 long long n4, float* __restrict__ d_nodes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float s = __ldg(g) * scale;
  float4 v = __ldg(reinterpret_cast<const float4*>(diff) + i);
  v.x *= s; v.y *= s; v.z *= s; v.w *= s;
  reinterpret_cast<float4*>(d_nodes)[i] = v;
}

// ---------------------------------------------------------------------------- PROTOTYPE / ADJ / ADJ_COMPLETE (one CTA)
struct ProtoSmem {
  float sr[TR_MAXK][TR_C];
  float tg[TR_MAXK][TR_C];
  float nsr[TR_MAXK], ntg[TR_MAXK];      // clamped row norms
  int present[TR_MAXK];
  float gs[TR_MAXK][TR_MAXK], gt[TR_MAXK][TR_MAXK];   // Gram matrices of the normalised rows
  float db[TR_MAXK][TR_MAXK];            // d(loss)/d(gt)
  float scal[8];
};

// rows of `tg` selected by `use_row`, others excluded (ADJ) ; completion handled by the caller writing sr into tg rows
__device__ void gram_normalised(const float (*x)[TR_C], const float* nrm, const int* use, int k, float (*g)[TR_MAXK]) {
  // one warp per (i, j) pair
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pair = warp; pair < k * k; pair += blockDim.x >> 5) {
    const int i = pair / k, j = pair % k;
    float s = 0.f;
    if (use[i] && use[j])
      for (int c = lane; c < TR_C; c += 32) s += (x[i][c] / nrm[i]) * (x[j][c] / nrm[j]);
    s = warp_sum(s);
    if (lane == 0) g[i][j] = s;
  }
}

// flags: bit 0 PROTOTYPE, bit 1 ADJ, bit 2 ADJ_COMPLETE.  losses[0..2] (0 where disabled); d_tg = d(sum of enabled losses)/d(tg_proto)
// (the caller scales by the upstream gradient and CON_LOSS_WEIGHT).
__global__ void __launch_bounds__(512) transfer_proto_kernel(const float* __restrict__ tg_proto, const float* __restrict__ proto, int p_iter,
                                                             int k, int flags, float eps, const float* __restrict__ add_in,
                                                             float* __restrict__ losses, float* __restrict__ d_tg) {
  extern __shared__ uint8_t sm_raw[];
  ProtoSmem& s = *reinterpret_cast<ProtoSmem*>(sm_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < k * TR_C; i += blockDim.x) {
    const int c = i / TR_C, j = i % TR_C;
    s.sr[c][j] = sr_value(proto, p_iter, c, j);
    s.tg[c][j] = __ldg(tg_proto + i);
    d_tg[i] = 0.f;
  }
  __syncthreads();
  for (int c = warp; c < k; c += nw) {   // present = (row sum != 0) (condgraph.py:469, 476); norms clamped at eps (sim_matrix)
    float sum = 0.f, q1 = 0.f, q2 = 0.f;
    for (int j = lane; j < TR_C; j += 32) {
      sum += s.tg[c][j];
      q1 += s.sr[c][j] * s.sr[c][j];
      q2 += s.tg[c][j] * s.tg[c][j];
    }
    sum = warp_sum(sum); q1 = warp_sum(q1); q2 = warp_sum(q2);
    if (lane == 0) {
      s.present[c] = sum != 0.f;
      s.nsr[c] = fmaxf(sqrtf(q1), eps);
      s.ntg[c] = fmaxf(sqrtf(q2), eps);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int c = 0; c < k; ++c) n += s.present[c];
    s.scal[0] = (float)n;
    losses[0] = losses[1] = losses[2] = 0.f;
  }
  __syncthreads();
  const float n_present = s.scal[0];

  // ---- PROTOTYPE: KL(softmax(sr[c]) || softmax(tg[c])) over present classes, mean over n_present * 256 elements
  if (flags & 1) {
    __shared__ float part[16];
    float acc = 0.f;
    for (int c = warp; c < k; c += nw) {
      if (!s.present[c]) continue;
      float x[8], t[8], p[8], q[8], mx, lse;
#pragma unroll
      for (int e = 0; e < 8; ++e) { x[e] = s.tg[c][lane * 8 + e]; t[e] = s.sr[c][lane * 8 + e]; }
      warp_softmax8(x, p, mx, lse);
      warp_softmax8(t, q, mx, lse);
      float l = 0.f;
      const float inv = 1.f / (n_present * (float)TR_C);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        l += q[e] * (logf(q[e]) - logf(p[e]));
        atomicAdd(&d_tg[c * TR_C + lane * 8 + e], (p[e] - q[e]) * inv);   // one writer per element in this phase
      }
      acc += warp_sum(l);
    }
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int i = 0; i < nw; ++i) tot += part[i];
      losses[0] = tot / (n_present * (float)TR_C);      // 0/0 = NaN when no class is present, like the reference's empty mean
    }
    __syncthreads();
  }

  // ---- ADJ / ADJ_COMPLETE: 1 - cos(vec(Gs), vec(Gt))
  for (int variant = 0; variant < 2; ++variant) {
    if (!(flags & (2 << variant))) continue;
    __shared__ int use[TR_MAXK];
    __shared__ int grad_row[TR_MAXK];
    if (variant == 1) {   // completion: absent target rows take the source rows (condgraph.py:484-486)
      for (int i = threadIdx.x; i < k * TR_C; i += blockDim.x) {
        const int c = i / TR_C;
        if (!s.present[c]) s.tg[c][i % TR_C] = s.sr[c][i % TR_C];
      }
    }
    if (threadIdx.x < k) {
      use[threadIdx.x] = variant == 1 ? 1 : s.present[threadIdx.x];
      grad_row[threadIdx.x] = s.present[threadIdx.x];
      if (variant == 1 && !s.present[threadIdx.x]) s.ntg[threadIdx.x] = s.nsr[threadIdx.x];
    }
    __syncthreads();
    gram_normalised(s.sr, s.nsr, use, k, s.gs);
    gram_normalised(s.tg, s.ntg, use, k, s.gt);
    __syncthreads();
    if (threadIdx.x == 0) {
      // F.cosine_similarity(a, b, dim=1, eps=1e-8): a.b / max(|a| |b|, eps)  (vectors of length n^2, unused pairs are zero)
      float ab = 0.f, aa = 0.f, bb = 0.f;
      for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
          ab += s.gs[i][j] * s.gt[i][j];
          aa += s.gs[i][j] * s.gs[i][j];
          bb += s.gt[i][j] * s.gt[i][j];
        }
      const float na = sqrtf(aa), nb = sqrtf(bb);
      const float den = fmaxf(na * nb, 1e-8f);
      const float c = ab / den;
      losses[1 + variant] = 1.f - c;
      s.scal[1] = den;
      s.scal[2] = c;
      s.scal[3] = bb;
    }
    __syncthreads();
    // d(1 - c)/d gt[i][j] = -(gs[i][j] / den - c * gt[i][j] / bb)
    for (int i = threadIdx.x; i < k * k; i += blockDim.x) {
      const int a = i / k, b = i % k;
      s.db[a][b] = (use[a] && use[b]) ? -(s.gs[a][b] / s.scal[1] - s.scal[2] * s.gt[a][b] / fmaxf(s.scal[3], 1e-30f)) : 0.f;
    }
    __syncthreads();
    // gt = That That^T, That[i] = tg[i] / ntg[i]:  dThat[i] = sum_j (db[i][j] + db[j][i]) That[j];
    // dtg[i] = (dThat[i] - That[i] (That[i] . dThat[i])) / ntg[i]   (norm above the clamp; a clamped row gets dThat / eps)
    for (int i = warp; i < k; i += nw) {
      if (!use[i] || !grad_row[i]) continue;
      float dth[8], th[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        dth[e] = 0.f;
        th[e] = s.tg[i][lane * 8 + e] / s.ntg[i];
      }
      for (int j = 0; j < k; ++j) {
        if (!use[j]) continue;
        const float w = s.db[i][j] + s.db[j][i];
#pragma unroll
        for (int e = 0; e < 8; ++e) dth[e] += w * (s.tg[j][lane * 8 + e] / s.ntg[j]);
      }
      float dot = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) dot += th[e] * dth[e];
      dot = warp_sum(dot);
      float q2 = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) q2 += s.tg[i][lane * 8 + e] * s.tg[i][lane * 8 + e];
      q2 = warp_sum(q2);
      const bool clamped = sqrtf(q2) < eps;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float g = clamped ? dth[e] / eps : (dth[e] - th[e] * dot) / s.ntg[i];
        atomicAdd(&d_tg[i * TR_C + lane * 8 + e], g);
      }
    }
    __syncthreads();
  }
  // losses[3] = everything enabled here (+ the NODES loss when the caller passes it): the scalar the module returns
  if (threadIdx.x == 0) losses[3] = ((losses[0] + losses[1]) + losses[2]) + (add_in ? __ldg(add_in) : 0.f);
}

// out = sum of the per-block double partials * scale -> fp32 scalar
__global__ void __launch_bounds__(32) transfer_finalize_kernel(const double* __restrict__ partials, int n, double scale, float* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += partials[i];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) *out = (float)(s * scale);
}

}  // namespace scan

using namespace scan;

extern "C" int32_t scan_transfer_nodes_num_partials(void) { return 2 * sm_count(); }

extern "C" int scan_transfer_nodes_fwd(const float* nodes, const int64_t* labels, const float* prototype, int32_t proto_iter, int32_t m,
                                       int32_t num_classes, float* diff, double* partials, float* loss, void* stream) {
  if (!nodes || !labels || !prototype || !diff || !partials || !loss || m < 1 || proto_iter < 1) return SCAN_EINVAL;
  if (num_classes < 1 || num_classes > TR_MAXK) return SCAN_ENOTSUP;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<long long>(ceil_div(m, 8), 2ll * sm_count());
  transfer_nodes_fwd_kernel<<<blocks, 256, 0, st>>>(nodes, (const long long*)labels, prototype, proto_iter, m, num_classes, diff, partials);
  SCAN_LAUNCH_CHECK("transfer_nodes_fwd_kernel");
  transfer_finalize_kernel<<<1, 32, 0, st>>>(partials, blocks, 1.0 / ((double)m * TR_C), loss);   // KLDivLoss reduction='mean'
  SCAN_LAUNCH_CHECK("transfer_finalize_kernel");
  return SCAN_OK;
}

extern "C" int scan_transfer_nodes_bwd(const float* diff, const float* d_loss, int32_t m, float* d_nodes, void* stream) {
  if (!diff || !d_loss || !d_nodes || m < 1) return SCAN_EINVAL;
  const long long n4 = (long long)m * TR_C / 4;
  transfer_nodes_bwd_kernel<<<(unsigned)ceil_div(n4, 256), 256, 0, (cudaStream_t)stream>>>(diff, d_loss, 1.f / ((float)m * TR_C), n4, d_nodes);
  SCAN_LAUNCH_CHECK("transfer_nodes_bwd_kernel");
  return SCAN_OK;
}

extern "C" int scan_transfer_proto(const float* tg_proto, const float* prototype, int32_t proto_iter, int32_t num_classes, int32_t flags,
                                   const float* add_in, float* losses4, float* d_tg_proto, void* stream) {
  if (!tg_proto || !prototype || !losses4 || !d_tg_proto || proto_iter < 1) return SCAN_EINVAL;
  if (num_classes < 1 || num_classes > TR_MAXK) return SCAN_ENOTSUP;
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr))
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(transfer_proto_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ProtoSmem)));
  transfer_proto_kernel<<<1, 512, sizeof(ProtoSmem), (cudaStream_t)stream>>>(tg_proto, prototype, proto_iter, num_classes, flags, 1e-8f,
                                                                             add_in, losses4, d_tg_proto);
  SCAN_LAUNCH_CHECK("transfer_proto_kernel");
  return SCAN_OK;
}
