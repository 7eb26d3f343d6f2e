// f2: FCOSLossComputation.__call__ (modeling/rpn/fcos/loss.py:168-230) fused: sigmoid focal classification loss over every
// location and class (csrc/cuda/SigmoidFocalLoss_cuda.cu:36-55 formulas, shared with losses.cu), IoU regression loss weighted by
// the centerness targets (layers/iou_loss.py:5-38) and BCE-with-logits centerness loss over the positive locations
// (compute_centerness_targets, loss.py:128-133), forward and backward.  The reference flattens and concatenates 15 NCHW maps
// (permute + reshape + cat), gathers the positives by `nonzero` (a host sync) and launches ~40 kernels; here the assignment
// kernel (scan_fcos_assign_reg) hands over labels + regression targets in the rows layout and ONE pass reads the head's NCHW
// maps in place: thread = location (consecutive threads = consecutive pixels: coalesced along the pixel axis of every map).
//   forward : per-CTA fp64 partials {focal sum, #pos, sum w, sum iou*w, sum iou, sum bce} -> fcos_loss_finalize_kernel
//   backward: d_cls, d_reg, d_ctr written in place of the three map sets (zero for non-positive locations of reg / ctr)
#include "common.cuh"

namespace scan {

struct FcosMaps {
  const float* cls[SCAN_MAX_LEVELS];   // [N, C, H, W] logits
  const float* reg[SCAN_MAX_LEVELS];   // [N, 4, H, W] (l, t, r, b) predictions (already exp()-ed by the head)
  const float* ctr[SCAN_MAX_LEVELS];   // [N, 1, H, W] logits
};
struct FcosGrads {
  float* cls[SCAN_MAX_LEVELS];
  float* reg[SCAN_MAX_LEVELS];
  float* ctr[SCAN_MAX_LEVELS];
};

struct FocalT {
  float p, log_p, log_1mp;
};
__device__ __forceinline__ FocalT fl_terms(float x) {   // same core as losses.cu::focal_terms
  const float e = expf(-fabsf(x));
  const float inv = 1.f / (1.f + e);
  const float l1pe = logf(1.f + e);
  FocalT t;
  t.p = (x >= 0.f ? 1.f : e) * inv;
  t.log_p = t.p >= 1.17549435e-38f ? fminf(x, 0.f) - l1pe : -87.33654475f;
  t.log_1mp = -fmaxf(x, 0.f) - l1pe;
  return t;
}
__device__ __forceinline__ float fl_pow(float v, float gamma, bool sq) { return sq ? v * v : powf(v, gamma); }

// centerness target of a positive location (loss.py:128-133)
__device__ __forceinline__ float ctr_target(const float4 t) {
  const float lr = fminf(t.x, t.z) / fmaxf(t.x, t.z);
  const float tb = fminf(t.y, t.w) / fmaxf(t.y, t.w);
  return sqrtf(lr * tb);
}

struct RowCoord {
  int l, n;
  long long hw, p;
};
__device__ __forceinline__ RowCoord row_coord(const Levels& lv, long long g) {
  RowCoord c;
  c.l = level_of_row(lv, g);
  c.hw = (long long)lv.h[c.l] * lv.w[c.l];
  const long long r = g - lv.row_off[c.l];
  c.n = (int)(r / c.hw);
  c.p = r - (long long)c.n * c.hw;
  return c;
}

constexpr int FL_NPART = 6;   // focal sum, #pos, sum w, sum iou*w, sum iou, sum bce

__global__ void __launch_bounds__(256) fcos_loss_fwd_kernel(Levels lv, FcosMaps mp, const int64_t* __restrict__ labels, const float4* __restrict__ reg_t,
                                                            int num_classes, float gamma, float alpha, double* __restrict__ partials) {
  __shared__ double red[8][FL_NPART];
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const bool sq = gamma == 2.f;
  double acc[FL_NPART] = {0, 0, 0, 0, 0, 0};
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < R; g += (long long)gridDim.x * blockDim.x) {
    const RowCoord c = row_coord(lv, g);
    const int t = (int)labels[g];
    const float* cls = mp.cls[c.l] + (long long)c.n * num_classes * c.hw + c.p;
    float f = 0.f;
    for (int d = 0; d < num_classes; ++d) {
      const FocalT ft = fl_terms(__ldg(cls + d * c.hw));
      if (t == d + 1) f += -alpha * fl_pow(1.f - ft.p, gamma, sq) * ft.log_p;
      else if (t >= 0) f += -(1.f - alpha) * fl_pow(ft.p, gamma, sq) * ft.log_1mp;
    }
    acc[0] += (double)f;
    if (t > 0) {
      const float4 tg = reg_t[g];
      const float* rp = mp.reg[c.l] + (long long)c.n * 4 * c.hw + c.p;
      const float pl = __ldg(rp), pt = __ldg(rp + c.hw), pr = __ldg(rp + 2 * c.hw), pb = __ldg(rp + 3 * c.hw);
      // layers/iou_loss.py:17-30
      const float t_area = (tg.x + tg.z) * (tg.y + tg.w);
      const float p_area = (pl + pr) * (pt + pb);
      const float wi = fminf(pl, tg.x) + fminf(pr, tg.z);
      const float hi = fminf(pb, tg.w) + fminf(pt, tg.y);
      const float inter = wi * hi;
      const float uni = t_area + p_area - inter;
      const float iou_l = -logf((inter + 1.f) / (uni + 1.f));
      const float w = ctr_target(tg);
      const float x = __ldg(mp.ctr[c.l] + (long long)c.n * c.hw + c.p);
      // BCEWithLogits: max(x, 0) - x t + log(1 + exp(-|x|))
      const float bce = fmaxf(x, 0.f) - x * w + logf(1.f + expf(-fabsf(x)));
      acc[1] += 1.0;
      acc[2] += (double)w;
      acc[3] += (double)(iou_l * w);
      acc[4] += (double)iou_l;
      acc[5] += (double)bce;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < FL_NPART; ++i) {
    const double v = warp_sum_d(acc[i]);
    if (lane == 0) red[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < FL_NPART) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partials[(long long)blockIdx.x * FL_NPART + threadIdx.x] = s;
  }
}

// sums[0..5] (fp64, kept for the backward) and the three losses of loss.py:207-230:
//   cls = focal / (n_pos + N); reg = sum(iou w) / sum(w) if sum(w) > 0 else mean(iou); ctr = mean(bce); both 0 without positives
//   (the reference returns `.sum()` of empty tensors there)
__global__ void __launch_bounds__(32) fcos_loss_finalize_kernel(const double* __restrict__ partials, int n_blocks, int n_images,
                                                                double* __restrict__ sums, float* __restrict__ losses) {
  double s[FL_NPART];
  for (int i = 0; i < FL_NPART; ++i) {
    double v = 0.0;
    for (int b = threadIdx.x; b < n_blocks; b += 32) v += partials[(long long)b * FL_NPART + i];
    s[i] = warp_sum_d(v);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < FL_NPART; ++i) sums[i] = s[i];
    losses[0] = (float)(s[0] / (s[1] + (double)n_images));
    if (s[1] > 0.0) {
      losses[1] = (float)(s[2] > 0.0 ? s[3] / s[2] : s[4] / s[1]);
      losses[2] = (float)(s[5] / s[1]);
    } else {
      losses[1] = 0.f;
      losses[2] = 0.f;
    }
  }
}

// d_losses[3] = upstream gradients of (cls, reg, ctr)
__global__ void __launch_bounds__(256) fcos_loss_bwd_kernel(Levels lv, FcosMaps mp, FcosGrads gr, const int64_t* __restrict__ labels,
                                                            const float4* __restrict__ reg_t, int num_classes, float gamma, float alpha,
                                                            int n_images, const double* __restrict__ sums, const float* __restrict__ d_losses) {
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const bool sq = gamma == 2.f;
  const double n_pos = sums[1], sw = sums[2];
  const float g_cls = (float)((double)__ldg(d_losses) / (n_pos + (double)n_images));
  const bool weighted = sw > 0.0;
  const float g_reg = n_pos > 0.0 ? (float)((double)__ldg(d_losses + 1) / (weighted ? sw : n_pos)) : 0.f;
  const float g_ctr = n_pos > 0.0 ? (float)((double)__ldg(d_losses + 2) / n_pos) : 0.f;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < R; g += (long long)gridDim.x * blockDim.x) {
    const RowCoord c = row_coord(lv, g);
    const int t = (int)labels[g];
    const long long cbase = (long long)c.n * num_classes * c.hw + c.p;
    for (int d = 0; d < num_classes; ++d) {
      const FocalT ft = fl_terms(__ldg(mp.cls[c.l] + cbase + d * c.hw));
      float v = 0.f;
      if (t == d + 1) v = -alpha * fl_pow(1.f - ft.p, gamma, sq) * (1.f - ft.p - ft.p * gamma * ft.log_p);
      else if (t >= 0) v = -(1.f - alpha) * fl_pow(ft.p, gamma, sq) * (ft.log_1mp * (1.f - ft.p) * gamma - ft.p);
      gr.cls[c.l][cbase + d * c.hw] = v * g_cls;
    }
    const long long rbase = (long long)c.n * 4 * c.hw + c.p;
    float dl = 0.f, dt = 0.f, dr = 0.f, db = 0.f, dc = 0.f;
    if (t > 0) {
      const float4 tg = reg_t[g];
      const float* rp = mp.reg[c.l] + rbase;
      const float pl = __ldg(rp), pt = __ldg(rp + c.hw), pr = __ldg(rp + 2 * c.hw), pb = __ldg(rp + 3 * c.hw);
      const float t_area = (tg.x + tg.z) * (tg.y + tg.w);
      const float wi = fminf(pl, tg.x) + fminf(pr, tg.z);
      const float hi = fminf(pb, tg.w) + fminf(pt, tg.y);
      const float inter = wi * hi;
      const float uni = t_area + (pl + pr) * (pt + pb) - inter;
      // L = -log(I + 1) + log(U + 1), U = At + Ap - I  ->  dL = -(1/(I+1) + 1/(U+1)) dI + 1/(U+1) dAp
      const float ci = -(1.f / (inter + 1.f) + 1.f / (uni + 1.f)), ca = 1.f / (uni + 1.f);
      // torch.min(a, b) backward: the smaller operand takes the gradient, a tie splits it evenly
      auto mind = [](float p, float q) { return p < q ? 1.f : (p == q ? 0.5f : 0.f); };
      const float w = weighted ? ctr_target(tg) : 1.f;
      const float s = g_reg * w;
      dl = s * (ci * mind(pl, tg.x) * hi + ca * (pt + pb));
      dr = s * (ci * mind(pr, tg.z) * hi + ca * (pt + pb));
      dt = s * (ci * mind(pt, tg.y) * wi + ca * (pl + pr));
      db = s * (ci * mind(pb, tg.w) * wi + ca * (pl + pr));
      const float x = __ldg(mp.ctr[c.l] + (long long)c.n * c.hw + c.p);
      dc = g_ctr * (1.f / (1.f + expf(-x)) - ctr_target(tg));
    }
    float* go = gr.reg[c.l] + rbase;
    go[0] = dl; go[c.hw] = dt; go[2 * c.hw] = dr; go[3 * c.hw] = db;
    gr.ctr[c.l][(long long)c.n * c.hw + c.p] = dc;
  }
}

}  // namespace scan

using namespace scan;

extern "C" int32_t scan_fcos_loss_num_partials(void) { return 4 * sm_count() * FL_NPART; }

static int fill_maps(const Levels& lv, const void* const* cls, const void* const* reg, const void* const* ctr, FcosMaps* m) {
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    const bool on = l < lv.n_levels;
    m->cls[l] = on ? (const float*)cls[l] : nullptr;
    m->reg[l] = on ? (const float*)reg[l] : nullptr;
    m->ctr[l] = on ? (const float*)ctr[l] : nullptr;
    if (on && (!m->cls[l] || !m->reg[l] || !m->ctr[l])) return SCAN_EINVAL;
  }
  return SCAN_OK;
}

extern "C" int scan_fcos_loss_fwd(const scan_levels_t* lvh, const void* const* cls_host, const void* const* reg_host, const void* const* ctr_host,
                                  const int64_t* labels, const float* reg_targets, int32_t num_classes, float gamma, float alpha,
                                  double* partials, double* sums6, float* losses3, void* stream) {
  Levels lv;
  int rc = make_levels(lvh, &lv);
  if (rc) return rc;
  if (!cls_host || !reg_host || !ctr_host || !labels || !reg_targets || !partials || !sums6 || !losses3 || num_classes < 1) return SCAN_EINVAL;
  FcosMaps m;
  if ((rc = fill_maps(lv, cls_host, reg_host, ctr_host, &m))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const int blocks = (int)std::min<long long>(ceil_div(R, 256), 4ll * sm_count());
  fcos_loss_fwd_kernel<<<blocks, 256, 0, st>>>(lv, m, labels, reinterpret_cast<const float4*>(reg_targets), num_classes, gamma, alpha, partials);
  SCAN_LAUNCH_CHECK("fcos_loss_fwd_kernel");
  fcos_loss_finalize_kernel<<<1, 32, 0, st>>>(partials, blocks, lv.n_images, sums6, losses3);
  SCAN_LAUNCH_CHECK("fcos_loss_finalize_kernel");
  return SCAN_OK;
}

extern "C" int scan_fcos_loss_bwd(const scan_levels_t* lvh, const void* const* cls_host, const void* const* reg_host, const void* const* ctr_host,
                                  const int64_t* labels, const float* reg_targets, int32_t num_classes, float gamma, float alpha,
                                  const double* sums6, const float* d_losses3, void* const* d_cls_host, void* const* d_reg_host,
                                  void* const* d_ctr_host, void* stream) {
  Levels lv;
  int rc = make_levels(lvh, &lv);
  if (rc) return rc;
  if (!cls_host || !reg_host || !ctr_host || !labels || !reg_targets || !sums6 || !d_losses3 || !d_cls_host || !d_reg_host || !d_ctr_host)
    return SCAN_EINVAL;
  FcosMaps m;
  if ((rc = fill_maps(lv, cls_host, reg_host, ctr_host, &m))) return rc;
  FcosGrads g;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    const bool on = l < lv.n_levels;
    g.cls[l] = on ? (float*)d_cls_host[l] : nullptr;
    g.reg[l] = on ? (float*)d_reg_host[l] : nullptr;
    g.ctr[l] = on ? (float*)d_ctr_host[l] : nullptr;
    if (on && (!g.cls[l] || !g.reg[l] || !g.ctr[l])) return SCAN_EINVAL;
  }
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const int blocks = (int)std::min<long long>(ceil_div(R, 256), 8ll * sm_count());
  fcos_loss_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(lv, m, g, labels, reinterpret_cast<const float4*>(reg_targets), num_classes,
                                                                 gamma, alpha, lv.n_images, sums6, d_losses3);
  SCAN_LAUNCH_CHECK("fcos_loss_bwd_kernel");
  return SCAN_OK;
}
