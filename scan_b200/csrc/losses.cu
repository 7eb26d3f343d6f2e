// K5: sigmoid focal loss (the reference's only native kernel next to the path, csrc/cuda/SigmoidFocalLoss_cuda.cu:21-101:
// the FORMULAS of :36-55 / :78-98 are the contract) and TEST.MODE map ensembling (modeling/rpn/fcos/fcos.py:162-169,
// inference.py:68).  HBM-bound streaming kernels.
//
// Design (not the reference's one-scalar-thread-per-(row, class) grid with a div/mod and a target re-read per element):
//   * thread = location row: the target is read ONCE per row, the C logits of the row with 128-bit loads (C = 8 for the
//     Cityscapes configs: two float4, C templated for 1 / 2 / 4 / 8, generic loop otherwise), results stored the same way;
//   * one exp and one log per element: with e = exp(-|x|), p = sigmoid(x) = (x >= 0 ? 1 : e) / (1 + e),
//     log(1 + exp(x - 2 x [x >= 0])) = log(1 + e) (the reference's stable log(1 - p) term) and log p = min(x, 0) - log(1 + e);
//     the reference's clamp log(max(p, FLT_MIN)) is kept for p < FLT_MIN; gamma == 2 (defaults.py:346) squares instead of powf;
//   * forward and backward share the per-element core (`focal_terms`), so the saved tensors are just logits + targets.
// The ensembling kernel processes ALL FPN levels in one launch (level descriptors by value); 'light' mode is a view, not a kernel.
#include "common.cuh"

namespace scan {

struct FocalTerms {
  float p, log_p, log_1mp;   // sigmoid(x), log(max(p, FLT_MIN)), log(1 - p) in the reference's stable form
};
__device__ __forceinline__ FocalTerms focal_terms(float x) {
  const float e = expf(-fabsf(x));
  const float inv = 1.f / (1.f + e);
  const float l1pe = logf(1.f + e);
  FocalTerms t;
  t.p = (x >= 0.f ? 1.f : e) * inv;
  t.log_p = t.p >= 1.17549435e-38f ? fminf(x, 0.f) - l1pe : -87.33654475f;   // logf(FLT_MIN)
  t.log_1mp = -fmaxf(x, 0.f) - l1pe;                                          // -x [x >= 0] - log(1 + exp(x - 2 x [x >= 0]))
  return t;
}
__device__ __forceinline__ float pow_gamma(float v, float gamma, bool square) { return square ? v * v : powf(v, gamma); }

// SigmoidFocalLoss_cuda.cu:36-55: loss = -[t == d+1] alpha (1-p)^g log p - [t >= 0, t != d+1] (1-alpha) p^g log(1-p)
__device__ __forceinline__ float focal_fwd_one(float x, int t, int d, float gamma, float alpha, bool sq) {
  const FocalTerms f = focal_terms(x);
  if (t == d + 1) return -alpha * pow_gamma(1.f - f.p, gamma, sq) * f.log_p;
  if (t >= 0) return -(1.f - alpha) * pow_gamma(f.p, gamma, sq) * f.log_1mp;
  return 0.f;
}
// SigmoidFocalLoss_cuda.cu:78-98
__device__ __forceinline__ float focal_bwd_one(float x, int t, int d, float gamma, float alpha, bool sq, float g) {
  const FocalTerms f = focal_terms(x);
  if (t == d + 1) return -alpha * pow_gamma(1.f - f.p, gamma, sq) * (1.f - f.p - f.p * gamma * f.log_p) * g;
  if (t >= 0) return -(1.f - alpha) * pow_gamma(f.p, gamma, sq) * (f.log_1mp * (1.f - f.p) * gamma - f.p) * g;
  return 0.f;
}

template <int C, bool BWD>
__global__ void __launch_bounds__(256) sigmoid_focal_rows_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                                 const float* __restrict__ d_losses, long long n_rows, int num_classes,
                                                                 float gamma, float alpha, float* __restrict__ out) {
  const bool sq = gamma == 2.f;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    const int t = __ldg(targets + r);
    if constexpr (C >= 4) {
#pragma unroll
      for (int v = 0; v < C / 4; ++v) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(logits + r * C) + v);
        float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
        if constexpr (BWD) g = __ldg(reinterpret_cast<const float4*>(d_losses + r * C) + v);
        float4 o;
        if constexpr (BWD) {
          o.x = focal_bwd_one(x.x, t, 4 * v + 0, gamma, alpha, sq, g.x);
          o.y = focal_bwd_one(x.y, t, 4 * v + 1, gamma, alpha, sq, g.y);
          o.z = focal_bwd_one(x.z, t, 4 * v + 2, gamma, alpha, sq, g.z);
          o.w = focal_bwd_one(x.w, t, 4 * v + 3, gamma, alpha, sq, g.w);
        } else {
          o.x = focal_fwd_one(x.x, t, 4 * v + 0, gamma, alpha, sq);
          o.y = focal_fwd_one(x.y, t, 4 * v + 1, gamma, alpha, sq);
          o.z = focal_fwd_one(x.z, t, 4 * v + 2, gamma, alpha, sq);
          o.w = focal_fwd_one(x.w, t, 4 * v + 3, gamma, alpha, sq);
        }
        reinterpret_cast<float4*>(out + r * C)[v] = o;
      }
    } else if constexpr (C == 2) {
      const float2 x = __ldg(reinterpret_cast<const float2*>(logits + r * 2));
      float2 g = make_float2(1.f, 1.f), o;
      if constexpr (BWD) g = __ldg(reinterpret_cast<const float2*>(d_losses + r * 2));
      if constexpr (BWD) {
        o.x = focal_bwd_one(x.x, t, 0, gamma, alpha, sq, g.x);
        o.y = focal_bwd_one(x.y, t, 1, gamma, alpha, sq, g.y);
      } else {
        o.x = focal_fwd_one(x.x, t, 0, gamma, alpha, sq);
        o.y = focal_fwd_one(x.y, t, 1, gamma, alpha, sq);
      }
      *reinterpret_cast<float2*>(out + r * 2) = o;
    } else {   // C == 0: generic class count (scalar accesses, still one target read and no div/mod per element)
      const int nc = C == 1 ? 1 : num_classes;
      for (int d = 0; d < nc; ++d) {
        const float x = __ldg(logits + r * nc + d);
        if constexpr (BWD)
          out[r * nc + d] = focal_bwd_one(x, t, d, gamma, alpha, sq, __ldg(d_losses + r * nc + d));
        else
          out[r * nc + d] = focal_fwd_one(x, t, d, gamma, alpha, sq);
      }
    }
  }
}

// TEST.MODE ensembling of every FPN level in one launch: out[l][n, c, p] over the K-1 foreground channels
struct EnsArgs {
  const float* cls[SCAN_MAX_LEVELS];
  const float* act[SCAN_MAX_LEVELS];
  float* out[SCAN_MAX_LEVELS];
  long long hw[SCAN_MAX_LEVELS];
  long long off[SCAN_MAX_LEVELS + 1];   // element offsets of the levels in the flat (n, c, p) index space
  int n_levels;
};
__global__ void __launch_bounds__(256) ensemble_levels_kernel(EnsArgs a, int n_images, int num_classes, int mode) {
  const int fg = num_classes - 1;
  const long long total = a.off[a.n_levels];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int l = 0;
#pragma unroll
    for (int j = 1; j < SCAN_MAX_LEVELS; ++j)
      if (j < a.n_levels && i >= a.off[j]) l = j;
    const long long e = i - a.off[l], hw = a.hw[l];
    const long long n = e / (fg * hw);
    const long long rem = e - n * fg * hw;   // c * hw + p
    float v;
    if (mode == 1) {
      v = __ldg(a.act[l] + n * num_classes * hw + hw + rem);
    } else {
      const float s = 1.f / (1.f + expf(-__ldg(a.cls[l] + e)));
      v = (mode == 2) ? (0.5f * s + 0.5f * __ldg(a.act[l] + n * num_classes * hw + hw + rem)) : s;
    }
    a.out[l][e] = v;
  }
}

static inline int stream_grid(long long total) {
  long long b = ceil_div(total, 256);
  long long cap = 16ll * sm_count();
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace scan

template <bool BWD>
static int launch_focal(const float* logits, const int32_t* targets, const float* d_losses, int64_t n_rows, int32_t num_classes,
                        float gamma, float alpha, float* out, cudaStream_t st) {
  using namespace scan;
  const int grid = stream_grid(n_rows);
  const bool vec = !((uintptr_t)logits & 15) && !((uintptr_t)out & 15) && !((uintptr_t)d_losses & 15);
  if (num_classes == 8 && vec)
    sigmoid_focal_rows_kernel<8, BWD><<<grid, 256, 0, st>>>(logits, targets, d_losses, n_rows, num_classes, gamma, alpha, out);
  else if (num_classes == 4 && vec)
    sigmoid_focal_rows_kernel<4, BWD><<<grid, 256, 0, st>>>(logits, targets, d_losses, n_rows, num_classes, gamma, alpha, out);
  else if (num_classes == 2 && vec)
    sigmoid_focal_rows_kernel<2, BWD><<<grid, 256, 0, st>>>(logits, targets, d_losses, n_rows, num_classes, gamma, alpha, out);
  else if (num_classes == 1)
    sigmoid_focal_rows_kernel<1, BWD><<<grid, 256, 0, st>>>(logits, targets, d_losses, n_rows, num_classes, gamma, alpha, out);
  else
    sigmoid_focal_rows_kernel<0, BWD><<<grid, 256, 0, st>>>(logits, targets, d_losses, n_rows, num_classes, gamma, alpha, out);
  SCAN_LAUNCH_CHECK("sigmoid_focal_rows_kernel");
  return SCAN_OK;
}

extern "C" int scan_sigmoid_focal_fwd(const float* logits, const int32_t* targets, int64_t n_rows, int32_t num_classes,
                                      float gamma, float alpha, float* losses, void* stream) {
  if (n_rows == 0) return SCAN_OK;  // SigmoidFocalLoss_cuda.cu:123-126
  if (!logits || !targets || !losses || n_rows < 0 || num_classes < 1) return SCAN_EINVAL;
  return launch_focal<false>(logits, targets, nullptr, n_rows, num_classes, gamma, alpha, losses, (cudaStream_t)stream);
}

extern "C" int scan_sigmoid_focal_bwd(const float* logits, const int32_t* targets, const float* d_losses, int64_t n_rows,
                                      int32_t num_classes, float gamma, float alpha, float* d_logits, void* stream) {
  if (n_rows == 0) return SCAN_OK;
  if (!logits || !targets || !d_losses || !d_logits || n_rows < 0 || num_classes < 1) return SCAN_EINVAL;
  return launch_focal<true>(logits, targets, d_losses, n_rows, num_classes, gamma, alpha, d_logits, (cudaStream_t)stream);
}

extern "C" int scan_ensemble_levels(const scan_levels_t* lvh, const void* const* cls_logits_host, const void* const* act_host,
                                    int32_t num_classes, int32_t mode, void* const* out_host, void* stream) {
  using namespace scan;
  Levels lv;
  int rc = make_levels(lvh, &lv);
  if (rc) return rc;
  if (!out_host || num_classes < 2 || mode < 0 || mode > 2) return SCAN_EINVAL;
  if ((mode != 1 && !cls_logits_host) || (mode != 0 && !act_host)) return SCAN_EINVAL;
  EnsArgs a;
  a.n_levels = lv.n_levels;
  long long off = 0;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    a.cls[l] = (l < lv.n_levels && cls_logits_host) ? (const float*)cls_logits_host[l] : nullptr;
    a.act[l] = (l < lv.n_levels && act_host) ? (const float*)act_host[l] : nullptr;
    a.out[l] = l < lv.n_levels ? (float*)out_host[l] : nullptr;
    a.hw[l] = l < lv.n_levels ? (long long)lv.h[l] * lv.w[l] : 0;
    a.off[l] = off;
    if (l < lv.n_levels) {
      if (!a.out[l] || (mode != 1 && !a.cls[l]) || (mode != 0 && !a.act[l])) return SCAN_EINVAL;
      off += (long long)lv.n_images * (num_classes - 1) * a.hw[l];
    }
  }
  a.off[SCAN_MAX_LEVELS] = off;
  for (int l = lv.n_levels; l <= SCAN_MAX_LEVELS; ++l) a.off[l] = off;
  ensemble_levels_kernel<<<stream_grid(off), 256, 0, (cudaStream_t)stream>>>(a, lv.n_images, num_classes, mode);
  SCAN_LAUNCH_CHECK("ensemble_levels_kernel");
  return SCAN_OK;
}
