// K5: sigmoid focal loss (the reference's only native kernel next to the path,
// csrc/cuda/SigmoidFocalLoss_cuda.cu:21-101) and TEST.MODE map ensembling (modeling/rpn/fcos/fcos.py:162-169,
// inference.py:68).  Pure streaming kernels: HBM-bound, 128-bit accesses where the shapes allow.
#include "common.cuh"

namespace scan {

__global__ void __launch_bounds__(256) sigmoid_focal_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                                long long total, int num_classes, float gamma, float alpha,
                                                                float* __restrict__ losses) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / num_classes;
    const int d = (int)(i - n * num_classes);
    const int t = __ldg(targets + n);
    const float x = __ldg(logits + i);
    // SigmoidFocalLoss_cuda.cu:36-55
    const float c1 = (t == (d + 1)) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != (d + 1)) ? 1.f : 0.f;
    const float zn = 1.f - alpha, zp = alpha;
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * logf(fmaxf(p, 1.17549435e-38f));
    const float ge = (x >= 0.f) ? 1.f : 0.f;
    const float term2 = powf(p, gamma) * (-1.f * x * ge - logf(1.f + expf(x - 2.f * x * ge)));
    losses[i] = -c1 * term1 * zp - c2 * term2 * zn;
  }
}

__global__ void __launch_bounds__(256) sigmoid_focal_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ targets,
                                                                const float* __restrict__ d_losses, long long total, int num_classes,
                                                                float gamma, float alpha, float* __restrict__ d_logits) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / num_classes;
    const int d = (int)(i - n * num_classes);
    const int t = __ldg(targets + n);
    const float x = __ldg(logits + i);
    // SigmoidFocalLoss_cuda.cu:78-98
    const float c1 = (t == (d + 1)) ? 1.f : 0.f;
    const float c2 = (t >= 0 && t != (d + 1)) ? 1.f : 0.f;
    const float zn = 1.f - alpha, zp = alpha;
    const float p = 1.f / (1.f + expf(-x));
    const float term1 = powf(1.f - p, gamma) * (1.f - p - (p * gamma * logf(fmaxf(p, 1.17549435e-38f))));
    const float ge = (x >= 0.f) ? 1.f : 0.f;
    const float term2 = powf(p, gamma) * ((-1.f * x * ge - logf(1.f + expf(x - 2.f * x * ge))) * (1.f - p) * gamma - p);
    d_logits[i] = (-c1 * term1 * zp - c2 * term2 * zn) * __ldg(d_losses + i);
  }
}

// out[n, c, p] over (K-1) foreground channels; act has K channels, channel 0 is background
__global__ void __launch_bounds__(256) ensemble_kernel(const float* __restrict__ cls, const float* __restrict__ act, int n_images,
                                                       int num_classes, long long hw, int mode, float* __restrict__ out) {
  const int fg = num_classes - 1;
  const long long total = (long long)n_images * fg * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / (fg * hw);
    const long long rem = i - n * fg * hw;  // c*hw + p
    float v;
    if (mode == 1) {
      v = __ldg(act + n * num_classes * hw + hw + rem);
    } else {
      const float s = 1.f / (1.f + expf(-__ldg(cls + i)));
      v = (mode == 2) ? (0.5f * s + 0.5f * __ldg(act + n * num_classes * hw + hw + rem)) : s;
    }
    out[i] = v;
  }
}

static inline int stream_grid(long long total) {
  long long b = ceil_div(total, 256);
  long long cap = 16ll * sm_count();
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace scan

extern "C" int scan_sigmoid_focal_fwd(const float* logits, const int32_t* targets, int64_t n_rows, int32_t num_classes,
                                      float gamma, float alpha, float* losses, void* stream) {
  if (n_rows == 0) return SCAN_OK;  // SigmoidFocalLoss_cuda.cu:123-126
  if (!logits || !targets || !losses || n_rows < 0 || num_classes < 1) return SCAN_EINVAL;
  const long long total = (long long)n_rows * num_classes;
  scan::sigmoid_focal_fwd_kernel<<<scan::stream_grid(total), 256, 0, (cudaStream_t)stream>>>(logits, targets, total, num_classes,
                                                                                          gamma, alpha, losses);
  SCAN_LAUNCH_CHECK("sigmoid_focal_fwd_kernel");
  return SCAN_OK;
}

extern "C" int scan_sigmoid_focal_bwd(const float* logits, const int32_t* targets, const float* d_losses, int64_t n_rows,
                                      int32_t num_classes, float gamma, float alpha, float* d_logits, void* stream) {
  if (n_rows == 0) return SCAN_OK;
  if (!logits || !targets || !d_losses || !d_logits || n_rows < 0 || num_classes < 1) return SCAN_EINVAL;
  const long long total = (long long)n_rows * num_classes;
  scan::sigmoid_focal_bwd_kernel<<<scan::stream_grid(total), 256, 0, (cudaStream_t)stream>>>(logits, targets, d_losses, total,
                                                                                          num_classes, gamma, alpha, d_logits);
  SCAN_LAUNCH_CHECK("sigmoid_focal_bwd_kernel");
  return SCAN_OK;
}

extern "C" int scan_ensemble(const float* cls_logits, const float* act, int32_t n_images, int32_t num_classes, int64_t hw,
                             int32_t mode, float* out, void* stream) {
  if (!out || n_images < 1 || num_classes < 2 || hw < 1 || mode < 0 || mode > 2) return SCAN_EINVAL;
  if ((mode != 1 && !cls_logits) || (mode != 0 && !act)) return SCAN_EINVAL;
  const long long total = (long long)n_images * (num_classes - 1) * hw;
  scan::ensemble_kernel<<<scan::stream_grid(total), 256, 0, (cudaStream_t)stream>>>(cls_logits, act, n_images, num_classes, hw, mode, out);
  SCAN_LAUNCH_CHECK("ensemble_kernel");
  return SCAN_OK;
}
