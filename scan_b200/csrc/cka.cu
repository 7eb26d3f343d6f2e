// f3: the thin kernels of the CKA discriminator FCOSDiscriminator_con (modeling/discriminator/
// fcos_head_discriminator_con.py:88-127; gradient reversal layer.py:6-24).  Its convolutions run on csrc/tower.cu:
//   * dis_tower ([Conv3x3 + GN(32) + ReLU] x num_convs, :20-33) = the head_in kernels;
//   * the per-class loop `for c: Conv3x3(cat(x, act_c): 257 -> 128) + ReLU + Conv3x3(128 -> 1)` (:47-63, :100-112) as TWO
//     launches for ALL classes: a block-structured [C * 128, 256 + C] weight over the two inputs (features | the C maps,
//     scan_conv3x3_rows2: the concatenation is never built), then a block-diagonal [C, C * 128] one.
// What is left are the class-map layout changes, the class-weighted BCE-with-logits (:113-120), column sums for the bias
// gradients and the gradient scaling of the reversal layer.  HBM-streaming kernels, one thread (or one float4) per element.
#include "common.cuh"

namespace scan {

struct ThinArgs {
  const float* nchw[SCAN_MAX_LEVELS];   // per level [N, k_total, H, W]
};

// NCHW channels [c0, c0 + k) -> rows [R, ld] columns [0, k); columns [k, ld) are zero-filled.  ld % 4 == 0.
// One thread per float4 of the output (consecutive threads = consecutive 16-byte pieces of a row: coalesced writes).
__global__ void __launch_bounds__(256) thin_pack_kernel(Levels lv, ThinArgs a, int k_total, int c0, int k, float* __restrict__ rows, int ld) {
  const int q = ld >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long g = t / q;
  if (g >= lv.row_off[lv.n_levels]) return;
  const int j4 = (int)(t - g * q) * 4;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (j4 < k) {
    const int l = level_of_row(lv, g);
    const long long hw = (long long)lv.h[l] * lv.w[l];
    const long long local = g - lv.row_off[l];
    const long long n = local / hw, p = local - n * hw;
    const float* src = a.nchw[l] + (n * k_total + c0) * hw + p;
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (j4 + e < k) v[e] = __ldg(src + (long long)(j4 + e) * hw);
  }
  *reinterpret_cast<float4*>(rows + g * ld + j4) = make_float4(v[0], v[1], v[2], v[3]);
}

// rows [R, ld] columns [0, k) * scale -> NCHW channels [c0, c0 + k) (the other channels are the caller's: zero-initialised)
__global__ void __launch_bounds__(256) thin_unpack_kernel(Levels lv, ThinArgs a, int k_total, int c0, int k, const float* __restrict__ rows,
                                                          int ld, float scale) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= lv.row_off[lv.n_levels]) return;
  const int l = level_of_row(lv, g);
  const long long hw = (long long)lv.h[l] * lv.w[l];
  const long long local = g - lv.row_off[l];
  const long long n = local / hw, p = local - n * hw;
  float* dst = const_cast<float*>(a.nchw[l]) + (n * k_total + c0) * hw + p;
  const float* src = rows + g * ld;
  for (int j = 0; j < k; ++j) dst[(long long)j * hw] = scale * __ldg(src + j);
}

// Inverse of the tap spread (tower.cu): out32[p, k] = sum_tap d[p - off(tap), k * 9 + tap] over the neighbours inside the image.
// d [R, ldd] holds, per pixel q, the K * 9 products "what q sends to its neighbour through tap" (one GEMM of the pixel rows
// against the [K * 9, C] weight slice); gathering them is the thin data gradient of a 3x3 convolution with K <= 14 inputs.
__global__ void __launch_bounds__(256) thin_gather_kernel(Levels lv, const float* __restrict__ d, int ldd, int k, float* __restrict__ out32) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long g = t >> 5;
  if (g >= lv.row_off[lv.n_levels]) return;
  const int kk = (int)(t & 31);
  float s = 0.f;
  if (kk < k) {
    const int l = level_of_row(lv, g);
    const int w = lv.w[l], h = lv.h[l];
    const long long local = g - lv.row_off[l];
    const int p = (int)(local % ((long long)h * w));
    const int y = p / w, x = p - y * w;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y - (tap / 3 - 1), xx = x - (tap % 3 - 1);
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) s += __ldg(d + (g + (long long)(yy - y) * w + (xx - x)) * ldd + kk * 9 + tap);
    }
  }
  out32[g * 32 + kk] = s;
}

__global__ void __launch_bounds__(256) scale_kernel(const float4* __restrict__ x, long long n4, float scale, float4* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    y[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
  }
}

// ---------------------------------------------------------------------------- column sums (bias gradients)
constexpr int CS_ROWS = 1024;   // rows per block

// partial[chunk][col] = sum over the chunk's rows of x[row, col]; 256 threads = 64 float4 columns x 4 row lanes
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, long long n_rows, int n_cols, int ld,
                                                             float* __restrict__ partial) {
  __shared__ float4 red[4][64];
  const int c4 = blockIdx.y * 64 + (threadIdx.x & 63), rl = threadIdx.x >> 6;
  const long long r0 = (long long)blockIdx.x * CS_ROWS, r1 = min(r0 + CS_ROWS, n_rows);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c4 * 4 < n_cols)
    for (long long r = r0 + rl; r < r1; r += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld) + c4);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  red[rl][threadIdx.x & 63] = s;
  __syncthreads();
  if (rl == 0 && c4 * 4 < n_cols) {
    const float4 a = red[0][threadIdx.x], b = red[1][threadIdx.x], c = red[2][threadIdx.x], d = red[3][threadIdx.x];
    float* o = partial + (long long)blockIdx.x * n_cols + c4 * 4;
    const float v[4] = {(a.x + b.x) + (c.x + d.x), (a.y + b.y) + (c.y + d.y), (a.z + b.z) + (c.z + d.z), (a.w + b.w) + (c.w + d.w)};
    for (int e = 0; e < 4; ++e)
      if (c4 * 4 + e < n_cols) o[e] = v[e];
  }
}
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int n_chunks, int n_cols, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cols) return;
  double s = 0.0;
  for (int i = 0; i < n_chunks; ++i) s += (double)__ldg(partial + (long long)i * n_cols + c);
  out[c] = (float)s;
}

// ---------------------------------------------------------------------------- class-weighted BCE with logits
// fcos_head_discriminator_con.py:113-121.  logits [R, ldl] columns [0, C), weights [R, ldw] columns [0, C) (the class maps).
//   C_total > 1:  loss = sum_c [ sum_p w bce(x, t) / sum_p w ] / C_total
//   C_total == 1: loss = mean_p bce(x, t)                                   (F.binary_cross_entropy_with_logits default)
__device__ __forceinline__ float bce_logits(float x, float t) { return fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x))); }

constexpr int BCE_MAX_C = 16;
constexpr int BCE_BLOCKS = 592;

__global__ void __launch_bounds__(256) bce_partial_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ w, int ldw,
                                                          long long n_rows, int n_cls, float target, int weighted,
                                                          double* __restrict__ partial /* [blocks][2 * BCE_MAX_C] */) {
  __shared__ double red[8][2 * BCE_MAX_C];
  double sl[BCE_MAX_C], sw[BCE_MAX_C];
#pragma unroll
  for (int c = 0; c < BCE_MAX_C; ++c) sl[c] = sw[c] = 0.0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < BCE_MAX_C; ++c)
      if (c < n_cls) {
        const float x = __ldg(logits + r * ldl + c);
        const float wt = weighted ? __ldg(w + r * ldw + c) : 1.f;
        sl[c] += (double)(wt * bce_logits(x, target));
        sw[c] += (double)wt;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < BCE_MAX_C; ++c) {
    const double a = warp_sum_d(sl[c]), b = warp_sum_d(sw[c]);
    if (lane == 0) {
      red[warp][c] = a;
      red[warp][BCE_MAX_C + c] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * BCE_MAX_C) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    partial[(long long)blockIdx.x * 2 * BCE_MAX_C + threadIdx.x] = s;
  }
}

// loss and the per-class gradient factor inv[c] = 1 / (sum_p w_c * C_total)  (or 1 / R for the unweighted mean)
__global__ void bce_finalize_kernel(const double* __restrict__ partial, int n_blocks, int n_cls, int weighted, float* __restrict__ loss,
                                    float* __restrict__ inv) {
  __shared__ double tot[2 * BCE_MAX_C];
  if (threadIdx.x < 2 * BCE_MAX_C) {
    double s = 0.0;
    for (int i = 0; i < n_blocks; ++i) s += partial[(long long)i * 2 * BCE_MAX_C + threadIdx.x];
    tot[threadIdx.x] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0;
    for (int c = 0; c < n_cls; ++c) {
      // the reference sums in fp32: loss_c = float(sum) / float(sum_w), then loss += loss_c / C_total
      const float lc = (float)tot[c] / (float)tot[BCE_MAX_C + c];
      l += (double)(lc / (float)n_cls);
      inv[c] = 1.f / ((float)tot[BCE_MAX_C + c] * (float)n_cls);
    }
    (void)weighted;
    *loss = (float)l;
  }
}

// d_logits[p, c] = d_loss * (sigmoid(x) - t) * w * inv[c]; written to dl32 [R, 32] (zero padded) and, if given, dl_wide [R, ld_wide]
__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ w, int ldw,
                                                      long long n_rows, int n_cls, float target, int weighted, const float* __restrict__ inv,
                                                      const float* __restrict__ d_loss, float* __restrict__ dl32, float* __restrict__ dl_wide,
                                                      int ld_wide) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const float dl = __ldg(d_loss);
  float v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    v[c] = 0.f;
    if (c < n_cls) {
      const float x = __ldg(logits + r * ldl + c);
      const float wt = weighted ? __ldg(w + r * ldw + c) : 1.f;
      v[c] = dl * (1.f / (1.f + expf(-x)) - target) * wt * __ldg(inv + c);
    }
  }
#pragma unroll
  for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dl32 + r * 32 + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  if (dl_wide) {
#pragma unroll
    for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dl_wide + r * ld_wide + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
  }
}

}  // namespace scan

using namespace scan;

static int thin_args(const Levels& lv, const void* const* ptrs, ThinArgs* a) {
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    a->nchw[l] = nullptr;
    if (l < lv.n_levels) {
      if (!ptrs[l]) return SCAN_EINVAL;
      a->nchw[l] = (const float*)ptrs[l];
    }
  }
  return SCAN_OK;
}

extern "C" int scan_thin_pack(const scan_levels_t* levels, const void* const* nchw_host, int32_t k_total, int32_t c0, int32_t k, float* rows,
                              int32_t ld, void* stream) {
  Levels lv;
  int rc = make_levels(levels, &lv);
  if (rc) return rc;
  ThinArgs a;
  if (!nchw_host || !rows || k < 1 || c0 < 0 || c0 + k > k_total || ld < k || (ld % 4) || ((uintptr_t)rows & 15)) return SCAN_EINVAL;
  rc = thin_args(lv, nchw_host, &a);
  if (rc) return rc;
  const long long R = lv.row_off[lv.n_levels];
  thin_pack_kernel<<<(unsigned)ceil_div(R * (ld / 4), 256), 256, 0, (cudaStream_t)stream>>>(lv, a, k_total, c0, k, rows, ld);
  SCAN_LAUNCH_CHECK("thin_pack_kernel");
  return SCAN_OK;
}

extern "C" int scan_thin_unpack(const scan_levels_t* levels, const float* rows, int32_t ld, int32_t k_total, int32_t c0, int32_t k, float scale,
                                void* const* nchw_host, void* stream) {
  Levels lv;
  int rc = make_levels(levels, &lv);
  if (rc) return rc;
  ThinArgs a;
  if (!nchw_host || !rows || k < 1 || c0 < 0 || c0 + k > k_total || ld < k) return SCAN_EINVAL;
  rc = thin_args(lv, (const void* const*)nchw_host, &a);
  if (rc) return rc;
  const long long R = lv.row_off[lv.n_levels];
  thin_unpack_kernel<<<(unsigned)ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(lv, a, k_total, c0, k, rows, ld, scale);
  SCAN_LAUNCH_CHECK("thin_unpack_kernel");
  return SCAN_OK;
}

extern "C" int scan_scale(const float* x, int64_t n, float scale, float* y, void* stream) {
  if (!x || !y || n < 0 || (n & 3) || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return SCAN_EINVAL;
  if (n == 0) return SCAN_OK;
  long long blocks = ceil_div(n / 4, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  scale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), n / 4, scale,
                                                                    reinterpret_cast<float4*>(y));
  SCAN_LAUNCH_CHECK("scale_kernel");
  return SCAN_OK;
}

extern "C" int64_t scan_colsum_workspace_bytes(int64_t n_rows, int32_t n_cols) {
  return (int64_t)ceil_div(n_rows, CS_ROWS) * n_cols * 4;
}

// out[c] = sum_r x[r, c] for c < n_cols (deterministic two-level sum); ld % 4 == 0, x 16-byte aligned
extern "C" int scan_colsum(const float* x, int64_t n_rows, int32_t n_cols, int32_t ld, float* out, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  if (!x || !out || !workspace || n_rows < 1 || n_cols < 1 || ld < n_cols || (ld % 4) || ((uintptr_t)x & 15)) return SCAN_EINVAL;
  const int chunks = (int)ceil_div(n_rows, CS_ROWS);
  if (workspace_bytes < (int64_t)chunks * n_cols * 4) return SCAN_EINVAL;
  colsum_partial_kernel<<<dim3((unsigned)chunks, (unsigned)ceil_div(n_cols, 256)), 256, 0, (cudaStream_t)stream>>>(x, n_rows, n_cols, ld,
                                                                                                                  (float*)workspace);
  SCAN_LAUNCH_CHECK("colsum_partial_kernel");
  colsum_final_kernel<<<(unsigned)ceil_div(n_cols, 256), 256, 0, (cudaStream_t)stream>>>((const float*)workspace, chunks, n_cols, out);
  SCAN_LAUNCH_CHECK("colsum_final_kernel");
  return SCAN_OK;
}

extern "C" int64_t scan_cka_bce_workspace_bytes(void) { return (int64_t)BCE_BLOCKS * 2 * BCE_MAX_C * 8; }

// loss [1], inv [16] (the per-class gradient factors scan_cka_bce_bwd needs); weights may be NULL when n_cls == 1 (plain mean)
extern "C" int scan_cka_bce_fwd(const float* logits, int32_t ldl, const float* weights, int32_t ldw, int64_t n_rows, int32_t n_cls, float target,
                                float* loss, float* inv, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!logits || !loss || !inv || !workspace || n_rows < 1 || n_cls < 1 || n_cls > BCE_MAX_C) return SCAN_EINVAL;
  if (workspace_bytes < scan_cka_bce_workspace_bytes()) return SCAN_EINVAL;
  const int weighted = n_cls > 1;
  if (weighted && !weights) return SCAN_EINVAL;
  int blocks = (int)ceil_div(n_rows, 256);
  if (blocks > BCE_BLOCKS) blocks = BCE_BLOCKS;
  bce_partial_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, ldl, weights, ldw, n_rows, n_cls, target, weighted, (double*)workspace);
  SCAN_LAUNCH_CHECK("bce_partial_kernel");
  bce_finalize_kernel<<<1, 64, 0, (cudaStream_t)stream>>>((const double*)workspace, blocks, n_cls, weighted, loss, inv);
  SCAN_LAUNCH_CHECK("bce_finalize_kernel");
  return SCAN_OK;
}

extern "C" int scan_cka_bce_bwd(const float* logits, int32_t ldl, const float* weights, int32_t ldw, int64_t n_rows, int32_t n_cls, float target,
                                const float* inv, const float* d_loss, float* dl32, float* dl_wide, int32_t ld_wide, void* stream) {
  if (!logits || !inv || !d_loss || !dl32 || n_rows < 1 || n_cls < 1 || n_cls > BCE_MAX_C || ((uintptr_t)dl32 & 15)) return SCAN_EINVAL;
  if (dl_wide && (ld_wide < 32 || (ld_wide % 4) || ((uintptr_t)dl_wide & 15))) return SCAN_EINVAL;
  const int weighted = n_cls > 1;
  if (weighted && !weights) return SCAN_EINVAL;
  bce_bwd_kernel<<<(unsigned)ceil_div(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(logits, ldl, weights, ldw, n_rows, n_cls, target, weighted,
                                                                                   inv, d_loss, dl32, dl_wide, ld_wide);
  SCAN_LAUNCH_CHECK("bce_bwd_kernel");
  return SCAN_OK;
}

// out32 [R, 32] (columns >= k zero) = tap gather of d [R, ldd] (columns k * 9 + tap), see thin_gather_kernel
extern "C" int scan_thin_gather(const scan_levels_t* levels, const float* d, int32_t ldd, int32_t k, float* out32, void* stream) {
  Levels lv;
  int rc = make_levels(levels, &lv);
  if (rc) return rc;
  if (!d || !out32 || k < 1 || k > 32 || ldd < 9 * k) return SCAN_EINVAL;
  const long long R = lv.row_off[lv.n_levels];
  thin_gather_kernel<<<(unsigned)ceil_div(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(lv, d, ldd, k, out32);
  SCAN_LAUNCH_CHECK("thin_gather_kernel");
  return SCAN_OK;
}
