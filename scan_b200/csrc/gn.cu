// f1: the GroupNorm(32) + ReLU of the head_in towers (condgraph.py:100-107 under :68-119) on the rows layout.
// cuDNN runs the 3x3 convolutions on channels-last tensors, i.e. on [pixels, 256] rows; these kernels normalise the five
// FPN levels of a call in ONE launch each and write the result straight into the contiguous [R, 256] rows matrix the rest
// of the path works on (no NCHW<->NHWC conversion, no separate ReLU pass, no pack).  HBM-bound streaming kernels:
//   forward   stats pass (read x) -> finalize (fp64) -> apply pass (read x, write y)
//   backward  reduce pass (read x, dy) -> finalize (fp64) -> apply pass (read x, dy, write dx); the ReLU mask is recomputed
//             from x with the forward's own operations, so y is never read back
// A block owns 128 consecutive pixels of one (level, image); a thread owns 4 channels (half a group of 8) of every 4th
// pixel, so each pixel row is one 1 KB coalesced access.  Partials are per block, combined in fp64: deterministic.
#include "common.cuh"

namespace scan {

constexpr int GN_ROWS = 128;     // pixels per block
constexpr int GN_C = 256;
constexpr int GN_G = 32;         // groups of 8 channels

struct GnLevels {
  const float* x[SCAN_MAX_LEVELS];     // conv output of level l, NHWC-dense [N * H_l * W_l, 256]
  const float* dy[SCAN_MAX_LEVELS];    // backward only: upstream gradient per level, same layout
  int chunks[SCAN_MAX_LEVELS];         // blocks per image of the level
  int blk_off[SCAN_MAX_LEVELS + 1];    // first block of the level
};

struct GnBlock {
  int l, n, chunk, hw, p0, np;   // level, image, chunk, pixels per image, first pixel, pixels in this block
  long long row0;                // row of the first pixel in the [R, 256] matrix
  int stat;                      // index of (l, n) in the stats arrays: (sum_{j<l} N) + n
};

__device__ __forceinline__ GnBlock gn_decode(const Levels& lv, const GnLevels& g, int b) {
  GnBlock r;
  int l = 0;
#pragma unroll
  for (int j = 1; j < SCAN_MAX_LEVELS; ++j)
    if (j < lv.n_levels && b >= g.blk_off[j]) l = j;
  const int rel = b - g.blk_off[l];
  r.l = l;
  r.n = rel / g.chunks[l];
  r.chunk = rel % g.chunks[l];
  r.hw = lv.h[l] * lv.w[l];
  r.p0 = r.chunk * GN_ROWS;
  r.np = min(GN_ROWS, r.hw - r.p0);
  r.row0 = lv.row_off[l] + (long long)r.n * r.hw + r.p0;
  r.stat = l * lv.n_images + r.n;
  return r;
}

// ---------------------------------------------------------------------------- forward
// cbias (may be null): the bias of the preceding convolution, added here so that the convolution runs bias-free and its bias
// gradient comes out of gn_bwd_apply_kernel instead of a separate full-tensor reduction
__device__ __forceinline__ float4 gn_bias4(const float* cbias, int c4) {
  return cbias ? __ldg(reinterpret_cast<const float4*>(cbias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ float4 gn_ldx(const float4* p, const float4& cb) {
  const float4 v = __ldg(p);
  return make_float4(v.x + cb.x, v.y + cb.y, v.z + cb.z, v.w + cb.w);
}

// y = relu(x * a + c) with a = rstd * gamma, c = beta - mean * a (the form torch's kernel uses; c and x * a + c one fused
// multiply-add each).  One definition with explicit roundings: the backward kernels recompute the ReLU mask [x * a + c > 0] from x instead of reading y back (two of seven
// HBM streams), which is only exact if forward and backward evaluate the very same operations.
__device__ __forceinline__ void gn_affine(float mean, float rstd, const float4& ga, const float4& be, float4& a, float4& c) {
  a = make_float4(__fmul_rn(rstd, ga.x), __fmul_rn(rstd, ga.y), __fmul_rn(rstd, ga.z), __fmul_rn(rstd, ga.w));
  c = make_float4(__fmaf_rn(-mean, a.x, be.x), __fmaf_rn(-mean, a.y, be.y), __fmaf_rn(-mean, a.z, be.z), __fmaf_rn(-mean, a.w, be.w));
}
__device__ __forceinline__ float4 gn_pre(const float4& v, const float4& a, const float4& c) {
  return make_float4(__fmaf_rn(v.x, a.x, c.x), __fmaf_rn(v.y, a.y, c.y), __fmaf_rn(v.z, a.z, c.z), __fmaf_rn(v.w, a.w, c.w));
}

__global__ void __launch_bounds__(256) gn_stats_kernel(Levels lv, GnLevels g, const float* __restrict__ cbias,
                                                       float* __restrict__ partial /* [blocks][32][2] */) {
  __shared__ float red[4][GN_G][2];
  const GnBlock b = gn_decode(lv, g, blockIdx.x);
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const float4 cb = gn_bias4(cbias, c4);
  const float4* x = reinterpret_cast<const float4*>(g.x[b.l] + ((long long)b.n * b.hw + b.p0) * GN_C) + c4;
  float s = 0.f, ss = 0.f;
  for (int p = rsub; p < b.np; p += 4) {
    const float4 v = gn_ldx(x + (long long)p * (GN_C / 4), cb);
    s += (v.x + v.y) + (v.z + v.w);
    ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);     // the two halves of the group
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  if ((c4 & 1) == 0) {
    red[rsub][c4 >> 1][0] = s;
    red[rsub][c4 >> 1][1] = ss;
  }
  __syncthreads();
  if (threadIdx.x < 2 * GN_G) {
    const int gi = threadIdx.x >> 1, k = threadIdx.x & 1;
    partial[((long long)blockIdx.x * GN_G + gi) * 2 + k] = (red[0][gi][k] + red[1][gi][k]) + (red[2][gi][k] + red[3][gi][k]);
  }
}

// one warp per (level, image, group): mean and 1/sqrt(var + eps), biased variance like torch's native_group_norm
__global__ void __launch_bounds__(256) gn_finalize_kernel(Levels lv, GnLevels g, const float* __restrict__ partial, float eps,
                                                          float* __restrict__ stats) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= lv.n_levels * lv.n_images * GN_G) return;
  const int gi = i % GN_G, ln = i / GN_G, l = ln / lv.n_images, n = ln % lv.n_images;
  const int b0 = g.blk_off[l] + n * g.chunks[l];
  double s = 0.0, ss = 0.0;
  for (int c = lane; c < g.chunks[l]; c += 32) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(partial) + (long long)(b0 + c) * GN_G + gi);
    s += (double)v.x;
    ss += (double)v.y;
  }
  s = warp_sum_d(s);
  ss = warp_sum_d(ss);
  if (lane == 0) {
    const double m = (double)lv.h[l] * lv.w[l] * 8.0;
    const double mean = s / m;
    const double var = fmax(ss / m - mean * mean, 0.0);
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(Levels lv, GnLevels g, const float* __restrict__ cbias,
                                                       const float* __restrict__ stats, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, float* __restrict__ y_rows) {
  const GnBlock b = gn_decode(lv, g, blockIdx.x);
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const float4 cb = gn_bias4(cbias, c4);
  const float mean = __ldg(stats + 2 * (b.stat * GN_G + (c4 >> 1)));
  const float rstd = __ldg(stats + 2 * (b.stat * GN_G + (c4 >> 1)) + 1);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  float4 a, c;
  gn_affine(mean, rstd, ga, be, a, c);
  const float4* x = reinterpret_cast<const float4*>(g.x[b.l] + ((long long)b.n * b.hw + b.p0) * GN_C) + c4;
  float4* y = reinterpret_cast<float4*>(y_rows + b.row0 * GN_C) + c4;
  for (int p = rsub; p < b.np; p += 4) {
    const float4 r = gn_pre(gn_ldx(x + (long long)p * (GN_C / 4), cb), a, c);
    y[(long long)p * (GN_C / 4)] = make_float4(fmaxf(r.x, 0.f), fmaxf(r.y, 0.f), fmaxf(r.z, 0.f), fmaxf(r.w, 0.f));
  }
}

// ---------------------------------------------------------------------------- backward
// per block and channel: A_c = sum dyr, B_c = sum dyr * xhat, with dyr = dy * [y > 0], xhat = (x - mean) * rstd; the mask
// [y > 0] is recomputed from x (gn_affine / gn_pre: the forward's own operations), y is not read
__global__ void __launch_bounds__(256) gn_bwd_reduce_kernel(Levels lv, GnLevels g, const float* __restrict__ cbias,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            float* __restrict__ partial /* [blocks][256][2] */) {
  __shared__ float red[4][GN_C][2];
  const GnBlock b = gn_decode(lv, g, blockIdx.x);
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const float4 cb = gn_bias4(cbias, c4);
  const float mean = __ldg(stats + 2 * (b.stat * GN_G + (c4 >> 1)));
  const float rstd = __ldg(stats + 2 * (b.stat * GN_G + (c4 >> 1)) + 1);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  float4 fa, fc;
  gn_affine(mean, rstd, ga, be, fa, fc);
  const long long base = ((long long)b.n * b.hw + b.p0) * GN_C;
  const float4* x = reinterpret_cast<const float4*>(g.x[b.l] + base) + c4;
  const float4* dy = reinterpret_cast<const float4*>(g.dy[b.l] + base) + c4;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
  for (int p = rsub; p < b.np; p += 4) {
    const long long o = (long long)p * (GN_C / 4);
    const float4 xv = gn_ldx(x + o, cb), dv = __ldg(dy + o);
    const float4 yv = gn_pre(xv, fa, fc);
    const float d0 = yv.x > 0.f ? dv.x : 0.f, d1 = yv.y > 0.f ? dv.y : 0.f, d2 = yv.z > 0.f ? dv.z : 0.f, d3 = yv.w > 0.f ? dv.w : 0.f;
    a[0] += d0; a[1] += d1; a[2] += d2; a[3] += d3;
    bb[0] += d0 * ((xv.x - mean) * rstd);
    bb[1] += d1 * ((xv.y - mean) * rstd);
    bb[2] += d2 * ((xv.z - mean) * rstd);
    bb[3] += d3 * ((xv.w - mean) * rstd);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    red[rsub][4 * c4 + i][0] = a[i];
    red[rsub][4 * c4 + i][1] = bb[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < GN_C * 2; i += 256) {
    const int c = i >> 1, k = i & 1;
    partial[(long long)blockIdx.x * GN_C * 2 + i] = (red[0][c][k] + red[1][c][k]) + (red[2][c][k] + red[3][c][k]);
  }
}

// one warp per output.  Group sums per (level, image, group): s1 = sum_c gamma_c A_c, s2 = sum_c gamma_c B_c -> gsum
// [L*N*32][2]; channel sums over everything: dgamma_c = sum B_c, dbeta_c = sum A_c.  fp64, fixed order: deterministic.
__global__ void __launch_bounds__(256) gn_bwd_finalize_kernel(Levels lv, GnLevels g, const float* __restrict__ partial,
                                                              const float* __restrict__ gamma, float* __restrict__ gsum,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int n_groups = lv.n_levels * lv.n_images * GN_G;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const float2* p2 = reinterpret_cast<const float2*>(partial);
  if (i < n_groups) {
    const int gi = i % GN_G, ln = i / GN_G, l = ln / lv.n_images, n = ln % lv.n_images;
    const int b0 = g.blk_off[l] + n * g.chunks[l];
    double s1 = 0.0, s2 = 0.0;
    for (int c = lane; c < g.chunks[l]; c += 32) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float2 v = __ldg(p2 + (long long)(b0 + c) * GN_C + gi * 8 + k);
        const double gm = (double)__ldg(gamma + gi * 8 + k);
        s1 += gm * (double)v.x;
        s2 += gm * (double)v.y;
      }
    }
    s1 = warp_sum_d(s1);
    s2 = warp_sum_d(s2);
    if (lane == 0) {
      gsum[2 * i] = (float)s1;
      gsum[2 * i + 1] = (float)s2;
    }
  } else if (i < n_groups + GN_C) {
    const int c = i - n_groups;
    const int n_blocks = g.blk_off[lv.n_levels];
    double sa = 0.0, sb = 0.0;
    for (int b = lane; b < n_blocks; b += 32) {
      const float2 v = __ldg(p2 + (long long)b * GN_C + c);
      sa += (double)v.x;
      sb += (double)v.y;
    }
    sa = warp_sum_d(sa);
    sb = warp_sum_d(sb);
    if (lane == 0) {
      dbeta[c] = (float)sa;
      dgamma[c] = (float)sb;
    }
  }
}

// dx, plus (when the convolution bias is folded in) per-block column sums of dx = the convolution's bias gradient partials
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(Levels lv, GnLevels g, const float* __restrict__ cbias,
                                                           const float* __restrict__ stats, const float* __restrict__ gsum,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ dx_rows, float* __restrict__ colsum /* [blocks][256] or null */) {
  __shared__ float red[4][GN_C];
  const GnBlock b = gn_decode(lv, g, blockIdx.x);
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const float4 cb = gn_bias4(cbias, c4);
  const int si = b.stat * GN_G + (c4 >> 1);
  const float mean = __ldg(stats + 2 * si), rstd = __ldg(stats + 2 * si + 1);
  const float inv_m = 1.f / ((float)b.hw * 8.f);
  const float k1 = __ldg(gsum + 2 * si) * inv_m, k2 = __ldg(gsum + 2 * si + 1) * inv_m;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  float4 fa, fc;
  gn_affine(mean, rstd, ga, be, fa, fc);
  const long long base = ((long long)b.n * b.hw + b.p0) * GN_C;
  const float4* x = reinterpret_cast<const float4*>(g.x[b.l] + base) + c4;
  const float4* dy = reinterpret_cast<const float4*>(g.dy[b.l] + base) + c4;
  float4* dx = reinterpret_cast<float4*>(dx_rows + b.row0 * GN_C) + c4;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = rsub; p < b.np; p += 4) {
    const long long o = (long long)p * (GN_C / 4);
    const float4 xv = gn_ldx(x + o, cb), dv = __ldg(dy + o);
    const float4 yv = gn_pre(xv, fa, fc);
    float4 r;
    // dx = rstd * (dyr * gamma - (s1 + xhat * s2) / m)
    r.x = rstd * ((yv.x > 0.f ? dv.x : 0.f) * ga.x - (k1 + (xv.x - mean) * rstd * k2));
    r.y = rstd * ((yv.y > 0.f ? dv.y : 0.f) * ga.y - (k1 + (xv.y - mean) * rstd * k2));
    r.z = rstd * ((yv.z > 0.f ? dv.z : 0.f) * ga.z - (k1 + (xv.z - mean) * rstd * k2));
    r.w = rstd * ((yv.w > 0.f ? dv.w : 0.f) * ga.w - (k1 + (xv.w - mean) * rstd * k2));
    dx[o] = r;
    cs.x += r.x; cs.y += r.y; cs.z += r.z; cs.w += r.w;
  }
  if (colsum) {
    red[rsub][4 * c4 + 0] = cs.x; red[rsub][4 * c4 + 1] = cs.y; red[rsub][4 * c4 + 2] = cs.z; red[rsub][4 * c4 + 3] = cs.w;
    __syncthreads();
    const int c = threadIdx.x;
    colsum[(long long)blockIdx.x * GN_C + c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
  }
}

// out[c] = sum over blocks of colsum[block][c]: one warp per channel, fp64, fixed order
__global__ void __launch_bounds__(256) gn_colsum_kernel(const float* __restrict__ colsum, int n_blocks, float* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= GN_C) return;
  double s = 0.0;
  for (int b = lane; b < n_blocks; b += 32) s += (double)__ldg(colsum + (long long)b * GN_C + c);
  s = warp_sum_d(s);
  if (lane == 0) out[c] = (float)s;
}

// ---------------------------------------------------------------------------- head_out epilogue: y = relu(u + v + bias)
// u = conv(features, W[:, :256]), v = conv(act maps, W[:, 256:]) (condgraph.py:379-384 without the concat); all levels per
// launch, output in the rows layout.  Backward: du = dv = dy * [y > 0], d_bias = column sums.
__global__ void __launch_bounds__(256) add_relu_kernel(Levels lv, GnLevels g /* x = u, dy = v */, const float* __restrict__ bias,
                                                       float* __restrict__ y_rows) {
  const GnBlock b = gn_decode(lv, g, blockIdx.x);
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const float4 cb = gn_bias4(bias, c4);
  const long long base = ((long long)b.n * b.hw + b.p0) * GN_C;
  const float4* u = reinterpret_cast<const float4*>(g.x[b.l] + base) + c4;
  const float4* v = g.dy[b.l] ? reinterpret_cast<const float4*>(g.dy[b.l] + base) + c4 : nullptr;
  float4* y = reinterpret_cast<float4*>(y_rows + b.row0 * GN_C) + c4;
  for (int p = rsub; p < b.np; p += 4) {
    const long long o = (long long)p * (GN_C / 4);
    float4 a = gn_ldx(u + o, cb);
    if (v) {
      const float4 w = __ldg(v + o);
      a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w;
    }
    y[o] = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
  }
}

__global__ void __launch_bounds__(256) add_relu_bwd_kernel(Levels lv, GnLevels g /* dy = upstream */, const float* __restrict__ y_rows,
                                                           float* __restrict__ d_rows, float* __restrict__ colsum) {
  __shared__ float red[4][GN_C];
  const GnBlock b = gn_decode(lv, g, blockIdx.x);
  const int c4 = threadIdx.x & 63, rsub = threadIdx.x >> 6;
  const long long base = ((long long)b.n * b.hw + b.p0) * GN_C;
  const float4* dy = reinterpret_cast<const float4*>(g.dy[b.l] + base) + c4;
  const float4* y = reinterpret_cast<const float4*>(y_rows + b.row0 * GN_C) + c4;
  float4* d = reinterpret_cast<float4*>(d_rows + b.row0 * GN_C) + c4;
  float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = rsub; p < b.np; p += 4) {
    const long long o = (long long)p * (GN_C / 4);
    const float4 dv = __ldg(dy + o), yv = __ldg(y + o);
    const float4 r = make_float4(yv.x > 0.f ? dv.x : 0.f, yv.y > 0.f ? dv.y : 0.f, yv.z > 0.f ? dv.z : 0.f, yv.w > 0.f ? dv.w : 0.f);
    d[o] = r;
    cs.x += r.x; cs.y += r.y; cs.z += r.z; cs.w += r.w;
  }
  if (colsum) {
    red[rsub][4 * c4 + 0] = cs.x; red[rsub][4 * c4 + 1] = cs.y; red[rsub][4 * c4 + 2] = cs.z; red[rsub][4 * c4 + 3] = cs.w;
    __syncthreads();
    const int c = threadIdx.x;
    colsum[(long long)blockIdx.x * GN_C + c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
  }
}

static int gn_build(const scan_levels_t* in, const void* const* x, const void* const* dy, Levels* lv, GnLevels* g,
                    bool need_x = true) {
  int rc = make_levels(in, lv);
  if (rc) return rc;
  int off = 0;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    g->blk_off[l] = off;
    g->x[l] = g->dy[l] = nullptr;
    g->chunks[l] = 1;
    if (l < lv->n_levels) {
      if ((need_x && (!x || !x[l])) || (dy && !dy[l])) return SCAN_EINVAL;
      g->x[l] = x ? (const float*)x[l] : nullptr;
      g->dy[l] = dy ? (const float*)dy[l] : nullptr;
      g->chunks[l] = (lv->h[l] * lv->w[l] + GN_ROWS - 1) / GN_ROWS;
      off += lv->n_images * g->chunks[l];
    }
  }
  g->blk_off[SCAN_MAX_LEVELS] = off;
  for (int l = lv->n_levels; l <= SCAN_MAX_LEVELS; ++l) g->blk_off[l] = off;
  return SCAN_OK;
}

static long long gn_blocks(const scan_levels_t* in) {
  long long off = 0;
  for (int l = 0; l < in->n_levels && l < SCAN_MAX_LEVELS; ++l)
    off += (long long)in->n_images * ((in->h[l] * in->w[l] + GN_ROWS - 1) / GN_ROWS);
  return off;
}

}  // namespace scan

extern "C" int64_t scan_gn_workspace_bytes(const scan_levels_t* lv) {
  if (!lv || lv->n_levels < 1 || lv->n_levels > SCAN_MAX_LEVELS) return 0;
  // backward partials dominate: [blocks][256][2] floats, + group sums, + dx column sums [blocks][256]
  return scan::gn_blocks(lv) * scan::GN_C * 3 * 4 + (long long)lv->n_levels * lv->n_images * scan::GN_G * 2 * 4 + 512;
}

extern "C" int scan_gn_relu_fwd(const scan_levels_t* lv_in, const void* const* x_levels_host, const float* conv_bias, const float* gamma,
                                const float* beta, float eps, float* y_rows, float* stats, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  using namespace scan;
  Levels lv;
  GnLevels g;
  int rc = gn_build(lv_in, x_levels_host, nullptr, &lv, &g);
  if (rc) return rc;
  if (!gamma || !beta || !y_rows || !stats || !workspace) return SCAN_EINVAL;
  if (workspace_bytes < scan_gn_workspace_bytes(lv_in)) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = g.blk_off[lv.n_levels];
  float* partial = (float*)workspace;
  gn_stats_kernel<<<blocks, 256, 0, st>>>(lv, g, conv_bias, partial);
  SCAN_LAUNCH_CHECK("gn_stats_kernel");
  const int n_stats = lv.n_levels * lv.n_images * GN_G;
  gn_finalize_kernel<<<(n_stats + 7) / 8, 256, 0, st>>>(lv, g, partial, eps, stats);
  SCAN_LAUNCH_CHECK("gn_finalize_kernel");
  gn_apply_kernel<<<blocks, 256, 0, st>>>(lv, g, conv_bias, stats, gamma, beta, y_rows);
  SCAN_LAUNCH_CHECK("gn_apply_kernel");
  return SCAN_OK;
}

// the apply pass alone, with statistics that came out of the convolution's epilogue (scan_conv3x3_rows_gn)
extern "C" int scan_gn_relu_apply(const scan_levels_t* lv_in, const void* const* x_levels_host, const float* conv_bias, const float* gamma,
                                  const float* beta, const float* stats, float* y_rows, void* stream) {
  using namespace scan;
  Levels lv;
  GnLevels g;
  int rc = gn_build(lv_in, x_levels_host, nullptr, &lv, &g);
  if (rc) return rc;
  if (!gamma || !beta || !y_rows || !stats) return SCAN_EINVAL;
  gn_apply_kernel<<<g.blk_off[lv.n_levels], 256, 0, (cudaStream_t)stream>>>(lv, g, conv_bias, stats, gamma, beta, y_rows);
  SCAN_LAUNCH_CHECK("gn_apply_kernel");
  return SCAN_OK;
}

extern "C" int scan_gn_relu_bwd(const scan_levels_t* lv_in, const void* const* x_levels_host, const void* const* dy_levels_host,
                                const float* conv_bias, const float* gamma, const float* beta, const float* stats, float* dx_rows,
                                float* dgamma, float* dbeta, float* d_conv_bias, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  using namespace scan;
  Levels lv;
  GnLevels g;
  int rc = gn_build(lv_in, x_levels_host, dy_levels_host, &lv, &g);
  if (rc) return rc;
  if (!dy_levels_host || !gamma || !beta || !stats || !dx_rows || !dgamma || !dbeta || !workspace) return SCAN_EINVAL;
  if ((conv_bias != nullptr) != (d_conv_bias != nullptr)) return SCAN_EINVAL;
  if (workspace_bytes < scan_gn_workspace_bytes(lv_in)) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = g.blk_off[lv.n_levels];
  float* partial = (float*)workspace;
  float* gsum = partial + (long long)blocks * GN_C * 2;
  float* colsum = gsum + (long long)lv.n_levels * lv.n_images * GN_G * 2 + 64;
  gn_bwd_reduce_kernel<<<blocks, 256, 0, st>>>(lv, g, conv_bias, stats, gamma, beta, partial);
  SCAN_LAUNCH_CHECK("gn_bwd_reduce_kernel");
  const int n_fin = lv.n_levels * lv.n_images * GN_G + GN_C;
  gn_bwd_finalize_kernel<<<(n_fin + 7) / 8, 256, 0, st>>>(lv, g, partial, gamma, gsum, dgamma, dbeta);
  SCAN_LAUNCH_CHECK("gn_bwd_finalize_kernel");
  gn_bwd_apply_kernel<<<blocks, 256, 0, st>>>(lv, g, conv_bias, stats, gsum, gamma, beta, dx_rows, d_conv_bias ? colsum : nullptr);
  SCAN_LAUNCH_CHECK("gn_bwd_apply_kernel");
  if (d_conv_bias) {
    gn_colsum_kernel<<<GN_C / 8, 256, 0, st>>>(colsum, blocks, d_conv_bias);
    SCAN_LAUNCH_CHECK("gn_colsum_kernel");
  }
  return SCAN_OK;
}

extern "C" int scan_add_relu_fwd(const scan_levels_t* lv_in, const void* const* u_levels_host, const void* const* v_levels_host,
                                 const float* bias, float* y_rows, void* stream) {
  using namespace scan;
  Levels lv;
  GnLevels g;
  int rc = gn_build(lv_in, u_levels_host, v_levels_host, &lv, &g);
  if (rc) return rc;
  if (!y_rows) return SCAN_EINVAL;
  add_relu_kernel<<<g.blk_off[lv.n_levels], 256, 0, (cudaStream_t)stream>>>(lv, g, bias, y_rows);
  SCAN_LAUNCH_CHECK("add_relu_kernel");
  return SCAN_OK;
}

extern "C" int scan_add_relu_bwd(const scan_levels_t* lv_in, const void* const* dy_levels_host, const float* y_rows, float* d_rows,
                                 float* d_bias, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace scan;
  Levels lv;
  GnLevels g;
  int rc = gn_build(lv_in, nullptr, dy_levels_host, &lv, &g, false);
  if (rc) return rc;
  if (!dy_levels_host || !y_rows || !d_rows || (d_bias && !workspace)) return SCAN_EINVAL;
  if (d_bias && workspace_bytes < scan_gn_workspace_bytes(lv_in)) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = g.blk_off[lv.n_levels];
  float* colsum = (float*)workspace;
  add_relu_bwd_kernel<<<blocks, 256, 0, st>>>(lv, g, y_rows, d_rows, d_bias ? colsum : nullptr);
  SCAN_LAUNCH_CHECK("add_relu_bwd_kernel");
  if (d_bias) {
    gn_colsum_kernel<<<GN_C / 8, 256, 0, st>>>(colsum, blocks, d_bias);
    SCAN_LAUNCH_CHECK("gn_colsum_kernel");
  }
  return SCAN_OK;
}
