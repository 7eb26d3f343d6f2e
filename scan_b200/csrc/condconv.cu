// K4b: conditional 1x1 convolution of the semantic conditioned kernels against every FPN pixel, fused with
// the class softmax / sigmoid, the activation-map store and the focal loss.
// Reference: modeling/rpn/fcos/condgraph.py:619-629 (dynamic_conv), :338-370 (get_act_loss),
// layers/sigmoid_focal_loss_wbg.py:7-64 (FocalLoss) and :148-177 (BCEFocalLoss).
//
// Forward = one skinny GEMM  logits[R, K] = rows[R, 256] . W[K, 256]^T  (K <= 16), HBM-bound (AI ~ 4 flop/B).
//   * persistent CTAs, one per SM; tile = 128 pixels x 256 channels; the kernel itself is condconv_ts.inl (TMA ring ->
//     converter warps -> hi/lo operand slots in TENSOR MEMORY -> tcgen05.mma kind::tf32, 3xTF32 error compensation);
//   * B = the conditioned kernels, zero-padded to N = 16 by TMA out-of-bounds fill, resident in shared memory
//     for the whole kernel, split once into hi + lo;
//   * epilogue warps: one TMEM lane = one pixel, so softmax / sigmoid, the NCHW activation-map store and the
//     focal-loss term are computed per thread in registers (tcgen05.ld 32x32b.x16).
// Backward = two streaming passes (d_logit, then d_rows + per-CTA d_weight partials; fp32 FFMA, HBM-bound).
// condconv_fwd_simt_kernel is the fp32 verification kernel the tests compare the tensor-core path with.
#include <stdlib.h>

#include "tc_common.cuh"

namespace scan {

constexpr int CC_C = 256;
constexpr int CC_BM = 128;
constexpr int CC_BK = 32;
constexpr int CC_KB = CC_C / CC_BK;  // 8 k-blocks per tile
constexpr int CC_N = 16;
constexpr int CC_STAGE_BYTES = CC_BM * CC_BK * 4;  // 16384
constexpr int CC_WBLK_BYTES = CC_N * CC_BK * 4;    // 2048
constexpr int CC_W_BYTES = CC_KB * CC_WBLK_BYTES;  // 16384
constexpr int CC_MAX_PARTIALS = 1024;

struct ActPtrs {
  float* p[SCAN_MAX_LEVELS];
};
struct ConstActPtrs {
  const float* p[SCAN_MAX_LEVELS];
};

constexpr uint32_t CC_IDESC = umma_idesc_tf32(CC_BM, CC_N);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------- shared epilogue
// z[0..K) are the logits of row g.  Writes the activation maps (NCHW, reference layout) and returns the
// un-normalised focal-loss term of the row.
__device__ __forceinline__ double act_epilogue(const Levels& lv, const ActPtrs& act, long long g, float (&z)[16], int K,
                                               int act_mode, const float* __restrict__ bias, const int64_t* __restrict__ labels,
                                               int* __restrict__ flags) {
  const int l = level_of_row(lv, g);
  const int hw = lv.h[l] * lv.w[l];
  const long long r = g - lv.row_off[l];
  const int n = (int)(r / hw);
  const int p = (int)(r - (long long)n * hw);
  float* out = act.p[l] + (long long)n * K * hw + p;
  if (bias) {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) z[k] += __ldg(bias + k);
  }
  double loss = 0.0;
  if (act_mode == 0) {
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) mx = fmaxf(mx, z[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        z[k] = expf(z[k] - mx);
        s += z[k];
      }
    const float inv = 1.f / s;
    const int t = labels ? (int)__ldg(labels + g) : -1;
    float pt = 1.f;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        const float pk = z[k] * inv;
        out[(long long)k * hw] = pk;
        if (k == t) pt = pk;
      }
    if (labels) {
      if (pt < 1e-15f) {  // sigmoid_focal_loss_wbg.py:50-52
        pt = 1e-15f;
        if (flags) atomicOr(flags, 1);
      }
      const float om = 1.f - pt;
      loss = (double)(-(om * om) * logf(pt));
    }
  } else {
    const int t = labels ? (int)__ldg(labels + g) : -1;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        const float pk = 1.f / (1.f + expf(-z[k]));
        out[(long long)k * hw] = pk;
        if (labels) {  // BCEFocalLoss, gamma 2, alpha 0.25, one-hot target over the 2 columns
          const float pc = fminf(fmaxf(pk, 0.00001f), 0.99999f);
          const float tk = (k == t) ? 1.f : 0.f;
          loss += (double)(-0.25f * (1.f - pc) * (1.f - pc) * tk * logf(pc) - 0.75f * pc * pc * (1.f - tk) * logf(1.f - pc));
        }
      }
  }
  return loss;
}

#include "condconv_ts.inl"

// ---------------------------------------------------------------------------- forward, fp32 FFMA (verification)
__global__ void __launch_bounds__(128) condconv_fwd_simt_kernel(Levels lv, const float* __restrict__ rows, const float* __restrict__ weight,
                                                                ActPtrs act, const float* __restrict__ bias,
                                                                const int64_t* __restrict__ labels, double* __restrict__ loss_partials,
                                                                int* __restrict__ flags, int K, int act_mode, int num_tiles) {
  __shared__ float ws[CC_N][CC_C];
  __shared__ float xs[128][33];
  __shared__ double red[4];
  for (int i = threadIdx.x; i < CC_N * CC_C; i += 128) ws[i / CC_C][i % CC_C] = (i / CC_C < K) ? weight[i] : 0.f;
  __syncthreads();
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  double loss = 0.0;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    float z[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) z[k] = 0.f;
    const long long g0 = (long long)tile * 128;
    for (int c0 = 0; c0 < CC_C; c0 += 32) {
      __syncthreads();
      for (int i = threadIdx.x; i < 128 * 32; i += 128) {
        const int r = i >> 5, c = i & 31;
        xs[r][c] = (g0 + r < R) ? rows[(g0 + r) * CC_C + c0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int c = 0; c < 32; ++c) {
        const float x = xs[threadIdx.x][c];
#pragma unroll
        for (int k = 0; k < 16; ++k) z[k] = fmaf(x, ws[k][c0 + c], z[k]);
      }
    }
    const long long g = g0 + threadIdx.x;
    if (g < R) loss += act_epilogue(lv, act, g, z, K, act_mode, bias, labels, flags);
  }
  if (loss_partials) {
    loss = warp_sum_d(loss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) loss_partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
  }
}

// ---------------------------------------------------------------------------- backward
constexpr int CB_ROWS = 64;     // rows per tile
constexpr int CB_THREADS = 256;  // one channel per thread

// d(logit) of one row from the saved activations, the upstream map gradient and the focal loss
__device__ __forceinline__ void row_dlogits(const Levels& lv, const ConstActPtrs& act, const ConstActPtrs& dact, long long g, int K,
                                            int act_mode, const int64_t* __restrict__ labels, float loss_scale, float* dz) {
  const int l = level_of_row(lv, g);
  const int hw = lv.h[l] * lv.w[l];
  const long long r = g - lv.row_off[l];
  const int n = (int)(r / hw);
  const int p = (int)(r - (long long)n * hw);
  const long long base = (long long)n * K * hw + p;
  const float* pa = act.p[l] + base;
  const float* ga = dact.p[l] ? dact.p[l] + base : nullptr;
  const int t = (labels && loss_scale != 0.f) ? (int)__ldg(labels + g) : -1;
  float pk[16], gk[16];
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    pk[k] = 0.f;
    gk[k] = 0.f;
    if (k < K) {
      pk[k] = __ldg(pa + (long long)k * hw);
      if (ga) gk[k] = __ldg(ga + (long long)k * hw);
      dot += pk[k] * gk[k];
    }
  }
  if (act_mode == 0) {
    float coef = 0.f;  // dL/dpt * pt
    if (t >= 0) {
      float pt = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k == t) pt = pk[k];
      if (pt >= 1e-15f) {  // clamp(min=1e-15) has zero gradient below the bound
        const float om = 1.f - pt;
        coef = loss_scale * (2.f * om * logf(pt) * pt - om * om);
      }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) dz[k] = pk[k] * (gk[k] - dot) + coef * ((k == t ? 1.f : 0.f) - pk[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        const float p_ = pk[k];
        float d = gk[k] * p_ * (1.f - p_);
        if (t >= 0 && p_ > 0.00001f && p_ < 0.99999f) {
          const float tk = (k == t) ? 1.f : 0.f;
          const float om = 1.f - p_;
          const float dldp = -0.25f * tk * (-2.f * om * logf(p_) + om * om / p_) - 0.75f * (1.f - tk) * (2.f * p_ * logf(om) - p_ * p_ / om);
          d += loss_scale * dldp * p_ * om;
        }
        dz[k] = d;
      }
  }
}

// ---- the product backward: two streaming passes -----------------------------------------------------------------
// pass A  d(logit) of every row from the saved maps, the upstream map gradients and the focal term -> dz [R, KP] (KP =
//         num_classes) in the workspace, plus per-block column sums (the bias gradient).  Thread = pixel, so the NCHW map
//         reads are coalesced along the pixel axis.
// pass B  one pass over the rows: d_rows = dz . W and per-CTA d_weight partials.  Thread = channel, the dz tile of 64 rows
//         is broadcast from shared memory with 128-bit loads, 8 independent row loads in flight per thread, 4 CTAs / SM.
//         Per (row, channel): 2 KP FMAs -- with KP = 9 that is below the fp32 issue budget of an HBM-bound stream.
__global__ void __launch_bounds__(256) condconv_dlogit_kernel(Levels lv, ConstActPtrs act, ConstActPtrs dact, const int64_t* __restrict__ labels,
                                                              float loss_scale, const float* __restrict__ d_loss, int K, int act_mode,
                                                              float* __restrict__ dzw, float* __restrict__ partial_b) {
  __shared__ float red[8][16];
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  if (d_loss) loss_scale *= __ldg(d_loss);  // d(total)/d(act_loss), a device scalar: no host sync
  float bsum[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) bsum[k] = 0.f;
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < R; g += (long long)gridDim.x * blockDim.x) {
    float dz[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) dz[k] = 0.f;
    row_dlogits(lv, act, dact, g, K, act_mode, labels, loss_scale, dz);
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < K) {
        dzw[g * K + k] = dz[k];
        bsum[k] += dz[k];
      }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float s = warp_sum(bsum[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial_b[blockIdx.x * 16 + threadIdx.x] = s;
  }
}

// ACC: d_rows already holds the gradient of a later consumer of the same rows (head_out's data gradient): add to it in place
template <int KP, bool ACC>
__global__ void __launch_bounds__(CB_THREADS, 4) condconv_bwd_rows_kernel(const float* __restrict__ rows, const float* __restrict__ weight,
                                                                         const float* __restrict__ dzw, long long R, int num_tiles,
                                                                         float* d_rows, float* __restrict__ partial_w) {
  constexpr int KS = (KP + 3) / 4 * 4;   // shared-memory row stride: whole float4s
  __shared__ __align__(16) float dzs[CB_ROWS * KS];
  const int c = threadIdx.x;
  float w[KP], acc[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    w[k] = __ldg(weight + k * CC_C + c);
    acc[k] = 0.f;
  }
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long g0 = (long long)tile * CB_ROWS;
    const int nrows = (int)min((long long)CB_ROWS, R - g0);
    __syncthreads();
    for (int i = threadIdx.x; i < CB_ROWS * KP; i += CB_THREADS) {
      const int r = i / KP, k = i - r * KP;
      dzs[r * KS + k] = r < nrows ? __ldg(dzw + g0 * KP + i) : 0.f;
    }
    __syncthreads();
    const float* xr = rows + g0 * CC_C + c;
    float* dr = d_rows + g0 * CC_C + c;
    for (int r0 = 0; r0 < CB_ROWS; r0 += 8) {
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = (r0 + j < nrows) ? __ldg(xr + (long long)(r0 + j) * CC_C) : 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* d4 = reinterpret_cast<const float4*>(dzs + (r0 + j) * KS);
        float dz[KS];
#pragma unroll
        for (int q = 0; q < KS / 4; ++q) {
          const float4 d = d4[q];
          dz[4 * q] = d.x; dz[4 * q + 1] = d.y; dz[4 * q + 2] = d.z; dz[4 * q + 3] = d.w;
        }
        float dx = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          dx = fmaf(dz[k], w[k], dx);
          acc[k] = fmaf(dz[k], x[j], acc[k]);
        }
        if (r0 + j < nrows) {
          if (ACC) dx += dr[(long long)(r0 + j) * CC_C];
          dr[(long long)(r0 + j) * CC_C] = dx;
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) partial_w[((long long)blockIdx.x * 16 + k) * CC_C + c] = acc[k];
}

// grid (K, 8): block (k, j) sums 32 channels of class k over the per-CTA partials with 8 row-lanes per channel and a fixed
// shared-memory tree (the one-block-per-class version took 27 us for 9.7 MB: 9 CTAs cannot pull bandwidth)
__global__ void __launch_bounds__(256) condconv_bwd_reduce_kernel(const float* __restrict__ partial_w, const float* __restrict__ partial_b,
                                                                  int n_parts, int n_parts_b, int K, float* __restrict__ d_weight,
                                                                  float* __restrict__ d_bias) {
  __shared__ float red[8][32];
  const int k = blockIdx.x;
  const int c = blockIdx.y * 32 + (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
  float s = 0.f;
  for (int i = lane_r; i < n_parts; i += 8) s += partial_w[((long long)i * 16 + k) * CC_C + c];
  red[lane_r][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane_r == 0) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) t += red[r][threadIdx.x];
    d_weight[k * CC_C + c] = t;
  }
  if (d_bias && blockIdx.y == 0 && threadIdx.x == 0) {
    float b = 0.f;
    for (int i = 0; i < n_parts_b; ++i) b += partial_b[i * 16 + k];
    d_bias[k] = b;
  }
}

// ---------------------------------------------------------------------------- host side
static EncodeTiledFn g_encode = nullptr;

int get_tensormap_encoder(EncodeTiledFn* out) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SCAN_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      set_cuda_error(cudaErrorUnknown, "cuTensorMapEncodeTiled entry point not found");
      return SCAN_ECUDA;
    }
    g_encode = (EncodeTiledFn)fn;
  }
  *out = g_encode;
  return SCAN_OK;
}

int make_rowmajor_map(CUtensorMap* m, const float* base, uint64_t n_rows, uint64_t n_cols, uint32_t box_rows) {
  EncodeTiledFn enc;
  int rc = get_tensormap_encoder(&enc);
  if (rc) return rc;
  cuuint64_t dims[2] = {n_cols, n_rows};
  cuuint64_t strides[1] = {n_cols * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed");
    return SCAN_ECUDA;
  }
  return SCAN_OK;
}

}  // namespace scan

extern "C" int32_t scan_condconv_num_partials(void) { return scan::CC_MAX_PARTIALS; }

extern "C" int scan_condconv_fwd(const scan_levels_t* lvh, const float* rows, const float* weight, const float* bias,
                                 int32_t num_classes, int32_t act_mode, void* const* act_nchw_host, const int64_t* labels,
                                 double* loss_partials, int32_t* flags, int32_t impl, void* stream) {
  using namespace scan;
  Levels lv;
  int rc = make_levels(lvh, &lv);
  if (rc) return rc;
  if (!rows || !weight || !act_nchw_host || num_classes < 1 || num_classes > SCAN_MAX_CLASSES) return SCAN_EINVAL;
  if (act_mode != 0 && act_mode != 1) return SCAN_EINVAL;
  if (labels && !loss_partials) return SCAN_EINVAL;
  // impl 0: tcgen05, A operands in tensor memory (product); 1: fp32 FFMA (verification kernel for the tests)
  if (((uintptr_t)rows & 15) || ((uintptr_t)weight & 15)) return SCAN_EINVAL;
  ActPtrs act;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    act.p[l] = l < lv.n_levels ? (float*)act_nchw_host[l] : nullptr;
    if (l < lv.n_levels && !act.p[l]) return SCAN_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  if (impl != 0) {  // the product kernel (impl 0) clears its own partial / flag slots
    if (loss_partials) SCAN_CUDA_CHECK(cudaMemsetAsync(loss_partials, 0, sizeof(double) * CC_MAX_PARTIALS, st));
    if (flags) SCAN_CUDA_CHECK(cudaMemsetAsync(flags, 0, sizeof(int32_t) * CC_MAX_PARTIALS, st));
  }
  const int num_tiles = (int)ceil_div(R, CC_BM);
  const int grid = std::min(std::min(num_tiles, sm_count()), CC_MAX_PARTIALS);
  if (impl == 1) {
    condconv_fwd_simt_kernel<<<std::min(num_tiles, CC_MAX_PARTIALS), 128, 0, st>>>(lv, rows, weight, act, bias, labels,
                                                                                  labels ? loss_partials : nullptr, flags,
                                                                                  num_classes, act_mode, num_tiles);
    SCAN_LAUNCH_CHECK("condconv_fwd_simt_kernel");
    return SCAN_OK;
  }
  if (impl != 0) return SCAN_EINVAL;
  CUtensorMap mx, mw;
  rc = make_rowmajor_map(&mx, rows, (uint64_t)R, CC_C, CC_BM);
  if (rc) return rc;
  rc = make_rowmajor_map(&mw, weight, (uint64_t)num_classes, CC_C, CC_N);
  if (rc) return rc;
  // product kernel: A operands in tensor memory (condconv_ts.inl)
  static unsigned long long ts_attr = 0;
  if (first_use_on_device(&ts_attr))
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(condconv_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
  condconv_fwd_ts_kernel<<<grid, TS_THREADS, TS_SMEM, st>>>(mx, mw, lv, act, bias, labels, labels ? loss_partials : nullptr, flags,
                                                           num_classes, act_mode, num_tiles, 0);
  SCAN_LAUNCH_CHECK("condconv_fwd_ts_kernel");
  return SCAN_OK;
}

extern "C" int64_t scan_condconv_bwd_workspace_bytes(const scan_levels_t* lvh, int32_t num_classes) {
  scan::Levels lv;
  if (scan::make_levels(lvh, &lv) || num_classes < 1 || num_classes > SCAN_MAX_CLASSES) return 0;
  const int64_t parts = 4ll * scan::sm_count();
  // d_weight partials [parts][16][256] | bias partials [parts][16] | dz [R][num_classes]
  return parts * 16 * scan::CC_C * 4 + parts * 16 * 4 + lv.row_off[SCAN_MAX_LEVELS] * num_classes * 4 + 512;
}

extern "C" int scan_condconv_bwd2(const scan_levels_t* lvh, const float* rows, const float* weight, int32_t num_classes,
                                  int32_t act_mode, const void* const* act_nchw_host, const void* const* d_act_nchw_host,
                                  const int64_t* labels, float loss_scale, const float* d_loss, float* d_rows, int32_t accumulate_rows,
                                  float* d_weight, float* d_bias, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace scan;
  Levels lv;
  int rc = make_levels(lvh, &lv);
  if (rc) return rc;
  if (!rows || !weight || !act_nchw_host || !d_rows || !d_weight || !workspace) return SCAN_EINVAL;
  if (num_classes < 1 || num_classes > SCAN_MAX_CLASSES || (act_mode != 0 && act_mode != 1)) return SCAN_EINVAL;
  if (workspace_bytes < scan_condconv_bwd_workspace_bytes(lvh, num_classes)) return SCAN_ECAPACITY;
  ConstActPtrs act, dact;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    act.p[l] = l < lv.n_levels ? (const float*)act_nchw_host[l] : nullptr;
    dact.p[l] = (l < lv.n_levels && d_act_nchw_host) ? (const float*)d_act_nchw_host[l] : nullptr;
    if (l < lv.n_levels && !act.p[l]) return SCAN_EINVAL;
  }
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const int num_tiles = (int)ceil_div(R, CB_ROWS);
  const int parts = std::min(num_tiles, 4 * sm_count());
  const int parts_b = (int)std::min<long long>(ceil_div(R, 256), 4 * sm_count());
  float* pw = (float*)workspace;
  float* pb = pw + (long long)4 * sm_count() * 16 * CC_C;
  float* dzw = pb + (long long)4 * sm_count() * 16 + 64;
  cudaStream_t st = (cudaStream_t)stream;
  condconv_dlogit_kernel<<<parts_b, 256, 0, st>>>(lv, act, dact, labels, loss_scale, d_loss, num_classes, act_mode, dzw, pb);
  SCAN_LAUNCH_CHECK("condconv_dlogit_kernel");
  switch (num_classes) {
#define SCAN_BWD_CASE(KP)                                                                                                \
  case KP:                                                                                                               \
    if (accumulate_rows)                                                                                                 \
      condconv_bwd_rows_kernel<KP, true><<<parts, CB_THREADS, 0, st>>>(rows, weight, dzw, R, num_tiles, d_rows, pw);    \
    else                                                                                                                 \
      condconv_bwd_rows_kernel<KP, false><<<parts, CB_THREADS, 0, st>>>(rows, weight, dzw, R, num_tiles, d_rows, pw);   \
    break;
    SCAN_BWD_CASE(1) SCAN_BWD_CASE(2) SCAN_BWD_CASE(3) SCAN_BWD_CASE(4) SCAN_BWD_CASE(5) SCAN_BWD_CASE(6) SCAN_BWD_CASE(7) SCAN_BWD_CASE(8)
    SCAN_BWD_CASE(9) SCAN_BWD_CASE(10) SCAN_BWD_CASE(11) SCAN_BWD_CASE(12) SCAN_BWD_CASE(13) SCAN_BWD_CASE(14) SCAN_BWD_CASE(15)
    SCAN_BWD_CASE(16)
#undef SCAN_BWD_CASE
  }
  SCAN_LAUNCH_CHECK("condconv_bwd_rows_kernel");
  condconv_bwd_reduce_kernel<<<dim3(num_classes, CC_C / 32), 256, 0, st>>>(pw, pb, parts, parts_b, num_classes, d_weight, d_bias);
  SCAN_LAUNCH_CHECK("condconv_bwd_reduce_kernel");
  return SCAN_OK;
}

extern "C" int scan_condconv_bwd(const scan_levels_t* lvh, const float* rows, const float* weight, int32_t num_classes,
                                 int32_t act_mode, const void* const* act_nchw_host, const void* const* d_act_nchw_host,
                                 const int64_t* labels, float loss_scale, const float* d_loss, float* d_rows, float* d_weight,
                                 float* d_bias, void* workspace, int64_t workspace_bytes, void* stream) {
  return scan_condconv_bwd2(lvh, rows, weight, num_classes, act_mode, act_nchw_host, d_act_nchw_host, labels, loss_scale, d_loss, d_rows, 0,
                            d_weight, d_bias, workspace, workspace_bytes, stream);
}
