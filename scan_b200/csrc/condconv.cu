// K4b: conditional 1x1 convolution of the semantic conditioned kernels against every FPN pixel, fused with
// the class softmax / sigmoid, the activation-map store and the focal loss.
// Reference: modeling/rpn/fcos/condgraph.py:619-629 (dynamic_conv), :338-370 (get_act_loss),
// layers/sigmoid_focal_loss_wbg.py:7-64 (FocalLoss) and :148-177 (BCEFocalLoss).
//
// Forward = one skinny GEMM  logits[R, K] = rows[R, 256] . W[K, 256]^T  (K <= 16), HBM-bound (AI ~ 4 flop/B).
//   * persistent CTAs, one per SM; tile = 128 pixels x 256 channels;
//   * A (pixels x channels, K-major) streamed by TMA in 8 k-blocks of [128 x 32] fp32 = 16 KB with the 128-byte
//     swizzle, 6-stage mbarrier ring (96 KB in flight per SM);
//   * 3xTF32 error compensation so the tensor-core path keeps fp32-level accuracy (a plain tf32 read truncates the
//     activations: measured 1e-3..5e-3 absolute error on the maps, coherent bias in the weight gradients): a
//     converter warpgroup splits every landed stage IN SHARED MEMORY into hi = rna_tf32(x) (in place) and
//     lo = x - hi (second buffer), element-wise and therefore swizzle-agnostic; the MMA warp issues
//     hi*Whi + hi*Wlo + lo*Whi (3 MMAs per k-step; the tensor pipe stays < 15 % busy);
//   * B = the conditioned kernels, zero-padded to N = 16 by TMA out-of-bounds fill, resident in shared memory
//     for the whole kernel, split once into hi + lo the same way;
//   * tcgen05.mma kind::tf32, M=128 N=16 K=8, fp32 accumulators in TMEM, 4-deep accumulator ring;
//   * epilogue warps: one TMEM lane = one pixel, so softmax / sigmoid, the NCHW activation-map store and the
//     focal-loss term are computed per thread in registers (tcgen05.ld 32x32b.x16).
// Backward = one pass over rows producing d_rows (dense write) and per-CTA partial d_weight (fp32 FFMA).
#include <stdlib.h>

#include "tc_common.cuh"

namespace scan {

constexpr int CC_C = 256;
constexpr int CC_BM = 128;
constexpr int CC_BK = 32;
constexpr int CC_KB = CC_C / CC_BK;  // 8 k-blocks per tile
constexpr int CC_N = 16;
constexpr int CC_STAGES = 6;
constexpr int CC_ACC = 4;
constexpr int CC_STAGE_BYTES = CC_BM * CC_BK * 4;  // 16384
constexpr int CC_WBLK_BYTES = CC_N * CC_BK * 4;    // 2048
constexpr int CC_W_BYTES = CC_KB * CC_WBLK_BYTES;  // 16384
constexpr int CC_SMEM = 1024 + 2 * CC_W_BYTES + 2 * CC_STAGES * CC_STAGE_BYTES + 1024;  // hi + lo per stage
constexpr int CC_CONV_GROUPS = 2;  // converter warpgroups; group c converts the k-blocks with index % CC_CONV_GROUPS == c
constexpr int CC_THREADS = 256 + 128 * CC_CONV_GROUPS;  // warps 0-3: TMA / MMA / TMEM alloc / idle, 4-7: epilogue, 8..: converters
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

// ---------------------------------------------------------------------------- forward, tcgen05
__global__ void __launch_bounds__(CC_THREADS, 1)
    condconv_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, Levels lv,
                           ActPtrs act, const float* __restrict__ bias, const int64_t* __restrict__ labels,
                           double* __restrict__ loss_partials, int* __restrict__ flags, int K, int act_mode, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* w_hi = smem;
  uint8_t* w_lo = smem + CC_W_BYTES;
  uint8_t* stages = smem + 2 * CC_W_BYTES;                    // hi parts (TMA destination, converted in place)
  uint8_t* stages_lo = stages + CC_STAGES * CC_STAGE_BYTES;   // lo parts
  uint64_t* bars = (uint64_t*)(stages_lo + CC_STAGES * CC_STAGE_BYTES);
  uint64_t* full_bar = bars;                         // [CC_STAGES] TMA -> converter
  uint64_t* empty_bar = bars + CC_STAGES;            // [CC_STAGES] MMA -> TMA
  uint64_t* ready_bar = bars + 2 * CC_STAGES;        // [CC_STAGES] converter -> MMA
  uint64_t* acc_full = bars + 3 * CC_STAGES;         // [CC_ACC]
  uint64_t* acc_empty = bars + 3 * CC_STAGES + CC_ACC;  // [CC_ACC]
  uint64_t* w_bar = bars + 3 * CC_STAGES + 2 * CC_ACC;
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);
  __shared__ double red[4];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];

  if (threadIdx.x == 0) {
    for (int i = 0; i < CC_STAGES; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 1);
      mbar_init(smem_u32(ready_bar + i), 128);
    }
    for (int i = 0; i < CC_ACC; ++i) {
      mbar_init(smem_u32(acc_full + i), 1);
      mbar_init(smem_u32(acc_empty + i), 128);
    }
    mbar_init(smem_u32(w_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {  // TMEM allocation: CC_ACC x 16 columns = 64
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(CC_ACC * CC_N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // conditioned kernels -> shared memory (raw fp32 via TMA, rows >= K zero-filled), then hi/lo split in place
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(w_bar), CC_W_BYTES);
    for (int kb = 0; kb < CC_KB; ++kb) tma_load_2d(smem_u32(w_hi + kb * CC_WBLK_BYTES), &tmap_w, smem_u32(w_bar), kb * CC_BK, 0);
  }
  mbar_wait(smem_u32(w_bar), 0);
  for (int i = threadIdx.x; i < CC_W_BYTES / 4; i += CC_THREADS) {
    const float wv = ((float*)w_hi)[i];
    uint32_t hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(wv));
    ((float*)w_hi)[i] = __uint_as_float(hi);
    ((float*)w_lo)[i] = wv - __uint_as_float(hi);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < CC_KB; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          mbar_expect_tx(smem_u32(full_bar + stage), CC_STAGE_BYTES);
          tma_load_2d(smem_u32(stages + stage * CC_STAGE_BYTES), &tmap_x, smem_u32(full_bar + stage), kb * CC_BK, tile * CC_BM);
          if (++stage == CC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(acc_empty + acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * CC_N;
        for (int kb = 0; kb < CC_KB; ++kb) {
          mbar_wait(smem_u32(ready_bar + stage), phase);
          tcgen05_fence_after();
          const uint32_t a_addr = smem_u32(stages + stage * CC_STAGE_BYTES);
          const uint32_t al_addr = smem_u32(stages_lo + stage * CC_STAGE_BYTES);
          const uint32_t bh_addr = smem_u32(w_hi + kb * CC_WBLK_BYTES);
          const uint32_t bl_addr = smem_u32(w_lo + kb * CC_WBLK_BYTES);
#pragma unroll
          for (int k = 0; k < CC_BK / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint64_t da = umma_desc_sw128(a_addr + k * 32);
            const uint64_t dbh = umma_desc_sw128(bh_addr + k * 32);
            umma_tf32(d, da, dbh, CC_IDESC, (kb | k) != 0);
            umma_tf32(d, da, umma_desc_sw128(bl_addr + k * 32), CC_IDESC, 1);
            umma_tf32(d, umma_desc_sw128(al_addr + k * 32), dbh, CC_IDESC, 1);
          }
          umma_commit(smem_u32(empty_bar + stage));  // frees the smem stage when these MMAs retire
          if (++stage == CC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(acc_full + acc));
        if (++acc == CC_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===== converter: fp32 stage -> tf32 hi (in place) + lo (second buffer); element-wise, swizzle-agnostic =====
    const int t = (threadIdx.x - 256) & 127;
    const int group = (threadIdx.x - 256) >> 7;
    long long it = 0;  // running k-block index of this CTA: stage = it % CC_STAGES, phase = (it / CC_STAGES) & 1
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < CC_KB; ++kb, ++it) {
        if ((int)(it % CC_CONV_GROUPS) != group) continue;
        const int stage = (int)(it % CC_STAGES);
        const uint32_t phase = (uint32_t)((it / CC_STAGES) & 1);
        mbar_wait(smem_u32(full_bar + stage), phase);
        float4* hi = reinterpret_cast<float4*>(stages + stage * CC_STAGE_BYTES);
        float4* lo = reinterpret_cast<float4*>(stages_lo + stage * CC_STAGE_BYTES);
#pragma unroll
        for (int j = 0; j < CC_STAGE_BYTES / 16 / 128; ++j) {
          const float4 v = hi[j * 128 + t];
          uint32_t hx, hy, hz, hw;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hx) : "f"(v.x));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hy) : "f"(v.y));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hz) : "f"(v.z));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hw) : "f"(v.w));
          const float4 h = make_float4(__uint_as_float(hx), __uint_as_float(hy), __uint_as_float(hz), __uint_as_float(hw));
          hi[j * 128 + t] = h;
          lo[j * 128 + t] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
        mbar_arrive(smem_u32(ready_bar + stage));
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp (4+q) owns TMEM lanes [32q, 32q+32) =====
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    double loss = 0.0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(acc_full + acc), acc_phase);
      tcgen05_fence_after();
      float z[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * CC_N, z);
      tcgen05_fence_before();
      mbar_arrive(smem_u32(acc_empty + acc));
      const long long g = (long long)tile * CC_BM + q * 32 + lane;
      if (g < R) loss += act_epilogue(lv, act, g, z, K, act_mode, bias, labels, flags);
      if (++acc == CC_ACC) { acc = 0; acc_phase ^= 1; }
    }
    if (loss_partials) {
      loss = warp_sum_d(loss);
      if (lane == 0) red[q] = loss;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && loss_partials) loss_partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CC_ACC * CC_N));
  }
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

__global__ void __launch_bounds__(CB_THREADS) condconv_bwd_kernel(Levels lv, const float* __restrict__ rows, const float* __restrict__ weight,
                                                                  ConstActPtrs act, ConstActPtrs dact, const int64_t* __restrict__ labels,
                                                                  float loss_scale, const float* __restrict__ d_loss, int K, int act_mode,
                                                                  int num_tiles,
                                                                  float* __restrict__ d_rows, float* __restrict__ partial_w,
                                                                  float* __restrict__ partial_b) {
  __shared__ __align__(16) float dzs[CB_ROWS][16];
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const int c = threadIdx.x;
  float w[16], acc[16];
  float bacc = 0.f;
  if (d_loss) loss_scale *= __ldg(d_loss);  // d(total)/d(act_loss), a device scalar: no host sync
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    w[k] = (k < K) ? __ldg(weight + k * CC_C + c) : 0.f;
    acc[k] = 0.f;
  }
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const long long g0 = (long long)tile * CB_ROWS;
    __syncthreads();
    if (threadIdx.x < CB_ROWS) {
      float dz[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) dz[k] = 0.f;
      if (g0 + threadIdx.x < R) row_dlogits(lv, act, dact, g0 + threadIdx.x, K, act_mode, labels, loss_scale, dz);
#pragma unroll
      for (int k = 0; k < 16; ++k) dzs[threadIdx.x][k] = dz[k];
    }
    __syncthreads();
    const int nrows = (int)min((long long)CB_ROWS, R - g0);
    if (threadIdx.x < 16) {
      float s = 0.f;
      for (int r = 0; r < nrows; ++r) s += dzs[r][threadIdx.x];
      bacc += s;
    }
#pragma unroll 8
    for (int r = 0; r < CB_ROWS; ++r) {
      if (r < nrows) {
        const float x = __ldg(rows + (g0 + r) * CC_C + c);
        const float4* d4 = reinterpret_cast<const float4*>(dzs[r]);
        float dx = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          if (k4 * 4 < K) {
            const float4 d = d4[k4];
            dx = fmaf(d.x, w[k4 * 4 + 0], dx);
            dx = fmaf(d.y, w[k4 * 4 + 1], dx);
            dx = fmaf(d.z, w[k4 * 4 + 2], dx);
            dx = fmaf(d.w, w[k4 * 4 + 3], dx);
            acc[k4 * 4 + 0] = fmaf(d.x, x, acc[k4 * 4 + 0]);
            acc[k4 * 4 + 1] = fmaf(d.y, x, acc[k4 * 4 + 1]);
            acc[k4 * 4 + 2] = fmaf(d.z, x, acc[k4 * 4 + 2]);
            acc[k4 * 4 + 3] = fmaf(d.w, x, acc[k4 * 4 + 3]);
          }
        }
        d_rows[(g0 + r) * CC_C + c] = dx;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) partial_w[((long long)blockIdx.x * 16 + k) * CC_C + c] = acc[k];
  if (threadIdx.x < 16) partial_b[blockIdx.x * 16 + threadIdx.x] = bacc;
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

template <int KP>
__global__ void __launch_bounds__(CB_THREADS, 4) condconv_bwd_rows_kernel(const float* __restrict__ rows, const float* __restrict__ weight,
                                                                         const float* __restrict__ dzw, long long R, int num_tiles,
                                                                         float* __restrict__ d_rows, float* __restrict__ partial_w) {
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
        if (r0 + j < nrows) dr[(long long)(r0 + j) * CC_C] = dx;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) partial_w[((long long)blockIdx.x * 16 + k) * CC_C + c] = acc[k];
}

__global__ void __launch_bounds__(256) condconv_bwd_reduce_kernel(const float* __restrict__ partial_w, const float* __restrict__ partial_b,
                                                                  int n_parts, int n_parts_b, int K, float* __restrict__ d_weight,
                                                                  float* __restrict__ d_bias) {
  const int k = blockIdx.x, c = threadIdx.x;
  float s = 0.f;
  for (int i = 0; i < n_parts; ++i) s += partial_w[((long long)i * 16 + k) * CC_C + c];
  d_weight[k * CC_C + c] = s;
  if (d_bias && c == 0) {
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

static int g_fwd_attr_set = 0;

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
  // impl 0: tcgen05, operands in tensor memory (product); 1: fp32 FFMA (verification); 2: tcgen05, operands in shared memory
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
  if (impl != 0 && impl != 2) return SCAN_EINVAL;
  CUtensorMap mx, mw;
  rc = make_rowmajor_map(&mx, rows, (uint64_t)R, CC_C, CC_BM);
  if (rc) return rc;
  rc = make_rowmajor_map(&mw, weight, (uint64_t)num_classes, CC_C, CC_N);
  if (rc) return rc;
  if (!g_fwd_attr_set) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(condconv_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CC_SMEM));
    g_fwd_attr_set = 1;
  }
  if (impl == 0) {  // product kernel: A operands in tensor memory (condconv_ts.inl)
    static int ts_attr = 0;
    if (!ts_attr) {
      SCAN_CUDA_CHECK(cudaFuncSetAttribute(condconv_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
      ts_attr = 1;
    }
    // layout experiment (tools/bench_kernels.py): SCAN_B200_CC_KMAJOR=1 -> `rows` holds the k-block-major copy [8][R][32]
    static const int kmajor = getenv("SCAN_B200_CC_KMAJOR") ? atoi(getenv("SCAN_B200_CC_KMAJOR")) : 0;
    if (kmajor) {
      rc = make_rowmajor_map(&mx, rows, (uint64_t)R * CC_KB, CC_BK, CC_BM);
      if (rc) return rc;
    }
    condconv_fwd_ts_kernel<<<grid, TS_THREADS, TS_SMEM, st>>>(mx, mw, lv, act, bias, labels, labels ? loss_partials : nullptr, flags,
                                                             num_classes, act_mode, num_tiles, kmajor ? R : 0);
    SCAN_LAUNCH_CHECK("condconv_fwd_ts_kernel");
    return SCAN_OK;
  }
  condconv_fwd_tc_kernel<<<grid, CC_THREADS, CC_SMEM, st>>>(mx, mw, lv, act, bias, labels, labels ? loss_partials : nullptr, flags,
                                                           num_classes, act_mode, num_tiles);
  SCAN_LAUNCH_CHECK("condconv_fwd_tc_kernel");
  return SCAN_OK;
}

extern "C" int64_t scan_condconv_bwd_workspace_bytes(const scan_levels_t* lvh, int32_t num_classes) {
  scan::Levels lv;
  if (scan::make_levels(lvh, &lv) || num_classes < 1 || num_classes > SCAN_MAX_CLASSES) return 0;
  const int64_t parts = 4ll * scan::sm_count();
  // d_weight partials [parts][16][256] | bias partials [parts][16] | dz [R][num_classes]
  return parts * 16 * scan::CC_C * 4 + parts * 16 * 4 + lv.row_off[SCAN_MAX_LEVELS] * num_classes * 4 + 512;
}

extern "C" int scan_condconv_bwd(const scan_levels_t* lvh, const float* rows, const float* weight, int32_t num_classes,
                                 int32_t act_mode, const void* const* act_nchw_host, const void* const* d_act_nchw_host,
                                 const int64_t* labels, float loss_scale, const float* d_loss, float* d_rows, float* d_weight,
                                 float* d_bias, void* workspace, int64_t workspace_bytes, void* stream) {
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
  static const int fused = getenv("SCAN_B200_CONDCONV_BWD_FUSED") ? atoi(getenv("SCAN_B200_CONDCONV_BWD_FUSED")) : 0;
  if (fused) {  // round-1 single-pass kernel (kept for comparison)
    condconv_bwd_kernel<<<parts, CB_THREADS, 0, st>>>(lv, rows, weight, act, dact, labels, loss_scale, d_loss, num_classes, act_mode,
                                                     num_tiles, d_rows, pw, pb);
    SCAN_LAUNCH_CHECK("condconv_bwd_kernel");
    condconv_bwd_reduce_kernel<<<num_classes, 256, 0, st>>>(pw, pb, parts, parts, num_classes, d_weight, d_bias);
    SCAN_LAUNCH_CHECK("condconv_bwd_reduce_kernel");
    return SCAN_OK;
  }
  condconv_dlogit_kernel<<<parts_b, 256, 0, st>>>(lv, act, dact, labels, loss_scale, d_loss, num_classes, act_mode, dzw, pb);
  SCAN_LAUNCH_CHECK("condconv_dlogit_kernel");
  switch (num_classes) {
#define SCAN_BWD_CASE(KP)                                                                                                \
  case KP:                                                                                                               \
    condconv_bwd_rows_kernel<KP><<<parts, CB_THREADS, 0, st>>>(rows, weight, dzw, R, num_tiles, d_rows, pw);           \
    break;
    SCAN_BWD_CASE(1) SCAN_BWD_CASE(2) SCAN_BWD_CASE(3) SCAN_BWD_CASE(4) SCAN_BWD_CASE(5) SCAN_BWD_CASE(6) SCAN_BWD_CASE(7) SCAN_BWD_CASE(8)
    SCAN_BWD_CASE(9) SCAN_BWD_CASE(10) SCAN_BWD_CASE(11) SCAN_BWD_CASE(12) SCAN_BWD_CASE(13) SCAN_BWD_CASE(14) SCAN_BWD_CASE(15)
    SCAN_BWD_CASE(16)
#undef SCAN_BWD_CASE
  }
  SCAN_LAUNCH_CHECK("condconv_bwd_rows_kernel");
  condconv_bwd_reduce_kernel<<<num_classes, 256, 0, st>>>(pw, pb, parts, parts_b, num_classes, d_weight, d_bias);
  SCAN_LAUNCH_CHECK("condconv_bwd_reduce_kernel");
  return SCAN_OK;
}
