// f1: the 3x3 convolutions of the GRAPHHead towers (modeling/rpn/fcos/condgraph.py:68-119: head_in =
// [Conv3x3 + GroupNorm(32) + ReLU] x 2, head_out = Conv3x3 + ReLU; called at :549 and :383) as an implicit GEMM on the 5th-gen
// tensor cores, over all FPN levels and images of a call in ONE launch, reading and writing the [R, C] rows layout (pixels of
// level 0 image 0, image 1, ..., level 1, ...; channels contiguous) the rest of the path works on.
//
//     Y[p, co] = sum_{tap} sum_{ci} X[p + off(tap), ci] * W[tap][co][ci]          (zero outside the image)
//
//   * no im2col and no padded copy: the A operand of tap (ky, kx) is the SAME activation tensor read through a 4-D TMA tensor
//     map (C, W, H, N) of the level with the box origin moved by (kx - 1, ky - 1); the TMA unit zero-fills what falls outside
//     the image, which is exactly the convolution's padding.  A tile is a th x tw rectangle of one image (th * tw <= 128, shape
//     chosen per level to minimise the tile count: 93 % of the MMA rows are real pixels for the 800 x 1344 pyramid);
//   * B = the weights repacked once per call to [9][Cout][Cin] (fprop) or [9 flipped][Cin][Cout] (data gradient: the same
//     kernel run on dY), K-major, 32-column (128-byte) boxes, 128-byte swizzle;
//   * K loop = 9 taps x Cin/32 chunks through an mbarrier ring; tcgen05.mma kind::tf32, fp32 accumulators in tensor memory;
//   * cta_group::2: the two CTAs of a cluster compute two pixel tiles against the same weights; each loads its own A tile and
//     HALF of the B tile, one M = 256 instruction issued by the leader reads both halves (half the weight traffic from L2 and
//     half the shared-memory operand reads per SM — the reason a 1-CTA tf32 GEMM cannot reach the pipe's peak);
//   * persistent CTAs, two accumulator buffers in TMEM (2 x 256 columns): the epilogue of tile i (TMEM -> registers -> bias /
//     addend / ReLU -> 128-byte row segments in HBM) overlaps the main loop of tile i + 1;
//   * arithmetic: single-pass TF32 (torch's and the reference's default for cuDNN convolutions, `allow_tf32 = True`) or, for
//     the parity runs, 3xTF32 (hi.hi + hi.lo + lo.hi on pre-split planes) which is fp32-accurate.
// The weight gradient is conv_wgrad_kernel below.
#include "tc_common.cuh"

namespace scan {

constexpr int CV_BM = 128;                       // pixels (MMA rows) per CTA tile
constexpr int CV_BK = 32;                        // fp32 channels per TMA box row = 128 bytes
constexpr int CV_A_BYTES = CV_BM * CV_BK * 4;    // 16 KB
constexpr int CV_THREADS = 256;                  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue

template <int CG, int BN>
struct ConvCfg {
  static constexpr int B_ROWS = BN / CG;                       // weight rows this CTA stages per k-chunk
  static constexpr int B_BYTES = B_ROWS * CV_BK * 4;
  static constexpr int STAGE_BYTES = CV_A_BYTES + B_BYTES;     // 48 KB (1 CTA) / 32 KB (pair)
  static constexpr int STAGES = (196 * 1024) / STAGE_BYTES > 8 ? 8 : (196 * 1024) / STAGE_BYTES;
  static constexpr int SMEM = 1024 + STAGES * STAGE_BYTES + 256;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator buffers
};

struct alignas(64) ConvMaps {
  CUtensorMap a[2][SCAN_MAX_LEVELS];   // [hi | lo plane][level]: 4-D (C, W, H, N), box (32, tw, th, 1)
  CUtensorMap a2[2][SCAN_MAX_LEVELS];  // optional second input tensor (its channels follow the first one's in the weight columns)
  CUtensorMap b[2];                    // [hi | lo plane]: 2-D [9 * rows_per_tap, Cin + Cin2], box (32, B_ROWS)
};

struct ConvLevel {
  int tile_off;          // first tile of the level
  int tiles_x, tiles_y;  // tiles per image
  int tw, th;            // tile rectangle
  int h, w;
  long long row_off;     // first row of the level in the rows layout
};

struct ConvArgs {
  int n_levels, n_images;
  ConvLevel lv[SCAN_MAX_LEVELS];
  int n_tiles;           // pixel tiles over all levels and images
  int n_blocks;          // BN-wide blocks of output channels
  int k_chunks;          // Cin / 32
  int k_chunks2;         // 32-channel chunks of the second input (0: none)
  int n_terms;           // 1: single-pass tf32, 3: 3xTF32
  int n_taps;            // 9: 3x3 convolution; 1: plain rows GEMM (the "1x1" case: no shifted reads)
  int seg_len;           // k-stages accumulated in tensor memory before the epilogue takes the partial sum (>= k_iters: one segment)
  int rows_per_tap;      // rows of one tap in the packed weight matrix (>= n_blocks * BN)
  float* out;            // [R, ldo]
  int ldo;
  int n_valid;           // valid output channels (<= n_blocks * BN)
  const float* bias;     // [n_valid] or null
  const float* addend;   // [R, ldo] or null: out = act(acc + bias + addend)
  const float* mask;     // [R, ldo] or null: out = mask > 0 ? out : 0 (the ReLU of a saved forward output, for data gradients)
  int relu;
  // GroupNorm statistics of the output as a by-product (BN = 256 only; null: off): every epilogue warp writes the sums and sums
  // of squares of (out + gn_bias) over its 32 pixels for each of the 32 groups of 8 channels -> gn_partial [tile][warp][32][2];
  // conv_gn_finalize_kernel combines them per (level, image) in fp64 in a fixed order.  Saves the statistics pass of
  // scan_gn_relu_fwd (one full read of the tensor); the epilogue has the time (it overlaps the next tile's main loop).
  float* gn_partial;
  const float* gn_bias;  // [256] or null: the convolution bias that scan_gn_relu adds on the fly
};

struct ConvTile {
  int l, n, x0, y0, tw, npix, h, w;
  long long row_base;    // row of pixel (0, 0) of the image
  uint32_t a_bytes;
};

__device__ __forceinline__ ConvTile conv_decode(const ConvArgs& g, int t) {
  int l = 0;
#pragma unroll
  for (int j = 1; j < SCAN_MAX_LEVELS; ++j)
    if (j < g.n_levels && t >= g.lv[j].tile_off) l = j;
  const ConvLevel& L = g.lv[l];
  const int rel = t - L.tile_off, per = L.tiles_x * L.tiles_y;
  const int n = rel / per, r = rel - n * per;
  const int ty = r / L.tiles_x, tx = r - ty * L.tiles_x;
  ConvTile o;
  o.l = l;
  o.n = n;
  o.x0 = tx * L.tw;
  o.y0 = ty * L.th;
  o.tw = L.tw;
  o.npix = L.tw * L.th;
  o.h = L.h;
  o.w = L.w;
  o.row_base = L.row_off + (long long)n * L.h * L.w;
  o.a_bytes = (uint32_t)o.npix * CV_BK * 4;
  return o;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int CG>
__device__ __forceinline__ void cv_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  if constexpr (CG == 1)
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
template <int CG>
__device__ __forceinline__ void cv_tma_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if constexpr (CG == 1)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void cv_tma_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
template <int CG>
__device__ __forceinline__ void cv_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) {
    if constexpr (CG == 1)
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
          : "memory");
  }
}
// completion of every MMA issued so far -> one arrival on the barrier at this offset (in both CTAs of the pair for CG = 2)
template <int CG>
__device__ __forceinline__ void cv_commit(uint32_t bar) {
  if (elect_one_sync()) {
    if constexpr (CG == 1)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                   "h"((uint16_t)3)
                   : "memory");
  }
}

__device__ __forceinline__ void cv_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------- forward / data-gradient kernel
template <int CG, int BN>
__global__ void __launch_bounds__(CV_THREADS, 1) conv3x3_kernel(const __grid_constant__ ConvMaps maps, const ConvArgs g) {
  using Cfg = ConvCfg<CG, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  uint64_t* bars = (uint64_t*)(stages + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                      // TMA -> MMA (lives in the leader CTA)
  uint64_t* empty_bar = bars + Cfg::STAGES;       // MMA -> TMA (one per CTA, signalled by a multicast commit)
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;    // [2] MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;             // [2] epilogue -> MMA (lives in the leader CTA)
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int pair = blockIdx.x / CG, n_pairs = gridDim.x / CG;
  const int n_tile_groups = (g.n_tiles + CG - 1) / CG;
  const int n_work = n_tile_groups * g.n_blocks;
  const int kc_all = g.k_chunks + g.k_chunks2;
  const int k_iters = g.n_terms * g.n_taps * kc_all;
  constexpr uint32_t IDESC = umma_idesc_tf32(CV_BM * CG, BN);

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(acc_full + i), 1);
      mbar_init(smem_u32(acc_empty + i), 4 * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(Cfg::TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
  }
  tcgen05_fence_before();
  __syncthreads();                               // CTA-level: the allocator's write of tmem_slot, the barrier inits
  if constexpr (CG == 2) cluster_sync_all();     // pair-level: the peer's barriers are initialised before any remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs of a pair: own A tile, own half of the weight tile) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = pair; w < n_work; w += n_pairs) {
        const int tg = w / g.n_blocks, nb = w - tg * g.n_blocks;
        const ConvTile t = conv_decode(g, min(tg * CG + (int)rank, g.n_tiles - 1));
        uint32_t bytes = t.a_bytes + Cfg::B_BYTES;
        if constexpr (CG == 2) bytes += conv_decode(g, min(tg * CG + (int)(rank ^ 1u), g.n_tiles - 1)).a_bytes + Cfg::B_BYTES;
        const int b_row0 = nb * BN + (int)rank * Cfg::B_ROWS;
        for (int term = 0; term < g.n_terms; ++term) {
          const CUtensorMap* ma = &maps.a[term == 2 ? 1 : 0][t.l];
          const CUtensorMap* ma2 = &maps.a2[term == 2 ? 1 : 0][t.l];
          const CUtensorMap* mb = &maps.b[term == 1 ? 1 : 0];
          for (int tap = 0; tap < g.n_taps; ++tap) {
            const int dy = g.n_taps == 9 ? tap / 3 - 1 : 0, dx = g.n_taps == 9 ? tap - (tap / 3) * 3 - 1 : 0;
            for (int kc = 0; kc < kc_all; ++kc) {
              mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
              uint32_t fb = smem_u32(full_bar + stage);
              if (rank == 0) mbar_expect_tx(fb, bytes);
              if constexpr (CG == 2) fb = mapa_shared(fb, 0);
              const uint32_t dst = smem_u32(stages + stage * Cfg::STAGE_BYTES);
              if (kc < g.k_chunks)
                cv_tma_4d<CG>(dst, ma, fb, kc * CV_BK, t.x0 + dx, t.y0 + dy, t.n);
              else
                cv_tma_4d<CG>(dst, ma2, fb, (kc - g.k_chunks) * CV_BK, t.x0 + dx, t.y0 + dy, t.n);
              cv_tma_2d<CG>(dst + CV_A_BYTES, mb, fb, kc * CV_BK, tap * g.rows_per_tap + b_row0);
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA): converged warp, one elected lane per instruction =====
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;
      for (int w = pair; w < n_work; w += n_pairs) {
        for (int i0 = 0; i0 < k_iters; i0 += g.seg_len, ++j) {
          const int buf = j & 1, i1 = min(i0 + g.seg_len, k_iters);
          mbar_wait(smem_u32(acc_empty + buf), ((uint32_t)(j >> 1) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t d = tmem_base + buf * BN;
          for (int i = i0; i < i1; ++i) {
            mbar_wait(smem_u32(full_bar + stage), phase);
            tcgen05_fence_after();
            const uint32_t a = smem_u32(stages + stage * Cfg::STAGE_BYTES), b = a + CV_A_BYTES;
#pragma unroll
            for (int k = 0; k < CV_BK / 8; ++k)
              cv_mma<CG>(d, umma_desc_sw128(a + k * 32), umma_desc_sw128(b + k * 32), IDESC, (i > i0) | (k != 0));
            cv_commit<CG>(smem_u32(empty_bar + stage));
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
          }
          cv_commit<CG>(smem_u32(acc_full + buf));
        }
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp (4 + q) owns TMEM lanes [32 q, 32 q + 32); thread = pixel =====
    const int q = warp - 4;
    const int m = q * 32 + lane;
    int j = 0;
    uint32_t ae = smem_u32(acc_empty);
    if constexpr (CG == 2) ae = mapa_shared(ae, 0);
    for (int w = pair; w < n_work; w += n_pairs) {
      const int tg = w / g.n_blocks, nb = w - tg * g.n_blocks;
      const int ti = tg * CG + (int)rank;
      const ConvTile t = conv_decode(g, min(ti, g.n_tiles - 1));
      const int py = m / t.tw, px = m - py * t.tw;
      const bool valid = ti < g.n_tiles && m < t.npix && t.y0 + py < t.h && t.x0 + px < t.w;
      const long long row = t.row_base + (long long)(t.y0 + py) * t.w + (t.x0 + px);
      float* out = g.out + (valid ? row : 0) * g.ldo + nb * BN;
      const float* add = g.addend ? g.addend + (valid ? row : 0) * g.ldo + nb * BN : nullptr;
      const float* msk = g.mask ? g.mask + (valid ? row : 0) * g.ldo + nb * BN : nullptr;
      // one pass per accumulation segment: the partial sums of a tile are combined in its output row with round-to-nearest adds
      // (this thread re-reads what it wrote), bias / addend / ReLU go on with the last one
      for (int i0 = 0; i0 < k_iters; i0 += g.seg_len, ++j) {
        const int buf = j & 1;
        const bool first = i0 == 0, last = i0 + g.seg_len >= k_iters;
        mbar_wait(smem_u32(acc_full + buf), (uint32_t)(j >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          cv_ld32(tl + c * 32, v);
          const int col = nb * BN + c * 32;
          const bool act = valid && col < g.n_valid;
          if (act) {
            if (col + 32 <= g.n_valid) {
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                float4 o = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
                float4* dst = reinterpret_cast<float4*>(out + c * 32 + e);
                if (!first) {
                  const float4 p = *dst;
                  o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                }
                if (last) {
                  if (g.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + col + e));
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                  }
                  if (add) {
                    const float4 p = __ldg(reinterpret_cast<const float4*>(add + c * 32 + e));
                    o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
                  }
                  if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                  if (msk) {
                    const float4 q4 = __ldg(reinterpret_cast<const float4*>(msk + c * 32 + e));
                    o.x = q4.x > 0.f ? o.x : 0.f; o.y = q4.y > 0.f ? o.y : 0.f; o.z = q4.z > 0.f ? o.z : 0.f; o.w = q4.w > 0.f ? o.w : 0.f;
                  }
                }
                *dst = o;
                v[e] = o.x; v[e + 1] = o.y; v[e + 2] = o.z; v[e + 3] = o.w;
              }
            } else {
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (col + e < g.n_valid) {
                  float o = v[e] + (first ? 0.f : out[c * 32 + e]);
                  if (last) {
                    o += (g.bias ? __ldg(g.bias + col + e) : 0.f) + (add ? __ldg(add + c * 32 + e) : 0.f);
                    o = g.relu ? fmaxf(o, 0.f) : o;
                    if (msk) o = __ldg(msk + c * 32 + e) > 0.f ? o : 0.f;
                  }
                  out[c * 32 + e] = o;
                }
            }
          }
          if constexpr (BN == 256) {
            if (g.gn_partial != nullptr && last) {   // warp-uniform: all 32 lanes take part in the shuffles
              // this chunk = 4 groups of 8 channels: {sum, sum of squares} per group over this thread's pixel ...
              float st[8];
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                float sm = 0.f, sq = 0.f;
#pragma unroll
                for (int e = 0; e < 8; e += 4) {
                  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (g.gn_bias) b = __ldg(reinterpret_cast<const float4*>(g.gn_bias + col + gq * 8 + e));
                  const float x0 = v[gq * 8 + e] + b.x, x1 = v[gq * 8 + e + 1] + b.y, x2 = v[gq * 8 + e + 2] + b.z,
                              x3 = v[gq * 8 + e + 3] + b.w;
                  sm += (x0 + x1) + (x2 + x3);
                  sq += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
                }
                st[2 * gq] = act ? sm : 0.f;          // a select, not a product: rows outside the image may hold anything
                st[2 * gq + 1] = act ? sq : 0.f;
              }
              // ... summed over the warp's 32 pixels by a halving butterfly (8 -> 4 -> 2 -> 1 values per lane, 9 shuffles):
              // lane bits 4, 3, 2 select which of the 8 values the lane ends up holding; fixed order, deterministic
              float r4[4], r2[2], r1;
              const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float keep = u16 ? st[i + 4] : st[i], send = u16 ? st[i] : st[i + 4];
                r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
              }
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const float keep = u8 ? r4[i + 2] : r4[i], send = u8 ? r4[i] : r4[i + 2];
                r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
              }
              {
                const float keep = u4 ? r2[1] : r2[0], send = u4 ? r2[0] : r2[1];
                r1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
              }
              r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
              r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
              const int idx = (u16 ? 4 : 0) + (u8 ? 2 : 0) + (u4 ? 1 : 0);
              if ((lane & 3) == 0 && ti < g.n_tiles) g.gn_partial[((long long)ti * 4 + q) * 64 + c * 8 + idx] = r1;
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(ae + buf * 8); else mbar_arrive(ae + buf * 8);
        }
      }
    }
  }
  tcgen05_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------- weight-gradient kernel
//     dW[tap][co][ci] = sum_p dY[p, co] * X[p + off(tap), ci]
// M = co, N = ci, K = pixels: both operands are read as they lie in memory (rows = pixels, channels contiguous), i.e. MN-major
// for the tensor core (instruction-descriptor bits 15 / 16, shared-memory layout type SWIZZLE_128B_BASE32B = the TMA mode
// SWIZZLE_128B_ATOM_32B; tools/mma_mnmajor_probe.cu), so no transposed copy of the activations is ever made.  A k-chunk is a
// ch x cw rectangle of exactly 32 pixels of one image (dY at the rectangle, X at the rectangle moved by the tap; TMA zero fill
// = the padding, and a rectangle hanging over the image edge contributes zeros).  CTA pairs (cta_group::2): CTA r stages the
// co half [128 r, 128 r + 128) of dY and the ci half of X, one M = 256, N = 256 instruction per 8 pixels.
// Work item = (pixel segment, tap): `seg_len` consecutive chunks accumulated in TMEM, then added into the pair's private slot
// for that tap (the owning thread re-reads its own previous value: deterministic).  Items are dealt round-robin with the tap
// fastest, so the nine taps of a segment run at the same time on different SMs and share the segment's rows in L2.  The
// slots are summed over pairs in a fixed order by conv_wgrad_reduce_kernel.
constexpr int WG_KB = 32;                         // pixels per k-chunk
constexpr int WG_OP_BYTES = 128 * WG_KB * 4;      // 128 channels x 32 pixels = 16 KB per operand per CTA
constexpr int WG_STAGE_BYTES = 2 * WG_OP_BYTES;
constexpr int WG_STAGES = 6;
constexpr int WG_SMEM = 1024 + WG_STAGES * WG_STAGE_BYTES + 256;
constexpr int WG_TILE = 256 * 256;                // floats of one (256 co x 256 ci) partial

struct alignas(64) WgMaps {
  CUtensorMap dy[2][SCAN_MAX_LEVELS];   // [hi | lo][level]: (32, W, H, C / 32, N), box (32, cw, ch, 4, 1), SWIZZLE_128B_ATOM_32B
  CUtensorMap x[2][SCAN_MAX_LEVELS];
};

struct WgLevel {
  int chunk_off, chunks_x, chunks_y, cw, ch;
};

struct WgArgs {
  int n_levels, n_images;
  WgLevel lv[SCAN_MAX_LEVELS];
  int n_chunks, seg_len, n_seg;
  int n_terms;
  int n_taps;               // 9: the convolution's taps (X read at shifted positions); 1: X is read in place
  int m_blocks, n_blocks;   // 256-wide blocks of co / (64 NB)-wide blocks of ci
  int n_items;              // n_seg * 9 * m_blocks * n_blocks
  int period, slots;        // distinct (tap, block) kinds a pair meets; slots = min(period, items per pair)
  float* partial;           // [pairs][slots][256][256]
};

__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t saddr) {
  // MN-major, SWIZZLE_128B_BASE32B (type 1): LBO = 4096 (next 32-channel block), SBO = 512
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(4096 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

// NB = 32-channel blocks of the X operand each CTA stages (4: the 256-wide convolution; 2: the 128-wide tap-spread class maps)
template <int NB>
__global__ void __launch_bounds__(CV_THREADS, 1) conv_wgrad_kernel(const __grid_constant__ WgMaps maps, const WgArgs g) {
  constexpr int WN = 64 * NB;                           // output columns of the pair's tile
  constexpr int X_BYTES = NB * 4096;
  constexpr int STAGE = WG_OP_BYTES + X_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  uint64_t* bars = (uint64_t*)(stages + WG_STAGES * STAGE);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* acc_full = bars + 2 * WG_STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int kinds = g.n_taps * g.m_blocks * g.n_blocks;
  constexpr uint32_t IDESC = umma_idesc_tf32(256, WN) | (1u << 15) | (1u << 16);

  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(acc_full + i), 1);
      mbar_init(smem_u32(acc_empty + i), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = pair; it < g.n_items; it += n_pairs) {
        const int kind = it % kinds, seg = it / kinds;
        const int tap = kind % g.n_taps, nb = (kind / g.n_taps) % g.n_blocks, mb = kind / (g.n_taps * g.n_blocks);
        const int dy = g.n_taps == 9 ? tap / 3 - 1 : 0, dx = g.n_taps == 9 ? tap - (tap / 3) * 3 - 1 : 0;
        const int c0 = seg * g.seg_len, c1 = min(c0 + g.seg_len, g.n_chunks);
        const int a_cb = mb * 8 + (int)rank * 4, b_cb = nb * 2 * NB + (int)rank * NB;     // first 32-channel block of this CTA's half
        // position of the segment's first chunk; stepped incrementally afterwards (one thread feeds a 512-cycle stage)
        int l0 = 0;
#pragma unroll
        for (int q = 1; q < SCAN_MAX_LEVELS; ++q)
          if (q < g.n_levels && c0 >= g.lv[q].chunk_off) l0 = q;
        const int rel0 = c0 - g.lv[l0].chunk_off, per0 = g.lv[l0].chunks_x * g.lv[l0].chunks_y;
        const int n0 = rel0 / per0, r0 = rel0 - n0 * per0;
        const int cy0 = r0 / g.lv[l0].chunks_x, cx0 = r0 - cy0 * g.lv[l0].chunks_x;
        for (int term = 0; term < g.n_terms; ++term) {
          const int pa = term == 2 ? 1 : 0, pb = term == 1 ? 1 : 0;
          int l = l0, n = n0, cy = cy0, cx = cx0;
          for (int c = c0; c < c1; ++c) {
            const WgLevel& L = g.lv[l];
            const int x0 = cx * L.cw, y0 = cy * L.ch;
            mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
            uint32_t fb = smem_u32(full_bar + stage);
            if (rank == 0) mbar_expect_tx(fb, 2 * STAGE);
            fb = mapa_shared(fb, 0);
            const uint32_t dst = smem_u32(stages + stage * STAGE);
            cv_tma_5d_pair(dst, &maps.dy[pa][l], fb, 0, x0, y0, a_cb, n);
            cv_tma_5d_pair(dst + WG_OP_BYTES, &maps.x[pb][l], fb, 0, x0 + dx, y0 + dy, b_cb, n);
            if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
            if (++cx == L.chunks_x) {
              cx = 0;
              if (++cy == L.chunks_y) {
                cy = 0;
                if (++n == g.n_images) { n = 0; ++l; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int j = 0;
      for (int it = pair; it < g.n_items; it += n_pairs, ++j) {
        const int seg = it / kinds;
        const int c0 = seg * g.seg_len, c1 = min(c0 + g.seg_len, g.n_chunks);
        const int k_iters = g.n_terms * (c1 - c0);
        const int buf = j & 1;
        mbar_wait(smem_u32(acc_empty + buf), ((uint32_t)(j >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + buf * WN;
        for (int i = 0; i < k_iters; ++i) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tcgen05_fence_after();
          const uint32_t a = smem_u32(stages + stage * STAGE), b = a + WG_OP_BYTES;
#pragma unroll
          for (int k = 0; k < WG_KB / 8; ++k) cv_mma<2>(d, umma_desc_mn32(a + k * 1024), umma_desc_mn32(b + k * 1024), IDESC, (i | k) != 0);
          cv_commit<2>(smem_u32(empty_bar + stage));
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        cv_commit<2>(smem_u32(acc_full + buf));
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const uint32_t ae = mapa_shared(smem_u32(acc_empty), 0);
    int j = 0;
    for (int it = pair; it < g.n_items; it += n_pairs, ++j) {
      const int buf = j & 1;
      const int slot = j % g.period;
      const bool first = j < g.period;
      float* out = g.partial + ((long long)pair * g.slots + slot) * (256 * WN) + (long long)((int)rank * 128 + q * 32 + lane) * WN;
      mbar_wait(smem_u32(acc_full + buf), (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
      const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + buf * WN;
#pragma unroll 1
      for (int c = 0; c < WN / 32; ++c) {
        float v[32];
        cv_ld32(tl + c * 32, v);
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float4 o = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
          float4* dst = reinterpret_cast<float4*>(out + c * 32 + e);
          if (!first) {
            const float4 p = *dst;
            o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
          }
          *dst = o;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(ae + buf * 8);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// d_w[co][ci][ky][kx] (element strides) = sum over the pairs that met kind (tap, nb, mb), pair order ascending.
// Block = 4 co rows x 64 float4 columns of one (mb, nb) tile; the (tap, pair) -> slot table is built once per block.
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ partial, int n_pairs, int slots, int period,
                                                                int n_items, int m_blocks, int n_blocks, long long s_co, long long s_ci,
                                                                long long s_ky, long long s_kx, float* __restrict__ d_w) {
  extern __shared__ signed char slot_of[];     // [9][n_pairs]
  const int kinds = 9 * m_blocks * n_blocks;
  const int mb = blockIdx.z / n_blocks, nb = blockIdx.z - mb * n_blocks;
  for (int i = threadIdx.x; i < 9 * n_pairs; i += blockDim.x) {     // (only row blockIdx.x of the table is used by this block)
    const int tap = i / n_pairs, p = i - tap * n_pairs;
    if (tap != (int)blockIdx.x) continue;
    const int kind = (mb * n_blocks + nb) * 9 + tap;
    int s = -1;
    for (int j = 0; j < period; ++j) {
      const int it = p + j * n_pairs;
      if (it >= n_items) break;
      if (it % kinds == kind) {
        s = j;
        break;
      }
    }
    slot_of[i] = (signed char)s;
  }
  __syncthreads();
  const int c4 = threadIdx.x & 63, row = blockIdx.y * 4 + (threadIdx.x >> 6);
  const long long off = (long long)row * 256 + c4 * 4;
  const int co = mb * 256 + row, ci = nb * 256 + c4 * 4;
  {
    const int tap = blockIdx.x;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < n_pairs; ++p) {
      const int sl = slot_of[tap * n_pairs + p];
      if (sl < 0) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(partial + ((long long)p * slots + sl) * WG_TILE + off));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    float* o = d_w + co * s_co + ci * s_ci + (tap / 3) * s_ky + (tap % 3) * s_kx;
    o[0] = s.x;
    o[s_ci] = s.y;
    o[2 * s_ci] = s.z;
    o[3 * s_ci] = s.w;
  }
}

// ---------------------------------------------------------------------------- thin weight gradient (the K class-map columns)
//     dW[co][256 + k][tap] = sum_p dY[p, co] * maps[p + off(tap), k]                     K * 9 <= 128
// The maps are first spread to S[p, k * 9 + tap] = maps[p + off(tap), k] (zero outside the image; [R, 128], 1/8 of dY's size),
// which turns the nine shifted products into ONE MN-major GEMM dY^T . S over the pixels: conv_wgrad_kernel<2> with a single
// "tap" -- dY is read once instead of nine times and the tensor core works on 128 useful columns instead of 9 of 256.
__global__ void __launch_bounds__(256) thin_spread_kernel(Levels lv, const float* __restrict__ maps32, int k, float* __restrict__ s) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long g = t >> 5;
  if (g >= lv.row_off[lv.n_levels]) return;
  const int c0 = (int)(t & 31) * 4;
  const int l = level_of_row(lv, g);
  const int w = lv.w[l], h = lv.h[l];
  const long long local = g - lv.row_off[l];
  const int hw = h * w;
  const int p = (int)(local % hw);
  const int y = p / w, x = p - y * w;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = c0 + e, kk = c / 9, tap = c - kk * 9;
    const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
    v[e] = (kk < k && yy >= 0 && yy < h && xx >= 0 && xx < w) ? __ldg(maps32 + (g + (long long)(yy - y) * w + (xx - x)) * 32 + kk) : 0.f;
  }
  *reinterpret_cast<float4*>(s + g * 128 + c0) = make_float4(v[0], v[1], v[2], v[3]);
}

// d_w[co][k][ky][kx] (element strides) = sum over the pairs' slot-0 tiles [256][128], pair order ascending
__global__ void __launch_bounds__(256) thin_wgrad_reduce_kernel(const float* __restrict__ partial, int n_pairs, int slots, int k,
                                                                long long s_co, long long s_k, long long s_ky, long long s_kx,
                                                                float* __restrict__ d_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // (co, col)
  const int co = i >> 7, col = i & 127;
  if (co >= 256 || col >= 9 * k) return;
  float s = 0.f;
  for (int p = 0; p < n_pairs; ++p) s += __ldg(partial + (long long)p * slots * (256 * 128) + co * 128 + col);
  const int kk = col / 9, tap = col - kk * 9;
  d_w[co * s_co + kk * s_k + (tap / 3) * s_ky + (tap % 3) * s_kx] = s;
}

// ---------------------------------------------------------------------------- operand preparation
// w[co][ci][ky][kx] with element strides s[4]  ->  packed[tap][row][col], rows / cols zero-padded to (rows_pad, cols_pad):
//   transpose = 0 (fprop):         row = co, col = ci, tap = ky * 3 + kx
//   transpose = 1 (data gradient): row = ci, col = co, tap = 8 - (ky * 3 + kx)        (the kernel rotated by 180 degrees)
// lo (may be null) receives the 3xTF32 residual plane rna_tf32(x - trunc_tf32(x)).
__device__ __forceinline__ float tf32_residual(float x) {
  const float d = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(d));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) conv_pack_weights_kernel(const float* __restrict__ w, long long s_co, long long s_ci,
                                                                long long s_ky, long long s_kx, int cout, int cin, int rows_pad,
                                                                int cols_pad, int transpose, float* __restrict__ hi,
                                                                float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = 9ll * rows_pad * cols_pad;
  if (i >= n) return;
  const int col = (int)(i % cols_pad), row = (int)((i / cols_pad) % rows_pad), tap = (int)(i / ((long long)cols_pad * rows_pad));
  const int t = transpose ? 8 - tap : tap;
  const int co = transpose ? col : row, ci = transpose ? row : col;
  float v = 0.f;
  if (co < cout && ci < cin) v = __ldg(w + co * s_co + ci * s_ci + (t / 3) * s_ky + (t % 3) * s_kx);
  hi[i] = v;
  if (lo) lo[i] = tf32_residual(v);
}

__global__ void __launch_bounds__(256) tf32_residual_kernel(const float4* __restrict__ x, long long n4, float4* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    lo[i] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
  }
}

// GroupNorm statistics from the convolution's epilogue partials: one warp per (level, image, group) sums the 4 * tiles-per-image
// {sum, sum of squares} pairs in fp64 (fixed order) -> stats [(l * N + n) * 32 + group][mean, 1 / sqrt(var + eps)], the array
// scan_gn_relu_fwd produces (biased variance like torch's native_group_norm)
struct ConvGnGeo {
  int n_levels, n_images;
  int tile_off[SCAN_MAX_LEVELS], per[SCAN_MAX_LEVELS], hw[SCAN_MAX_LEVELS];
};

__global__ void __launch_bounds__(256) conv_gn_finalize_kernel(ConvGnGeo geo, const float* __restrict__ partial, float eps,
                                                               float* __restrict__ stats) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= geo.n_levels * geo.n_images * 32) return;
  const int gi = i % 32, ln = i / 32, l = ln / geo.n_images, n = ln % geo.n_images;
  const long long first = ((long long)geo.tile_off[l] + (long long)n * geo.per[l]) * 4;
  const int count = geo.per[l] * 4;
  double s = 0.0, ss = 0.0;
  for (int j = lane; j < count; j += 32) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(partial) + (first + j) * 32 + gi);
    s += (double)v.x;
    ss += (double)v.y;
  }
  s = warp_sum_d(s);
  ss = warp_sum_d(ss);
  if (lane == 0) {
    const double m = (double)geo.hw[l] * 8.0;
    const double mean = s / m;
    const double var = fmax(ss / m - mean * mean, 0.0);
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// ---------------------------------------------------------------------------- host side
// tile rectangle of a level: th * tw <= 128 minimising the tile count (ties: the wider one, longer contiguous runs)
static void conv_pick_tile(int h, int w, int* th_out, int* tw_out) {
  long long best = -1;
  int bt = 1, bw = 1;
  for (int tw = 1; tw <= 128; ++tw) {
    const int th = 128 / tw;
    if (th < 1) break;
    const int th_eff = th > h ? h : th, tw_eff = tw > w ? w : tw;
    const long long tiles = ceil_div(h, th_eff) * ceil_div(w, tw_eff);
    if (best < 0 || tiles < best || (tiles == best && tw_eff > bw)) {
      best = tiles;
      bt = th_eff;
      bw = tw_eff;
    }
  }
  *th_out = bt;
  *tw_out = bw;
}

static int conv_make_a_map(CUtensorMap* m, const float* base, int c, int w, int h, int n, int tw, int th) {
  EncodeTiledFn enc;
  int rc = get_tensormap_encoder(&enc);
  if (rc) return rc;
  if (((uintptr_t)base & 15) || (c % 32)) return SCAN_EINVAL;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)tw, (cuuint32_t)th, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed (conv activations)");
    return SCAN_ECUDA;
  }
  return SCAN_OK;
}

static int conv_make_b_map(CUtensorMap* m, const float* base, long long rows, int cols, int box_rows) {
  EncodeTiledFn enc;
  int rc = get_tensormap_encoder(&enc);
  if (rc) return rc;
  if (((uintptr_t)base & 15) || (cols % 32)) return SCAN_EINVAL;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed (conv weights)");
    return SCAN_ECUDA;
  }
  return SCAN_OK;
}

template <int CG, int BN>
static int conv_launch(const ConvMaps& maps, const ConvArgs& g, cudaStream_t st) {
  using Cfg = ConvCfg<CG, BN>;
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr))
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(conv3x3_kernel<CG, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  const int n_work = (int)ceil_div(g.n_tiles, CG) * g.n_blocks;
  int pairs = sm_count() / CG;
  if (pairs > n_work) pairs = n_work;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * CG));
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SCAN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv3x3_kernel<CG, BN>, maps, g));
  return SCAN_OK;
}

}  // namespace scan

using namespace scan;

// ---------------------------------------------------------------------------- C ABI
extern "C" int64_t scan_conv3x3_packed_floats(int32_t rows, int32_t cols) {
  const long long rp = (rows + 255) / 256 * 256, cp = (cols + 31) / 32 * 32;
  return 9ll * rp * cp;
}

extern "C" int scan_conv3x3_pack_weights(const float* w, int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx, int cout, int cin,
                                         int transpose, float* packed_hi, float* packed_lo, void* stream) {
  if (!w || !packed_hi || cout < 1 || cin < 1) return SCAN_EINVAL;
  const int rows = transpose ? cin : cout, cols = transpose ? cout : cin;
  const int rp = (rows + 255) / 256 * 256, cp = (cols + 31) / 32 * 32;
  const long long n = 9ll * rp * cp;
  conv_pack_weights_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, s_co, s_ci, s_ky, s_kx, cout, cin, rp, cp,
                                                                                          transpose, packed_hi, packed_lo);
  SCAN_LAUNCH_CHECK("conv_pack_weights_kernel");
  return SCAN_OK;
}

extern "C" int scan_tf32_residual(const float* x, int64_t n, float* lo, void* stream) {
  if (!x || !lo || n < 0 || (n & 3) || ((uintptr_t)x & 15) || ((uintptr_t)lo & 15)) return SCAN_EINVAL;
  if (n == 0) return SCAN_OK;
  long long blocks = ceil_div(n / 4, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  tf32_residual_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), n / 4,
                                                                            reinterpret_cast<float4*>(lo));
  SCAN_LAUNCH_CHECK("tf32_residual_kernel");
  return SCAN_OK;
}

// y_rows[R, ldo] = act(conv3x3([x_rows[R, cin] | x2_rows[R, cin2]]; packed weights) + bias + addend), optionally masked.
// `packed` is scan_conv3x3_pack_weights' output for n_out output channels (rows padded to 256) and cin + cin2 input channels
// (cin, cin2 multiples of 32; x2_rows may be NULL).  x_lo / packed_lo (and x2_lo) non-null selects 3xTF32.  cta_group = 1 or 2.
static int conv_rows_impl(const scan_levels_t* levels, const float* x_rows, const float* x_lo, int cin, const float* x2_rows,
                          const float* x2_lo, int cin2, const float* packed, const float* packed_lo, int n_out, const float* bias,
                          const float* addend, const float* mask, int relu, float* y_rows, int ldo, int cta_group, int n_taps,
                          void* stream, float* gn_partial = nullptr, const float* gn_bias = nullptr, ConvGnGeo* gn_geo = nullptr) {
  Levels lv;
  int rc = make_levels(levels, &lv);
  if (rc) return rc;
  if (!x_rows || !packed || !y_rows || cin < 32 || (cin % 32) || n_out < 1 || ldo < n_out || (ldo % 4) || ((uintptr_t)y_rows & 15))
    return SCAN_EINVAL;
  if ((x_lo == nullptr) != (packed_lo == nullptr)) return SCAN_EINVAL;
  if (x2_rows && (cin2 < 32 || (cin2 % 32) || (x_lo == nullptr) != (x2_lo == nullptr))) return SCAN_EINVAL;
  if (!x2_rows) cin2 = 0;
  if (cta_group != 1 && cta_group != 2) return SCAN_EINVAL;
  if ((bias && ((uintptr_t)bias & 15)) || (addend && ((uintptr_t)addend & 15)) || (mask && ((uintptr_t)mask & 15))) return SCAN_EINVAL;
  ConvMaps maps;
  ConvArgs g = {};
  g.n_levels = lv.n_levels;
  g.n_images = lv.n_images;
  int tiles = 0;
  for (int l = 0; l < lv.n_levels; ++l) {
    ConvLevel& L = g.lv[l];
    conv_pick_tile(lv.h[l], lv.w[l], &L.th, &L.tw);
    L.h = lv.h[l];
    L.w = lv.w[l];
    L.tiles_x = (int)ceil_div(L.w, L.tw);
    L.tiles_y = (int)ceil_div(L.h, L.th);
    L.tile_off = tiles;
    L.row_off = lv.row_off[l];
    tiles += lv.n_images * L.tiles_x * L.tiles_y;
    if (gn_geo) {
      gn_geo->tile_off[l] = L.tile_off;
      gn_geo->per[l] = L.tiles_x * L.tiles_y;
      gn_geo->hw[l] = L.h * L.w;
    }
    rc = conv_make_a_map(&maps.a[0][l], x_rows + lv.row_off[l] * cin, cin, L.w, L.h, lv.n_images, L.tw, L.th);
    if (rc) return rc;
    maps.a[1][l] = maps.a[0][l];
    if (x_lo) {
      rc = conv_make_a_map(&maps.a[1][l], x_lo + lv.row_off[l] * cin, cin, L.w, L.h, lv.n_images, L.tw, L.th);
      if (rc) return rc;
    }
    maps.a2[0][l] = maps.a[0][l];
    maps.a2[1][l] = maps.a[1][l];
    if (x2_rows) {
      rc = conv_make_a_map(&maps.a2[0][l], x2_rows + lv.row_off[l] * cin2, cin2, L.w, L.h, lv.n_images, L.tw, L.th);
      if (rc) return rc;
      maps.a2[1][l] = maps.a2[0][l];
      if (x2_lo) {
        rc = conv_make_a_map(&maps.a2[1][l], x2_lo + lv.row_off[l] * cin2, cin2, L.w, L.h, lv.n_images, L.tw, L.th);
        if (rc) return rc;
      }
    }
  }
  for (int l = lv.n_levels; l < SCAN_MAX_LEVELS; ++l) {
    g.lv[l] = g.lv[lv.n_levels - 1];
    g.lv[l].tile_off = 0x7fffffff;
    maps.a[0][l] = maps.a[0][0];
    maps.a[1][l] = maps.a[1][0];
    maps.a2[0][l] = maps.a2[0][0];
    maps.a2[1][l] = maps.a2[1][0];
  }
  const int rows_per_tap = (n_out + 255) / 256 * 256;     // the packed weight pads every tap to a multiple of 256 rows
  const int bn = n_out <= 32 ? 32 : 256;                  // thin outputs (class maps): N = 32 MMAs, 48 cycles instead of 128
  g.n_tiles = tiles;
  g.n_blocks = (n_out + bn - 1) / bn;
  g.k_chunks = cin / 32;
  g.k_chunks2 = cin2 / 32;
  g.n_terms = x_lo ? 3 : 1;
  g.n_taps = n_taps;
  // single-pass TF32: one accumulation per tile (what cuDNN does).  3xTF32: the tensor core TRUNCATES when it adds into its
  // accumulator, so the fp32-accurate mode hands the partial sum to the epilogue every 4 k-stages (16 MMAs:
  // a bias of at most 16 x 2^-24, below the rounding noise of an fp32 FFMA convolution; epilogue-bound, ~4x slower, parity only)
  g.seg_len = x_lo ? 4 : g.n_terms * n_taps * (g.k_chunks + g.k_chunks2);
  g.rows_per_tap = rows_per_tap;
  g.out = y_rows;
  g.ldo = ldo;
  g.n_valid = n_out;
  g.bias = bias;
  g.addend = addend;
  g.mask = mask;
  g.relu = relu;
  if (gn_partial && (n_out != 256 || ((uintptr_t)gn_partial & 7) || (gn_bias && ((uintptr_t)gn_bias & 15)))) return SCAN_EINVAL;
  g.gn_partial = gn_partial;
  g.gn_bias = gn_bias;
  if (gn_geo) {
    gn_geo->n_levels = lv.n_levels;
    gn_geo->n_images = lv.n_images;
  }
  rc = conv_make_b_map(&maps.b[0], packed, (long long)n_taps * rows_per_tap, cin + cin2, bn / cta_group);
  if (rc) return rc;
  maps.b[1] = maps.b[0];
  if (packed_lo) {
    rc = conv_make_b_map(&maps.b[1], packed_lo, (long long)n_taps * rows_per_tap, cin + cin2, bn / cta_group);
    if (rc) return rc;
  }
  if (bn == 32) return cta_group == 2 ? conv_launch<2, 32>(maps, g, (cudaStream_t)stream) : conv_launch<1, 32>(maps, g, (cudaStream_t)stream);
  return cta_group == 2 ? conv_launch<2, 256>(maps, g, (cudaStream_t)stream) : conv_launch<1, 256>(maps, g, (cudaStream_t)stream);
}

extern "C" int scan_conv3x3_rows2(const scan_levels_t* levels, const float* x_rows, const float* x_lo, int cin, const float* x2_rows,
                                  const float* x2_lo, int cin2, const float* packed, const float* packed_lo, int n_out, const float* bias,
                                  const float* addend, const float* mask, int relu, float* y_rows, int ldo, int cta_group, void* stream) {
  return conv_rows_impl(levels, x_rows, x_lo, cin, x2_rows, x2_lo, cin2, packed, packed_lo, n_out, bias, addend, mask, relu, y_rows, ldo,
                        cta_group, 9, stream);
}

// y_rows [R, ldo] (first n_out columns) = act(x_rows [R, cin] . w^T + bias): the same persistent CTA-pair kernel with ONE tap and no
// shifted reads (a 1x1 convolution / rows GEMM).  w [rows padded to 256, cin] row-major, zero rows beyond n_out; w_lo: 3xTF32.
extern "C" int scan_conv1x1_rows(const scan_levels_t* levels, const float* x_rows, const float* x_lo, int32_t cin, const float* w,
                                 const float* w_lo, int32_t n_out, const float* bias, int32_t relu, float* y_rows, int32_t ldo,
                                 int32_t cta_group, void* stream) {
  return conv_rows_impl(levels, x_rows, x_lo, cin, nullptr, nullptr, 0, w, w_lo, n_out, bias, nullptr, nullptr, relu, y_rows, ldo,
                        cta_group, 1, stream);
}

extern "C" int scan_conv3x3_rows(const scan_levels_t* levels, const float* x_rows, const float* x_lo, int cin, const float* packed,
                                 const float* packed_lo, int n_out, const float* bias, const float* addend, int relu, float* y_rows, int ldo,
                                 int cta_group, void* stream) {
  return scan_conv3x3_rows2(levels, x_rows, x_lo, cin, nullptr, nullptr, 0, packed, packed_lo, n_out, bias, addend, nullptr, relu, y_rows,
                            ldo, cta_group, stream);
}

// The tower convolution in front of a GroupNorm(32): y_rows [R, 256] = conv3x3(x_rows) WITHOUT bias, plus the GroupNorm statistics
// of (y + gn_bias) as a by-product of the epilogue -> stats [L * N * 32][2] = what scan_gn_relu_fwd would compute in its
// statistics pass (scan_gn_relu_apply then normalises with them).  workspace: scan_conv3x3_gn_workspace_bytes(levels).
extern "C" int64_t scan_conv3x3_gn_workspace_bytes(const scan_levels_t* levels) {
  Levels lv;
  if (make_levels(levels, &lv)) return -1;
  long long tiles = 0;
  for (int l = 0; l < lv.n_levels; ++l) {
    int th, tw;
    conv_pick_tile(lv.h[l], lv.w[l], &th, &tw);
    tiles += (long long)lv.n_images * ceil_div(lv.w[l], tw) * ceil_div(lv.h[l], th);
  }
  return tiles * 4 * 64 * 4 + 256;
}

extern "C" int scan_conv3x3_rows_gn(const scan_levels_t* levels, const float* x_rows, const float* x_lo, int32_t cin, const float* packed,
                                    const float* packed_lo, const float* gn_bias, float eps, float* y_rows, float* stats, int32_t cta_group,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
  if (!stats || !workspace || workspace_bytes < scan_conv3x3_gn_workspace_bytes(levels)) return SCAN_EINVAL;
  float* partial = (float*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  ConvGnGeo geo = {};
  int rc = conv_rows_impl(levels, x_rows, x_lo, cin, nullptr, nullptr, 0, packed, packed_lo, 256, nullptr, nullptr, nullptr, 0, y_rows, 256,
                          cta_group, 9, stream, partial, gn_bias, &geo);
  if (rc) return rc;
  const int n_stats = geo.n_levels * geo.n_images * 32;
  conv_gn_finalize_kernel<<<(n_stats + 7) / 8, 256, 0, (cudaStream_t)stream>>>(geo, partial, eps, stats);
  SCAN_LAUNCH_CHECK("conv_gn_finalize_kernel");
  return SCAN_OK;
}

// ---------------------------------------------------------------------------- weight gradient: host side
namespace scan {

static void wg_pick_chunk(int h, int w, int* ch_out, int* cw_out) {
  long long best = -1;
  int bh = 1, bw = 32;
  for (int cw = 32; cw >= 1; cw >>= 1) {        // ties: the wider rectangle (longer contiguous runs)
    const int ch = 32 / cw;
    const long long n = ceil_div(h, ch) * ceil_div(w, cw);
    if (best < 0 || n < best) {
      best = n;
      bh = ch;
      bw = cw;
    }
  }
  *ch_out = bh;
  *cw_out = bw;
}

static int wg_gcd(int a, int b) { return b == 0 ? a : wg_gcd(b, a % b); }

// the static schedule shared by the workspace query and the launch
static int wg_plan(const Levels& lv, int cin, int cout, int precise, WgArgs* g, int* pairs_out, int n_taps = 9, int x_width = 256) {
  if (cin < x_width || (cin % x_width) || cout < 256 || (cout % 256)) return SCAN_EINVAL;
  g->n_levels = lv.n_levels;
  g->n_images = lv.n_images;
  int chunks = 0;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    WgLevel& L = g->lv[l];
    if (l < lv.n_levels) {
      wg_pick_chunk(lv.h[l], lv.w[l], &L.ch, &L.cw);
      L.chunks_x = (int)ceil_div(lv.w[l], L.cw);
      L.chunks_y = (int)ceil_div(lv.h[l], L.ch);
      L.chunk_off = chunks;
      chunks += lv.n_images * L.chunks_x * L.chunks_y;
    } else {
      L = g->lv[lv.n_levels - 1];
      L.chunk_off = 0x7fffffff;
    }
  }
  g->n_chunks = chunks;
  g->n_terms = precise ? 3 : 1;
  g->n_taps = n_taps;
  g->m_blocks = cout / 256;
  g->n_blocks = cin / x_width;
  const int kinds = n_taps * g->m_blocks * g->n_blocks;
  int pairs = sm_count() / 2;
  if (pairs < 1) pairs = 1;
  // segments: accumulation chains of at most 512 k-chunks (8 x 3 terms = 96 MMAs in the fp32-accurate mode: the tensor core
  // truncates when it adds into the accumulator), and a segment count that fills the last wave of pairs
  const int max_chain = precise ? 8 : 512;
  const int n_min = (int)ceil_div(chunks, max_chain);
  int best_n = n_min;
  double best_eff = -1.0;
  for (int n = n_min; n <= n_min + 2 * pairs; ++n) {
    const long long items = (long long)kinds * n;
    const double eff = (double)items / (double)(ceil_div(items, pairs) * pairs);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best_n = n;
    }
    if (eff >= 0.97) break;
  }
  g->seg_len = (int)ceil_div(chunks, best_n);
  g->n_seg = (int)ceil_div(chunks, g->seg_len);
  g->n_items = g->n_seg * kinds;
  if (pairs > g->n_items) pairs = g->n_items;
  const int step = pairs % kinds;
  g->period = step == 0 ? 1 : kinds / wg_gcd(step, kinds);
  const int per_pair = (int)ceil_div(g->n_items, pairs);
  g->slots = g->period < per_pair ? g->period : per_pair;
  *pairs_out = pairs;
  return SCAN_OK;
}

static int wg_make_map(CUtensorMap* m, const float* base, int c, int w, int h, int n, int cw, int ch, int box_blocks = 4) {
  EncodeTiledFn enc;
  int rc = get_tensormap_encoder(&enc);
  if (rc) return rc;
  if ((uintptr_t)base & 15) return SCAN_EINVAL;
  // the channel axis split into (32, C / 32) with the block index as 4th dimension: one box = 4 blocks x (ch x cw pixels) x 32
  // channels lands as [block][pixel][32], the MN-major operand layout
  cuuint64_t dims[5] = {32, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(c / 32), (cuuint64_t)n};
  cuuint64_t strides[4] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, 128, (cuuint64_t)h * w * c * 4};
  cuuint32_t box[5] = {32, (cuuint32_t)cw, (cuuint32_t)ch, (cuuint32_t)box_blocks, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed (conv wgrad)");
    return SCAN_ECUDA;
  }
  return SCAN_OK;
}

}  // namespace scan

extern "C" int64_t scan_conv3x3_wgrad_workspace_bytes(const scan_levels_t* levels, int32_t cin, int32_t cout, int32_t precise) {
  Levels lv;
  if (make_levels(levels, &lv)) return -1;
  WgArgs g = {};
  int pairs = 0;
  if (wg_plan(lv, cin, cout, precise, &g, &pairs)) return -1;
  return (int64_t)pairs * g.slots * WG_TILE * 4;
}

// d_w[co][ci][ky][kx] (element strides s_*) = sum_p dy_rows[p, co] * x_rows[p + off(ky, kx), ci]; cin, cout multiples of 256.
// x_lo / dy_lo both non-NULL: 3xTF32.
extern "C" int scan_conv3x3_wgrad(const scan_levels_t* levels, const float* x_rows, const float* x_lo, int32_t cin, const float* dy_rows,
                                  const float* dy_lo, int32_t cout, float* d_w, int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
  Levels lv;
  int rc = make_levels(levels, &lv);
  if (rc) return rc;
  if (!x_rows || !dy_rows || !d_w || !workspace || (x_lo == nullptr) != (dy_lo == nullptr)) return SCAN_EINVAL;
  WgArgs g = {};
  int pairs = 0;
  rc = wg_plan(lv, cin, cout, x_lo != nullptr, &g, &pairs);
  if (rc) return rc;
  if (workspace_bytes < (int64_t)pairs * g.slots * WG_TILE * 4 || ((uintptr_t)workspace & 15)) return SCAN_EINVAL;
  g.partial = (float*)workspace;
  WgMaps maps;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    const int ll = l < lv.n_levels ? l : 0;
    const WgLevel& L = g.lv[ll];
    rc = wg_make_map(&maps.dy[0][l], dy_rows + lv.row_off[ll] * cout, cout, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch);
    if (rc) return rc;
    rc = wg_make_map(&maps.x[0][l], x_rows + lv.row_off[ll] * cin, cin, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch);
    if (rc) return rc;
    if (x_lo) {
      rc = wg_make_map(&maps.dy[1][l], dy_lo + lv.row_off[ll] * cout, cout, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch);
      if (rc) return rc;
      rc = wg_make_map(&maps.x[1][l], x_lo + lv.row_off[ll] * cin, cin, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch);
      if (rc) return rc;
    } else {
      maps.dy[1][l] = maps.dy[0][l];
      maps.x[1][l] = maps.x[0][l];
    }
  }
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr))
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = WG_SMEM;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SCAN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<4>, maps, g));
  conv_wgrad_reduce_kernel<<<dim3(9, 64, (unsigned)(g.m_blocks * g.n_blocks)), 256, 9 * pairs, (cudaStream_t)stream>>>(
      g.partial, pairs, g.slots, g.period, g.n_items, g.m_blocks, g.n_blocks, s_co, s_ci, s_ky, s_kx, d_w);
  SCAN_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
  return SCAN_OK;
}

extern "C" int64_t scan_thin_wgrad_workspace_bytes(const scan_levels_t* levels, int32_t precise) {
  Levels lv;
  if (make_levels(levels, &lv)) return -1;
  WgArgs g = {};
  int pairs = 0;
  if (wg_plan(lv, 128, 256, precise, &g, &pairs, 1, 128)) return -1;
  const int64_t R = lv.row_off[lv.n_levels];
  return (int64_t)pairs * g.slots * 256 * 128 * 4 + R * 128 * 4 * (precise ? 2 : 1) + 1024;
}

// d_w[co][k][ky][kx] (element strides s_*; the map columns of head_out's weight gradient) from dy_rows [R, 256] and the class maps
// maps32 [R, 32] (columns [0, k), 9 k <= 128).  dy_lo non-NULL: 3xTF32.
extern "C" int scan_thin_wgrad(const scan_levels_t* levels, const float* maps32, const float* dy_rows, const float* dy_lo, int32_t k,
                               float* d_w, int64_t s_co, int64_t s_k, int64_t s_ky, int64_t s_kx, void* workspace, int64_t workspace_bytes,
                               void* stream) {
  Levels lv;
  int rc = make_levels(levels, &lv);
  if (rc) return rc;
  if (!maps32 || !dy_rows || !d_w || !workspace || k < 1 || 9 * k > 128 || ((uintptr_t)workspace & 15)) return SCAN_EINVAL;
  const int precise = dy_lo != nullptr;
  if (workspace_bytes < scan_thin_wgrad_workspace_bytes(levels, precise)) return SCAN_EINVAL;
  WgArgs g = {};
  int pairs = 0;
  rc = wg_plan(lv, 128, 256, precise, &g, &pairs, 1, 128);
  if (rc) return rc;
  const long long R = lv.row_off[lv.n_levels];
  cudaStream_t st = (cudaStream_t)stream;
  g.partial = (float*)workspace;
  float* s = g.partial + (long long)pairs * g.slots * 256 * 128;
  float* s_lo = precise ? s + R * 128 : nullptr;
  thin_spread_kernel<<<(unsigned)ceil_div(R * 32, 256), 256, 0, st>>>(lv, maps32, k, s);
  SCAN_LAUNCH_CHECK("thin_spread_kernel");
  if (precise) {
    long long blocks = ceil_div(R * 32, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    tf32_residual_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(s), R * 32, reinterpret_cast<float4*>(s_lo));
    SCAN_LAUNCH_CHECK("tf32_residual_kernel");
  }
  WgMaps maps;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    const int ll = l < lv.n_levels ? l : 0;
    const WgLevel& L = g.lv[ll];
    rc = wg_make_map(&maps.dy[0][l], dy_rows + lv.row_off[ll] * 256, 256, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch, 4);
    if (rc) return rc;
    rc = wg_make_map(&maps.x[0][l], s + lv.row_off[ll] * 128, 128, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch, 2);
    if (rc) return rc;
    maps.dy[1][l] = maps.dy[0][l];
    maps.x[1][l] = maps.x[0][l];
    if (precise) {
      rc = wg_make_map(&maps.dy[1][l], dy_lo + lv.row_off[ll] * 256, 256, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch, 4);
      if (rc) return rc;
      rc = wg_make_map(&maps.x[1][l], s_lo + lv.row_off[ll] * 128, 128, lv.w[ll], lv.h[ll], lv.n_images, L.cw, L.ch, 2);
      if (rc) return rc;
    }
  }
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr))
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(conv_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(pairs * 2));
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = WG_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  SCAN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<2>, maps, g));
  thin_wgrad_reduce_kernel<<<(256 * 128) / 256, 256, 0, st>>>(g.partial, pairs, g.slots, k, s_co, s_k, s_ky, s_kx, d_w);
  SCAN_LAUNCH_CHECK("thin_wgrad_reduce_kernel");
  return SCAN_OK;
}
