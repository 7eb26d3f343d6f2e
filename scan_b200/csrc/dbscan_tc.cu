// K2 (distance phase): pairwise squared distances of the DBSCAN point set on the tensor cores.
//
// Gram tiles G = P_i . P_j^T (128 x 128, reduction 256) with tcgen05.mma kind::tf32: both operands are K-major row
// blocks of the same [n, 256] point matrix, streamed by TMA ([128 x 32] fp32 boxes, 128-byte swizzle) through a
// 6-stage mbarrier ring; fp32 accumulators live in TMEM (2 x 128 columns, double-buffered so the epilogue of one
// tile overlaps the MMAs of the next).  Only tiles with bi <= bj are computed; the epilogue writes the bit block and
// its transpose (warp ballots), so the full symmetric adjacency matrix is produced from half of the flops.
//
// Exactness: operands are pre-split into tf32 hi + lo parts (db_split_kernel) and the Gram entry is accumulated as
// hi.hi + hi.lo + lo.hi (3xTF32, fp32 accumulate): ~1e-6 relative.  It is only trusted outside the band
// |d2 - eps^2| > 2.5e-5 (|p_i|^2 + |p_j|^2) (worst-case fp32 accumulation bound over 256 terms); inside the band the
// pair is re-evaluated exactly as sklearn does it (float64 accumulation of the fp32 inputs, scan::db_exact_within).
// (A single-tf32 Gram needs a 2.2e-3 band: measured 13 ms of divergent fp64 rechecks at n = 36 k -- round-1 run 5.)
// Labels therefore stay bit-exact with sklearn (tests/test_gpu_kernels.py) while the bulk of the n^2 x 256
// arithmetic runs at tensor-core speed.
#include "dbscan_common.cuh"
#include "tc_common.cuh"

namespace scan {

constexpr int GT = 128;                       // tile edge (points)
constexpr int GK = 32;                        // channels per stage
constexpr int G_STAGES = 3;
constexpr int G_BOX_BYTES = GT * GK * 4;        // 16 KB
constexpr int G_STAGE_BYTES = 4 * G_BOX_BYTES;  // A_hi, A_lo, B_hi, B_lo boxes: 64 KB
constexpr int G_SMEM = 1024 + G_STAGES * G_STAGE_BYTES + 1024;
constexpr int G_THREADS = 256;
constexpr uint32_t G_IDESC = umma_idesc_tf32(GT, GT);

__device__ __forceinline__ void tile_of(long long t, int nt, int& bi, int& bj) {
  // t enumerates (bi, bj), bi <= bj, row by row: row bi starts at bi*nt - bi*(bi-1)/2
  double x = (2.0 * nt + 1.0 - sqrt((2.0 * nt + 1.0) * (2.0 * nt + 1.0) - 8.0 * (double)t)) * 0.5;
  int b = (int)x;
  if (b < 0) b = 0;
  if (b > nt - 1) b = nt - 1;
  while ((long long)b * nt - (long long)b * (b - 1) / 2 > t) --b;
  while ((long long)(b + 1) * nt - (long long)(b + 1) * b / 2 <= t) ++b;
  bi = b;
  bj = b + (int)(t - ((long long)b * nt - (long long)b * (b - 1) / 2));
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
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

__global__ void __launch_bounds__(G_THREADS, 1)
    db_adj_tc_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ points, const float* __restrict__ sq,
                     const int* info, int n_fixed, int dim, float eps2f, double eps2, long long wpr, uint32_t* __restrict__ adj,
                     int* info_w) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = smem;
  uint64_t* bars = (uint64_t*)(stages + G_STAGES * G_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + G_STAGES;
  uint64_t* acc_full = bars + 2 * G_STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * G_STAGES + 2;  // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * G_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int nt = (n + GT - 1) / GT;
  const long long total = (long long)nt * (nt + 1) / 2;
  const int kblocks = dim / GK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < G_STAGES; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(acc_full + i), 1);
      mbar_init(smem_u32(acc_empty + i), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * GT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        int bi, bj;
        tile_of(t, nt, bi, bj);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          mbar_expect_tx(smem_u32(full_bar + stage), G_STAGE_BYTES);
          uint8_t* st = stages + stage * G_STAGE_BYTES;
          const uint32_t fb = smem_u32(full_bar + stage);
          tma_load_2d(smem_u32(st), &tmap, fb, kb * GK, bi * GT);                          // A hi
          tma_load_2d(smem_u32(st + G_BOX_BYTES), &tmap, fb, dim + kb * GK, bi * GT);      // A lo
          tma_load_2d(smem_u32(st + 2 * G_BOX_BYTES), &tmap, fb, kb * GK, bj * GT);        // B hi
          tma_load_2d(smem_u32(st + 3 * G_BOX_BYTES), &tmap, fb, dim + kb * GK, bj * GT);  // B lo
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        mbar_wait(smem_u32(acc_empty + acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * GT;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tcgen05_fence_after();
          const uint32_t ah = smem_u32(stages + stage * G_STAGE_BYTES);
          const uint32_t al = ah + G_BOX_BYTES, bh = ah + 2 * G_BOX_BYTES, bl = ah + 3 * G_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < GK / 8; ++k) {
            const uint64_t dah = umma_desc_sw128(ah + k * 32), dbh = umma_desc_sw128(bh + k * 32);
            umma_tf32(d, umma_desc_sw128(al + k * 32), dbh, G_IDESC, (kb | k) != 0);
            umma_tf32(d, dah, umma_desc_sw128(bl + k * 32), G_IDESC, 1);
            umma_tf32(d, dah, dbh, G_IDESC, 1);
          }
          umma_commit(smem_u32(empty_bar + stage));
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(smem_u32(acc_full + acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    int n_re = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x) {
      int bi, bj;
      tile_of(t, nt, bi, bj);
      const int i0 = bi * GT, j0 = bj * GT;
      const int i = i0 + q * 32 + lane;
      const float si = i < n ? __ldg(sq + i) : 0.f;
      mbar_wait(smem_u32(acc_full + acc), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < GT; c0 += 32) {
        float g[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GT + c0, g);
        uint32_t word = 0;
        // |p_j|^2 of the 32 columns of this block: ONE coalesced load per lane, broadcast by shuffle below (a dependent
        // global load per column cost 38 k cycles per tile: ncu source view, profiles/r01_dbscan_adj.txt)
        const float sj_lane = (j0 + c0 + lane < n) ? __ldg(sq + j0 + c0 + lane) : 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int j = j0 + c0 + c;
          bool within = false;
          const float sj = __shfl_sync(0xffffffffu, sj_lane, c);
          if (i < n && j < n) {
            if (i == j) {
              within = true;
            } else {
              const float d2 = si + sj - 2.f * g[c];
              const float tol = 2.5e-5f * (si + sj) + 1e-7f * eps2f;
              if (fabsf(d2 - eps2f) <= tol) {
                within = db_exact_within(points + (long long)i * dim, points + (long long)j * dim, dim, eps2);
                ++n_re;
              } else {
                within = d2 < eps2f;
              }
            }
          }
          word |= (within ? 1u : 0u) << c;
          if (bi != bj) {  // transposed block: bit (row j, column i); one ballot = the word of row j for this warp's 32 rows
            const uint32_t tw = __ballot_sync(0xffffffffu, within);
            if (lane == c && j < n) adj[(long long)j * wpr + (i0 >> 5) + q] = tw;
          }
        }
        if (i < n) adj[(long long)i * wpr + ((j0 + c0) >> 5)] = word;
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(acc_empty + acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    n_re = (int)warp_sum((float)n_re);
    if (lane == 0 && n_re) atomicAdd(info_w + 5, n_re);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * GT));
  }
}

static int g_adj_attr = 0;

int launch_db_adj_tc(const float* points, const float* points_hl, const float* sq, const int* info, int n_fixed, int cap, int dim,
                     float eps2f, double eps2, long long wpr, uint32_t* adj, int* info_w, cudaStream_t st) {
  if (dim % GK || ((uintptr_t)points_hl & 15) || wpr % 4) return SCAN_EINVAL;
  CUtensorMap map;
  int rc = make_rowmajor_map(&map, points_hl, (uint64_t)cap, (uint64_t)(2 * dim), GT);
  if (rc) return rc;
  if (!g_adj_attr) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(db_adj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    g_adj_attr = 1;
  }
  db_adj_tc_kernel<<<sm_count(), G_THREADS, G_SMEM, st>>>(map, points, sq, info, n_fixed, dim, eps2f, eps2, wpr, adj, info_w);
  SCAN_LAUNCH_CHECK("db_adj_tc_kernel");
  return SCAN_OK;
}

}  // namespace scan
