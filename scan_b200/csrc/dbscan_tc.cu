// K2 (distance phase): pairwise squared distances of the DBSCAN point set on the tensor cores.
//
// Gram tiles G = P_i . P_j^T (128 x 128, reduction 256) with tcgen05.mma kind::tf32 straight from the fp32 point
// matrix (K-major rows, TMA [128 x 32] boxes, 128-byte swizzle).  The kernel is L2-bandwidth bound, not MMA bound
// (a 128x128x256 tile needs 256 KB of operands for 8.4 MFLOP), so the schedule maximises operand reuse:
//   * a CTA works on a UNIT = one row block bi and a run of up to DB_CHUNK column blocks bj >= bi; the A tile (128 points
//     x 256 dims = 128 KB) is loaded ONCE per unit and stays resident in shared memory, only B tiles stream through a
//     4-stage ring -> 128 KB of L2 traffic per tile instead of 256 KB (and instead of 512 KB for a 3xTF32 variant:
//     measured 5.1 ms at n = 36 k, round-1 run 7);
//   * only tiles with bj >= bi are computed; db_mirror_kernel transposes the bit blocks afterwards;
//   * fp32 accumulators in TMEM, 2 x 128 columns, so the epilogue of one tile overlaps the MMAs of the next.
//
// Exactness: the tensor core truncates the fp32 operands to tf32 (<= 2^-10 relative per operand), so a Gram entry is
// only trusted outside the band |d2 - eps^2| > 2.2e-3 (|p_i|^2 + |p_j|^2); pairs inside the band get a provisional bit and
// are appended to a work list; db_recheck_kernel re-evaluates them exactly as sklearn does (float64 accumulation of the
// fp32 inputs, one warp per pair: 8 dims per lane, shuffle reduction) and flips the bits that were wrong.  (History: a
// per-thread recheck inside the epilogue serialised 256-step fp64 loops behind one lane, 13 ms; a warp-cooperative recheck
// inside the epilogue still made every tile wait for its slowest warp's chain of dependent loads, 1.2 ms; deferred: 0.73 +
// 0.18 ms.)  Labels stay bit-exact with sklearn (tests/test_gpu_kernels.py::test_dbscan_*).
#include "dbscan_common.cuh"
#include "tc_common.cuh"

namespace scan {

constexpr int GT = 128;   // tile edge (points)
constexpr int GK = 32;    // channels per TMA box
constexpr int G_DIM = 256;
constexpr int G_KB = G_DIM / GK;                 // 8 k-blocks
constexpr int G_BOX_BYTES = GT * GK * 4;         // 16 KB
constexpr int G_A_BYTES = G_KB * G_BOX_BYTES;    // 128 KB resident A tile
constexpr int G_STAGES = 4;                      // B ring (6 stages measured: no change)
constexpr int G_SMEM = 1024 + G_A_BYTES + G_STAGES * G_BOX_BYTES + 1024;
constexpr int G_EPI_WARPS = 16;                 // epilogue warp e: TMEM lane quarter e % 4, 32-column block e / 4
constexpr int G_THREADS = 128 + 32 * G_EPI_WARPS;
constexpr int DB_CHUNK = 48;                     // column blocks per unit
constexpr uint32_t G_IDESC = umma_idesc_tf32(GT, GT);

// packed fp32x2 arithmetic (sm_100: FADD2 / FFMA2)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
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

// Work enumeration shared by the three roles: units (bi, [j_begin, j_end)) in row order, unit u belongs to CTA u % grid.
struct UnitIter {
  int nt, bi, jb, u;
  __device__ UnitIter(int nt_) : nt(nt_), bi(0), jb(0), u(0) {}
  // advances to the next unit owned by this CTA; returns false when exhausted
  __device__ bool next(int& row, int& j0, int& j1) {
    while (bi < nt) {
      const int mine = (u % (int)gridDim.x) == (int)blockIdx.x;
      row = bi;
      j0 = bi + jb;
      j1 = min(nt, j0 + DB_CHUNK);
      ++u;
      jb += DB_CHUNK;
      if (bi + jb >= nt) { ++bi; jb = 0; }
      if (mine) return true;
    }
    return false;
  }
};

// exact sklearn test by the whole warp (256-d rows): each lane takes 8 dims with four independent 128-bit loads (one
// memory latency instead of eight), fp64 accumulation, shuffle reduction
__device__ __noinline__ bool warp_exact_within(const float* __restrict__ a, const float* __restrict__ b, double eps2, int lane) {
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  const float4 x0 = __ldg(a4 + lane), x1 = __ldg(a4 + 32 + lane);
  const float4 y0 = __ldg(b4 + lane), y1 = __ldg(b4 + 32 + lane);
  const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  const float ys[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
  double sa = 0.0, sb = 0.0, ab = 0.0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const double x = (double)xs[e], y = (double)ys[e];
    sa = fma(x, x, sa);
    sb = fma(y, y, sb);
    ab = fma(x, y, ab);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    ab += __shfl_xor_sync(0xffffffffu, ab, o);
  }
  double d2 = sa + sb - 2.0 * ab;
  if (d2 < 0.0) d2 = 0.0;
  return d2 <= eps2;
}

__global__ void __launch_bounds__(G_THREADS, 1)
    db_adj_tc_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ points, const float* __restrict__ sq,
                     const int* info, int n_fixed, float eps2f, double eps2, long long wpr, uint32_t* __restrict__ adj, int* info_w,
                     unsigned long long* __restrict__ re_list) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float4 colv[G_EPI_WARPS][16];   // per epilogue warp: 16 column pairs (sq_c, sq_c+1, 2.2e-3 sq_c, 2.2e-3 sq_c+1)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tile = smem;
  uint8_t* stages = smem + G_A_BYTES;
  uint64_t* bars = (uint64_t*)(stages + G_STAGES * G_BOX_BYTES);
  uint64_t* b_full = bars;                         // [G_STAGES]
  uint64_t* b_empty = bars + G_STAGES;             // [G_STAGES]
  uint64_t* acc_full = bars + 2 * G_STAGES;        // [2]
  uint64_t* acc_empty = bars + 2 * G_STAGES + 2;   // [2]
  uint64_t* a_full = bars + 2 * G_STAGES + 4;
  uint64_t* a_free = bars + 2 * G_STAGES + 5;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * G_STAGES + 6);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int nt = (n + GT - 1) / GT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < G_STAGES; ++i) {
      mbar_init(smem_u32(b_full + i), 1);
      mbar_init(smem_u32(b_empty + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(acc_full + i), 1);
      mbar_init(smem_u32(acc_empty + i), 32 * G_EPI_WARPS);
    }
    mbar_init(smem_u32(a_full), 1);
    mbar_init(smem_u32(a_free), 1);
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
    // ===== TMA producer =====
    if (lane == 0) {
      UnitIter it(nt);
      int row, j0, j1, stage = 0, units = 0;
      uint32_t phase = 0;
      while (it.next(row, j0, j1)) {
        if (units > 0) mbar_wait(smem_u32(a_free), (uint32_t)((units - 1) & 1));  // MMAs of the previous unit retired
        mbar_expect_tx(smem_u32(a_full), G_A_BYTES);
        for (int kb = 0; kb < G_KB; ++kb) tma_load_2d(smem_u32(a_tile + kb * G_BOX_BYTES), &tmap, smem_u32(a_full), kb * GK, row * GT);
        for (int bj = j0; bj < j1; ++bj) {
          for (int kb = 0; kb < G_KB; ++kb) {
            mbar_wait(smem_u32(b_empty + stage), phase ^ 1);
            mbar_expect_tx(smem_u32(b_full + stage), G_BOX_BYTES);
            tma_load_2d(smem_u32(stages + stage * G_BOX_BYTES), &tmap, smem_u32(b_full + stage), kb * GK, bj * GT);
            if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
          }
        }
        ++units;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: all 32 lanes run the loop, every tcgen05 instruction goes out from one elected lane =====
    {
      UnitIter it(nt);
      int row, j0, j1, stage = 0, acc = 0, units = 0;
      uint32_t phase = 0, acc_phase = 0;
      while (it.next(row, j0, j1)) {
        mbar_wait(smem_u32(a_full), (uint32_t)(units & 1));
        for (int bj = j0; bj < j1; ++bj) {
          mbar_wait(smem_u32(acc_empty + acc), acc_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t d = tmem_base + acc * GT;
          for (int kb = 0; kb < G_KB; ++kb) {
            mbar_wait(smem_u32(b_full + stage), phase);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(a_tile + kb * G_BOX_BYTES);
            const uint32_t b_addr = smem_u32(stages + stage * G_BOX_BYTES);
#pragma unroll
            for (int k = 0; k < GK / 8; ++k)
              if (elect_one_sync())
                umma_tf32(d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), G_IDESC, (kb | k) != 0);
            if (elect_one_sync()) umma_commit(smem_u32(b_empty + stage));
            if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
          }
          if (elect_one_sync()) umma_commit(smem_u32(acc_full + acc));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit(smem_u32(a_free));  // arrives when every MMA reading this A tile has retired
        ++units;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===== epilogue: 16 warps; warp e owns TMEM lanes [32q, 32q+32) (q = e % 4: rows of the tile) and the 32-column
    // block c0 = 32 * (e / 4).  One warp per scheduler is latency-bound (IPC ~0.1, ncu round-1 run 9): four per
    // scheduler hide the ALU / shuffle latencies.  Only the block itself is written; db_mirror_kernel adds the
    // transposed blocks afterwards. =====
    const int e = warp - 4;
    const int q = e & 3;
    const int c0 = (e >> 2) * 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    UnitIter it(nt);
    int row, jb0, jb1;
    while (it.next(row, jb0, jb1)) {
      const int i0 = row * GT;
      const int i = i0 + q * 32 + lane;
      const float si = i < n ? __ldg(sq + i) : 0.f;
      for (int bj = jb0; bj < jb1; ++bj) {
        const int j0 = bj * GT;
        const float sj_lane = (j0 + c0 + lane < n) ? __ldg(sq + j0 + c0 + lane) : 0.f;
        mbar_wait(smem_u32(acc_full + acc), acc_phase);
        tcgen05_fence_after();
        float g[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GT + c0, g);
        tcgen05_fence_before();
        mbar_arrive(smem_u32(acc_empty + acc));   // the accumulator block is in registers
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        // pass 1: d = d2 - eps^2 = ((si - eps^2) + sj) - 2 g ; within <=> d < 0 ; unsure <=> |d| < tol = (2.2e-3 si + k) + 2.2e-3 sj.
        // The epilogue is the issue-bound part of this kernel (ncu: the MMA warp never waits for operands, the epilogue warps
        // ran ~21 instructions per entry): the per-column terms come from shared memory two columns per 128-bit broadcast
        // load, the arithmetic is packed fp32x2 (FADD2 / FFMA2), and the two bit masks are assembled from SIGN bits with a
        // shift-in (entries exactly on a boundary are inside the band and re-evaluated exactly anyway).
        __syncwarp();
        {
          float* slot = reinterpret_cast<float*>(&colv[e][lane >> 1]);
          slot[lane & 1] = sj_lane;
          slot[2 + (lane & 1)] = 2.2e-3f * sj_lane;
        }
        __syncwarp();
        uint32_t word = 0, unsure = 0;
        const unsigned long long a2 = pack2(si - eps2f, si - eps2f);
        const float bq = fmaf(2.2e-3f, si, 1e-6f * eps2f);
        const unsigned long long b2 = pack2(bq, bq), m2 = pack2(-2.f, -2.f);
#pragma unroll
        for (int cp = 0; cp < 16; ++cp) {
          const float4 v = colv[e][cp];
          const unsigned long long t2 = add2(a2, pack2(v.x, v.y));
          const unsigned long long d2 = fma2(m2, pack2(g[2 * cp], g[2 * cp + 1]), t2);
          const unsigned long long tol2 = add2(b2, pack2(v.z, v.w));
          float d0, d1, t0, t1;
          unpack2(d2, d0, d1);
          unpack2(tol2, t0, t1);
          const float e0 = fabsf(d0) - t0, e1 = fabsf(d1) - t1;
          word = (word >> 1) | (__float_as_uint(d0) & 0x80000000u);
          unsure = (unsure >> 1) | (__float_as_uint(e0) & 0x80000000u);
          word = (word >> 1) | (__float_as_uint(d1) & 0x80000000u);
          unsure = (unsure >> 1) | (__float_as_uint(e1) & 0x80000000u);
        }
        // validity: columns beyond n, rows beyond n; the diagonal is always "within" and never rechecked
        const int ncol = n - (j0 + c0);
        const uint32_t colmask = ncol >= 32 ? 0xffffffffu : (ncol <= 0 ? 0u : ((1u << ncol) - 1u));
        const uint32_t rowmask = (i < n) ? colmask : 0u;
        const int dcol = i - (j0 + c0);                      // column of the diagonal element inside this block
        const uint32_t diag = (dcol >= 0 && dcol < 32) ? (1u << dcol) : 0u;
        word = (word | diag) & rowmask;
        unsure = unsure & rowmask & ~diag;
        // pass 2: the in-band pairs need sklearn's float64 test.  Doing it here puts a chain of dependent global loads into
        // the tile loop, and every tile waits for its slowest warp (measured: 16 pairs per tile on average tripled the tile
        // time).  The pairs are appended to a work list instead (one atomic per warp and tile) and db_recheck_kernel flips
        // the provisional bits afterwards with full parallelism; only when the list is full are they evaluated inline.
        {
          const int cnt = __popc(unsure);
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          if (total) {
            int base = 0;
            if (lane == 0) base = atomicAdd(info_w + 5, total);
            base = __shfl_sync(0xffffffffu, base, 0);
            int slot = base + incl - cnt;
            uint32_t bits = unsure, keep = 0;
            while (bits) {
              const int c = __ffs(bits) - 1;
              bits &= bits - 1;
              if (slot < DB_RE_CAP) re_list[slot] = ((unsigned long long)(uint32_t)i << 32) | (uint32_t)(j0 + c0 + c);
              else keep |= 1u << c;
              ++slot;
            }
            unsure = keep;
          }
        }
        uint32_t lanes = __ballot_sync(0xffffffffu, unsure != 0u);
        while (lanes) {   // overflow of the work list only
          const int src = __ffs(lanes) - 1;
          lanes &= lanes - 1;
          uint32_t bits = __shfl_sync(0xffffffffu, unsure, src);
          const float* pi = points + (long long)(i0 + q * 32 + src) * G_DIM;
          while (bits) {
            const int c = __ffs(bits) - 1;
            bits &= bits - 1;
            const bool r = warp_exact_within(pi, points + (long long)(j0 + c0 + c) * G_DIM, eps2, lane);
            if (lane == src) word = (word & ~(1u << c)) | ((r ? 1u : 0u) << c);
          }
        }
        if (i < n) adj[(long long)i * wpr + ((j0 + c0) >> 5)] = word;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * GT));
  }
}

// exact float64 re-evaluation of the deferred in-band pairs: one warp per pair, flips the provisional bit when it is wrong
__global__ void __launch_bounds__(256) db_recheck_kernel(const float* __restrict__ points, const int* __restrict__ info_w, double eps2,
                                                         long long wpr, const unsigned long long* __restrict__ re_list,
                                                         uint32_t* __restrict__ adj) {
  const int n_re = min(info_w[5], DB_RE_CAP);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < n_re; p += gridDim.x * wpb) {
    const unsigned long long e = re_list[p];
    const int i = (int)(e >> 32), j = (int)(e & 0xffffffffu);
    const bool r = warp_exact_within(points + (long long)i * G_DIM, points + (long long)j * G_DIM, eps2, lane);
    if (lane == 0) {
      uint32_t* w = adj + (long long)i * wpr + (j >> 5);
      const uint32_t bit = 1u << (j & 31);
      if ((((*w) & bit) != 0u) != r) atomicXor(w, bit);
    }
  }
}

// 32 x 32 bit-matrix transpose across a warp (lane r holds row r, bit c = column c): five butterfly steps that swap the
// off-diagonal sub-blocks of size j (recursive block transpose; ~45 instructions instead of 32 ballots + selects)
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    const bool hi = (lane & j) != 0;
    const uint32_t low = hi ? y : x, high = hi ? x : y;
    const uint32_t t = ((low >> j) ^ high) & m;
    x = hi ? (x ^ t) : (x ^ (t << j));
  }
  return x;
}

// adjacency is symmetric: the tile kernel wrote the 32x32 bit blocks (I, J) of tiles with tile(I) <= tile(J); this adds
// the transposed blocks (J, I) for tile(I) < tile(J).  One warp per block: lane r loads row r's word, 32 ballots
// transpose it, lane c stores row c of the mirrored block.  Memory-bound pass over n^2/16 bytes.
__global__ void __launch_bounds__(128) db_mirror_kernel(const int* info, int n_fixed, long long wpr, uint32_t* __restrict__ adj) {
  // one CTA per 128 x 128 bit tile strictly above the tile diagonal: 128-bit row accesses on both sides (the per-block
  // version read and wrote one 4-byte word per 32-byte sector: 269 us at n = 38 k), 32 x 32 ballot transposes in between
  __shared__ uint32_t tin[128][5], tout[128][5];      // 5-word pitch: conflict-free column access
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int nt = (n + 127) >> 7;
  const long long total = (long long)nt * nt;
  const int r = threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int TI = (int)(t / nt), TJ = (int)(t - (long long)TI * nt);
    if (TI >= TJ) continue;
    const int gi = TI * 128 + r;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gi < n) v = *reinterpret_cast<const uint4*>(adj + (long long)gi * wpr + TJ * 4);
    tin[r][0] = v.x; tin[r][1] = v.y; tin[r][2] = v.z; tin[r][3] = v.w;
    __syncthreads();
#pragma unroll
    for (int jb = 0; jb < 4; ++jb) {     // block (rows w*32.., word jb) -> block (rows jb*32.., word w)
      tout[jb * 32 + lane][w] = warp_transpose32(tin[w * 32 + lane][jb], lane);
    }
    __syncthreads();
    const int gj = TJ * 128 + r;
    if (gj < n) *reinterpret_cast<uint4*>(adj + (long long)gj * wpr + TI * 4) = make_uint4(tout[r][0], tout[r][1], tout[r][2], tout[r][3]);
    __syncthreads();
  }
}

static unsigned long long g_adj_attr = 0;

int launch_db_adj_tc(const float* points, const float* sq, const int* info, int n_fixed, int cap, int dim, float eps2f, double eps2,
                     long long wpr, uint32_t* adj, int* info_w, unsigned long long* re_list, cudaStream_t st) {
  if (dim != G_DIM) return SCAN_ENOTSUP;  // callers fall back to the FFMA tile kernel for other widths
  if (((uintptr_t)points & 15) || wpr % 4) return SCAN_EINVAL;
  CUtensorMap map;
  int rc = make_rowmajor_map(&map, points, (uint64_t)cap, (uint64_t)dim, GT);
  if (rc) return rc;
  if (first_use_on_device(&g_adj_attr)) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(db_adj_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
  }
  db_adj_tc_kernel<<<sm_count(), G_THREADS, G_SMEM, st>>>(map, points, sq, info, n_fixed, eps2f, eps2, wpr, adj, info_w, re_list);
  SCAN_LAUNCH_CHECK("db_adj_tc_kernel");
  db_recheck_kernel<<<8 * sm_count(), 256, 0, st>>>(points, info_w, eps2, wpr, re_list, adj);
  SCAN_LAUNCH_CHECK("db_recheck_kernel");
  db_mirror_kernel<<<16 * sm_count(), 128, 0, st>>>(info, n_fixed, wpr, adj);
  SCAN_LAUNCH_CHECK("db_mirror_kernel");
  return SCAN_OK;
}

}  // namespace scan
