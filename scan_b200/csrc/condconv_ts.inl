// Included by condconv.cu (inside namespace scan).  Forward kernel, "TS" variant: the tf32 hi/lo operands of the
// activations live in TENSOR MEMORY instead of shared memory (tcgen05.mma with A from TMEM), so that shared memory holds
// nothing but the raw fp32 TMA ring: 12 stages x 16 KB = 192 KB in flight per SM (the SS variant keeps hi + lo per stage
// in shared memory and is limited to 6 stages = 96 KB, measured 44 % of HBM peak: occupancy-limited ring).
//
//   TMA -> smem stage (fp32, 128B swizzle) -> converter thread r reads ITS pixel row (8 x 16 B chunks, chunk c of row r
//   sits at physical chunk c ^ (r & 7)), releases the stage, splits into hi = rna_tf32(x), lo = x - hi and stores both
//   with tcgen05.st.32x32b.x32 into an operand slot (64 TMEM columns: hi | lo; lane = pixel row) -> MMA warp:
//   D[128 x 16] += A_tmem[128 x 8] . W_smem[16 x 8]^T for (hi,Whi), (hi,Wlo), (lo,Whi).
// TMEM map (512 columns allocated): [0,64) 4 accumulator slots x 16, [64, 64 + 64*TS_OPS) operand slots.

constexpr int TS_STAGES = 12;
constexpr int TS_OPS = 4;
// Accumulators: a chain of MMAs into ONE accumulator is serialised by the tensor-pipe latency (96 dependent N=16 MMAs per
// tile capped both operand variants at ~3.5 TB/s, round-1 run 12).  Per k-step the kernel therefore issues only two MMAs,
//   Da[128 x 32] += A_hi . [Whi ; Wlo]^T   (N = 32: hi*Whi and hi*Wlo at once)      Db[128 x 16] += A_lo . Whi^T
// and alternates between two accumulator sets by k-step parity: 4 independent, interleaved chains of 16 MMAs.  The epilogue adds
// the six 16-column groups.  One accumulator slot = Da0 | Da1 | Db0 | Db1 = 96 columns.
constexpr int TS_ACC = 2;
constexpr int TS_ACC_COLS = 96;
constexpr int TS_OP_COL0 = TS_ACC * TS_ACC_COLS;  // 192
constexpr uint32_t TS_IDESC32 = umma_idesc_tf32(CC_BM, 32);
constexpr int TS_SMEM = 1024 + 2 * CC_W_BYTES + TS_STAGES * CC_STAGE_BYTES + 1024;
constexpr int TS_THREADS = 512;  // warps 0-3 TMA / MMA / alloc / idle, 4-7 epilogue, 8-11 and 12-15 converters (alternate k-blocks)

// issued by one elected lane of the (converged) MMA warp
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  if (elect_one_sync()) asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}

__global__ void __launch_bounds__(TS_THREADS, 1)
    condconv_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, Levels lv,
                           ActPtrs act, const float* __restrict__ bias, const int64_t* __restrict__ labels,
                           double* __restrict__ loss_partials, int* __restrict__ flags, int K, int act_mode, int num_tiles,
                           long long kmajor_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* w_cat = smem;                                   // per k-block 4 KB: [Whi (16 rows) | Wlo (16 rows)]: one N = 32 operand
  uint8_t* stages = smem + 2 * CC_W_BYTES;
  uint64_t* bars = (uint64_t*)(stages + TS_STAGES * CC_STAGE_BYTES);
  uint64_t* full_bar = bars;                              // [TS_STAGES] TMA -> converter
  uint64_t* empty_bar = bars + TS_STAGES;                 // [TS_STAGES] converter -> TMA (128 arrivals)
  uint64_t* op_full = bars + 2 * TS_STAGES;               // [TS_OPS] converter -> MMA (128 arrivals)
  uint64_t* op_empty = bars + 2 * TS_STAGES + TS_OPS;     // [TS_OPS] MMA commit -> converter
  uint64_t* acc_full = bars + 2 * TS_STAGES + 2 * TS_OPS; // [TS_ACC]
  uint64_t* acc_empty = acc_full + TS_ACC;                // [TS_ACC]
  uint64_t* w_bar = acc_empty + TS_ACC;                   // TMA -> splitter threads
  uint64_t* w_ready = w_bar + 1;                          // splitter threads (warps >= 4) -> MMA
  uint32_t* tmem_slot = (uint32_t*)(w_ready + 1);
  __shared__ double red[4];
  __shared__ int flag_s;

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];

  if (threadIdx.x == 0) {
    for (int i = 0; i < TS_STAGES; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(empty_bar + i), 128);
    }
    for (int i = 0; i < TS_OPS; ++i) {
      mbar_init(smem_u32(op_full + i), 128);
      mbar_init(smem_u32(op_empty + i), 1);
    }
    for (int i = 0; i < TS_ACC; ++i) {
      mbar_init(smem_u32(acc_full + i), 1);
      mbar_init(smem_u32(acc_empty + i), 128);
    }
    mbar_init(smem_u32(w_bar), 1);
    mbar_init(smem_u32(w_ready), TS_THREADS - 128);
    flag_s = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(w_bar), CC_W_BYTES);
    for (int kb = 0; kb < CC_KB; ++kb) tma_load_2d(smem_u32(w_cat + kb * 2 * CC_WBLK_BYTES), &tmap_w, smem_u32(w_bar), kb * CC_BK, 0);
  }
  // The activation stream starts immediately (warp 0 below); the kernel weights are split by the epilogue / converter
  // warps while the first stages are in flight, and only the MMA warp waits for them (w_ready).
  if (warp >= 4) {
    mbar_wait(smem_u32(w_bar), 0);
    for (int i = threadIdx.x - 128; i < CC_W_BYTES / 4; i += TS_THREADS - 128) {
      float* blk = (float*)(w_cat + (i / (CC_WBLK_BYTES / 4)) * 2 * CC_WBLK_BYTES);
      const int e = i % (CC_WBLK_BYTES / 4);
      const float wv = blk[e];
      uint32_t hi;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(wv));
      blk[e] = __uint_as_float(hi);
      blk[CC_WBLK_BYTES / 4 + e] = wv - __uint_as_float(hi);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(smem_u32(w_ready));
  }
  if (blockIdx.x == 0) {  // no host-side memsets: unused partial / flag slots are cleared here
    for (int i = gridDim.x + threadIdx.x; i < CC_MAX_PARTIALS; i += TS_THREADS) {
      if (loss_partials) loss_partials[i] = 0.0;
      if (flags) flags[i] = 0;
    }
  }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < CC_KB; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          mbar_expect_tx(smem_u32(full_bar + stage), CC_STAGE_BYTES);
          if (kmajor_rows)  // experiment: k-block-major tiled activations [8][R][32]: every box is 16 KB contiguous
            tma_load_2d(smem_u32(stages + stage * CC_STAGE_BYTES), &tmap_x, smem_u32(full_bar + stage), 0,
                        (int)(kb * kmajor_rows + (long long)tile * CC_BM));
          else
            tma_load_2d(smem_u32(stages + stage * CC_STAGE_BYTES), &tmap_x, smem_u32(full_bar + stage), kb * CC_BK, tile * CC_BM);
          if (++stage == TS_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {   // all 32 lanes run the issue loop; every tcgen05.mma / commit goes out from one elected lane
      int op = 0, acc = 0;
      uint32_t op_phase = 0, acc_phase = 0;
      mbar_wait(smem_u32(w_ready), 0);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(acc_empty + acc), acc_phase ^ 1);
        tcgen05_fence_after();
        for (int kb = 0; kb < CC_KB; ++kb) {
          const uint32_t da0 = tmem_base + acc * TS_ACC_COLS;       // Da[k & 1] at +32 * (k & 1)
          const uint32_t db0 = tmem_base + acc * TS_ACC_COLS + 64;  // Db[k & 1] at +16 * (k & 1)
          mbar_wait(smem_u32(op_full + op), op_phase);
          tcgen05_fence_after();
          const uint32_t a_hi = tmem_base + TS_OP_COL0 + op * 64;
          const uint32_t a_lo = a_hi + 32;
          const uint32_t b_addr = smem_u32(w_cat + kb * 2 * CC_WBLK_BYTES);
#pragma unroll
          for (int k = 0; k < CC_BK / 8; ++k) {
            const uint64_t dwb = umma_desc_sw128(b_addr + k * 32);
            // chain set = k & 1: consecutive MMAs go Da0, Db0, Da1, Db1, ... so four chains are in flight at any time
            const uint32_t first = (kb | (k >> 1)) != 0;   // the first MMA of each chain overwrites
            umma_tf32_ts(da0 + (k & 1) * 32, a_hi + k * 8, dwb, TS_IDESC32, first);
            umma_tf32_ts(db0 + (k & 1) * 16, a_lo + k * 8, dwb, CC_IDESC, first);
          }
          if (elect_one_sync()) umma_commit(smem_u32(op_empty + op));
          if (++op == TS_OPS) { op = 0; op_phase ^= 1; }
        }
        if (elect_one_sync()) umma_commit(smem_u32(acc_full + acc));
        if (++acc == TS_ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 8) {
    // ===== converter: smem row -> registers -> hi/lo -> TMEM operand slot =====
    const int group = (warp - 8) >> 2;           // 0 or 1: alternate k-blocks
    const int r = ((warp & 3) << 5) + lane;      // pixel row of the tile == TMEM lane this warp may access
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    long long it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < CC_KB; ++kb, ++it) {
        if ((int)(it & 1) != group) continue;
        const int stage = (int)(it % TS_STAGES);
        const uint32_t sphase = (uint32_t)((it / TS_STAGES) & 1);
        const int op = (int)(it % TS_OPS);
        const uint32_t ophase = (uint32_t)((it / TS_OPS) & 1);
        mbar_wait(smem_u32(full_bar + stage), sphase);
        const uint8_t* row = stages + stage * CC_STAGE_BYTES + r * 128;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(row + ((c ^ (r & 7)) << 4));
          const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t h;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x[e]));
            hi[c * 4 + e] = h;
            lo[c * 4 + e] = __float_as_uint(x[e] - __uint_as_float(h));
          }
        }
        mbar_arrive(smem_u32(empty_bar + stage));  // the row is in registers: the TMA may refill the stage
        mbar_wait(smem_u32(op_empty + op), ophase ^ 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + lane_base + TS_OP_COL0 + op * 64;
        tmem_st32(taddr, hi);
        tmem_st32(taddr + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tcgen05_fence_before();
        mbar_arrive(smem_u32(op_full + op));
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    double loss = 0.0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(acc_full + acc), acc_phase);
      tcgen05_fence_after();
      float z[16], part[16];
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TS_ACC_COLS;
      tmem_ld16(tacc, z);                      // Da0: hi*Whi (even k-blocks)
#pragma unroll
      for (int grp = 1; grp < 6; ++grp) {      // Da0 hi*Wlo, Da1 hi*Whi, Da1 hi*Wlo, Db0 lo*Whi, Db1 lo*Whi
        tmem_ld16(tacc + grp * 16, part);
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) z[kk] += part[kk];
      }
      tcgen05_fence_before();
      mbar_arrive(smem_u32(acc_empty + acc));
      const long long g = (long long)tile * CC_BM + q * 32 + lane;
      if (g < R) loss += act_epilogue(lv, act, g, z, K, act_mode, bias, labels, &flag_s);
      if (++acc == TS_ACC) { acc = 0; acc_phase ^= 1; }
    }
    if (loss_partials) {
      loss = warp_sum_d(loss);
      if (lane == 0) red[q] = loss;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (loss_partials) loss_partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
    if (flags) flags[blockIdx.x] = flag_s;
  }
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}
