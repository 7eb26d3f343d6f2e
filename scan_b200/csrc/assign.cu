// K1: node generation -- FCOS ground-truth assignment, deterministic node sampling, row gather/scatter.
// Reference: modeling/rpn/fcos/loss.py:262-343 (assignment), :430-458 and :497-516 (sampling),
// modeling/rpn/fcos/condgraph.py:631-655 (locations).  All integer results are bit-exact.
#include "common.cuh"

namespace scan {

// ------------------------------------------------------------------------------------------------
// FCOS assignment.  One thread per row g (level-first, image-major, y, x).  fp32 arithmetic with
// explicit round-to-nearest intrinsics so that no FMA contraction can change a comparison.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fcos_assign_kernel(Levels lv, const float* __restrict__ boxes,
                                                          const int64_t* __restrict__ box_labels,
                                                          const int32_t* __restrict__ box_count, int g_max,
                                                          int64_t* __restrict__ labels_out, float4* __restrict__ reg_out) {
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= R) return;
  const int l = level_of_row(lv, g);
  const int hw = lv.h[l] * lv.w[l];
  const long long r = g - lv.row_off[l];
  const int n = (int)(r / hw);
  const int p = (int)(r - (long long)n * hw);
  const int yi = p / lv.w[l], xi = p - yi * lv.w[l];
  const int s = lv.stride[l];
  // condgraph.py:642-655: arange(0, w*s, s) + s//2, exact in fp32
  const float x = (float)(xi * s) + (float)(s / 2);
  const float y = (float)(yi * s) + (float)(s / 2);
  // loss.py:263-269 object_sizes_of_interest (INF = 1e8), per level index
  const float INF = 100000000.0f;
  float lo, hi;
  switch (l) {
    case 0: lo = -1.f; hi = 64.f; break;
    case 1: lo = 64.f; hi = 128.f; break;
    case 2: lo = 128.f; hi = 256.f; break;
    case 3: lo = 256.f; hi = 512.f; break;
    default: lo = 512.f; hi = INF; break;
  }
  const int G = box_count[n];
  const float4* b4 = reinterpret_cast<const float4*>(boxes) + (long long)n * g_max;
  float best = INF;
  int best_i = 0;
  for (int i = 0; i < G; ++i) {
    const float4 b = __ldg(b4 + i);
    const float dl = __fsub_rn(x, b.x), dt = __fsub_rn(y, b.y);
    const float dr = __fsub_rn(b.z, x), db = __fsub_rn(b.w, y);
    const float mn = fminf(fminf(dl, dt), fminf(dr, db));
    const float mx = fmaxf(fmaxf(dl, dt), fmaxf(dr, db));
    // structures/bounding_box.py:229-230: (x1 - x0 + 1) * (y1 - y0 + 1)
    float area = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
    const bool ok = (mn > 0.f) && (mx >= lo) && (mx <= hi);
    if (!ok) area = INF;
    if (area < best) {  // strict: first minimum wins (torch.min tie rule)
      best = area;
      best_i = i;
    }
  }
  int64_t lab = 0;
  if (best != INF) lab = __ldg(box_labels + (long long)n * g_max + best_i);
  labels_out[g] = lab;
  if (reg_out) {   // loss.py:111: reg_targets of the chosen box (box 0 when nothing matched: argmin of an all-INF row)
    const float4 b = __ldg(b4 + best_i);
    reg_out[g] = make_float4(__fsub_rn(x, b.x), __fsub_rn(y, b.y), __fsub_rn(b.z, x), __fsub_rn(b.w, y));
  }
}

// ------------------------------------------------------------------------------------------------
// Node sampling: count -> scan -> fill -> negative selection.  Everything stays on the device; the
// caller reads `meta` when it needs M.
// ------------------------------------------------------------------------------------------------
constexpr int SB = 1024;  // rows per block (= threads per block)

__device__ __forceinline__ bool is_pos(int mode, const int64_t* labels, const uint8_t* mask, long long g) {
  return mode == 0 ? (labels[g] > 0) : (mask[g] != 0);
}

// block_pos[b * SCAN_MAX_LEVELS + l] = number of positive rows of level l inside block b
__global__ void __launch_bounds__(SB) sample_count_kernel(Levels lv, int mode, const int64_t* __restrict__ labels,
                                                          const uint8_t* __restrict__ mask, int* __restrict__ block_pos) {
  __shared__ int cnt[SCAN_MAX_LEVELS];
  if (threadIdx.x < SCAN_MAX_LEVELS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const long long g = (long long)blockIdx.x * SB + threadIdx.x;
  if (g < R && is_pos(mode, labels, mask, g)) atomicAdd(&cnt[level_of_row(lv, g)], 1);
  __syncthreads();
  if (threadIdx.x < SCAN_MAX_LEVELS) block_pos[blockIdx.x * SCAN_MAX_LEVELS + threadIdx.x] = cnt[threadIdx.x];
}

// floor(np.linspace(0, stop, num))[j] in float64, numpy semantics (loss.py:448, 503)
__device__ __forceinline__ long long floor_linspace(long long stop, long long num, long long j) {
  if (num <= 1) return 0;
  if (j == num - 1) return stop;  // numpy forces the last sample to `stop`
  const double step = (double)stop / (double)(num - 1);
  return (long long)floor(__dmul_rn((double)j, step));
}

// single block: exclusive scan of block_pos over blocks per level, then the meta record
__global__ void __launch_bounds__(1024) sample_scan_kernel(Levels lv, int mode, int with_bg, int n_blocks, int cap,
                                                           int* __restrict__ block_pos, scan_sample_meta_t* __restrict__ meta) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  __shared__ int total[SCAN_MAX_LEVELS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int l = 0; l < lv.n_levels; ++l) {
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
      const int b = base + threadIdx.x;
      int v = (b < n_blocks) ? block_pos[b * SCAN_MAX_LEVELS + l] : 0;
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) warp_tot[wid] = incl;
      __syncthreads();
      if (wid == 0) {
        int w = warp_tot[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int t = __shfl_up_sync(0xffffffffu, wi, o);
          if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;  // exclusive prefix of warp totals
      }
      __syncthreads();
      const int excl = carry + warp_tot[wid] + incl - v;
      if (b < n_blocks) block_pos[b * SCAN_MAX_LEVELS + l] = excl;
      __syncthreads();
      if (threadIdx.x == 1023) carry = excl + v;
      __syncthreads();
    }
    if (threadIdx.x == 0) total[l] = carry;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int err = 0;
    long long off = 0;
    for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
      int np = 0, nn = 0, ns = 0;
      if (l < lv.n_levels) {
        np = total[l];
        nn = (int)(lv.row_off[l + 1] - lv.row_off[l]) - np;
        if (mode == 0) {
          ns = with_bg ? ((np > nn) ? nn : np) : 0;  // loss.py:445-449
        } else {
          ns = np;                                   // loss.py:503
          if (np > 0 && nn == 0) { err = 1; ns = 0; }
        }
      }
      meta->n_pos[l] = np;
      meta->n_neg[l] = nn;
      meta->n_neg_sel[l] = ns;
      meta->neg_off[l] = (int)off;
      off += ns;
    }
    meta->n_neg_nodes = (int)off;
    for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
      meta->pos_off[l] = (int)off;
      off += meta->n_pos[l];
    }
    if (off > cap) err = 2;
    meta->n_nodes = (int)off;
    meta->error = err;
    meta->reserved = 0;
  }
}

// positives go straight to their node slot; negatives are listed per level (neg_list[row_off[l] + rank])
__global__ void __launch_bounds__(SB) sample_fill_kernel(Levels lv, int mode, const int64_t* __restrict__ labels,
                                                         const uint8_t* __restrict__ mask, const int64_t* __restrict__ plabel,
                                                         const int* __restrict__ block_pos_excl,
                                                         const scan_sample_meta_t* __restrict__ meta,
                                                         int* __restrict__ neg_list, int32_t* __restrict__ node_rows,
                                                         int64_t* __restrict__ node_labels) {
  __shared__ int warp_cnt[SCAN_MAX_LEVELS][32];
  if (meta->error == 2) return;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  const long long g = (long long)blockIdx.x * SB + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const bool valid = g < R;
  const int l = valid ? level_of_row(lv, g) : -1;
  const bool pos = valid && is_pos(mode, labels, mask, g);
  // rank of this row among the block's positives of the same level
  int my_rank = 0;
  for (int j = 0; j < lv.n_levels; ++j) {
    const unsigned bal = __ballot_sync(0xffffffffu, pos && l == j);
    if (lane == 0) warp_cnt[j][wid] = __popc(bal);
    if (l == j) my_rank = __popc(bal & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (!valid) return;
  int before = 0;
  for (int w2 = 0; w2 < wid; ++w2) before += warp_cnt[l][w2];
  const int pos_rank = block_pos_excl[blockIdx.x * SCAN_MAX_LEVELS + l] + before + my_rank;  // positives of level l before g
  if (pos) {
    const int slot = meta->pos_off[l] + pos_rank;
    node_rows[slot] = (int32_t)g;
    node_labels[slot] = mode == 0 ? labels[g] : plabel[g];
  } else {
    const long long neg_rank = (g - lv.row_off[l]) - pos_rank;
    neg_list[lv.row_off[l] + neg_rank] = (int)g;
  }
}

__global__ void __launch_bounds__(256) sample_neg_kernel(Levels lv, int mode, const scan_sample_meta_t* __restrict__ meta,
                                                         const int* __restrict__ neg_list, int32_t* __restrict__ node_rows,
                                                         int64_t* __restrict__ node_labels) {
  if (meta->error == 2) return;
  const int n_neg_nodes = meta->n_neg_nodes;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_neg_nodes; i += gridDim.x * blockDim.x) {
    int l = 0;
    for (int j = 1; j < lv.n_levels; ++j)
      if (i >= meta->neg_off[j]) l = j;
    const long long j = i - meta->neg_off[l];
    const long long nn = meta->n_neg[l], ns = meta->n_neg_sel[l], np = meta->n_pos[l];
    long long r;
    if (mode == 0 && np > nn) {
      r = j;  // all negatives, in order
    } else {
      r = floor_linspace(nn - 2, ns, j);
      if (r < 0) r += nn;  // python negative indexing (n_neg == 1)
    }
    node_rows[i] = neg_list[lv.row_off[l] + r];
    node_labels[i] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Row gather / scatter-add.  One warp per node; C/4 float4 per row, 128-bit coalesced.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ rows, const int32_t* __restrict__ idx,
                                                          int n, int c4, float4* __restrict__ out) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
    const float4* src = rows + (long long)__ldg(idx + i) * c4;
    float4* dst = out + (long long)i * c4;
    for (int j = lane; j < c4; j += 32) dst[j] = __ldg(src + j);
  }
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float4* __restrict__ d_nodes, const int32_t* __restrict__ idx,
                                                               int n, int c4, float* __restrict__ d_rows) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < n; i += gridDim.x * warps_per_block) {
    float* dst = d_rows + (long long)__ldg(idx + i) * c4 * 4;
    const float4* src = d_nodes + (long long)i * c4;
    for (int j = lane; j < c4; j += 32) {
      const float4 v = __ldg(src + j);
      atomicAdd(dst + 4 * j + 0, v.x);
      atomicAdd(dst + 4 * j + 1, v.y);
      atomicAdd(dst + 4 * j + 2, v.z);
      atomicAdd(dst + 4 * j + 3, v.w);
    }
  }
}

}  // namespace scan

extern "C" int scan_fcos_assign(const scan_levels_t* lvh, const float* boxes, const int64_t* box_labels,
                                const int32_t* box_count, int32_t g_max, int64_t* labels_out, void* stream) {
  scan::Levels lv;
  int rc = scan::make_levels(lvh, &lv);
  if (rc) return rc;
  if (!boxes || !box_labels || !box_count || !labels_out || g_max < 1) return SCAN_EINVAL;
  if (lv.n_levels > 5) return SCAN_ENOTSUP;  // the reference defines 5 size ranges (loss.py:263-269)
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  scan::fcos_assign_kernel<<<(unsigned)scan::ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(
      lv, boxes, box_labels, box_count, g_max, labels_out, nullptr);
  SCAN_LAUNCH_CHECK("fcos_assign_kernel");
  return SCAN_OK;
}

// the same assignment for FCOSLossComputation (loss.py:40-126), which also needs the (l, t, r, b) regression targets [R, 4]
extern "C" int scan_fcos_assign_reg(const scan_levels_t* lvh, const float* boxes, const int64_t* box_labels,
                                    const int32_t* box_count, int32_t g_max, int64_t* labels_out, float* reg_targets_out,
                                    void* stream) {
  scan::Levels lv;
  int rc = scan::make_levels(lvh, &lv);
  if (rc) return rc;
  if (!boxes || !box_labels || !box_count || !labels_out || !reg_targets_out || g_max < 1) return SCAN_EINVAL;
  if (lv.n_levels > 5) return SCAN_ENOTSUP;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  scan::fcos_assign_kernel<<<(unsigned)scan::ceil_div(R, 256), 256, 0, (cudaStream_t)stream>>>(
      lv, boxes, box_labels, box_count, g_max, labels_out, reinterpret_cast<float4*>(reg_targets_out));
  SCAN_LAUNCH_CHECK("fcos_assign_kernel");
  return SCAN_OK;
}

extern "C" int64_t scan_sample_workspace_bytes(int64_t n_rows) {
  const int64_t n_blocks = (n_rows + scan::SB - 1) / scan::SB;
  return n_blocks * SCAN_MAX_LEVELS * 4 + n_rows * 4 + 256;
}

extern "C" int scan_sample_nodes(const scan_levels_t* lvh, int32_t mode, int32_t with_bg, const int64_t* labels,
                                 const uint8_t* pos_mask, const int64_t* plabel, int32_t* node_rows,
                                 int64_t* node_labels, int32_t cap, scan_sample_meta_t* meta, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  scan::Levels lv;
  int rc = scan::make_levels(lvh, &lv);
  if (rc) return rc;
  if (mode == 0 ? !labels : (!pos_mask || !plabel)) return SCAN_EINVAL;
  if (!node_rows || !node_labels || !meta || !workspace || cap < 1) return SCAN_EINVAL;
  const long long R = lv.row_off[SCAN_MAX_LEVELS];
  if (workspace_bytes < scan_sample_workspace_bytes(R)) return SCAN_ECAPACITY;
  const int n_blocks = (int)scan::ceil_div(R, scan::SB);
  int* block_pos = (int*)workspace;
  int* neg_list = block_pos + ((long long)n_blocks * SCAN_MAX_LEVELS + 63) / 64 * 64;
  cudaStream_t st = (cudaStream_t)stream;
  scan::sample_count_kernel<<<n_blocks, scan::SB, 0, st>>>(lv, mode, labels, pos_mask, block_pos);
  SCAN_LAUNCH_CHECK("sample_count_kernel");
  scan::sample_scan_kernel<<<1, 1024, 0, st>>>(lv, mode, with_bg, n_blocks, cap, block_pos, meta);
  SCAN_LAUNCH_CHECK("sample_scan_kernel");
  scan::sample_fill_kernel<<<n_blocks, scan::SB, 0, st>>>(lv, mode, labels, pos_mask, plabel, block_pos, meta,
                                                          neg_list, node_rows, node_labels);
  SCAN_LAUNCH_CHECK("sample_fill_kernel");
  scan::sample_neg_kernel<<<2 * scan::sm_count(), 256, 0, st>>>(lv, mode, meta, neg_list, node_rows, node_labels);
  SCAN_LAUNCH_CHECK("sample_neg_kernel");
  return SCAN_OK;
}

extern "C" int scan_gather_rows(const float* rows, const int32_t* node_rows, int32_t n_nodes, int32_t channels,
                                float* out, void* stream) {
  if (n_nodes == 0) return SCAN_OK;
  if (!rows || !node_rows || !out || n_nodes < 0 || channels < 4 || channels % 4) return SCAN_EINVAL;
  const int blocks = (int)std::min<long long>(scan::ceil_div(n_nodes, 8), 8ll * scan::sm_count());
  scan::gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)rows, node_rows, n_nodes,
                                                                    channels / 4, (float4*)out);
  SCAN_LAUNCH_CHECK("gather_rows_kernel");
  return SCAN_OK;
}

extern "C" int scan_scatter_add_rows(const float* d_nodes, const int32_t* node_rows, int32_t n_nodes,
                                     int32_t channels, float* d_rows, void* stream) {
  if (n_nodes == 0) return SCAN_OK;
  if (!d_nodes || !node_rows || !d_rows || n_nodes < 0 || channels < 4 || channels % 4) return SCAN_EINVAL;
  const int blocks = (int)std::min<long long>(scan::ceil_div(n_nodes, 8), 8ll * scan::sm_count());
  scan::scatter_add_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)d_nodes, node_rows, n_nodes,
                                                                         channels / 4, d_rows);
  SCAN_LAUNCH_CHECK("scatter_add_rows_kernel");
  return SCAN_OK;
}
