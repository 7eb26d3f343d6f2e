// f4: the FCOS post-processor behind TEST.MODE 'common' / 'light' / 'precision'
// (modeling/rpn/fcos/inference.py:54-194 FCOSPostProcessor, structures/boxlist_ops.py:9-31 boxlist_nms + :58-74 remove_small_boxes,
// structures/bounding_box.py:214-224 clip_to_image, csrc/cuda/nms.cu:13-67 IoU / suppression rule).
// Input: the per-level class PROBABILITY maps the ensembling kernel produced (scan_ensemble_levels), the regression and
// centerness maps of the FCOS head, all NCHW and read in place.  The reference runs a Python loop over levels x images with
// boolean-mask indexing, nonzero, topk, per-class nonzero + NMS (host round trips everywhere) and a CPU kthvalue.  Here: four
// launches for all levels and images, no host synchronisation until the caller reads the per-image detection counts.
//
//   pp_select_kernel   block = (level, image): candidates = (location, class) with prob > thresh in the reference's order
//                      (location-major, class-minor); more than PRE_NMS_TOP_N -> the k-th largest score prob * sigmoid(ctr) by
//                      a 4-pass radix select on the float bits (ties at the threshold resolved by candidate order); ordered
//                      compaction with box decode, clip_to_image and the min-size filter fused in.
//   pp_sort_kernel     block = image: gathers the level segments, bitonic sort of 64-bit keys (label | ~score bits | index) in
//                      shared memory -> per class, score-descending order (the order the greedy NMS needs).
//   pp_mask_kernel     suppression bit matrix (same class, IoU > NMS_TH, +1 pixel convention), 64 columns per block.
//   pp_reduce_kernel   block = image: one warp walks the rows greedily (removed-set in registers, 3 words per lane), then the
//                      DETECTIONS_PER_IMG cap (keep score >= the top_n-th largest, like the kthvalue threshold of :181-187) and
//                      the output order of the reference: class ascending, original candidate order inside a class.
#include "common.cuh"

namespace scan {

constexpr int PP_MAX_K = 1024;       // PRE_NMS_TOP_N capacity per (image, level)
constexpr int PP_SORT_N = 8192;      // detections per image the sort handles (5 levels x 1000 = 5000 <= 8192)
constexpr int PP_WORDS = PP_SORT_N / 64;

struct PpMaps {
  const float* prob[SCAN_MAX_LEVELS];   // [N, C, H, W] probabilities
  const float* reg[SCAN_MAX_LEVELS];    // [N, 4, H, W]
  const float* ctr[SCAN_MAX_LEVELS];    // [N, 1, H, W] logits
};

__device__ __forceinline__ float pp_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

// block-wide exclusive scan of one int per thread (1024 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int w = warp_tot[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    warp_tot[lane] = wi - w;
    if (lane == 31) *total = wi;
  }
  __syncthreads();
  const int r = warp_tot[wid] + incl - v;
  __syncthreads();
  return r;
}

// one block per (level, image)
__global__ void __launch_bounds__(1024) pp_select_kernel(Levels lv, PpMaps mp, int num_classes, float thresh, int top_n, float min_size,
                                                         const int* __restrict__ image_hw, float* __restrict__ seg_box,
                                                         float* __restrict__ seg_score, int* __restrict__ seg_label, int* __restrict__ seg_count) {
  __shared__ int hist[256];
  __shared__ int warp_tot[32];
  __shared__ int s_total, s_bin, s_krem;
  const int l = blockIdx.x, n = blockIdx.y;
  if (l >= lv.n_levels) return;
  const int hw = lv.h[l] * lv.w[l], C = num_classes;
  const long long n_el = (long long)hw * C;                  // candidate index e = loc * C + c (inference.py:66-74)
  const float* prob = mp.prob[l] + (long long)n * C * hw;
  const float* ctr = mp.ctr[l] + (long long)n * hw;
  const float* reg = mp.reg[l] + (long long)n * 4 * hw;
  auto score_of = [&](long long e, bool& cand) -> float {
    const int loc = (int)(e / C), c = (int)(e - (long long)loc * C);
    const float p = __ldg(prob + (long long)c * hw + loc);
    cand = p > thresh;
    return p * pp_sigmoid(__ldg(ctr + loc));
  };
  // ---- pass 0: candidate count
  int cnt = 0;
  for (long long e = threadIdx.x; e < n_el; e += 1024) {
    bool cand;
    score_of(e, cand);
    cnt += cand;
  }
  int total;
  block_excl_scan(cnt, warp_tot, &s_total);
  total = s_total;
  // ---- radix select of the top_n-th largest score among the candidates (only when there are more than top_n)
  unsigned int prefix = 0, prefix_mask = 0;
  int k_rem = top_n;
  const bool select = total > top_n;
  if (select) {
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      if (threadIdx.x < 256) hist[threadIdx.x] = 0;
      __syncthreads();
      for (long long e = threadIdx.x; e < n_el; e += 1024) {
        bool cand;
        const unsigned int b = __float_as_uint(score_of(e, cand));
        if (cand && (b & prefix_mask) == prefix) atomicAdd(&hist[(b >> shift) & 255], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int acc = 0, bin = 255;
        for (; bin > 0; --bin) {
          if (acc + hist[bin] >= k_rem) break;
          acc += hist[bin];
        }
        s_bin = bin;
        s_krem = k_rem - acc;
      }
      __syncthreads();
      prefix |= (unsigned int)s_bin << shift;
      prefix_mask |= 255u << shift;
      k_rem = s_krem;
      __syncthreads();
    }
  }
  // prefix = bits of the top_n-th largest score, k_rem = how many candidates EQUAL to it are taken (in candidate order)
  // ---- ordered compaction, chunk by chunk (1024 elements per round keeps the candidate order)
  int out_base = 0, eq_base = 0;
  const long long seg = ((long long)n * lv.n_levels + l) * PP_MAX_K;
  const float iw = (float)image_hw[2 * n + 1], ih = (float)image_hw[2 * n];
  const int s = lv.stride[l];
  for (long long e0 = 0; e0 < n_el; e0 += 1024) {
    const long long e = e0 + threadIdx.x;
    bool cand = false;
    float sc = 0.f;
    if (e < n_el) sc = score_of(e, cand);
    const unsigned int bits = __float_as_uint(sc);
    const bool gt = cand && (!select || bits > prefix);
    const bool eq = cand && select && bits == prefix;
    int tot_eq;
    const int eq_rank = block_excl_scan(eq ? 1 : 0, warp_tot, &s_total);
    tot_eq = s_total;
    const bool take = gt || (eq && eq_base + eq_rank < k_rem);
    // decode + clip + min-size filter (inference.py:105-118): a box removed here still consumed its top-k slot
    float x1 = 0.f, y1 = 0.f, x2 = 0.f, y2 = 0.f;
    bool keep = false;
    int c = 0;
    if (take) {
      const int loc = (int)(e / C);
      c = (int)(e - (long long)loc * C);
      const int yi = loc / lv.w[l], xi = loc - yi * lv.w[l];
      const float px = (float)(xi * s) + (float)(s / 2), py = (float)(yi * s) + (float)(s / 2);
      x1 = px - __ldg(reg + loc);
      y1 = py - __ldg(reg + hw + loc);
      x2 = px + __ldg(reg + 2 * hw + loc);
      y2 = py + __ldg(reg + 3 * hw + loc);
      x1 = fminf(fmaxf(x1, 0.f), iw - 1.f);
      y1 = fminf(fmaxf(y1, 0.f), ih - 1.f);
      x2 = fminf(fmaxf(x2, 0.f), iw - 1.f);
      y2 = fminf(fmaxf(y2, 0.f), ih - 1.f);
      keep = (x2 - x1 + 1.f >= min_size) && (y2 - y1 + 1.f >= min_size);
    }
    int tot_keep;
    const int pos = block_excl_scan(keep ? 1 : 0, warp_tot, &s_total);
    tot_keep = s_total;
    if (keep && out_base + pos < PP_MAX_K) {
      const long long o = seg + out_base + pos;
      reinterpret_cast<float4*>(seg_box)[o] = make_float4(x1, y1, x2, y2);
      seg_score[o] = sqrtf(sc);
      seg_label[o] = c + 1;
    }
    out_base += tot_keep;
    eq_base += tot_eq;
  }
  if (threadIdx.x == 0) seg_count[n * lv.n_levels + l] = min(out_base, PP_MAX_K);
}

// one block per image: concatenate the level segments, sort by (label asc, score desc, index asc)
__global__ void __launch_bounds__(1024) pp_sort_kernel(int n_levels, const float* __restrict__ seg_box, const float* __restrict__ seg_score,
                                                       const int* __restrict__ seg_label, const int* __restrict__ seg_count,
                                                       float* __restrict__ s_box, float* __restrict__ s_score, int* __restrict__ s_label,
                                                       int* __restrict__ s_orig, int* __restrict__ n_det) {
  extern __shared__ unsigned long long keys[];   // [PP_SORT_N]
  __shared__ int off[SCAN_MAX_LEVELS + 1];
  const int n = blockIdx.x;
  if (threadIdx.x == 0) {
    int o = 0;
    for (int l = 0; l < n_levels; ++l) {
      off[l] = o;
      o += seg_count[n * n_levels + l];
    }
    off[n_levels] = min(o, PP_SORT_N);
    n_det[n] = off[n_levels];
  }
  __syncthreads();
  const int total = off[n_levels];
  for (int i = threadIdx.x; i < PP_SORT_N; i += 1024) {
    unsigned long long k = ~0ull;
    if (i < total) {
      int l = 0;
      while (l + 1 < n_levels && i >= off[l + 1]) ++l;
      const long long src = ((long long)n * n_levels + l) * PP_MAX_K + (i - off[l]);
      const unsigned int sb = ~__float_as_uint(seg_score[src]);
      k = ((unsigned long long)(unsigned int)seg_label[src] << 48) | ((unsigned long long)sb << 16) | (unsigned long long)i;
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= PP_SORT_N; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < PP_SORT_N / 2; t += 1024) {
        const int lo = (t / stride) * stride * 2 + (t % stride), hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a > b) == up) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < total; i += 1024) {
    const int orig = (int)(keys[i] & 0xFFFF);
    int l = 0;
    while (l + 1 < n_levels && orig >= off[l + 1]) ++l;
    const long long src = ((long long)n * n_levels + l) * PP_MAX_K + (orig - off[l]);
    const long long dst = (long long)n * PP_SORT_N + i;
    reinterpret_cast<float4*>(s_box)[dst] = reinterpret_cast<const float4*>(seg_box)[src];
    s_score[dst] = seg_score[src];
    s_label[dst] = seg_label[src];
    s_orig[dst] = orig;
  }
}

// csrc/cuda/nms.cu:13-21
__device__ __forceinline__ float pp_iou(const float4 a, const float4 b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float width = fmaxf(right - left + 1.f, 0.f), height = fmaxf(bottom - top + 1.f, 0.f);
  const float inter = width * height;
  const float sa = (a.z - a.x + 1.f) * (a.w - a.y + 1.f), sb = (b.z - b.x + 1.f) * (b.w - b.y + 1.f);
  return inter / (sa + sb - inter);
}

// grid (PP_WORDS, N): block (w, n) fills mask[n][i][w] for every row i: bit b set <=> column j = 64 w + b is suppressed by row i
__global__ void __launch_bounds__(256) pp_mask_kernel(const float* __restrict__ s_box, const int* __restrict__ s_label, const int* __restrict__ n_det,
                                                      float nms_thresh, unsigned long long* __restrict__ mask) {
  __shared__ float4 cb[64];
  __shared__ int cl[64];
  const int w = blockIdx.x, n = blockIdx.y;
  const int total = n_det[n];
  const int j0 = w * 64;
  if (j0 >= total) return;
  const float4* box = reinterpret_cast<const float4*>(s_box) + (long long)n * PP_SORT_N;
  const int* lab = s_label + (long long)n * PP_SORT_N;
  if (threadIdx.x < 64) {
    const int j = j0 + threadIdx.x;
    cb[threadIdx.x] = j < total ? box[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    cl[threadIdx.x] = j < total ? lab[j] : -1;
  }
  __syncthreads();
  const int i_end = min(total, j0 + 64);     // only rows before the last column of this word can suppress it
  for (int i = threadIdx.x; i < i_end; i += 256) {
    const float4 a = box[i];
    const int la = lab[i];
    unsigned long long bits = 0;
#pragma unroll 8
    for (int b = 0; b < 64; ++b) {
      const int j = j0 + b;
      if (j > i && cl[b] == la && pp_iou(a, cb[b]) > nms_thresh) bits |= 1ull << b;
    }
    mask[((long long)n * PP_SORT_N + i) * PP_WORDS + w] = bits;
  }
}

// one block per image
__global__ void __launch_bounds__(1024) pp_reduce_kernel(const float* __restrict__ s_box, const float* __restrict__ s_score, const int* __restrict__ s_label,
                                                         const int* __restrict__ s_orig, const int* __restrict__ n_det,
                                                         const unsigned long long* __restrict__ mask, int post_top_n,
                                                         float* __restrict__ out_box, float* __restrict__ out_score, int* __restrict__ out_label,
                                                         int* __restrict__ out_count) {
  extern __shared__ unsigned char sm[];
  unsigned char* keep = sm;                                     // [PP_SORT_N]
  __shared__ int warp_tot[32];
  __shared__ int s_total;
  __shared__ float s_thr;
  const int n = blockIdx.x;
  const int total = n_det[n];
  const int words = (total + 63) / 64;
  for (int i = threadIdx.x; i < PP_SORT_N; i += 1024) keep[i] = 0;
  __syncthreads();
  if (threadIdx.x < 32) {
    // greedy walk in (class, score-descending) order; lane L owns the removed-set words L, L + 32, L + 64, L + 96
    const int lane = threadIdx.x;
    unsigned long long remv[4] = {0, 0, 0, 0};
    for (int i = 0; i < total; ++i) {
      const int wd = i >> 6;
      const unsigned long long word = __shfl_sync(0xffffffffu, remv[wd >> 5], wd & 31);
      if (!((word >> (i & 63)) & 1ull)) {
        if (lane == 0) keep[i] = 1;
        const unsigned long long* row = mask + ((long long)n * PP_SORT_N + i) * PP_WORDS;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int wq = lane + 32 * q;
          if (wq >= wd && wq < words) remv[q] |= row[wq];     // words before wd were never written for this row (j > i only)
        }
      }
    }
  }
  __syncthreads();
  // ---- DETECTIONS_PER_IMG cap (inference.py:177-187): threshold = the post_top_n-th largest kept score, keep score >= it
  int my_cnt = 0;
  for (int i = threadIdx.x; i < total; i += 1024) my_cnt += keep[i];
  block_excl_scan(my_cnt, warp_tot, &s_total);
  const int n_keep = s_total;
  if (threadIdx.x == 0) s_thr = -1.f;
  __syncthreads();
  const float* score = s_score + (long long)n * PP_SORT_N;
  if (post_top_n > 0 && n_keep > post_top_n) {
    for (int i = threadIdx.x; i < total; i += 1024) {
      if (!keep[i]) continue;
      const float v = score[i];
      int gt = 0, eq = 0;
      for (int j = 0; j < total; ++j)
        if (keep[j]) {
          const float u = score[j];
          gt += u > v;
          eq += u == v;
        }
      if (gt <= post_top_n - 1 && post_top_n - 1 < gt + eq) s_thr = v;   // every thread that gets here writes the same value
    }
    __syncthreads();
    const float thr = s_thr;
    for (int i = threadIdx.x; i < total; i += 1024)
      if (keep[i] && score[i] < thr) keep[i] = 0;
  }
  __syncthreads();
  // ---- emit in the reference's order: class ascending, original candidate order inside a class (boxlist[keep] keeps the
  // order of the concatenated per-level lists; csrc/cuda/nms.cu:123-129 returns the kept indices sorted ascending)
  const int* lab = s_label + (long long)n * PP_SORT_N;
  const int* orig = s_orig + (long long)n * PP_SORT_N;
  int n_out = 0;
  for (int i = threadIdx.x; i < total; i += 1024) {
    if (!keep[i]) continue;
    const int li = lab[i], oi = orig[i];
    int pos = 0;
    for (int j = 0; j < total; ++j)
      if (keep[j]) {
        const int lj = lab[j];
        pos += (lj < li) || (lj == li && orig[j] < oi);
      }
    const long long dst = (long long)n * PP_SORT_N + pos;
    reinterpret_cast<float4*>(out_box)[dst] = reinterpret_cast<const float4*>(s_box)[(long long)n * PP_SORT_N + i];
    out_score[dst] = score[i];
    out_label[dst] = li;
    ++n_out;
  }
  block_excl_scan(n_out, warp_tot, &s_total);
  if (threadIdx.x == 0) out_count[n] = s_total;
}

}  // namespace scan

using namespace scan;

extern "C" int64_t scan_postprocess_workspace_bytes(int32_t n_images, int32_t n_levels) {
  const long long seg = (long long)n_images * n_levels * PP_MAX_K;
  const long long per = (long long)n_images * PP_SORT_N;
  return seg * (16 + 4 + 4) + (long long)n_images * n_levels * 4 + per * (16 + 4 + 4 + 4) + n_images * 4 + per * PP_WORDS * 8 + 4096;
}

// out_box [N, 8192, 4], out_score [N, 8192], out_label [N, 8192] int32, out_count [N] int32; image_hw [N, 2] int32 (h, w) on the device
extern "C" int scan_postprocess(const scan_levels_t* lvh, const void* const* prob_host, const void* const* reg_host, const void* const* ctr_host,
                                int32_t num_classes, const int32_t* image_hw, float pre_nms_thresh, int32_t pre_nms_top_n, float nms_thresh,
                                int32_t post_top_n, float min_size, float* out_box, float* out_score, int32_t* out_label, int32_t* out_count,
                                void* workspace, int64_t workspace_bytes, void* stream) {
  Levels lv;
  int rc = make_levels(lvh, &lv);
  if (rc) return rc;
  if (!prob_host || !reg_host || !ctr_host || !image_hw || !out_box || !out_score || !out_label || !out_count || !workspace) return SCAN_EINVAL;
  if (num_classes < 1 || num_classes > 255 || pre_nms_top_n < 1 || pre_nms_top_n > PP_MAX_K) return SCAN_ENOTSUP;
  if ((long long)lv.n_levels * pre_nms_top_n > PP_SORT_N) return SCAN_ENOTSUP;
  if (workspace_bytes < scan_postprocess_workspace_bytes(lv.n_images, lv.n_levels)) return SCAN_ECAPACITY;
  PpMaps mp;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    const bool on = l < lv.n_levels;
    mp.prob[l] = on ? (const float*)prob_host[l] : nullptr;
    mp.reg[l] = on ? (const float*)reg_host[l] : nullptr;
    mp.ctr[l] = on ? (const float*)ctr_host[l] : nullptr;
    if (on && (!mp.prob[l] || !mp.reg[l] || !mp.ctr[l])) return SCAN_EINVAL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int N = lv.n_images, L = lv.n_levels;
  char* p = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  auto take = [&](long long bytes) {
    char* r = p;
    p = (char*)(((uintptr_t)(p + bytes) + 255) & ~(uintptr_t)255);
    return r;
  };
  const long long seg = (long long)N * L * PP_MAX_K, per = (long long)N * PP_SORT_N;
  float* seg_box = (float*)take(seg * 16);
  float* seg_score = (float*)take(seg * 4);
  int* seg_label = (int*)take(seg * 4);
  int* seg_count = (int*)take((long long)N * L * 4);
  float* s_box = (float*)take(per * 16);
  float* s_score = (float*)take(per * 4);
  int* s_label = (int*)take(per * 4);
  int* s_orig = (int*)take(per * 4);
  int* n_det = (int*)take(N * 4);
  unsigned long long* mask = (unsigned long long*)take(per * PP_WORDS * 8);
  if (p > (char*)workspace + workspace_bytes) return SCAN_ECAPACITY;
  pp_select_kernel<<<dim3(L, N), 1024, 0, st>>>(lv, mp, num_classes, pre_nms_thresh, pre_nms_top_n, min_size, image_hw, seg_box, seg_score,
                                               seg_label, seg_count);
  SCAN_LAUNCH_CHECK("pp_select_kernel");
  static unsigned long long attr = 0;
  if (first_use_on_device(&attr)) {
    SCAN_CUDA_CHECK(cudaFuncSetAttribute(pp_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SORT_N * 8));
  }
  pp_sort_kernel<<<N, 1024, PP_SORT_N * 8, st>>>(L, seg_box, seg_score, seg_label, seg_count, s_box, s_score, s_label, s_orig, n_det);
  SCAN_LAUNCH_CHECK("pp_sort_kernel");
  pp_mask_kernel<<<dim3(PP_WORDS, N), 256, 0, st>>>(s_box, s_label, n_det, nms_thresh, mask);
  SCAN_LAUNCH_CHECK("pp_mask_kernel");
  pp_reduce_kernel<<<N, 1024, PP_SORT_N, st>>>(s_box, s_score, s_label, s_orig, n_det, mask, post_top_n, out_box, out_score, out_label, out_count);
  SCAN_LAUNCH_CHECK("pp_reduce_kernel");
  return SCAN_OK;
}
