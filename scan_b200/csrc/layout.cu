// NCHW <-> rows transposition (the reference's features[l].permute(0,2,3,1).reshape(-1,C), loss.py:440).
// HBM-bound: every element is read once and written once; 64x64 fp32 tiles staged in shared memory so both
// the global read (along pixels) and the global write (along channels) are 256-byte coalesced rows.
#include "common.cuh"

namespace scan {

struct PackArgs {
  const float* nchw[SCAN_MAX_LEVELS];
  long long tile_off[SCAN_MAX_LEVELS + 1];  // first tile index of each level
  int ptiles[SCAN_MAX_LEVELS];              // pixel tiles per image of the level
  const float* rows_lv[SCAN_MAX_LEVELS];    // unpack only: per-level [N*H*W, C] inputs (null: one rows matrix in `rows_in`)
};

constexpr int TP = 64;  // pixels per tile
constexpr int TC = 64;  // channels per tile

template <bool kUnpack, bool kAccumulate>
__global__ void __launch_bounds__(256) pack_kernel(Levels lv, PackArgs args, int channels, float* rows_out,
                                                   const float* rows_in, float* const* /*unused*/) {
  __shared__ float tile[TC][TP + 1];
  long long t = blockIdx.x;
  int l = 0;
#pragma unroll
  for (int j = 1; j < SCAN_MAX_LEVELS; ++j)
    if (j < lv.n_levels && t >= args.tile_off[j]) l = j;
  t -= args.tile_off[l];
  const int ctiles = channels / TC;
  const int hw = lv.h[l] * lv.w[l];
  const int ct = (int)(t % ctiles);
  long long r = t / ctiles;
  const int pt = (int)(r % args.ptiles[l]);
  const int n = (int)(r / args.ptiles[l]);
  const int p0 = pt * TP, c0 = ct * TC;
  float* nchw = const_cast<float*>(args.nchw[l]) + ((long long)n * channels + c0) * hw;
  const long long row0 = lv.row_off[l] + (long long)n * hw + p0;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  if (!kUnpack) {
#pragma unroll 4
    for (int c = ty; c < TC; c += 4) {
      int p = p0 + tx;
      tile[c][tx] = (p < hw) ? __ldg(nchw + (long long)c * hw + p) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = ty; pp < TP; pp += 4) {
      if (p0 + pp < hw) rows_out[(row0 + pp) * channels + c0 + tx] = tile[tx][pp];
    }
  } else {
    // separate per-level inputs (the gradients cuDNN hands back level by level): rebase so that the rows index still applies
    const float* src = args.rows_lv[l] ? args.rows_lv[l] - lv.row_off[l] * channels : rows_in;
#pragma unroll 4
    for (int pp = ty; pp < TP; pp += 4) {
      tile[tx][pp] = (p0 + pp < hw) ? __ldg(src + (row0 + pp) * channels + c0 + tx) : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = ty; c < TC; c += 4) {
      int p = p0 + tx;
      if (p < hw) {
        float* dst = nchw + (long long)c * hw + p;
        if (kAccumulate) *dst += tile[c][tx]; else *dst = tile[c][tx];
      }
    }
  }
}

static int build_args(const Levels& lv, const void* const* ptrs, int channels, PackArgs* a, long long* total) {
  if (channels % TC != 0) return SCAN_EINVAL;
  long long off = 0;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    a->tile_off[l] = off;
    if (l < lv.n_levels) {
      if (!ptrs[l]) return SCAN_EINVAL;
      a->nchw[l] = (const float*)ptrs[l];
      a->rows_lv[l] = nullptr;
      a->ptiles[l] = (lv.h[l] * lv.w[l] + TP - 1) / TP;
      off += (long long)lv.n_images * a->ptiles[l] * (channels / TC);
    } else {
      a->nchw[l] = nullptr;
      a->rows_lv[l] = nullptr;
      a->ptiles[l] = 1;
    }
  }
  a->tile_off[SCAN_MAX_LEVELS] = off;
  *total = off;
  return SCAN_OK;
}

}  // namespace scan

extern "C" int scan_pack_rows(const scan_levels_t* lvh, const void* const* nchw_host, int32_t channels,
                              float* rows, void* stream) {
  scan::Levels lv;
  int rc = scan::make_levels(lvh, &lv);
  if (rc) return rc;
  if (!nchw_host || !rows) return SCAN_EINVAL;
  scan::PackArgs a;
  long long total;
  rc = scan::build_args(lv, nchw_host, channels, &a, &total);
  if (rc) return rc;
  scan::pack_kernel<false, false><<<(unsigned)total, 256, 0, (cudaStream_t)stream>>>(lv, a, channels, rows, nullptr, nullptr);
  SCAN_LAUNCH_CHECK("pack_kernel");
  return SCAN_OK;
}

extern "C" int scan_unpack_rows(const scan_levels_t* lvh, const float* rows, int32_t channels,
                                void* const* nchw_host, int32_t accumulate, void* stream) {
  scan::Levels lv;
  int rc = scan::make_levels(lvh, &lv);
  if (rc) return rc;
  if (!nchw_host || !rows) return SCAN_EINVAL;
  scan::PackArgs a;
  long long total;
  rc = scan::build_args(lv, (const void* const*)nchw_host, channels, &a, &total);
  if (rc) return rc;
  if (accumulate)
    scan::pack_kernel<true, true><<<(unsigned)total, 256, 0, (cudaStream_t)stream>>>(lv, a, channels, nullptr, rows, nullptr);
  else
    scan::pack_kernel<true, false><<<(unsigned)total, 256, 0, (cudaStream_t)stream>>>(lv, a, channels, nullptr, rows, nullptr);
  SCAN_LAUNCH_CHECK("unpack_kernel");
  return SCAN_OK;
}

// the same with one NHWC-dense [N*H_l*W_l, C] input per level (HOST array of device pointers): one launch for all levels
extern "C" int scan_unpack_levels(const scan_levels_t* lvh, const void* const* rows_levels_host, int32_t channels,
                                  void* const* nchw_host, void* stream) {
  scan::Levels lv;
  int rc = scan::make_levels(lvh, &lv);
  if (rc) return rc;
  if (!nchw_host || !rows_levels_host) return SCAN_EINVAL;
  scan::PackArgs a;
  long long total;
  rc = scan::build_args(lv, (const void* const*)nchw_host, channels, &a, &total);
  if (rc) return rc;
  for (int l = 0; l < lv.n_levels; ++l) {
    if (!rows_levels_host[l]) return SCAN_EINVAL;
    a.rows_lv[l] = (const float*)rows_levels_host[l];
  }
  scan::pack_kernel<true, false><<<(unsigned)total, 256, 0, (cudaStream_t)stream>>>(lv, a, channels, nullptr, nullptr, nullptr);
  SCAN_LAUNCH_CHECK("unpack_kernel");
  return SCAN_OK;
}
