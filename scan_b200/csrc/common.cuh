// Shared helpers of the scan_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/scan_b200.h"

namespace scan {

// Device-side copy of scan_levels_t plus derived row offsets; passed by value to kernels.
struct Levels {
  int n_levels;
  int n_images;
  int h[SCAN_MAX_LEVELS];
  int w[SCAN_MAX_LEVELS];
  int stride[SCAN_MAX_LEVELS];
  long long row_off[SCAN_MAX_LEVELS + 1];  // row_off[l] = N * sum_{j<l} H_j W_j ; row_off[n_levels] = R
};

inline int make_levels(const scan_levels_t* in, Levels* out) {
  if (!in || in->n_levels < 1 || in->n_levels > SCAN_MAX_LEVELS || in->n_images < 1) return SCAN_EINVAL;
  out->n_levels = in->n_levels;
  out->n_images = in->n_images;
  long long off = 0;
  for (int l = 0; l < SCAN_MAX_LEVELS; ++l) {
    if (l < in->n_levels) {
      if (in->h[l] < 1 || in->w[l] < 1) return SCAN_EINVAL;
      out->h[l] = in->h[l];
      out->w[l] = in->w[l];
      out->stride[l] = in->stride[l];
      out->row_off[l] = off;
      off += (long long)in->n_images * in->h[l] * in->w[l];
    } else {
      out->h[l] = out->w[l] = out->stride[l] = 0;
      out->row_off[l] = off;
    }
  }
  out->row_off[SCAN_MAX_LEVELS] = off;
  for (int l = in->n_levels; l <= SCAN_MAX_LEVELS; ++l) out->row_off[l] = off;
  if (off >= (1ll << 31)) return SCAN_EINVAL;  // rows are indexed with int32
  return SCAN_OK;
}

__device__ __forceinline__ int level_of_row(const Levels& lv, long long g) {
  int l = 0;
#pragma unroll
  for (int j = 1; j < SCAN_MAX_LEVELS; ++j)
    if (j < lv.n_levels && g >= lv.row_off[j]) l = j;
  return l;
}

void set_cuda_error(cudaError_t e, const char* where);
int sm_count();
// true the first time it is called on the current device for this mask (per-device one-shot work such as
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), which is a per-device setting)
inline bool first_use_on_device(unsigned long long* mask) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

#define SCAN_CUDA_CHECK(expr)                                  \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) {                                   \
      scan::set_cuda_error(_e, #expr);                         \
      return SCAN_ECUDA;                                       \
    }                                                          \
  } while (0)

#define SCAN_LAUNCH_CHECK(name)                                \
  do {                                                         \
    cudaError_t _e = cudaGetLastError();                       \
    if (_e != cudaSuccess) {                                   \
      scan::set_cuda_error(_e, name);                          \
      return SCAN_ECUDA;                                       \
    }                                                          \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

static inline long long ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// Counter-based Bernoulli source of the attention dropout mask: 32-bit hash of (seed, chunk, query, key).  All attention
// kernels (forward and backward, every implementation) must produce THIS value so that masks agree.  Callers compare it
// with a threshold, i.e. they consume the HIGH bits: multiply - xorshift - multiply leaves those depending on every
// input bit, and a trailing xorshift (which only touches low bits) is omitted on purpose.
//   h = mix(pre(seed, chunk) ^ i * ATTN_DROP_CI ^ j * ATTN_DROP_CJ)
// The tensor-core kernels keep the loop-invariant part in a register and step the other by a constant add, which leaves
// 6 integer ops per element.
constexpr uint32_t ATTN_DROP_CI = 0xC2B2AE35u;
constexpr uint32_t ATTN_DROP_CJ = 0x27D4EB2Fu;
__device__ __forceinline__ uint32_t attn_drop_pre(uint64_t seed, uint32_t chunk) {
  return (uint32_t)seed ^ ((uint32_t)(seed >> 32) * 0x9E3779B9u) ^ (chunk * 0x85EBCA6Bu);
}
__device__ __forceinline__ uint32_t attn_drop_mix(uint32_t h) {
  h *= 0x7FEB352Du;
  h ^= h >> 15;
  h *= 0x846CA68Bu;
  return h;
}
__device__ __forceinline__ uint32_t attn_drop_hash(uint64_t seed, uint32_t chunk, uint32_t i, uint32_t j) {
  return attn_drop_mix(attn_drop_pre(seed, chunk) ^ (i * ATTN_DROP_CI) ^ (j * ATTN_DROP_CJ));
}

}  // namespace scan
