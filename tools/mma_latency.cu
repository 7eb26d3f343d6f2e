// Micro-benchmark: issue-to-retire time of tcgen05.mma kind::tf32 chains (dependent vs. interleaved accumulators, SS vs TS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scan_b200/csrc tools/mma_latency.cu -o gpurun_out/mma_latency
#include <cstdio>

#include "tc_common.cuh"
using namespace scan;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

struct Case {
  int n_mma, n_chains, n, ts, col_stride;
};

template <int NMMA, int CHAINS, int N, int TS, int STRIDE>
__device__ __forceinline__ long long run_case(uint32_t tm, uint32_t a_s, uint32_t b_s, uint32_t bar, uint32_t& phase) {
  constexpr uint32_t idesc = umma_idesc_tf32(128, N);
  long long dt = 0;
  for (int rep = 0; rep < 3; ++rep) {
    const long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < NMMA; ++i) {
      const uint32_t d = tm + 256 + (i % CHAINS) * STRIDE;
      const uint64_t bd = umma_desc_sw128(b_s + (i & 3) * 32);
      if (TS)
        mma_ts(d, tm + (i & 7) * 8, bd, idesc, 1);
      else
        umma_tf32(d, umma_desc_sw128(a_s + (i & 3) * 32), bd, idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, phase);
    phase ^= 1;
    dt = clock64() - t0;
  }
  return dt;
}

#define CASES(X)                                                                                                          \
  X(1, 1, 128, 0, 128) X(32, 1, 128, 0, 128) X(32, 2, 128, 0, 128) X(32, 2, 128, 0, 64) X(1, 1, 64, 0, 64) X(32, 1, 64, 0, 64)  \
  X(32, 2, 64, 0, 64) X(32, 4, 64, 0, 64) X(32, 2, 64, 0, 32) X(1, 1, 32, 0, 32) X(32, 1, 32, 0, 32) X(32, 2, 32, 0, 32)        \
  X(32, 4, 32, 0, 32) X(32, 8, 32, 0, 32) X(1, 1, 128, 1, 128) X(32, 1, 128, 1, 128) X(32, 2, 128, 1, 128) X(32, 2, 128, 1, 64) \
  X(1, 1, 64, 1, 64) X(32, 1, 64, 1, 64) X(32, 2, 64, 1, 64) X(32, 4, 64, 1, 64) X(32, 2, 64, 1, 32) X(1, 1, 32, 1, 32)         \
  X(32, 1, 32, 1, 32) X(32, 2, 32, 1, 32) X(32, 4, 32, 1, 32) X(32, 8, 32, 1, 32) X(1, 1, 16, 1, 16) X(32, 1, 16, 1, 16)        \
  X(32, 4, 16, 1, 16) X(64, 1, 128, 0, 128) X(64, 2, 128, 0, 128)

__global__ void __launch_bounds__(128) lat_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((float*)smem)[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 16384);
    int c = 0;
#define X(a, b, n, t, s) out[c++] = run_case<a, b, n, t, s>(tm, a_s, b_s, smem_u32(&bar), phase);
    CASES(X)
#undef X
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
  }
}

int main() {
  Case h[64];
  int n = 0;
#define X(a, b, nn, t, s) h[n++] = {a, b, nn, t, s};
  CASES(X)
#undef X
  long long* o;
  cudaMalloc(&o, 64 * 8);
  cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  lat_kernel<<<1, 128, 64 * 1024>>>(o);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error %s\n", cudaGetErrorString(e));
    return 1;
  }
  long long r[64];
  cudaMemcpy(r, o, n * 8, cudaMemcpyDeviceToHost);
  for (int i = 0; i < n; ++i)
    printf("%s N=%3d mmas=%2d chains=%d stride=%3d : %6lld cycles (%.1f / mma)\n", h[i].ts ? "TS" : "SS", h[i].n, h[i].n_mma,
           h[i].n_chains, h[i].col_stride, r[i], (double)r[i] / h[i].n_mma);
  return 0;
}
