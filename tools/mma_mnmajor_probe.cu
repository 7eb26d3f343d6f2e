// Probe: does tcgen05.mma kind::tf32 accept an MN-major B operand (128-byte swizzle)?  If it does, the attention backward can
// feed K / Q / dO row tiles to both GEMM orientations and drop the transposed operand planes.
// D[128 x 64] = A[128 x 32] . B[32 x 64], A K-major, B stored "rows = k, 64 n-values per row" (MN-major), integer-valued
// inputs so tf32 is exact.  Prints the number of mismatching entries against a host reference.
// RESULT (B200, round 1): variant 3 is bit-exact -- layout type 1 (SWIZZLE_128B_BASE32B), Swizzle<2,5,2>, LBO = n-block stride,
// SBO = 512; variants 1, 2 (plain SWIZZLE_128B), 4, 5 and 6, 7 (no swizzle) are wrong.  Usage: mma_mnmajor_probe.bin [variant]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I scan_b200/csrc tools/mma_mnmajor_probe.cu \
//        scan_b200/csrc/build/core.o -o tools/mma_mnmajor_probe.bin -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"
using namespace scan;

constexpr int PM = 128, PN = 64, PK = 32;

__device__ __forceinline__ uint32_t swz128(uint32_t byte_off) { return byte_off ^ (((byte_off >> 7) & 7u) << 4); }

// SWIZZLE_128B descriptors.  K-major: LBO unused, SBO = 1024 (8-row groups).  MN-major: LBO = byte stride between 32-element
// blocks along N, SBO = byte stride between groups of 8 k-rows.
__device__ __forceinline__ uint64_t desc_any(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout_type << 61);
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) { return desc_any(saddr, lbo, sbo, 2); }
// CUTLASS UMMA::LayoutType::SWIZZLE_128B_BASE32B = Swizzle<2,5,2>: 32-byte chunks, bits [5,7) ^= bits [7,9)
__device__ __forceinline__ uint32_t swz128_base32(uint32_t byte_off) { return byte_off ^ (((byte_off >> 7) & 3u) << 5); }

// MN-major variants under test (B[k][n], 64 n-values per k-row):
//  1: SW128,         rows of 128 B, LBO = n-block stride 4096, SBO = k-group stride 1024
//  2: SW128,         LBO / SBO swapped
//  3: SW128_BASE32B, LBO 4096, SBO 512 (4-row k atoms)      4: SW128_BASE32B, LBO 4096, SBO 1024
//  5: SW128_BASE32B, LBO 512,  SBO 4096                     6: no swizzle (interleave): 8 k-rows x 16 B core matrices,
//                                                              n-cores 128 B apart (SBO), k-groups 2048 B apart (LBO)
//  7: as 6 with LBO / SBO swapped
struct Variant { uint32_t type, lbo, sbo, kstep; int swz, arrange; };
__host__ __device__ inline Variant variant_of(int v) {
  switch (v) {
    case 1: return {2, 4096, 1024, 1024, 1, 0};
    case 2: return {2, 1024, 4096, 1024, 1, 0};
    case 3: return {1, 4096, 512, 1024, 2, 0};
    case 4: return {1, 4096, 1024, 1024, 2, 0};
    case 5: return {1, 512, 4096, 1024, 2, 0};
    case 6: return {0, 2048, 128, 2048, 0, 1};
    default: return {0, 128, 2048, 2048, 0, 1};
  }
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d,
                                                    int b_mn_major) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a_s = smem;            // 128 rows x 128 B, K-major SW128
  uint8_t* b_s = smem + 16384;    // MN-major: [n_blk (2)][k_grp (4)][8 k-rows x 128 B]; K-major: 64 rows (n) x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < PM * PK; i += blockDim.x) {
    const int r = i / PK, c = i % PK;
    *(float*)(a_s + swz128(r * 128 + c * 4)) = a[i];
  }
  for (int i = threadIdx.x; i < PK * PN; i += blockDim.x) {
    const int k = i / PN, n = i % PN;
    uint32_t off;
    if (b_mn_major) {
      const Variant vt = variant_of(b_mn_major);
      if (vt.arrange == 0) {
        const uint32_t in_atom = (k % 8) * 128 + (n % 32) * 4;
        off = (n / 32) * 4096 + (k / 8) * 1024 + (vt.swz == 1 ? swz128(in_atom) : swz128_base32(in_atom));
      } else {
        off = (k / 8) * 2048 + (n / 4) * 128 + (k % 8) * 16 + (n % 4) * 4;     // 16 n-cores of 8 x 16 B per k-group
      }
    } else
      off = swz128(n * 128 + k * 4);     // K-major control: row = n, 32 k-values per 128-byte row
    *(float*)(b_s + off) = b[i];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32) {
    // instruction descriptor: bit 16 = B is MN-major
    const uint32_t idesc = umma_idesc_tf32(PM, PN) | (b_mn_major ? (1u << 16) : 0u);
    for (int ks = 0; ks < PK / 8; ++ks) {
      const uint64_t da = desc_sw128(smem_u32(a_s) + ks * 32, 16, 1024);
      const Variant vt = variant_of(b_mn_major);
      const uint64_t db = b_mn_major ? desc_any(smem_u32(b_s) + ks * vt.kstep, vt.lbo, vt.sbo, vt.type)
                                     : desc_sw128(smem_u32(b_s) + ks * 32, 16, 1024);
      if (elect_one_sync()) umma_tf32(tm, da, db, idesc, ks != 0);
    }
    if (elect_one_sync()) umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tcgen05_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < PN; c0 += 16) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(tm + ((uint32_t)(warp * 32) << 16) + c0)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) d[(warp * 32 + lane) * PN + c0 + i] = __uint_as_float(r[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64));
  }
}

int main(int argc, char** argv) {
  std::vector<float> a(PM * PK), b(PK * PN), want(PM * PN, 0.f), got(PM * PN);
  srand(7);
  for (auto& x : a) x = (float)(rand() % 9 - 4);
  for (auto& x : b) x = (float)(rand() % 9 - 4);
  for (int m = 0; m < PM; ++m)
    for (int n = 0; n < PN; ++n) {
      float s = 0.f;
      for (int k = 0; k < PK; ++k) s += a[m * PK + k] * b[k * PN + n];
      want[m * PN + n] = s;
    }
  float *da, *db, *dd;
  cudaMalloc(&da, a.size() * 4);
  cudaMalloc(&db, b.size() * 4);
  cudaMalloc(&dd, got.size() * 4);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  for (int mn = 0; mn < 8; ++mn) {
    if (only >= 0 && mn != only) continue;
    cudaMemset(dd, 0, got.size() * 4);
    probe_kernel<<<1, 128, 40 * 1024>>>(da, db, dd, mn);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("variant %d: CUDA error %s\n", mn, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(got.data(), dd, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t i = 0; i < got.size(); ++i) bad += got[i] != want[i];
    printf("variant %d (%s): %d / %d entries differ from the host reference\n", mn, mn ? "MN-major B" : "K-major control", bad, PM * PN);
  }
  return 0;
}
