// Pieces shared by the SIMT and the tensor-core DBSCAN distance kernels.
#pragma once
#include "common.cuh"

namespace scan {

// exact sklearn test: float64 accumulation of the fp32 inputs
static __device__ __noinline__ bool db_exact_within(const float* __restrict__ a, const float* __restrict__ b, int dim, double eps2) {
  double sa = 0.0, sb = 0.0, ab = 0.0;
  for (int d = 0; d < dim; ++d) {
    const double x = (double)__ldg(a + d), y = (double)__ldg(b + d);
    sa = fma(x, x, sa);
    sb = fma(y, y, sb);
    ab = fma(x, y, ab);
  }
  double d2 = sa + sb - 2.0 * ab;
  if (d2 < 0.0) d2 = 0.0;
  return d2 <= eps2;
}


// capacity of the deferred re-evaluation list of the tensor-core distance kernel (pairs inside the tf32 error band);
// pairs beyond it are re-evaluated inline
constexpr int DB_RE_CAP = 1 << 21;

int launch_db_adj_tc(const float* points, const float* sq, const int* info, int n_fixed, int cap, int dim, float eps2f, double eps2,
                     long long wpr, uint32_t* adj, int* info_w, unsigned long long* re_list, cudaStream_t st);

}  // namespace scan
