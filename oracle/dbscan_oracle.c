/* CPU ORACLE (test infrastructure, not product code): plain-C restatement of
 * sklearn.cluster.DBSCAN(eps, min_samples, metric='euclidean', algorithm='brute').fit_predict on float32 input,
 * the third-party call the reference makes at fcos_core/modeling/rpn/fcos/loss.py:416 (sklearn 1.9.0 in this image;
 * the reference does not pin a version).
 *
 * sklearn semantics restated (sklearn/neighbors/_base.py radius_neighbors -> EuclideanRadiusNeighbors32:
 * float64 accumulation of the float32 inputs, d2 = |x|^2 + |y|^2 - 2 x.y clamped at 0, neighbour iff d2 <= eps^2,
 * the point itself included; sklearn/cluster/_dbscan_inner.pyx:17-43: clusters are grown from core points in
 * index order, a non-core point joins the first cluster that reaches it):
 *   core_i   = #neighbours >= min_samples
 *   clusters = connected components of the core-core neighbour graph, numbered by their smallest member index
 *   border   = smallest cluster number among the point's core neighbours; noise = -1
 * O(n^2 d) time, O(n^2 / 8) memory (bit adjacency), OpenMP over rows.
 * Pinned against sklearn itself by tests/test_oracle_vs_reference.py::test_dbscan_c_oracle_matches_sklearn.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static long uf_find(long* parent, long a) {
  while (parent[a] != a) {
    parent[a] = parent[parent[a]];
    a = parent[a];
  }
  return a;
}

int dbscan_oracle(const float* x, long n, long d, double eps, int min_samples, int* labels) {
  if (n <= 0) return 0;
  const double eps2 = eps * eps;
  const long wpr = (n + 63) / 64;
  double* xd = (double*)malloc(sizeof(double) * (size_t)n * (size_t)d);
  double* sq = (double*)malloc(sizeof(double) * (size_t)n);
  uint64_t* adj = (uint64_t*)calloc((size_t)n * (size_t)wpr, sizeof(uint64_t));
  long* count = (long*)calloc((size_t)n, sizeof(long));
  long* parent = (long*)malloc(sizeof(long) * (size_t)n);
  int* cid = (int*)malloc(sizeof(int) * (size_t)n);
  if (!xd || !sq || !adj || !count || !parent || !cid) return -1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    double s = 0.0;
    for (long k = 0; k < d; ++k) {
      const double v = (double)x[i * d + k];
      xd[i * d + k] = v;
      s += v * v;
    }
    sq[i] = s;
    parent[i] = i;
  }
#pragma omp parallel for schedule(dynamic, 8)
  for (long i = 0; i < n; ++i) {
    const double* a = xd + i * d;
    long c = 0;
    for (long j = 0; j < n; ++j) {
      const double* b = xd + j * d;
      double dot = 0.0;
#pragma omp simd reduction(+ : dot)
      for (long k = 0; k < d; ++k) dot += a[k] * b[k];
      double d2 = sq[i] + sq[j] - 2.0 * dot;
      if (d2 < 0.0) d2 = 0.0;
      if (d2 <= eps2 || i == j) {
        adj[i * wpr + (j >> 6)] |= (uint64_t)1 << (j & 63);
        ++c;
      }
    }
    count[i] = c;
  }
  for (long i = 0; i < n; ++i) {
    if (count[i] < min_samples) continue;
    for (long w = 0; w <= (i >> 6); ++w) {
      uint64_t bits = adj[i * wpr + w];
      while (bits) {
        const long j = (w << 6) + __builtin_ctzll(bits);
        bits &= bits - 1;
        if (j < i && count[j] >= min_samples) {
          long ra = uf_find(parent, i), rb = uf_find(parent, j);
          if (ra != rb) {
            if (ra < rb) parent[rb] = ra; else parent[ra] = rb;
          }
        }
      }
    }
  }
  int next = 0;
  for (long i = 0; i < n; ++i) {
    cid[i] = -1;
    if (count[i] >= min_samples && uf_find(parent, i) == i) cid[i] = next++;  /* roots in index order */
  }
  for (long i = 0; i < n; ++i) {
    if (count[i] >= min_samples) {
      labels[i] = cid[uf_find(parent, i)];
      continue;
    }
    int best = -1;
    for (long w = 0; w < wpr; ++w) {
      uint64_t bits = adj[i * wpr + w];
      while (bits) {
        const long j = (w << 6) + __builtin_ctzll(bits);
        bits &= bits - 1;
        if (count[j] >= min_samples) {
          const int c = cid[uf_find(parent, j)];
          if (best < 0 || c < best) best = c;
        }
      }
    }
    labels[i] = best;
  }
  free(xd); free(sq); free(adj); free(count); free(parent); free(cid);
  return 0;
}
