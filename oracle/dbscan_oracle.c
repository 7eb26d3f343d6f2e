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
  /* symmetric, cache-blocked pair loop: (BI x BJ) blocks with bj <= bi; a pair is evaluated once and both adjacency
   * halves are set.  2 x 4 register tile of float64 dot products, 4 lanes each (the summation order differs from BLAS'
   * dgemm in the last float64 bits only: |error| ~ 1e-13 relative, far below any float32 input spacing around eps^2).
   * Bits are set with atomic ORs (the mirrored half lands in other threads' rows). */
  enum { BI = 32, BJ = 32 };
  typedef double v4d __attribute__((vector_size(32), aligned(8)));
  const long nbi = (n + BI - 1) / BI;
#pragma omp parallel for schedule(dynamic, 1)
  for (long bi = nbi - 1; bi >= 0; --bi) {
    const long i0 = bi * BI;
    for (long j0 = 0; j0 < i0 + BI && j0 < n; j0 += BJ) {
      for (long i = i0; i < i0 + BI && i < n; i += 2) {
        const double* a0 = xd + i * d;
        const double* a1 = xd + (i + 1 < n ? i + 1 : i) * d;
        for (long j = j0; j < j0 + BJ && j < n && j <= i + 1; j += 4) {
          const double* bp[4];
          for (int t = 0; t < 4; ++t) bp[t] = xd + (j + t < n ? j + t : j) * d;
          v4d acc[2][4];
          for (int r = 0; r < 2; ++r)
            for (int t = 0; t < 4; ++t) acc[r][t] = (v4d){0.0, 0.0, 0.0, 0.0};
          long k = 0;
          for (; k + 4 <= d; k += 4) {
            const v4d va0 = *(const v4d*)(a0 + k), va1 = *(const v4d*)(a1 + k);
            for (int t = 0; t < 4; ++t) {
              const v4d vb = *(const v4d*)(bp[t] + k);
              acc[0][t] += va0 * vb;
              acc[1][t] += va1 * vb;
            }
          }
          for (int r = 0; r < 2; ++r) {
            const long ii = i + r;
            if (ii >= n) continue;
            const double* ar = r ? a1 : a0;
            for (int t = 0; t < 4; ++t) {
              const long jj = j + t;
              if (jj >= n || jj > ii) continue;
              double dot = (acc[r][t][0] + acc[r][t][1]) + (acc[r][t][2] + acc[r][t][3]);
              for (long kk = k; kk < d; ++kk) dot += ar[kk] * bp[t][kk];
              double d2 = sq[ii] + sq[jj] - 2.0 * dot;
              if (d2 < 0.0) d2 = 0.0;
              if (d2 <= eps2 || ii == jj) {
#pragma omp atomic
                adj[ii * wpr + (jj >> 6)] |= (uint64_t)1 << (jj & 63);
                if (jj != ii) {
#pragma omp atomic
                  adj[jj * wpr + (ii >> 6)] |= (uint64_t)1 << (ii & 63);
                }
              }
            }
          }
        }
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    long c = 0;
    for (long w = 0; w < wpr; ++w) c += __builtin_popcountll(adj[i * wpr + w]);
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
