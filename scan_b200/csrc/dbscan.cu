// K2: DBSCAN target-node sampling.
// Reference: modeling/rpn/fcos/loss.py:397-423 (DBSCAN_batch_cpu) calling sklearn.cluster.DBSCAN(eps, n_jobs=-1)
// (un-pinned third-party dependency; sklearn 1.9.0 semantics restated in oracle/condgraph_oracle.py:dbscan_labels):
// brute-force radius neighbours with float64 distance evaluation |x|^2+|y|^2-2x.y (clamped at 0) <= eps^2,
// neighbourhood includes the point itself, core <=> >= min_samples neighbours, clusters grown in index order.
// Deterministic restatement used here: components of the core-core graph by lock-free union-find (root = smallest
// index), cluster id = rank of the root, border point -> smallest cluster id among its core neighbours, else -1.
//
// Kernels: ordered compaction of the selected (n, cls, y, x) entries; point build (feature row x activation);
// 64x64 pairwise-distance tiles in shared memory (fp32 FFMA Gram, fp64 re-evaluation inside a band around eps^2)
// producing a bit adjacency matrix; popcount neighbour counts; warp-per-row union-find; root ranking; labelling.
#include <stdlib.h>

#include "dbscan_common.cuh"

namespace scan {

constexpr int DB_SB = 1024;

struct DbWs {
  int* block_cnt;      // [n_blocks]
  int* sel_flat;       // [cap] flat entry index ((n*CLS+cls)*H+y)*W+x
  float* sel_act;      // [cap]
  float* points;       // [cap, dim]
  float* sq;           // [cap]
  uint32_t* adj;       // [cap, wpr]
  int* count;          // [cap]
  int* parent;         // [cap]
  int* cid;            // [cap] cluster id of roots / scan buffer
  int* block_cnt2;     // [cap/1024 + 1]
  unsigned long long* re_list;   // [DB_RE_CAP] deferred exact re-evaluations: (i << 32) | j
  long long wpr;       // words per adjacency row
};

static inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

static long long dbws_layout(long long cap, long long n_entries, int dim, char* base, DbWs* ws) {
  long long off = 0;
  auto take = [&](long long bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const long long nb = (n_entries + DB_SB - 1) / DB_SB + 1;
  const long long wpr = align_up((cap + 31) / 32, 4);  // whole 128-column tiles
  DbWs w;
  w.block_cnt = (int*)take(nb * 4);
  w.sel_flat = (int*)take(cap * 4);
  w.sel_act = (float*)take(cap * 4);
  w.points = (float*)take(cap * dim * 4);
  w.sq = (float*)take(cap * 4);
  w.adj = (uint32_t*)take(cap * wpr * 4);
  w.count = (int*)take(cap * 4);
  w.parent = (int*)take(cap * 4);
  w.cid = (int*)take(cap * 4);
  w.block_cnt2 = (int*)take((cap / DB_SB + 2) * 4);
  w.re_list = (unsigned long long*)take((long long)DB_RE_CAP * 8);
  w.wpr = wpr;
  if (ws) *ws = w;
  return off;
}

// info: [0] n_points [1] n_clusters [2] n_noise [3] skipped [4] error [5] n_recheck [6] any_nonzero [7] scratch
// ---------------------------------------------------------------------------- selection (ordered compaction)
__global__ void __launch_bounds__(DB_SB) db_count_kernel(const float* __restrict__ act, int n_images, int K, int hw, float thr,
                                                         int* __restrict__ block_cnt) {
  __shared__ int wc[32];
  const long long E = (long long)n_images * (K - 1) * hw;
  const long long e = (long long)blockIdx.x * DB_SB + threadIdx.x;
  bool sel = false;
  if (e < E) {
    const long long n = e / ((long long)(K - 1) * hw);
    const long long rem = e - n * (K - 1) * hw;
    sel = act[n * K * hw + hw + rem] > thr;  // channel 0 is background: act[:, 1:]
  }
  const unsigned b = __ballot_sync(0xffffffffu, sel);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int i = 0; i < 32; ++i) s += wc[i];
    block_cnt[blockIdx.x] = s;
  }
}

// single block exclusive scan of cnt[0..n) in place; total -> *total_out (optionally clamped reporting)
__global__ void __launch_bounds__(1024) db_scan_kernel(int* __restrict__ cnt, int n, int* __restrict__ total_out) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int b = base + threadIdx.x;
    const int v = b < n ? cnt[b] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      int w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const int excl = carry + warp_tot[wid] + incl - v;
    if (b < n) cnt[b] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(DB_SB) db_fill_kernel(const float* __restrict__ act, int n_images, int K, int hw, float thr,
                                                        const int* __restrict__ block_excl, int cap, int* info,
                                                        int* __restrict__ sel_flat, float* __restrict__ sel_act) {
  __shared__ int wc[32];
  if (info[0] > cap) {  // capacity exceeded: flag and drop everything
    if (blockIdx.x == 0 && threadIdx.x == 0) info[4] = 1;
    return;
  }
  const long long E = (long long)n_images * (K - 1) * hw;
  const long long e = (long long)blockIdx.x * DB_SB + threadIdx.x;
  bool sel = false;
  float a = 0.f;
  if (e < E) {
    const long long n = e / ((long long)(K - 1) * hw);
    const long long rem = e - n * (K - 1) * hw;
    a = act[n * K * hw + hw + rem];
    sel = a > thr;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, sel);
  if (lane == 0) wc[wid] = __popc(b);
  __syncthreads();
  if (sel) {
    int before = 0;
    for (int i = 0; i < wid; ++i) before += wc[i];
    const int slot = block_excl[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
    sel_flat[slot] = (int)e;
    sel_act[slot] = a;
  }
}

// one warp per point: p = fl32(row * act); also |p|^2 in fp32 and the any-nonzero flag (loss.py:415)
__global__ void __launch_bounds__(256) db_points_kernel(const float* __restrict__ rows_level, int K, int hw, int dim,
                                                        const int* info, const int* __restrict__ sel_flat,
                                                        const float* __restrict__ sel_act, float* __restrict__ points,
                                                        float* __restrict__ sq, int* info_w) {
  const int n = info[4] ? 0 : info[0];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const long long e = sel_flat[i];
    const long long img = e / ((long long)(K - 1) * hw);
    const long long pix = e % hw;
    const float a = sel_act[i];
    const float4* src = reinterpret_cast<const float4*>(rows_level + (img * hw + pix) * dim);
    float4* dst = reinterpret_cast<float4*>(points + (long long)i * dim);
    float s = 0.f;
    bool nz = false;
    for (int j = lane; j < dim / 4; j += 32) {
      float4 v = __ldg(src + j);
      v.x = __fmul_rn(v.x, a); v.y = __fmul_rn(v.y, a); v.z = __fmul_rn(v.z, a); v.w = __fmul_rn(v.w, a);
      dst[j] = v;
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      nz |= (v.x != 0.f) | (v.y != 0.f) | (v.z != 0.f) | (v.w != 0.f);
    }
    s = warp_sum(s);
    if (lane == 0) sq[i] = s;
    if (__any_sync(0xffffffffu, nz) && lane == 0) atomicOr(info_w + 6, 1);
  }
}

// squared norms only (stand-alone clustering entry point)
__global__ void __launch_bounds__(256) db_sqnorm_kernel(const float* __restrict__ points, int n, int dim, float* __restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const float4* src = reinterpret_cast<const float4*>(points + (long long)i * dim);
    float s = 0.f;
    for (int j = lane; j < dim / 4; j += 32) {
      const float4 v = __ldg(src + j);
      s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    s = warp_sum(s);
    if (lane == 0) sq[i] = s;
  }
}

// ---------------------------------------------------------------------------- pairwise distances -> adjacency bits
constexpr int DT = 64;    // tile edge
constexpr int DKC = 32;   // reduction chunk
constexpr int DLD = 36;   // padded leading dimension (floats)

__global__ void __launch_bounds__(256) db_adj_kernel(const float* __restrict__ points, const float* __restrict__ sq, const int* info,
                                                     int n_fixed, int dim, float eps2f, double eps2, long long wpr,
                                                     uint32_t* __restrict__ adj, int* info_w) {
  __shared__ __align__(16) float As[DT * DLD];
  __shared__ __align__(16) float Bs[DT * DLD];
  __shared__ uint8_t flags[DT][DT + 4];
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int nt = (n + DT - 1) / DT;
  const long long total = (long long)nt * nt;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int bi = (int)(t / nt), bj = (int)(t % nt);
    const int i0 = bi * DT, j0 = bj * DT;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
    for (int k0 = 0; k0 < dim; k0 += DKC) {
      __syncthreads();
      for (int i = threadIdx.x; i < DT * (DKC / 4); i += 256) {
        const int r = i >> 3, c4 = i & 7;
        float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
        if (i0 + r < n) va = __ldg(reinterpret_cast<const float4*>(points + (long long)(i0 + r) * dim + k0) + c4);
        if (j0 + r < n) vb = __ldg(reinterpret_cast<const float4*>(points + (long long)(j0 + r) * dim + k0) + c4);
        *reinterpret_cast<float4*>(As + r * DLD + c4 * 4) = va;
        *reinterpret_cast<float4*>(Bs + r * DLD + c4 * 4) = vb;
      }
      __syncthreads();
#pragma unroll
      for (int d = 0; d < DKC; d += 4) {
        float4 av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(As + (ty * 4 + a) * DLD + d);
#pragma unroll
        for (int c = 0; c < 4; ++c) bv[c] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * c) * DLD + d);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[a][c] = fmaf(av[a].x, bv[c].x, acc[a][c]);
            acc[a][c] = fmaf(av[a].y, bv[c].y, acc[a][c]);
            acc[a][c] = fmaf(av[a].z, bv[c].z, acc[a][c]);
            acc[a][c] = fmaf(av[a].w, bv[c].w, acc[a][c]);
          }
      }
    }
    int n_re = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ty * 4 + a;
      const float si = i < n ? __ldg(sq + i) : 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx + 16 * c;
        bool within = false;
        if (i < n && j < n) {
          if (i == j) {
            within = true;
          } else {
            const float sj = __ldg(sq + j);
            const float d2 = si + sj - 2.f * acc[a][c];
            const float tol = 1e-4f * (si + sj + eps2f);
            if (fabsf(d2 - eps2f) <= tol) {
              within = db_exact_within(points + (long long)i * dim, points + (long long)j * dim, dim, eps2);
              ++n_re;
            } else {
              within = d2 < eps2f;
            }
          }
        }
        flags[ty * 4 + a][tx + 16 * c] = within ? 1 : 0;
      }
    }
    if (n_re) atomicAdd(info_w + 5, n_re);
    __syncthreads();
    if (threadIdx.x < 128) {
      const int r = threadIdx.x >> 1, h = threadIdx.x & 1;
      if (i0 + r < n) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) word |= (uint32_t)flags[r][h * 32 + b] << b;
        adj[(long long)(i0 + r) * wpr + (j0 >> 5) + h] = word;
      }
    }
  }
}

// ---------------------------------------------------------------------------- counts, union-find, labels
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;  // path halving (benign race: only ever points closer to the root)
    x = p;
    p = gp;
  }
  return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  int ra = uf_find(parent, a), rb = uf_find(parent, b);
  while (ra != rb) {
    if (ra < rb) { const int t = ra; ra = rb; rb = t; }  // ra > rb: hang the larger root below the smaller
    const int old = atomicCAS(parent + ra, ra, rb);
    if (old == ra) return;
    ra = uf_find(parent, old);
    rb = uf_find(parent, rb);
  }
}

// warp per row: neighbour count (incl. self) -> core flag in count's sign convention: count[i] = #neighbours
__global__ void __launch_bounds__(256) db_count_rows_kernel(const uint32_t* __restrict__ adj, const int* info, int n_fixed,
                                                            long long wpr, int* __restrict__ count, int* __restrict__ parent) {
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nw = (n + 31) >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    int c = 0;
    for (int w = lane; w < nw; w += 32) c += __popc(__ldg(adj + (long long)i * wpr + w));
    c = (int)warp_sum((float)c);
    if (lane == 0) {
      count[i] = c;
      parent[i] = i;
    }
  }
}

// Connected components of the core-core graph in three cheap passes instead of one union per edge (a single dense
// cluster of n points has n^2/2 edges):
//   1. every core point hooks onto its SMALLEST-index core neighbour (one union per point, early exit);
//   2. flatten (parent[i] = root);
//   3. every core point re-scans its neighbours j < i and unions only where the flattened roots differ -- one
//      coalesced-ish load of parent[j] per edge, real unions are rare after pass 1.
__global__ void __launch_bounds__(256) db_union_min_kernel(const uint32_t* __restrict__ adj, const int* info, int n_fixed, long long wpr,
                                                           int min_samples, const int* __restrict__ count, int* __restrict__ parent) {
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    if (count[i] < min_samples) continue;
    const int nw = (i >> 5) + 1;
    int found = 0x7fffffff;
    for (int w0 = 0; w0 < nw && found == 0x7fffffff; w0 += 32) {
      const int w = w0 + lane;
      int cand = 0x7fffffff;
      if (w < nw) {
        uint32_t bits = __ldg(adj + (long long)i * wpr + w);
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int j = (w << 5) + b;
          if (j < i && count[j] >= min_samples) { cand = j; break; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
      found = cand;
    }
    if (lane == 0 && found != 0x7fffffff) uf_union(parent, i, found);
  }
}

__global__ void __launch_bounds__(256) db_flatten_kernel(const int* info, int n_fixed, int* __restrict__ parent) {
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = i;
    while (true) {
      const int p = parent[r];
      if (p == r) break;
      r = p;
    }
    parent[i] = r;  // only ever written with an ancestor: safe against concurrent readers
  }
}

// Dense regimes (the trained-like maps of the benchmark: ~40 k points, one giant cluster) have ~n^2/2 set adjacency bits and
// nearly all of them join two points that the min-hook pass has already put into the same set.  Testing each bit costs two
// dependent random loads (count[j], parent[j]), so the pass below first filters whole 32-bit words against a mask of the
// points that can be skipped for rows of the DOMINANT set: non-core points and core points whose (flattened) root is the
// dominant root.  Two points with equal roots after the flatten stay in one set for ever, so the filter never drops a
// necessary union; a poor choice of the dominant root only costs speed.
constexpr int DB_PICK = 256;    // sampled core points that vote for the dominant root
__global__ void __launch_bounds__(DB_PICK) db_pick_root_kernel(const int* info, int n_fixed, int min_samples, const int* __restrict__ count,
                                                               const int* __restrict__ parent, int* __restrict__ rstar) {
  __shared__ int cand[DB_PICK];
  __shared__ int best_v[DB_PICK / 32], best_c[DB_PICK / 32];
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int t = threadIdx.x;
  const long long idx = n >= DB_PICK ? (long long)t * n / DB_PICK : (t < n ? t : -1);
  int c = -1;
  if (idx >= 0 && count[idx] >= min_samples) c = parent[idx];
  cand[t] = c;
  __syncthreads();
  int v = 0;
  if (c >= 0)
    for (int k = 0; k < DB_PICK; ++k) v += (cand[k] == c);
  // block argmax of (votes, candidate)
  int bv = v, bc = c;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ov = __shfl_xor_sync(0xffffffffu, bv, o), oc = __shfl_xor_sync(0xffffffffu, bc, o);
    if (ov > bv || (ov == bv && oc > bc)) { bv = ov; bc = oc; }
  }
  if ((t & 31) == 0) { best_v[t >> 5] = bv; best_c[t >> 5] = bc; }
  __syncthreads();
  if (t == 0) {
    for (int k = 1; k < DB_PICK / 32; ++k)
      if (best_v[k] > bv || (best_v[k] == bv && best_c[k] > bc)) { bv = best_v[k]; bc = best_c[k]; }
    rstar[0] = bv > 0 ? bc : -1;
  }
}

__global__ void __launch_bounds__(256) db_dom_mask_kernel(const int* info, int n_fixed, int cap, int min_samples,
                                                          const int* __restrict__ count, const int* __restrict__ parent,
                                                          const int* __restrict__ rstar, uint32_t* __restrict__ dmask) {
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int rs = rstar[0];
  const int words = (cap + 31) / 32;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < words * 32; j += gridDim.x * blockDim.x) {
    const bool skip = j >= n || count[j] < min_samples || parent[j] == rs;
    const unsigned b = __ballot_sync(0xffffffffu, skip);
    if ((threadIdx.x & 31) == 0) dmask[j >> 5] = b;
  }
}

__global__ void __launch_bounds__(256) db_union_rest_kernel(const uint32_t* __restrict__ adj, const int* info, int n_fixed, long long wpr,
                                                            int min_samples, const int* __restrict__ count, int* parent,
                                                            const int* __restrict__ rstar, const uint32_t* __restrict__ dmask) {
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int rs = rstar[0];
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    if (count[i] < min_samples) continue;
    const int ri = parent[i];  // flattened root at the start of this pass
    const bool dom = ri == rs;
    const int nw = (i >> 5) + 1;
    for (int w = lane; w < nw; w += 32) {
      uint32_t bits = __ldg(adj + (long long)i * wpr + w);
      if (dom) bits &= ~__ldg(dmask + w);
      while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int j = (w << 5) + b;
        if (j < i && count[j] >= min_samples && parent[j] != ri) uf_union(parent, i, j);
      }
    }
  }
}

// flatten + root flags: cid[i] = 1 if i is a core root else 0 (then scanned)
__global__ void __launch_bounds__(DB_SB) db_roots_kernel(const int* info, int n_fixed, int min_samples, const int* __restrict__ count,
                                                         int* __restrict__ parent, int* __restrict__ cid, int* __restrict__ block_cnt) {
  __shared__ int wc[32];
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int i = blockIdx.x * DB_SB + threadIdx.x;
  bool root = false;
  if (i < n && count[i] >= min_samples) {
    const int r = uf_find(parent, i);
    root = (r == i);
  }
  if (i < n) cid[i] = root ? 1 : 0;
  const unsigned b = __ballot_sync(0xffffffffu, root);
  if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int k = 0; k < 32; ++k) s += wc[k];
    block_cnt[blockIdx.x] = s;
  }
}

// cid[i] (root flag) -> rank of the root among roots in index order
__global__ void __launch_bounds__(DB_SB) db_rank_kernel(const int* info, int n_fixed, const int* __restrict__ block_excl,
                                                        int* __restrict__ cid) {
  __shared__ int wc[32];
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int i = blockIdx.x * DB_SB + threadIdx.x;
  const bool root = (i < n) && cid[i] == 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, root);
  if (lane == 0) wc[wid] = __popc(b);
  __syncthreads();
  if (i < n) {
    int before = 0;
    for (int k = 0; k < wid; ++k) before += wc[k];
    cid[i] = root ? block_excl[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u)) : -1;
  }
}

// warp per point: sklearn label
__global__ void __launch_bounds__(256) db_label_kernel(const uint32_t* __restrict__ adj, const int* info, int n_fixed, long long wpr,
                                                       int min_samples, const int* __restrict__ count, int* __restrict__ parent,
                                                       const int* __restrict__ cid, int* __restrict__ labels, int* info_w) {
  const int n = n_fixed >= 0 ? n_fixed : (info[4] ? 0 : info[0]);
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int nw = (n + 31) >> 5;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    int lab;
    if (count[i] >= min_samples) {
      lab = cid[uf_find(parent, i)];
    } else {
      int best = 0x7fffffff;
      for (int w = lane; w < nw; w += 32) {
        uint32_t bits = __ldg(adj + (long long)i * wpr + w);
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int j = (w << 5) + b;
          if (j < n && count[j] >= min_samples) best = min(best, cid[uf_find(parent, j)]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
      lab = best == 0x7fffffff ? -1 : best;
    }
    if (lane == 0) {
      labels[i] = lab;
      if (lab < 0) atomicAdd(info_w + 2, 1);
    }
  }
}

// location mask (loss.py:417-421) and pseudo labels (loss.py:500)
__global__ void __launch_bounds__(256) db_mask_kernel(const int* info, const int* __restrict__ sel_flat, const int* __restrict__ labels,
                                                      int K, int hw, uint8_t* __restrict__ pos_mask) {
  const int n = info[4] ? 0 : info[0];
  const bool skipped = info[6] == 0;  // all selected points are exactly zero: clustering skipped, all entries stay 1
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int lab = skipped ? -1 : labels[i];
    if (lab != 0) {  // noise (-1 -> 1) and clusters >= 1 count; cluster 0 is dropped
      const long long e = sel_flat[i];
      const long long img = e / ((long long)(K - 1) * hw);
      pos_mask[img * hw + e % hw] = 1;
    }
  }
}

__global__ void __launch_bounds__(256) db_plabel_kernel(const float* __restrict__ act, int n_images, int K, int hw, int64_t* __restrict__ plabel) {
  const long long total = (long long)n_images * hw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / hw, p = i - n * hw;
    const float* a = act + n * K * hw + p;
    float best = a[hw];
    int bi = 1;
    for (int c = 2; c < K; ++c) {
      const float v = a[(long long)c * hw];
      if (v > best) { best = v; bi = c; }
    }
    plabel[i] = bi;  // argmax over act[:,1:] + 1
  }
}

__global__ void db_finish_kernel(int* info, const int* n_roots, int standalone) {
  info[1] = *n_roots;
  info[3] = (!standalone && info[6] == 0) ? 1 : 0;
}

static int cluster_points(const DbWs& ws, const float* points, const float* sq, const int* info, int n_fixed, int cap, int dim, double eps,
                          int min_samples, int* labels, int* info_w, cudaStream_t st) {
  const int sms = sm_count();
  const float eps2f = (float)(eps * eps);
  static const int simt = getenv("SCAN_B200_DBSCAN_SIMT") ? atoi(getenv("SCAN_B200_DBSCAN_SIMT")) : 0;
  if (simt) {  // fp32 FFMA verification kernel (bring-up / tests only)
    db_adj_kernel<<<4 * sms, 256, 0, st>>>(points, sq, info, n_fixed, dim, eps2f, eps * eps, ws.wpr, ws.adj, info_w);
    SCAN_LAUNCH_CHECK("db_adj_kernel");
  } else {
    int rc = launch_db_adj_tc(points, sq, info, n_fixed, cap, dim, eps2f, eps * eps, ws.wpr, ws.adj, info_w, ws.re_list, st);
    if (rc == SCAN_ENOTSUP) {  // point width other than 256: FFMA tiles
      db_adj_kernel<<<4 * sms, 256, 0, st>>>(points, sq, info, n_fixed, dim, eps2f, eps * eps, ws.wpr, ws.adj, info_w);
      SCAN_LAUNCH_CHECK("db_adj_kernel");
      rc = SCAN_OK;
    }
    if (rc) return rc;
  }
  db_count_rows_kernel<<<4 * sms, 256, 0, st>>>(ws.adj, info, n_fixed, ws.wpr, ws.count, ws.parent);
  SCAN_LAUNCH_CHECK("db_count_rows_kernel");
  db_union_min_kernel<<<4 * sms, 256, 0, st>>>(ws.adj, info, n_fixed, ws.wpr, min_samples, ws.count, ws.parent);
  SCAN_LAUNCH_CHECK("db_union_min_kernel");
  db_flatten_kernel<<<2 * sms, 256, 0, st>>>(info, n_fixed, ws.parent);
  SCAN_LAUNCH_CHECK("db_flatten_kernel");
  // scratch: the cluster-id array is not in use before db_roots_kernel; word 0 of block_cnt2 holds the dominant root
  uint32_t* dmask = (uint32_t*)ws.cid;
  db_pick_root_kernel<<<1, DB_PICK, 0, st>>>(info, n_fixed, min_samples, ws.count, ws.parent, ws.block_cnt2);
  SCAN_LAUNCH_CHECK("db_pick_root_kernel");
  db_dom_mask_kernel<<<2 * sms, 256, 0, st>>>(info, n_fixed, cap, min_samples, ws.count, ws.parent, ws.block_cnt2, dmask);
  SCAN_LAUNCH_CHECK("db_dom_mask_kernel");
  db_union_rest_kernel<<<4 * sms, 256, 0, st>>>(ws.adj, info, n_fixed, ws.wpr, min_samples, ws.count, ws.parent, ws.block_cnt2, dmask);
  SCAN_LAUNCH_CHECK("db_union_rest_kernel");
  const int nb = (cap + DB_SB - 1) / DB_SB;
  db_roots_kernel<<<nb, DB_SB, 0, st>>>(info, n_fixed, min_samples, ws.count, ws.parent, ws.cid, ws.block_cnt2);
  SCAN_LAUNCH_CHECK("db_roots_kernel");
  db_scan_kernel<<<1, 1024, 0, st>>>(ws.block_cnt2, nb, info_w + 7);
  SCAN_LAUNCH_CHECK("db_scan_kernel");
  db_rank_kernel<<<nb, DB_SB, 0, st>>>(info, n_fixed, ws.block_cnt2, ws.cid);
  SCAN_LAUNCH_CHECK("db_rank_kernel");
  db_label_kernel<<<4 * sms, 256, 0, st>>>(ws.adj, info, n_fixed, ws.wpr, min_samples, ws.count, ws.parent, ws.cid, labels, info_w);
  SCAN_LAUNCH_CHECK("db_label_kernel");
  return SCAN_OK;
}

}  // namespace scan

extern "C" int64_t scan_dbscan_workspace_bytes(int64_t cap_points) {
  // entries bound for the block counters: a level never has more than 2^31 entries; size for cap-independent part generously
  return scan::dbws_layout(cap_points, (long long)1 << 26, 256, nullptr, nullptr);
}

extern "C" int scan_dbscan_level(const float* rows_level, const float* act_nchw, int32_t n_images, int32_t num_classes, int32_t h,
                                 int32_t w, float thr, double eps, int32_t min_samples, int32_t cap_points, uint8_t* pos_mask,
                                 int64_t* plabel, int32_t* point_labels, int32_t* info, void* workspace, int64_t workspace_bytes,
                                 void* stream) {
  using namespace scan;
  if (!rows_level || !act_nchw || !pos_mask || !plabel || !point_labels || !info || !workspace) return SCAN_EINVAL;
  if (n_images < 1 || num_classes < 2 || h < 1 || w < 1 || cap_points < 1 || min_samples < 1) return SCAN_EINVAL;
  const int hw = h * w;
  const long long E = (long long)n_images * (num_classes - 1) * hw;
  if (E > (1ll << 26)) return SCAN_ENOTSUP;  // block counters are sized for 2^26 entries per level
  DbWs ws;
  const long long need = dbws_layout(cap_points, (long long)1 << 26, 256, (char*)workspace, &ws);
  if (workspace_bytes < need) return SCAN_ECAPACITY;
  cudaStream_t st = (cudaStream_t)stream;
  SCAN_CUDA_CHECK(cudaMemsetAsync(info, 0, 8 * sizeof(int32_t), st));
  SCAN_CUDA_CHECK(cudaMemsetAsync(pos_mask, 0, (size_t)n_images * hw, st));
  const int nb = (int)ceil_div(E, DB_SB);
  const int sms = sm_count();
  db_count_kernel<<<nb, DB_SB, 0, st>>>(act_nchw, n_images, num_classes, hw, thr, ws.block_cnt);
  SCAN_LAUNCH_CHECK("db_count_kernel");
  db_scan_kernel<<<1, 1024, 0, st>>>(ws.block_cnt, nb, info);
  SCAN_LAUNCH_CHECK("db_scan_kernel");
  db_fill_kernel<<<nb, DB_SB, 0, st>>>(act_nchw, n_images, num_classes, hw, thr, ws.block_cnt, cap_points, info, ws.sel_flat, ws.sel_act);
  SCAN_LAUNCH_CHECK("db_fill_kernel");
  db_points_kernel<<<4 * sms, 256, 0, st>>>(rows_level, num_classes, hw, 256, info, ws.sel_flat, ws.sel_act, ws.points, ws.sq, info);
  SCAN_LAUNCH_CHECK("db_points_kernel");
  int rc = cluster_points(ws, ws.points, ws.sq, info, -1, cap_points, 256, eps, min_samples, point_labels, info, st);
  if (rc) return rc;
  db_mask_kernel<<<2 * sms, 256, 0, st>>>(info, ws.sel_flat, point_labels, num_classes, hw, pos_mask);
  SCAN_LAUNCH_CHECK("db_mask_kernel");
  db_plabel_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)n_images * hw, 256), 8ll * sms), 256, 0, st>>>(act_nchw, n_images,
                                                                                                                 num_classes, hw, plabel);
  SCAN_LAUNCH_CHECK("db_plabel_kernel");
  db_finish_kernel<<<1, 1, 0, st>>>(info, info + 7, 0);
  SCAN_LAUNCH_CHECK("db_finish_kernel");
  return SCAN_OK;
}

extern "C" int scan_dbscan_points(const float* points, int32_t n, int32_t dim, double eps, int32_t min_samples, int32_t* labels,
                                  int32_t* info, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace scan;
  if (!info) return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  SCAN_CUDA_CHECK(cudaMemsetAsync(info, 0, 8 * sizeof(int32_t), st));
  if (n == 0) return SCAN_OK;
  if (!points || !labels || !workspace || n < 0 || dim < 4 || dim % DKC || min_samples < 1) return SCAN_EINVAL;
  DbWs ws;
  const long long need = dbws_layout(n, (long long)1 << 26, 256, (char*)workspace, &ws);
  if (workspace_bytes < need) return SCAN_ECAPACITY;
  db_sqnorm_kernel<<<4 * sm_count(), 256, 0, st>>>(points, n, dim, ws.sq);
  SCAN_LAUNCH_CHECK("db_sqnorm_kernel");
  int rc = cluster_points(ws, points, ws.sq, info, n, n, dim, eps, min_samples, labels, info, st);
  if (rc) return rc;
  db_finish_kernel<<<1, 1, 0, st>>>(info, info + 7, 1);
  SCAN_LAUNCH_CHECK("db_finish_kernel");
  return SCAN_OK;
}
