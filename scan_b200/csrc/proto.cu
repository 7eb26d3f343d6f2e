// K3b: per-class prototype sums and the incremental paradigm EMA.
// Reference: modeling/rpn/fcos/condgraph.py:395-398 (class means of the aggregated nodes) and :558-617
// (update_prototype, update_prototype_nx1, update_prototype_nx1_rnn).
// The packed [K, C+1] sum|count buffer is what the multi-GPU prototype all-reduce operates on (§8e).
#include "common.cuh"

namespace scan {

// Each block owns a contiguous chunk of nodes; thread t owns channel t (C <= 1024) of every class
// accumulator, so the shared-memory accumulation is conflict- and atomic-free; one global atomicAdd per
// (class, channel) per block at the end.
__global__ void class_sums_kernel(const float* __restrict__ nodes, const int64_t* __restrict__ labels, int m, int c,
                                  int k, int shift, int nodes_per_block, float* __restrict__ packed) {
  extern __shared__ float acc[];  // [k][c] then counts [k]
  float* cnt = acc + k * c;
  for (int i = threadIdx.x; i < k * c + k; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int i0 = blockIdx.x * nodes_per_block;
  const int i1 = min(m, i0 + nodes_per_block);
  for (int i = i0; i < i1; ++i) {
    const int cls = (int)(labels[i] - shift);
    if (cls < 0 || cls >= k) continue;
    for (int j = threadIdx.x; j < c; j += blockDim.x) acc[cls * c + j] += nodes[(long long)i * c + j];
    if (threadIdx.x == 0) cnt[cls] += 1.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < k * c; i += blockDim.x) {
    const float v = acc[i];
    if (v != 0.f) atomicAdd(packed + (i / c) * (c + 1) + (i % c), v);
  }
  if (threadIdx.x < k && cnt[threadIdx.x] != 0.f) atomicAdd(packed + threadIdx.x * (c + 1) + c, cnt[threadIdx.x]);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
  return t;
}

// one block per class
__global__ void __launch_bounds__(256) proto_update_kernel(const float* __restrict__ packed, int c, int p, int slot, int shift,
                                                           int cosine_on, float momentum, float* __restrict__ proto,
                                                           float* __restrict__ batch_out) {
  __shared__ float red[32];
  const int cls = blockIdx.x;
  const float count = packed[cls * (c + 1) + c];
  float* pr = proto + (long long)cls * c * p;
  float sb = 0.f, dot = 0.f, no = 0.f, nb = 0.f;
  for (int j = threadIdx.x; j < c; j += blockDim.x) {
    const float b = count > 0.f ? packed[cls * (c + 1) + j] / count : 0.f;
    batch_out[cls * c + j] = b;
    const float o = pr[j * p + slot];
    sb += b;
    dot += o * b;
    no += o * o;
    nb += b * b;
  }
  sb = block_sum(sb, red);
  const bool exist = sb != 0.f;  // prototype_batch.sum(-1).bool(), condgraph.py:560,573,589
  float m = momentum;
  if (cosine_on) {
    dot = block_sum(dot, red);
    no = block_sum(no, red);
    nb = block_sum(nb, red);
    // torch.cosine_similarity: x/max(|x|,eps) . y/max(|y|,eps), eps = 1e-8
    m = dot / (fmaxf(sqrtf(no), 1e-8f) * fmaxf(sqrtf(nb), 1e-8f));
  }
  __syncthreads();
  for (int j = threadIdx.x; j < c; j += blockDim.x) {
    if (shift)
      for (int i = 0; i < p - 1; ++i) pr[j * p + i] = pr[j * p + i + 1];  // condgraph.py:597-598, all classes
    if (exist) {
      const float b = batch_out[cls * c + j];
      const float o = pr[j * p + slot];
      pr[j * p + slot] = o * m + b * (1.f - m);
    }
  }
}

}  // namespace scan

extern "C" int scan_class_sums(const float* nodes, const int64_t* labels, int32_t m, int32_t channels, int32_t num_classes,
                               int32_t label_shift, float* packed_sums, void* stream) {
  if (!packed_sums || num_classes < 1 || num_classes > SCAN_MAX_CLASSES || channels < 1 || m < 0) return SCAN_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  SCAN_CUDA_CHECK(cudaMemsetAsync(packed_sums, 0, sizeof(float) * num_classes * (channels + 1), st));
  if (m == 0) return SCAN_OK;
  if (!nodes || !labels) return SCAN_EINVAL;
  const size_t smem = sizeof(float) * (num_classes * channels + num_classes);
  if (smem > 48 * 1024) return SCAN_ENOTSUP;
  int blocks = 2 * scan::sm_count();
  int npb = (m + blocks - 1) / blocks;
  if (npb < 8) npb = 8;
  blocks = (m + npb - 1) / npb;
  scan::class_sums_kernel<<<blocks, 256, smem, st>>>(nodes, labels, m, channels, num_classes, label_shift, npb, packed_sums);
  SCAN_LAUNCH_CHECK("class_sums_kernel");
  return SCAN_OK;
}

extern "C" int scan_proto_update(const float* packed_sums, int32_t num_classes, int32_t channels, int32_t proto_iter,
                                 int32_t slot, int32_t shift, int32_t cosine_on, float momentum, float* prototype,
                                 float* proto_batch_out, void* stream) {
  if (!packed_sums || !prototype || !proto_batch_out || num_classes < 1 || channels < 1 || proto_iter < 1 || slot < 0 ||
      slot >= proto_iter)
    return SCAN_EINVAL;
  scan::proto_update_kernel<<<num_classes, 256, 0, (cudaStream_t)stream>>>(packed_sums, channels, proto_iter, slot, shift,
                                                                        cosine_on, momentum, prototype, proto_batch_out);
  SCAN_LAUNCH_CHECK("proto_update_kernel");
  return SCAN_OK;
}
