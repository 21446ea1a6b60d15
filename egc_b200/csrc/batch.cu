// Mini-batch plumbing either side of the layer, sm_100a:
//   * collation of many small graphs into one block-diagonal graph - what PyG's DataLoader / Batch.from_data_list does
//     on the CPU for the reference (experiments/zinc/configs.py:36-45,60-67, experiments/cifar/configs.py:42-53):
//     node ids of graph g are shifted by the number of nodes before it, `batch[i]` = graph of node i;
//   * graph readout - global_mean_pool / global_add_pool / global_max_pool (experiments/zinc/models.py:46-53,73) - as a
//     segmented reduction over the contiguous node ranges of a collated batch, with its backward.
// These steps are launch-latency bound at the reference's batch sizes (128 graphs, 3 k - 15 k nodes): one small
// kernel each, no atomics, deterministic.
#include "common.cuh"

namespace egc {

// first g in [0, n] with ptr[g + 1] > i, i.e. the segment that holds element i (ptr non-decreasing, ptr[0] = 0)
__device__ __forceinline__ int segment_of(const int32_t* __restrict__ ptr, int n_seg, int64_t i) {
  int lo = 0, hi = n_seg;                       // invariant: ptr[lo] <= i, answer in [lo, hi)
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(ptr + mid) <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_collate_edges(const int64_t* __restrict__ src_local, const int64_t* __restrict__ dst_local,
                                const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ node_ptr, int n_graphs,
                                int64_t n_edges, int64_t* __restrict__ src_out, int64_t* __restrict__ dst_out,
                                int32_t* __restrict__ flags) {
  const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int g = segment_of(edge_ptr, n_graphs, e);
  const int64_t off = __ldg(node_ptr + g), cnt = __ldg(node_ptr + g + 1) - off;
  const int64_t s = src_local[e], d = dst_local[e];
  if (s < 0 || d < 0 || s >= cnt || d >= cnt) atomicOr(flags, 1);      // local id outside its graph
  src_out[e] = s + off;
  dst_out[e] = d + off;
}

__global__ void k_batch_vector(const int32_t* __restrict__ node_ptr, int n_graphs, int64_t n_nodes, int64_t* __restrict__ batch) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  batch[i] = segment_of(node_ptr, n_graphs, i);
}

// ptr[g] = number of nodes with batch id < g (batch sorted ascending); flags bit 0: unsorted, bit 1: id out of range
__global__ void k_segment_ptr(const int64_t* __restrict__ batch, int64_t n, int n_graphs, int32_t* __restrict__ ptr,
                              int32_t* __restrict__ flags) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int64_t prev = i == 0 ? -1 : batch[i - 1];
  const int64_t cur = i == n ? n_graphs : batch[i];
  if (i < n && (cur < 0 || cur >= n_graphs)) { atomicOr(flags, 2); return; }
  if (prev < -1 || prev >= n_graphs) return;                 // flagged by the thread that owns element i - 1
  if (cur < prev) { atomicOr(flags, 1); return; }
  for (int64_t g = prev + 1; g <= cur; ++g) ptr[g] = static_cast<int32_t>(i);    // boundaries (empty graphs included)
}

// mode 0 sum, 1 mean (count clamped to 1), 2 max (empty segment -> 0, arg = first winner, -1 when empty)
template <int MODE>
__global__ void k_segment_pool_fwd(const float* __restrict__ x, const int32_t* __restrict__ ptr, int f,
                                   float* __restrict__ out, int32_t* __restrict__ arg) {
  const int g = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= f) return;
  const int b = __ldg(ptr + g), e = __ldg(ptr + g + 1);
  float acc = MODE == 2 ? -INFINITY : 0.f;
  int best = -1;
  for (int i = b; i < e; ++i) {
    const float v = __ldg(x + static_cast<int64_t>(i) * f + c);
    if (MODE == 2) { if (v > acc || best < 0) { acc = v; best = i; } }
    else acc += v;
  }
  if (MODE == 1) acc = acc / static_cast<float>(max(e - b, 1));
  if (MODE == 2 && best < 0) acc = 0.f;
  out[static_cast<int64_t>(g) * f + c] = acc;
  if (MODE == 2 && arg != nullptr) arg[static_cast<int64_t>(g) * f + c] = best;
}

template <int MODE>
__global__ void k_segment_pool_bwd(const float* __restrict__ d_out, const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ arg, int f, float* __restrict__ d_x) {
  const int g = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= f) return;
  const int b = __ldg(ptr + g), e = __ldg(ptr + g + 1);
  float go = __ldg(d_out + static_cast<int64_t>(g) * f + c);
  if (MODE == 1) go = go / static_cast<float>(max(e - b, 1));
  const int best = MODE == 2 ? __ldg(arg + static_cast<int64_t>(g) * f + c) : -1;
  for (int i = b; i < e; ++i)
    d_x[static_cast<int64_t>(i) * f + c] = MODE == 2 ? (i == best ? go : 0.f) : go;
}

}  // namespace egc

using namespace egc;

extern "C" {

int egc_collate_edges(const int64_t* src_local, const int64_t* dst_local, const int32_t* edge_ptr, const int32_t* node_ptr,
                      int32_t n_graphs, int64_t n_edges, int64_t n_nodes, int64_t* src_out, int64_t* dst_out,
                      int64_t* batch_out, int32_t* flags, void* stream) {
  EGC_REQUIRE(n_graphs >= 0 && n_edges >= 0 && n_nodes >= 0, "egc_collate_edges: negative size");
  EGC_REQUIRE(edge_ptr && node_ptr && flags, "egc_collate_edges: null pointer");
  EGC_REQUIRE(n_edges == 0 || (src_local && dst_local && src_out && dst_out), "egc_collate_edges: null edge arrays");
  cudaStream_t st = as_stream(stream);
  EGC_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
  if (n_edges > 0) {
    LaunchScope ls("k_collate_edges", st);
    k_collate_edges<<<ceil_div(n_edges, 256), 256, 0, st>>>(src_local, dst_local, edge_ptr, node_ptr, n_graphs, n_edges,
                                                             src_out, dst_out, flags);
  }
  EGC_LAUNCH_CHECK("k_collate_edges");
  if (batch_out != nullptr && n_nodes > 0) {
    LaunchScope ls("k_batch_vector", st);
    k_batch_vector<<<ceil_div(n_nodes, 256), 256, 0, st>>>(node_ptr, n_graphs, n_nodes, batch_out);
  }
  EGC_LAUNCH_CHECK("k_batch_vector");
  return EGC_OK;
}

int egc_segment_ptr(const int64_t* batch, int64_t n, int32_t n_graphs, int32_t* ptr, int32_t* flags, void* stream) {
  EGC_REQUIRE(n >= 0 && n_graphs >= 0 && n < (int64_t{1} << 31), "egc_segment_ptr: bad sizes");
  EGC_REQUIRE(ptr && flags && (n == 0 || batch), "egc_segment_ptr: null pointer");
  cudaStream_t st = as_stream(stream);
  EGC_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
  {
    LaunchScope ls("k_segment_ptr", st);
    k_segment_ptr<<<ceil_div(n + 1, 256), 256, 0, st>>>(batch, n, n_graphs, ptr, flags);
  }
  EGC_LAUNCH_CHECK("k_segment_ptr");
  return EGC_OK;
}

int egc_segment_pool_fwd(const float* x, const int32_t* ptr, int32_t n_graphs, int32_t f, int32_t mode, float* out,
                         int32_t* arg, void* stream) {
  EGC_REQUIRE(n_graphs >= 0 && f >= 0 && mode >= 0 && mode <= 2, "egc_segment_pool_fwd: bad arguments");
  if (n_graphs == 0 || f == 0) return EGC_OK;
  EGC_REQUIRE(x && ptr && out, "egc_segment_pool_fwd: null pointer");
  cudaStream_t st = as_stream(stream);
  const dim3 grid(n_graphs, ceil_div(f, 128));
  {
    LaunchScope ls("k_segment_pool_fwd", st);
    if (mode == 0) k_segment_pool_fwd<0><<<grid, 128, 0, st>>>(x, ptr, f, out, arg);
    else if (mode == 1) k_segment_pool_fwd<1><<<grid, 128, 0, st>>>(x, ptr, f, out, arg);
    else k_segment_pool_fwd<2><<<grid, 128, 0, st>>>(x, ptr, f, out, arg);
  }
  EGC_LAUNCH_CHECK("k_segment_pool_fwd");
  return EGC_OK;
}

int egc_segment_pool_bwd(const float* d_out, const int32_t* ptr, const int32_t* arg, int32_t n_graphs, int32_t f,
                         int32_t mode, float* d_x, void* stream) {
  EGC_REQUIRE(n_graphs >= 0 && f >= 0 && mode >= 0 && mode <= 2, "egc_segment_pool_bwd: bad arguments");
  if (n_graphs == 0 || f == 0) return EGC_OK;
  EGC_REQUIRE(d_out && ptr && d_x && (mode != 2 || arg), "egc_segment_pool_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  const dim3 grid(n_graphs, ceil_div(f, 128));
  {
    LaunchScope ls("k_segment_pool_bwd", st);
    if (mode == 0) k_segment_pool_bwd<0><<<grid, 128, 0, st>>>(d_out, ptr, arg, f, d_x);
    else if (mode == 1) k_segment_pool_bwd<1><<<grid, 128, 0, st>>>(d_out, ptr, arg, f, d_x);
    else k_segment_pool_bwd<2><<<grid, 128, 0, st>>>(d_out, ptr, arg, f, d_x);
  }
  EGC_LAUNCH_CHECK("k_segment_pool_bwd");
  return EGC_OK;
}

}  // extern "C"
