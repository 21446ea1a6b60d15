// Graph preparation kernels: edge_index -> target-major CSR with PyG self-loop semantics,
// torch_sparse fill_diag, gcn_norm weights, CSC transpose (csr2csc) and the long-row chunk plan.
// All integer work; results are bit-exact with the reference (see include/egc_b200.h for the
// reference call sites each entry point replaces).
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace egc {

#define EGC_META_ERRFLAGS 7  // bit0: node id out of range, bit1: CSR row not sorted by column

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// largest r with rowptr[r] <= e  (row that owns nnz position e)
__device__ __forceinline__ int row_of_nnz(const int32_t* __restrict__ rowptr, int n_rows, int e) {
  int lo = 0, hi = n_rows;  // invariant: rowptr[lo] <= e < rowptr[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void k_max_id(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                         int* __restrict__ max_id) {
  long long m = -1;
  for (int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; e < n_edges;
       e += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    long long s = src[e], d = dst[e];
    m = max(m, max(s, d));
  }
  int mi = static_cast<int>(min(m, static_cast<long long>(INT32_MAX)));
  for (int o = 16; o > 0; o >>= 1) mi = max(mi, __shfl_xor_sync(kFull, mi, o));
  if ((threadIdx.x & 31) == 0 && mi >= 0) atomicMax(max_id, mi);
}

// keys: target id (or n_nodes = "dropped"), vals: source id.  Loops are appended after the edges so a
// stable sort leaves them last inside every row, exactly like add_remaining_self_loops + scatter.
__global__ void k_make_edge_keys(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, int64_t n_edges,
                                 int n_nodes, int loops_mode, const int* __restrict__ max_id,
                                 int32_t* __restrict__ keys, int32_t* __restrict__ vals, int32_t* __restrict__ meta) {
  const int n_loops = loops_mode == EGC_LOOPS_ALL_NODES ? n_nodes
                    : loops_mode == EGC_LOOPS_UP_TO_MAX_ID ? min(*max_id + 1, n_nodes) : 0;
  const int64_t total = n_edges + (loops_mode != EGC_LOOPS_NONE ? n_nodes : 0);
  for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
       t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int key, val;
    if (t < n_edges) {
      long long s = src[t], d = dst[t];
      bool bad = s < 0 || d < 0 || s >= n_nodes || d >= n_nodes;
      if (bad) atomicOr(meta + EGC_META_ERRFLAGS, 1);
      bool drop = bad || (loops_mode != EGC_LOOPS_NONE && s == d);
      key = drop ? n_nodes : static_cast<int>(d);
      val = bad ? 0 : static_cast<int>(s);
    } else {
      int i = static_cast<int>(t - n_edges);
      key = i < n_loops ? i : n_nodes;
      val = i;
    }
    keys[t] = key;
    vals[t] = val;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) meta[EGC_META_N_LOOPS] = n_loops;
}

__global__ void k_ptr_from_sorted_keys(const int32_t* __restrict__ keys_sorted, int n_total, int n_rows,
                                       int32_t* __restrict__ ptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n_rows) ptr[i] = lower_bound_i32(keys_sorted, n_total, i);
}

// nnz, longest row, number of long rows and of chunks
__global__ void k_row_stats(const int32_t* __restrict__ ptr, int n_rows, int32_t* __restrict__ meta) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int deg = 0;
  if (i < n_rows) deg = ptr[i + 1] - ptr[i];
  int is_long = deg > EGC_CHUNK_EDGES ? 1 : 0;
  int chunks = is_long ? (deg + EGC_CHUNK_EDGES - 1) / EGC_CHUNK_EDGES : 0;
  int mx = deg;
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(kFull, mx, o));
    is_long += __shfl_xor_sync(kFull, is_long, o);
    chunks += __shfl_xor_sync(kFull, chunks, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (mx > 0) atomicMax(meta + EGC_META_MAX_DEG, mx);
    if (is_long) atomicAdd(meta + EGC_META_N_LONG, is_long);
    if (chunks) atomicAdd(meta + EGC_META_N_CHUNKS, chunks);
  }
  if (i == 0) meta[EGC_META_NNZ] = ptr[n_rows];
}

// ---------------------------------------------------------------------------------------------
// fill_diag on an int64 CSR (one warp per row)
// ---------------------------------------------------------------------------------------------
__global__ void k_filldiag_count(const int64_t* __restrict__ rowptr_in, const int64_t* __restrict__ col_in,
                                 int n_dst, int n_src, int n_diag, int fill, int32_t* __restrict__ cnt,
                                 int32_t* __restrict__ meta) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row > n_dst) return;
  if (row == n_dst) {
    if (lane == 0) cnt[row] = 0;
    return;
  }
  int64_t b = rowptr_in[row], e = rowptr_in[row + 1];
  int kept = 0;
  bool bad = false, unsorted = false;
  for (int64_t p = b + lane; p < e; p += 32) {
    long long c = col_in[p];
    if (c < 0 || c >= n_src) bad = true;
    if (fill && p > b && col_in[p - 1] > c) unsorted = true;   // order only matters for the diagonal insert
    kept += (!fill || c != row) ? 1 : 0;
  }
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(kFull, kept, o);
  if (lane == 0) cnt[row] = kept + ((fill && row < n_diag) ? 1 : 0);
  if (bad) atomicOr(meta + EGC_META_ERRFLAGS, 1);
  if (unsorted) atomicOr(meta + EGC_META_ERRFLAGS, 2);
}

__global__ void k_filldiag_fill(const int64_t* __restrict__ rowptr_in, const int64_t* __restrict__ col_in,
                                const float* __restrict__ val_in, int n_dst, int n_diag, int fill,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                float* __restrict__ val_out) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  const int64_t b = rowptr_in[row], e = rowptr_in[row + 1];
  const int base = rowptr[row];
  const bool insert = fill && row < n_diag;
  int kept_before = 0, n_less = 0;
  for (int64_t p0 = b; p0 < e; p0 += 32) {
    int64_t p = p0 + lane;
    bool valid = p < e;
    long long c = valid ? col_in[p] : 0;
    bool keep = valid && (!fill || c != row);
    bool less = keep && c < row;
    unsigned km = __ballot_sync(kFull, keep), lm = __ballot_sync(kFull, less);
    if (keep) {
      int rank = kept_before + __popc(km & ((1u << lane) - 1));
      int pos = base + rank + ((insert && c > row) ? 1 : 0);
      col[pos] = static_cast<int32_t>(c);
      if (val_out) val_out[pos] = val_in ? val_in[p] : 1.0f;
    }
    kept_before += __popc(km);
    n_less += __popc(lm);
  }
  if (insert && lane == 0) {
    col[base + n_less] = row;
    if (val_out) val_out[base + n_less] = 1.0f;
  }
}

// ---------------------------------------------------------------------------------------------
// gcn_norm: degree, deg^-1/2, per-nnz weight (one warp per row)
// ---------------------------------------------------------------------------------------------
__global__ void k_degree(const int32_t* __restrict__ rowptr, const float* __restrict__ value, int n_dst,
                         float* __restrict__ deg, float* __restrict__ dis) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  int b = rowptr[row], e = rowptr[row + 1];
  float d;
  if (value == nullptr) {
    d = static_cast<float>(e - b);   // sum of unit weights, exact
  } else {
    d = 0.f;
    for (int p = b + lane; p < e; p += 32) d += value[p];
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(kFull, d, o);
  }
  if (lane == 0) {
    // torch CPU pow(-0.5) == 1/sqrt (both correctly rounded here: -prec-div / -prec-sqrt defaults)
    float r = 1.0f / sqrtf(d);
    if (isinf(r)) r = 0.f;           // masked_fill(dis == inf, 0)
    if (deg) deg[row] = d;
    dis[row] = r;
  }
}

__global__ void k_symnorm_values(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 const float* __restrict__ value, const float* __restrict__ dis, int n_dst,
                                 float* __restrict__ val_sym) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= n_dst) return;
  int b = rowptr[row], e = rowptr[row + 1];
  float di = dis[row];
  for (int p = b + lane; p < e; p += 32) {
    float v = value ? value[p] : 1.0f;
    val_sym[p] = __fmul_rn(__fmul_rn(v, di), __ldg(dis + col[p]));
  }
}

// ---------------------------------------------------------------------------------------------
// transpose helpers
// ---------------------------------------------------------------------------------------------
__global__ void k_iota(int32_t* __restrict__ out, int n) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = t;
}

__global__ void k_rows_of_perm(const int32_t* __restrict__ rowptr, int n_rows, const int32_t* __restrict__ perm,
                               int n, int32_t* __restrict__ rowidx) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) rowidx[t] = row_of_nnz(rowptr, n_rows, perm[t]);
}

// ---------------------------------------------------------------------------------------------
// chunk plan
// ---------------------------------------------------------------------------------------------
__global__ void k_plan_flags(const int32_t* __restrict__ ptr, int n_rows, int32_t* __restrict__ is_long,
                             int32_t* __restrict__ n_chunks) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows) return;
  int deg = i < n_rows ? ptr[i + 1] - ptr[i] : 0;
  int l = deg > EGC_CHUNK_EDGES ? 1 : 0;
  is_long[i] = l;
  n_chunks[i] = l ? (deg + EGC_CHUNK_EDGES - 1) / EGC_CHUNK_EDGES : 0;
}

__global__ void k_plan_fill(const int32_t* __restrict__ ptr, int n_rows, const int32_t* __restrict__ long_off,
                            const int32_t* __restrict__ chunk_off, int32_t* __restrict__ long_rows,
                            int32_t* __restrict__ long_chunk_ptr, int32_t* __restrict__ chunk_row,
                            int32_t* __restrict__ chunk_begin) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_rows) return;
  if (i == n_rows) {
    long_chunk_ptr[long_off[i]] = chunk_off[i];   // sentinel = total number of chunks
    return;
  }
  int b = ptr[i], deg = ptr[i + 1] - b;
  if (deg <= EGC_CHUNK_EDGES) return;
  int li = long_off[i], c0 = chunk_off[i];
  long_rows[li] = i;
  long_chunk_ptr[li] = c0;
  int nc = (deg + EGC_CHUNK_EDGES - 1) / EGC_CHUNK_EDGES;
  for (int c = 0; c < nc; ++c) {
    chunk_row[c0 + c] = i;
    chunk_begin[c0 + c] = b + c * EGC_CHUNK_EDGES;
  }
}

static int key_bits(int n_keys_max) {  // keys in [0, n_keys_max]
  int bits = 1;
  while (bits < 31 && (1ll << bits) <= n_keys_max) ++bits;
  return bits;
}

static size_t radix_temp_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                  static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), n);
  return bytes;
}

static size_t scan_temp_bytes(int n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr), n);
  return bytes;
}

}  // namespace egc

using namespace egc;

extern "C" {

size_t egc_csr_from_edges_workspace_bytes(int64_t n_edges, int32_t n_nodes) {
  int64_t total = n_edges + n_nodes;
  if (total <= 0 || total >= INT32_MAX) return 0;
  size_t b = 0;
  b += 3 * align_up(static_cast<size_t>(total) * 4, 256);
  b += 256;  // max id
  b += align_up(radix_temp_bytes(static_cast<int>(total)), 256);
  return b + 256;
}

int egc_csr_from_edges(const int64_t* src, const int64_t* dst, int64_t n_edges, int32_t n_nodes, int32_t loops,
                       int32_t* rowptr, int32_t* col, int32_t* meta, void* workspace, size_t workspace_bytes,
                       void* stream) {
  EGC_REQUIRE(n_edges >= 0 && n_nodes > 0, "egc_csr_from_edges: n_edges=%lld n_nodes=%d", (long long)n_edges, n_nodes);
  EGC_REQUIRE(loops >= 0 && loops <= 2, "egc_csr_from_edges: loops mode %d", loops);
  EGC_REQUIRE(n_edges + n_nodes < INT32_MAX, "egc_csr_from_edges: graph too large for int32 indices");
  EGC_REQUIRE(rowptr && col && meta && workspace && (n_edges == 0 || (src && dst)), "egc_csr_from_edges: null pointer");
  EGC_REQUIRE(workspace_bytes >= egc_csr_from_edges_workspace_bytes(n_edges, n_nodes),
              "egc_csr_from_edges: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int total = static_cast<int>(n_edges + (loops != EGC_LOOPS_NONE ? n_nodes : 0));
  const int cap = static_cast<int>(n_edges + n_nodes);
  Carver ws(workspace);
  int32_t* keys_in = ws.take<int32_t>(cap);
  int32_t* vals_in = ws.take<int32_t>(cap);
  int32_t* keys_out = ws.take<int32_t>(cap);
  int* max_id = ws.take<int>(1);
  size_t temp_bytes = radix_temp_bytes(cap);
  void* temp = ws.take<char>(temp_bytes);

  EGC_CUDA(cudaMemsetAsync(meta, 0, EGC_META_SLOTS * sizeof(int32_t), st));
  EGC_CUDA(cudaMemsetAsync(max_id, 0xff, sizeof(int), st));  // -1
  const int threads = 256;
  const int grid_e = std::max(1, std::min(ceil_div(std::max<int64_t>(n_edges, 1), threads), sm_count() * 16));
  if (loops == EGC_LOOPS_UP_TO_MAX_ID && n_edges > 0) {
    {
      LaunchScope egc_ls_("k_max_id", st);
      k_max_id<<<grid_e, threads, 0, st>>>(src, dst, n_edges, max_id);
    }
    EGC_LAUNCH_CHECK("k_max_id");
  }
  if (total > 0) {
    const int grid_t = std::max(1, std::min(ceil_div(total, threads), sm_count() * 16));
    {
      LaunchScope egc_ls_("k_make_edge_keys", st);
      k_make_edge_keys<<<grid_t, threads, 0, st>>>(src, dst, n_edges, n_nodes, loops, max_id, keys_in, vals_in, meta);
    }
    EGC_LAUNCH_CHECK("k_make_edge_keys");
    EGC_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, col, total, 0,
                                             key_bits(n_nodes), st));
  }
  {
    LaunchScope egc_ls_("k_ptr_from_sorted_keys", st);
    k_ptr_from_sorted_keys<<<ceil_div(n_nodes + 1, threads), threads, 0, st>>>(keys_out, total, n_nodes, rowptr);
  }
  EGC_LAUNCH_CHECK("k_ptr_from_sorted_keys");
  {
    LaunchScope egc_ls_("k_row_stats", st);
    k_row_stats<<<ceil_div(n_nodes, threads), threads, 0, st>>>(rowptr, n_nodes, meta);
  }
  EGC_LAUNCH_CHECK("k_row_stats");
  return EGC_OK;
}

size_t egc_csr_fill_diag_workspace_bytes(int32_t n_dst) {
  if (n_dst <= 0) return 0;
  return align_up(static_cast<size_t>(n_dst + 1) * 4, 256) + align_up(scan_temp_bytes(n_dst + 1), 256) + 256;
}

int egc_csr_fill_diag(const int64_t* rowptr_in, const int64_t* col_in, const float* value_in, int32_t n_dst,
                      int32_t n_src, int32_t fill_diag, int32_t* rowptr, int32_t* col, float* value_out,
                      int32_t* meta, void* workspace, size_t workspace_bytes, void* stream) {
  EGC_REQUIRE(n_dst > 0 && n_src > 0, "egc_csr_fill_diag: n_dst=%d n_src=%d", n_dst, n_src);
  EGC_REQUIRE(rowptr_in && col_in && rowptr && col && meta && workspace, "egc_csr_fill_diag: null pointer");
  EGC_REQUIRE(value_in == nullptr || value_out != nullptr, "egc_csr_fill_diag: value_out required with value_in");
  EGC_REQUIRE(workspace_bytes >= egc_csr_fill_diag_workspace_bytes(n_dst), "egc_csr_fill_diag: workspace too small");
  cudaStream_t st = as_stream(stream);
  Carver ws(workspace);
  int32_t* cnt = ws.take<int32_t>(n_dst + 1);
  size_t temp_bytes = scan_temp_bytes(n_dst + 1);
  void* temp = ws.take<char>(temp_bytes);
  const int n_diag = std::min(n_dst, n_src);
  const int threads = 256, rows_per_block = threads / 32;
  EGC_CUDA(cudaMemsetAsync(meta, 0, EGC_META_SLOTS * sizeof(int32_t), st));
  {
    LaunchScope egc_ls_("k_filldiag_count", st);
    k_filldiag_count<<<ceil_div(n_dst + 1, rows_per_block), threads, 0, st>>>(rowptr_in, col_in, n_dst, n_src, n_diag,
                                                                            fill_diag, cnt, meta);
  }
  EGC_LAUNCH_CHECK("k_filldiag_count");
  EGC_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, cnt, rowptr, n_dst + 1, st));
  {
    LaunchScope egc_ls_("k_filldiag_fill", st);
    k_filldiag_fill<<<ceil_div(n_dst, rows_per_block), threads, 0, st>>>(rowptr_in, col_in, value_in, n_dst, n_diag,
                                                                       fill_diag, rowptr, col, value_out);
  }
  EGC_LAUNCH_CHECK("k_filldiag_fill");
  {
    LaunchScope egc_ls_("k_row_stats", st);
    k_row_stats<<<ceil_div(n_dst, threads), threads, 0, st>>>(rowptr, n_dst, meta);
  }
  EGC_LAUNCH_CHECK("k_row_stats");
  return EGC_OK;
}

int egc_symnorm_weights(const int32_t* rowptr, const int32_t* col, const float* value, int32_t n_dst, float* deg,
                        float* dis, float* val_sym, void* stream) {
  EGC_REQUIRE(n_dst > 0 && rowptr && col && dis && val_sym, "egc_symnorm_weights: bad arguments");
  cudaStream_t st = as_stream(stream);
  const int threads = 256, rows_per_block = threads / 32;
  {
    LaunchScope egc_ls_("k_degree", st);
    k_degree<<<ceil_div(n_dst, rows_per_block), threads, 0, st>>>(rowptr, value, n_dst, deg, dis);
  }
  EGC_LAUNCH_CHECK("k_degree");
  {
    LaunchScope egc_ls_("k_symnorm_values", st);
    k_symnorm_values<<<ceil_div(n_dst, rows_per_block), threads, 0, st>>>(rowptr, col, value, dis, n_dst, val_sym);
  }
  EGC_LAUNCH_CHECK("k_symnorm_values");
  return EGC_OK;
}

size_t egc_csr_transpose_workspace_bytes(int32_t nnz, int32_t n_dst, int32_t n_src) {
  (void)n_dst; (void)n_src;
  if (nnz <= 0) return 256;
  return 2 * align_up(static_cast<size_t>(nnz) * 4, 256) + align_up(radix_temp_bytes(nnz), 256) + 256;
}

int egc_csr_transpose(const int32_t* rowptr, const int32_t* col, int32_t nnz, int32_t n_dst, int32_t n_src,
                      int32_t* colptr, int32_t* rowidx, int32_t* csr2csc, int32_t* meta, void* workspace,
                      size_t workspace_bytes, void* stream) {
  EGC_REQUIRE(nnz >= 0 && n_dst > 0 && n_src > 0, "egc_csr_transpose: bad sizes");
  EGC_REQUIRE(rowptr && colptr && meta && workspace && (nnz == 0 || (col && rowidx && csr2csc)),
              "egc_csr_transpose: null pointer");
  EGC_REQUIRE(workspace_bytes >= egc_csr_transpose_workspace_bytes(nnz, n_dst, n_src),
              "egc_csr_transpose: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int threads = 256;
  EGC_CUDA(cudaMemsetAsync(meta, 0, EGC_META_SLOTS * sizeof(int32_t), st));
  Carver ws(workspace);
  int32_t* iota = ws.take<int32_t>(std::max(nnz, 1));
  int32_t* keys_out = ws.take<int32_t>(std::max(nnz, 1));
  if (nnz > 0) {
    size_t temp_bytes = radix_temp_bytes(nnz);
    void* temp = ws.take<char>(temp_bytes);
    {
      LaunchScope egc_ls_("k_iota", st);
      k_iota<<<ceil_div(nnz, threads), threads, 0, st>>>(iota, nnz);
    }
    EGC_LAUNCH_CHECK("k_iota");
    EGC_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, col, keys_out, iota, csr2csc, nnz, 0,
                                             key_bits(n_src), st));
    {
      LaunchScope egc_ls_("k_rows_of_perm", st);
      k_rows_of_perm<<<ceil_div(nnz, threads), threads, 0, st>>>(rowptr, n_dst, csr2csc, nnz, rowidx);
    }
    EGC_LAUNCH_CHECK("k_rows_of_perm");
  }
  {
    LaunchScope egc_ls_("k_ptr_from_sorted_keys", st);
    k_ptr_from_sorted_keys<<<ceil_div(n_src + 1, threads), threads, 0, st>>>(keys_out, nnz, n_src, colptr);
  }
  EGC_LAUNCH_CHECK("k_ptr_from_sorted_keys");
  {
    LaunchScope egc_ls_("k_row_stats", st);
    k_row_stats<<<ceil_div(n_src, threads), threads, 0, st>>>(colptr, n_src, meta);
  }
  EGC_LAUNCH_CHECK("k_row_stats");
  return EGC_OK;
}

size_t egc_plan_build_workspace_bytes(int32_t n_rows) {
  if (n_rows <= 0) return 0;
  return 4 * align_up(static_cast<size_t>(n_rows + 1) * 4, 256) + align_up(scan_temp_bytes(n_rows + 1), 256) + 256;
}

int egc_plan_build(const int32_t* rowptr, int32_t n_rows, int32_t n_long, int32_t n_chunks, int32_t* long_rows,
                   int32_t* long_chunk_ptr, int32_t* chunk_row, int32_t* chunk_begin, void* workspace,
                   size_t workspace_bytes, void* stream) {
  EGC_REQUIRE(n_rows > 0 && n_long >= 0 && n_chunks >= 0, "egc_plan_build: bad sizes");
  if (n_long == 0) return EGC_OK;
  EGC_REQUIRE(rowptr && long_rows && long_chunk_ptr && chunk_row && chunk_begin && workspace,
              "egc_plan_build: null pointer");
  EGC_REQUIRE(workspace_bytes >= egc_plan_build_workspace_bytes(n_rows), "egc_plan_build: workspace too small");
  cudaStream_t st = as_stream(stream);
  Carver ws(workspace);
  int32_t* is_long = ws.take<int32_t>(n_rows + 1);
  int32_t* nchunk = ws.take<int32_t>(n_rows + 1);
  int32_t* long_off = ws.take<int32_t>(n_rows + 1);
  int32_t* chunk_off = ws.take<int32_t>(n_rows + 1);
  size_t temp_bytes = scan_temp_bytes(n_rows + 1);
  void* temp = ws.take<char>(temp_bytes);
  const int threads = 256;
  {
    LaunchScope egc_ls_("k_plan_flags", st);
    k_plan_flags<<<ceil_div(n_rows + 1, threads), threads, 0, st>>>(rowptr, n_rows, is_long, nchunk);
  }
  EGC_LAUNCH_CHECK("k_plan_flags");
  EGC_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, is_long, long_off, n_rows + 1, st));
  EGC_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, nchunk, chunk_off, n_rows + 1, st));
  {
    LaunchScope egc_ls_("k_plan_fill", st);
    k_plan_fill<<<ceil_div(n_rows + 1, threads), threads, 0, st>>>(rowptr, n_rows, long_off, chunk_off, long_rows,
                                                                 long_chunk_ptr, chunk_row, chunk_begin);
  }
  EGC_LAUNCH_CHECK("k_plan_fill");
  return EGC_OK;
}

}  // extern "C"
