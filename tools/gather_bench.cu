// Micro-benchmark: what gather bandwidth can a B200 sustain for random fixed-size rows?
// (design input for the SpMM kernels: L2-resident vs HBM-resident tables, row width, loads in flight)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/gather_bench tools/gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

template <int U, int ROWF4>   // ROWF4 = float4 per row handled by one group of lanes (32 -> 512 B rows)
__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ table, const int* __restrict__ idx,
                                                int n_idx_per_warp, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5);
  constexpr int G = ROWF4 >= 32 ? 32 : ROWF4, NG = 32 / G, PASS = ROWF4 >= 32 ? ROWF4 / 32 : 1;
  const int g = lane / G, piece = lane % G;
  const int* my = idx + (size_t)gw * n_idx_per_warp;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int e0 = 0; e0 < n_idx_per_warp; e0 += U * NG) {
    int j[U];
#pragma unroll
    for (int u = 0; u < U; ++u) j[u] = __ldg(my + e0 + u * NG + g);
    float4 x[U][PASS];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < PASS; ++q) x[u][q] = __ldg(table + (size_t)j[u] * ROWF4 + q * 32 + piece);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < PASS; ++q) { acc.x += x[u][q].x; acc.y += x[u][q].y; acc.z += x[u][q].z; acc.w += x[u][q].w; }
  }
  out[(size_t)gw * 32 + lane] = acc;
}

template <int U, int ROWF4>
void run(const char* label, size_t table_bytes, int warps_total, int n_per_warp, int max_blocks_per_sm) {
  size_t rows = table_bytes / (ROWF4 * 16);
  float4* table; int* idx; float4* out;
  cudaMalloc(&table, rows * ROWF4 * 16);
  cudaMemset(table, 0, rows * ROWF4 * 16);
  std::vector<int> h((size_t)warps_total * n_per_warp);
  unsigned long long s = 88172645463325252ull;
  for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int)(s % rows); }
  cudaMalloc(&idx, h.size() * 4);
  cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, (size_t)warps_total * 32 * 16);
  // limit occupancy with dynamic smem
  int smem = max_blocks_per_sm > 0 ? (200 * 1024 / max_blocks_per_sm) : 0;
  cudaFuncSetAttribute(k_gather<U, ROWF4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(a);
    k_gather<U, ROWF4><<<warps_total / 8, 256, smem>>>(table, idx, n_per_warp, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (it > 0 && ms < best) best = ms;
  }
  double bytes = (double)h.size() * ROWF4 * 16;
  printf("%-28s table=%6.0f MB row=%4d B U=%2d occ<=%d blocks/SM : %7.3f ms  %7.1f GB/s  (%s)\n", label,
         table_bytes / 1e6, ROWF4 * 16, U, max_blocks_per_sm, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(table); cudaFree(idx); cudaFree(out);
}

int main() {
  const int W = 148 * 8 * 16;          // warps
  // 512 B rows (arxiv EGC-M basis row), L2-resident vs not
  run<8, 32>("512B rows", 64ull << 20, W, 128, 8);
  run<8, 32>("512B rows", 87ull << 20, W, 128, 8);
  run<8, 32>("512B rows", 87ull << 20, W, 128, 4);
  run<8, 32>("512B rows", 87ull << 20, W, 128, 2);
  run<4, 32>("512B rows", 87ull << 20, W, 128, 8);
  run<16, 32>("512B rows", 87ull << 20, W, 128, 8);
  run<8, 32>("512B rows", 260ull << 20, W, 128, 8);
  run<8, 32>("512B rows", 1024ull << 20, W, 128, 8);
  // 1536 B rows (three interleaved streams of the backward CSC pass)
  run<4, 96>("1536B rows", 260ull << 20, W, 64, 8);
  run<4, 96>("1536B rows", 65ull << 20, W, 64, 8);
  // 256 B rows (mag EGC-S), 128 B rows (feature slabs)
  run<8, 16>("256B rows", 188ull << 20, W, 256, 8);
  run<8, 16>("256B rows", 64ull << 20, W, 256, 8);
  run<8, 8>("128B rows", 64ull << 20, W, 512, 8);
  run<8, 8>("128B rows", 260ull << 20, W, 512, 8);
  return 0;
}
