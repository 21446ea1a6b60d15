// Decodes how tcgen05.mma (kind::tf32, no-swizzle descriptors) addresses shared memory for K-major and
// MN-major operands: A (or B) is filled with its own word index, the other operand is an identity, so
// D reveals which smem word was used for every (row, k).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -Iinclude -Iegc_b200/csrc -o build/umma_probe tools/umma_probe.cu
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
using namespace egc;

struct Cfg { int a_major, b_major; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int mode; /*0: decode A, 1: decode B*/ };

__global__ void k_probe(Cfg c, float* out /*[128][16]*/) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* A = reinterpret_cast<float*>(smem);            // 16 KB region
  float* B = reinterpret_cast<float*>(smem + 16384);    // 16 KB region
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096; i += blockDim.x) { A[i] = 0.f; B[i] = 0.f; }
  __syncthreads();
  if (c.mode == 0) {
    for (int i = tid; i < 2048; i += blockDim.x) A[i] = float(i);          // A = word index
    // B = identity over k (K-major known-good layout: piece p=k/4, row n: off = p*b_lbo + n*16 + (k%4)*4), N=16
    if (tid < 8) { int k = tid, n = tid; *reinterpret_cast<float*>(smem + 16384 + (k / 4) * c.b_lbo + n * 16 + (k % 4) * 4) = 1.f; }
  } else {
    for (int i = tid; i < 2048; i += blockDim.x) B[i] = float(i);
    if (tid < 8) { int k = tid, m = tid; *reinterpret_cast<float*>(smem + (k / 4) * c.a_lbo + m * 16 + (k % 4) * 4) = 1.f; }
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 32);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(c.a_major) << 15) | (uint32_t(c.b_major) << 16) |
                     (uint32_t(16 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    umma_tf32(tb, make_desc(smem_u32(A), c.a_lbo, c.a_sbo), make_desc(smem_u32(B), c.b_lbo, c.b_sbo), idesc, 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t r[16];
  tmem_ld16(tb + (uint32_t(warp * 32) << 16), r);
  tmem_ld_wait();
  for (int j = 0; j < 16; ++j) out[tid * 16 + j] = __uint_as_float(r[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 32);
}

int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  std::vector<float> h(128 * 16);
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  // K-major reference strides for the identity operand: N(or M) rows: lbo = rows*16, sbo = 128
  Cfg cfgs[] = {
    {0, 0, 2048, 128, 256, 128, 0},     // A K-major sanity: expect word(m,k) = (k/4)*512 + m*4 + k%4
    {1, 0, 4096, 128, 256, 128, 0},     // A MN-major, LBO=4096 SBO=128
    {1, 0, 128, 4096, 256, 128, 0},     // swapped
    {1, 0, 4096, 256, 256, 128, 0},     // SBO=256
    {1, 0, 512, 128, 256, 128, 0},
    {0, 1, 2048, 128, 4096, 128, 1},    // B MN-major (N=16): decode B
    {0, 1, 2048, 128, 128, 4096, 1},
  };
  for (auto& c : cfgs) {
    cudaMemset(d, 0, 128 * 16 * 4);
    k_probe<<<1, 128, 36864>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    printf("a_major=%d b_major=%d a(lbo=%u,sbo=%u) b(lbo=%u,sbo=%u) mode=%d : %s\n", c.a_major, c.b_major, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, c.mode, cudaGetErrorString(e));
    if (c.mode == 0) {
      for (int m : {0, 1, 2, 3, 4, 5, 8, 9, 32, 127}) { printf("  m=%3d:", m); for (int k = 0; k < 8; ++k) printf(" %6.0f", h[m * 16 + k]); printf("\n"); }
    } else {
      for (int k = 0; k < 8; ++k) { printf("  k=%d:", k); for (int n = 0; n < 16; ++n) printf(" %5.0f", h[k * 16 + n]); printf("\n"); }
    }
  }
  return 0;
}
