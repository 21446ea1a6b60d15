// Instantiates the G = 32 family (basis rows of 65..128 floats) of the fast forward aggregation kernel.
#include <algorithm>

#include "aggregate_fast.cuh"

namespace egc {
int launch_aggregate_fast_g32(const AggParams& p, int mask, bool arg, int smem_bytes, cudaStream_t st) {
  return launch_fast_family<32>(p, mask, arg, smem_bytes, st);
}
}  // namespace egc
