// Instantiates the G = 16 family (basis rows of 33..64 floats) of the fast forward aggregation kernel.
#include <algorithm>

#include "aggregate_fast.cuh"

namespace egc {
int launch_aggregate_fast_g16(const AggParams& p, int mask, bool arg, int smem_bytes, cudaStream_t st) {
  return launch_fast_family<16>(p, mask, arg, smem_bytes, st);
}
}  // namespace egc
