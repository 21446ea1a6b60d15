// Instantiates the VEC=4 family of the fused forward aggregation kernel (one TU per family: parallel compiles).
#include "aggregate_impl.cuh"

namespace egc {
int launch_aggregate_v4(const AggParams& p, int mask, bool linw, bool arg, int smem_bytes, cudaStream_t st) {
  return launch_family<4>(p, mask, linw, arg, smem_bytes, st);
}
}  // namespace egc
