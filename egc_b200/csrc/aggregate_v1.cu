// Instantiates the VEC=1 family of the fused forward aggregation kernel (one TU per family: parallel compiles).
#include "aggregate_impl.cuh"

namespace egc {
int launch_aggregate_v1(const AggParams& p, int mask, bool linw, bool arg, int smem_bytes, cudaStream_t st) {
  return launch_family<1>(p, mask, linw, arg, smem_bytes, st);
}
}  // namespace egc
