// Instantiates the VEC=1 backward pass-1 family of k_aggregate (one TU per family: parallel compiles).
#include "aggregate_impl.cuh"

namespace egc {
int launch_aggregate_bwd_v1(const AggParams& p, int mask, bool linw, int smem_bytes, cudaStream_t st) {
  return launch_family<1, true>(p, mask, linw, smem_bytes, st);
}
}  // namespace egc
