// Instantiates the fast forward aggregation kernel for the specialised layer shapes (EGC_STATIC_CFGS).
#include <algorithm>

#include "aggregate_fast.cuh"

namespace egc {
int launch_aggregate_fast_static(int cfg_index, const AggParams& p, bool arg, int smem_bytes, cudaStream_t st) {
  switch (cfg_index) {
#define X(I, ...) case I: return launch_fast_arg<StaticCfg<__VA_ARGS__>>(p, arg, smem_bytes, st);
    EGC_STATIC_CFGS(X)
#undef X
  }
  set_error("aggregate: unknown static configuration %d", cfg_index);
  return EGC_ERR_UNSUPPORTED;
}
}  // namespace egc
