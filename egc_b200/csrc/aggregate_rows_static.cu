// Instantiates the row-block forward aggregation kernel (aggregate_rows.cuh) for the specialised layer shapes.
#include <algorithm>

#include "aggregate_rows.cuh"

namespace egc {
int launch_aggregate_rows_static(int cfg_index, const AggParams& p, bool arg, int* task_counter, cudaStream_t st) {
  switch (cfg_index) {
#define X(I, ...) case I: return launch_rows_arg<StaticCfg<__VA_ARGS__>>(p, arg, task_counter, st);
    EGC_STATIC_CFGS(X)
#undef X
  }
  set_error("aggregate: unknown static configuration %d", cfg_index);
  return EGC_ERR_UNSUPPORTED;
}
int rows_kernel_smem_bytes(const AggParams& p) { return rows_smem_bytes(p); }
}  // namespace egc
