// Instantiates the G = 32 family of the row-block forward aggregation kernel (aggregate_rows.cuh) for layer shapes
// without a specialised configuration.
#include <algorithm>

#include "aggregate_rows.cuh"

namespace egc {
int launch_aggregate_rows_g32(const AggParams& p, int mask, bool arg, int* task_counter, cudaStream_t st) {
  return launch_rows_family<32>(p, mask, arg, task_counter, st);
}
}  // namespace egc
