// tcgen05 (kind::tf32) projections - placeholder until the tensor-core kernels land:
// reports "unsupported" so EGC_GEMM_AUTO resolves to the exact-fp32 FFMA tiles.
#include "project.cuh"

namespace egc {

bool project_tc_supported(int, int, int, int) { return false; }

int project_fwd_tc(const float*, const float*, const float*, const float*, int, int, int, int, int, float*, float*,
                   int, cudaStream_t) {
  set_error("tensor-core projection not built");
  return EGC_ERR_UNSUPPORTED;
}

size_t project_bwd_tc_workspace(int, int, int, int) { return 0; }

int project_bwd_tc(const float*, const float*, const float*, const float*, const float*, int, int, int, int, float*,
                   float*, float*, float*, int, void*, size_t, cudaStream_t) {
  set_error("tensor-core projection not built");
  return EGC_ERR_UNSUPPORTED;
}

}  // namespace egc
