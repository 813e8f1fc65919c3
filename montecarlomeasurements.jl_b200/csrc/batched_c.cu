// batched_c.cu — complex per-particle small-matrix kernels for n = 1..8 (see batched_kernels.cuh)
#include "batched_kernels.cuh"

namespace mcm {
int lac_run_small(int dtype, int op, int nmat, const LacArgs& g, int real_input, size_t n) {
  return dtype == MCM_F64 ? lac_by_op<double, 1, 8>(op, nmat, g, real_input, n) : lac_by_op<float, 1, 8>(op, nmat, g, real_input, n);
}
}  // namespace mcm
