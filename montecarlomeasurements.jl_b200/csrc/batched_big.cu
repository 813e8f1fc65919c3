// batched_big.cu — real per-particle small-matrix kernels for n = 9..16 (see batched_kernels.cuh)
#include "batched_kernels.cuh"

namespace mcm {
int la_run_big(int dtype, int op, int nmat, const LaArgs& g, int m, size_t n) {
  return dtype == MCM_F64 ? la_by_op<double, 9, 16>(op, nmat, g, m, n) : la_by_op<float, 9, 16>(op, nmat, g, m, n);
}
}  // namespace mcm
