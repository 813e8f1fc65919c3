// batched.cu — C entry points and the real kernels for n = 1..8 (the other sizes: batched_big.cu, batched_c*.cu);
// per-particle small-matrix operations on device.
// The reference maps a host closure over the particle index (ℝⁿ2ℝⁿ_function / ℝⁿ2ℂⁿ_function,
// src/particles.jl:241-282; eigvals :395-406; exp / exp! / lyap :603-607; `\` on matrices via bymap): N calls into
// LAPACK on n x n matrices gathered one scalar per particle-variable.  Here entry (r, c) of the uncertain matrix is a
// length-N device vector, thread i owns particle i, loads its n x n matrix with n^2 coalesced reads (adjacent
// threads = adjacent particles of the same entry), runs the algorithm of mcm_small_linalg.cuh in registers / local
// memory and writes every output entry coalesced.  Bound: HBM for n <= 3 (8 (n^2 + outputs) bytes per particle),
// FP64 pipe beyond.
#include "batched_kernels.cuh"

namespace mcm {
int la_run_small(int dtype, int op, int nmat, const LaArgs& g, int m, size_t n) {
  return dtype == MCM_F64 ? la_by_op<double, 1, 8>(op, nmat, g, m, n) : la_by_op<float, 1, 8>(op, nmat, g, m, n);
}
}  // namespace mcm

using namespace mcm;

extern "C" int mcm_batched_linalg(int op, int dtype, int nmat, int m, const void* const* A, const void* const* B,
                                  void* const* out, void* const* out_im, size_t n) {
  MCM_NEED_INIT();
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: bad dtype %d", dtype);
  if (nmat < 1 || nmat > LA_MAXN) return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg: matrix size must be 1..%d", LA_MAXN);
  if (!A || !out) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: null table");
  const bool needs_b = op == MCM_LA_SOLVE || op == MCM_LA_LYAP;
  if (op == MCM_LA_LYAP) m = nmat;
  if (needs_b && (!B || m < 1 || nmat * m > LA_MAXPTR)) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: bad right-hand side");
  if (op == MCM_LA_EIGVALS && !out_im) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: eigvals needs out_im");
  int nout = 0;
  switch (op) {
    case MCM_LA_SOLVE: nout = nmat * m; break;
    case MCM_LA_INV: case MCM_LA_EXP: case MCM_LA_LYAP: nout = nmat * nmat; break;
    case MCM_LA_DET: nout = 1; break;
    case MCM_LA_EIGVALS: case MCM_LA_EIGVALS_SYM: case MCM_LA_SVDVALS: nout = nmat; break;
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg: unknown operation %d", op);
  }
  LaArgs g{};
  for (int e = 0; e < nmat * nmat; ++e) {
    if (!A[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: A[%d] is null", e);
    g.A[e] = A[e];
  }
  if (needs_b)
    for (int e = 0; e < nmat * m; ++e) {
      if (!B[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: B[%d] is null", e);
      g.B[e] = B[e];
    }
  for (int e = 0; e < nout; ++e) {
    if (!out[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: out[%d] is null", e);
    g.out[e] = out[e];
  }
  if (op == MCM_LA_EIGVALS)
    for (int e = 0; e < nmat; ++e) {
      if (!out_im[e]) return fail(MCM_ERR_INVALID, "mcm_batched_linalg: out_im[%d] is null", e);
      g.out_im[e] = out_im[e];
    }
  if (n == 0) return MCM_OK;
  return nmat <= 8 ? la_run_small(dtype, op, nmat, g, m, n) : la_run_big(dtype, op, nmat, g, m, n);
}

extern "C" int mcm_batched_linalg_complex(int op, int dtype, int nmat, const void* const* A_re, const void* const* A_im,
                                          void* const* out_re, void* const* out_im, size_t n) {
  MCM_NEED_INIT();
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: bad dtype %d", dtype);
  if (nmat < 1 || nmat > LA_MAXN) return fail(MCM_ERR_UNSUPPORTED, "mcm_batched_linalg_complex: matrix size must be 1..%d", LA_MAXN);
  if (!A_re || !out_re) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: null table");
  int nout = 0;
  switch (op) {
    case MCM_LAC_EXP: case MCM_LAC_LOG: nout = nmat * nmat; break;
    case MCM_LAC_DET: nout = 1; break;
    case MCM_LAC_EIGVALS: case MCM_LAC_SVDVALS: nout = nmat; break;
    default: return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: unknown operation %d", op);
  }
  if (op != MCM_LAC_SVDVALS && !out_im) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: this operation needs out_im");
  LacArgs g{};
  for (int e = 0; e < nmat * nmat; ++e) {
    if (!A_re[e] || (A_im && !A_im[e])) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: A[%d] is null", e);
    g.Are[e] = A_re[e];
    g.Aim[e] = A_im ? A_im[e] : nullptr;
  }
  for (int e = 0; e < nout; ++e) {
    if (!out_re[e] || (op != MCM_LAC_SVDVALS && !out_im[e])) return fail(MCM_ERR_INVALID, "mcm_batched_linalg_complex: out[%d] is null", e);
    g.ore[e] = out_re[e];
    g.oim[e] = op != MCM_LAC_SVDVALS ? out_im[e] : nullptr;
  }
  if (n == 0) return MCM_OK;
  const int real_input = A_im ? 0 : 1;
  if (nmat <= 8) return lac_run_small(dtype, op, nmat, g, real_input, n);
  return nmat <= 12 ? lac_run_mid(dtype, op, nmat, g, real_input, n) : lac_run_big(dtype, op, nmat, g, real_input, n);
}
