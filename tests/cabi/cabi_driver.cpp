// cabi_driver.cpp — a plain C++ caller of the C ABI (no Python, no torch): what a Julia `ccall` or any other FFI
// sees.  Exercises README quick-start arithmetic (README.md:28-35: a = pi +- 0.1; sin(a)*a^2 + exp(-a)) through
// single primitives and through one fused tape, statistics, quantiles and the covariance of two variables.
// Exit code 0 = all checks passed; 3 = no device (the library must refuse to run, there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/mcm_b200.h"

#define CK(x)                                                                       \
  do {                                                                              \
    int rc__ = (x);                                                                 \
    if (rc__ != MCM_OK) { std::printf("FAIL %s -> %d: %s\n", #x, rc__, mcm_last_error()); return 1; } \
  } while (0)
#define EXPECT(cond)                                                  \
  do {                                                                \
    if (!(cond)) { std::printf("CHECK FAILED: %s (line %d)\n", #cond, __LINE__); return 1; } \
  } while (0)

int main() {
  std::printf("libmcm_b200 version %d\n", mcm_version());
  if (mcm_init(0) != MCM_OK) {
    std::printf("mcm_init failed (expected without a GPU): %s\n", mcm_last_error());
    double dummy[6];
    // every other entry point must refuse as well
    if (mcm_moments(MCM_F64, dummy, 1, dummy) != MCM_ERR_NOT_INIT) { std::printf("entry point did not refuse\n"); return 1; }
    return 3;
  }
  const size_t N = 2000;   // DEFAULT_NUM_PARTICLES (MonteCarloMeasurements.jl:47)
  void *a = nullptr, *t1 = nullptr, *t2 = nullptr, *y = nullptr, *yf = nullptr;
  CK(mcm_alloc(&a, N * 8)); CK(mcm_alloc(&t1, N * 8)); CK(mcm_alloc(&t2, N * 8)); CK(mcm_alloc(&y, N * 8)); CK(mcm_alloc(&yf, N * 8));
  // a = pi +- 0.1 : systematic sample in grid order
  CK(mcm_systematic_normal(MCM_F64, a, N, 3.14159265358979323846, 0.1, nullptr, 0, N));
  double m[6];
  CK(mcm_moments(MCM_F64, a, N, m));
  const double std_a = std::sqrt(m[2] / (m[0] - 1.0));
  std::printf("pmean(a) = %.15g  pstd(a) = %.17g\n", m[1], std_a);
  EXPECT(std::fabs(std_a - 0.09999231528930486) < 1e-12);          // README.md:31-35 known answer
  // y = sin(a) * a^2 + exp(-a), one launch per primitive (register_primitive.jl bodies)
  CK(mcm_unary(MCM_U_SIN, MCM_F64, a, t1, N));
  CK(mcm_powi(MCM_F64, a, 2, t2, N));
  CK(mcm_binary(MCM_B_MUL, MCM_F64, t1, 0.0, t2, 0.0, t1, N));
  CK(mcm_unary(MCM_U_NEG, MCM_F64, a, t2, N));
  CK(mcm_unary(MCM_U_EXP, MCM_F64, t2, t2, N));
  CK(mcm_binary(MCM_B_ADD, MCM_F64, t1, 0.0, t2, 0.0, y, N));
  // the same expression as ONE fused tape
  mcm_tape_ins tape[8];
  std::memset(tape, 0, sizeof(tape));
  auto ins = [&](int i, int op, int dst, int x, int b) { tape[i].op = (uint8_t)op; tape[i].dst = (uint8_t)dst; tape[i].a = (uint8_t)x; tape[i].b = (uint8_t)b; };
  ins(0, MCM_TAPE_LOAD, 0, 0, 0);
  ins(1, MCM_TAPE_UNARY + MCM_U_SIN, 1, 0, 0);
  ins(2, MCM_TAPE_POWI, 2, 0, 2);
  ins(3, MCM_TAPE_BINARY + MCM_B_MUL, 1, 1, 2);
  ins(4, MCM_TAPE_UNARY + MCM_U_NEG, 2, 0, 0);
  ins(5, MCM_TAPE_UNARY + MCM_U_EXP, 2, 2, 0);
  ins(6, MCM_TAPE_BINARY + MCM_B_ADD, 1, 1, 2);
  ins(7, MCM_TAPE_STORE, 0, 1, 0);
  const void* inputs[1] = {a};
  void* outputs[1] = {yf};
  double st[6];
  CK(mcm_chain(MCM_F64, tape, 8, 8, 8, 0, nullptr, 0, inputs, 1, outputs, 1, N, 1, st));
  std::vector<double> hy(N), hyf(N);
  CK(mcm_download(hy.data(), y, N * 8));
  CK(mcm_download(hyf.data(), yf, N * 8));
  EXPECT(std::memcmp(hy.data(), hyf.data(), N * 8) == 0);            // fused == unfused, bit for bit
  CK(mcm_moments(MCM_F64, y, N, m));
  // fused statistics == separate reduction up to the order of summation (different grids): 1e-12 relative
  EXPECT(std::fabs(m[1] - st[1]) <= 1e-13 * std::fabs(m[1]) && std::fabs(m[2] - st[2]) <= 1e-12 * m[2] && m[0] == st[0] &&
         m[3] == st[3] && m[4] == st[4]);
  std::printf("y = sin(a)*a^2 + exp(-a): pmean = %.15g  pstd = %.15g\n", m[1], std::sqrt(m[2] / (m[0] - 1.0)));
  // quantiles: median of a symmetric sample is the centre
  const double q[3] = {0.025, 0.5, 0.975};
  double qa[3];
  CK(mcm_quantile(MCM_F64, a, N, q, 3, qa));
  EXPECT(std::fabs(qa[1] - 3.14159265358979323846) < 1e-14 && qa[0] < qa[1] && qa[1] < qa[2]);
  // pcov of (a, y)
  const void* vars[2] = {a, y};
  void* dC = nullptr;
  CK(mcm_alloc(&dC, 4 * 8));
  CK(mcm_gram(MCM_F64, vars, 2, N, 1, nullptr, (double*)dC, nullptr));
  double Cm[4];
  CK(mcm_download(Cm, dC, sizeof(Cm)));
  EXPECT(std::fabs(Cm[0] - std_a * std_a) < 1e-15 && Cm[1] == Cm[2]);
  // numerical domain errors raise the sticky flag instead of throwing
  CK(mcm_fill(MCM_F64, t1, N, -1.0));
  CK(mcm_unary(MCM_U_SQRT, MCM_F64, t1, t2, N));
  uint32_t bits = 0;
  CK(mcm_check_flags(&bits, 1));
  EXPECT(bits & MCM_FLAG_DOMAIN);
  // decision layer: nearest particle to pi is the median of the sorted sample; with_nominal moves it to index 0
  uint64_t nearest = 0;
  CK(mcm_argmin_absdiff(MCM_F64, a, N, 3.14159265358979323846, &nearest));
  EXPECT(nearest == N / 2 - 1 || nearest == N / 2);
  CK(mcm_shuffle_insert(MCM_F64, a, N, nearest, 3.0, t1));
  double first = 0.0;
  CK(mcm_download(&first, t1, 8));
  EXPECT(first == 3.0);
  // per-particle 2 x 2 matrices [[a, 1], [0, a]]: determinant a^2, inverse (0,0) entry 1/a
  CK(mcm_fill(MCM_F64, t1, N, 1.0));
  CK(mcm_fill(MCM_F64, t2, N, 0.0));
  const void* Am[4] = {a, t1, t2, a};
  void* det = nullptr;
  CK(mcm_alloc(&det, N * 8));
  void* outs1[1] = {det};
  CK(mcm_batched_linalg(MCM_LA_DET, MCM_F64, 2, 0, Am, nullptr, outs1, nullptr, N));
  CK(mcm_binary(MCM_B_MUL, MCM_F64, a, 0.0, a, 0.0, yf, N));
  CK(mcm_binary(MCM_B_SUB, MCM_F64, det, 0.0, yf, 0.0, det, N));
  CK(mcm_moments(MCM_F64, det, N, m));
  EXPECT(std::fabs(m[3]) < 1e-14 && std::fabs(m[4]) < 1e-14);
  // Complex{Particles}: (a + i)(a - i) = a^2 + 1 exactly in the real part, 0 in the imaginary part
  CK(mcm_complex_op(MCM_C_MUL, MCM_F64, a, nullptr, 0.0, 1.0, a, nullptr, 0.0, -1.0, det, t2, N));
  CK(mcm_moments(MCM_F64, t2, N, m));
  EXPECT(m[3] == 0.0 && m[4] == 0.0);
  CK(mcm_free(det));
  // bad arguments are rejected on the host with a message
  EXPECT(mcm_unary(9999, MCM_F64, a, t1, N) != MCM_OK && std::strlen(mcm_last_error()) > 0);
  CK(mcm_free(a)); CK(mcm_free(t1)); CK(mcm_free(t2)); CK(mcm_free(y)); CK(mcm_free(yf)); CK(mcm_free(dC));
  CK(mcm_finalize());
  std::printf("C ABI driver: all checks passed\n");
  return 0;
}
