// chain.cu — fused evaluation of a whole chain of registered primitives.
//
// The reference evaluates `sin(a)*a^2 + exp(-a)` as five separate broadcasts,
// each allocating a Vector{T}(N) (src/register_primitive.jl:44-58,94-97).  At
// N = 1e8 that is 10 x 16-24 B/particle of HBM traffic for the canonical chain.
// Here the chain is a tape for a tiny register machine (8 value slots per
// particle, all in registers): one launch reads each input once, writes each
// requested output once, and optionally folds the result into the one-pass
// moments of mcm_moments.cuh (pmean/pstd, src/particles.jl:302-323) so that
// y never has to exist in HBM when only statistics are wanted.
// A loop section [loop_begin, loop_end) x reps allows in-register time
// stepping (RK4 on the pendulum of examples/ode_sim.jl:20-25).
#include <string.h>
#include "mcm_math.cuh"
#include "mcm_moments.cuh"

namespace mcm {

struct ChainParams {
  mcm_tape_ins tape[MCM_TAPE_MAX_INS];
  double consts[MCM_TAPE_MAX_CONST];
  const void* in[MCM_TAPE_MAX_IN];
  void* out[MCM_TAPE_MAX_OUT];
  unsigned long long n;
  int ntape, loop_begin, loop_end, reps;
  int stats_slot;
};

constexpr int S = MCM_TAPE_SLOTS;

template <typename T, int V>
__device__ __forceinline__ void slot_get(const T (&r)[S][V], const ChainParams& p, int idx, T (&x)[V]) {
  if (idx & MCM_TAPE_CONST) {
    const T c = (T)p.consts[idx & (MCM_TAPE_MAX_CONST - 1)];
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = c;
    return;
  }
  switch (idx & (S - 1)) {
#define MCM_G(K) \
  case K: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = r[K][v]; break;
    MCM_G(0) MCM_G(1) MCM_G(2) MCM_G(3) MCM_G(4) MCM_G(5) MCM_G(6) MCM_G(7)
#undef MCM_G
  }
}

template <typename T, int V>
__device__ __forceinline__ void slot_put(T (&r)[S][V], int idx, const T (&x)[V]) {
  switch (idx & (S - 1)) {
#define MCM_P(K) \
  case K: _Pragma("unroll") for (int v = 0; v < V; ++v) r[K][v] = x[v]; break;
    MCM_P(0) MCM_P(1) MCM_P(2) MCM_P(3) MCM_P(4) MCM_P(5) MCM_P(6) MCM_P(7)
#undef MCM_P
  }
}

template <typename T, int V>
__device__ __forceinline__ void exec_range(T (&r)[S][V], const ChainParams& p, int pc0, int pc1, size_t j,
                                           int nvalid, bool& bad) {
  using P = Pack<T>;
  for (int pc = pc0; pc < pc1; ++pc) {
    const mcm_tape_ins ins = p.tape[pc];
    const int op = ins.op;
    T x[V], y[V], z[V];
    if (op == MCM_TAPE_LOAD) {
      const T* src = static_cast<const T*>(p.in[ins.a & (MCM_TAPE_MAX_IN - 1)]);
      if (nvalid == V) {
        const P t = ld_stream(reinterpret_cast<const P*>(src) + j);
#pragma unroll
        for (int v = 0; v < V; ++v) x[v] = t.v[v];
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v) x[v] = (v < nvalid) ? src[j * V + v] : T(1);
      }
      slot_put<T, V>(r, ins.dst, x);
      continue;
    }
    slot_get<T, V>(r, p, ins.a, x);
    if (op == MCM_TAPE_STORE) {
      T* dst = static_cast<T*>(p.out[ins.dst & (MCM_TAPE_MAX_OUT - 1)]);
      if (dst != nullptr) {
        if (nvalid == V) {
          P t;
#pragma unroll
          for (int v = 0; v < V; ++v) t.v[v] = x[v];
          st_stream(reinterpret_cast<P*>(dst) + j, t);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v)
            if (v < nvalid) dst[j * V + v] = x[v];
        }
      }
      continue;
    }
    if (op == MCM_TAPE_MOV) {
      slot_put<T, V>(r, ins.dst, x);
      continue;
    }
    if (op == MCM_TAPE_POWI) {
      const int k = (int)(signed char)ins.b;
#pragma unroll
      for (int v = 0; v < V; ++v) x[v] = math::m_powi(x[v], k);
      slot_put<T, V>(r, ins.dst, x);
      continue;
    }
    if (op >= MCM_TAPE_BINARY) slot_get<T, V>(r, p, ins.b, y);
    if (op >= MCM_TAPE_TERNARY) slot_get<T, V>(r, p, ins.c, z);

#define MCM_U(OP)                                                                        \
  case MCM_TAPE_UNARY + OP:                                                              \
    _Pragma("unroll") for (int v = 0; v < V; ++v) {                                      \
      const T t = math::unary_apply<OP, T>(x[v]);                                        \
      if (math::UnaryCanDomainError<OP>::value) bad |= (t != t) && (x[v] == x[v]) && (v < nvalid); \
      x[v] = t;                                                                          \
    }                                                                                    \
    break;
#define MCM_B(OP)                                                                        \
  case MCM_TAPE_BINARY + OP:                                                             \
    _Pragma("unroll") for (int v = 0; v < V; ++v) {                                      \
      const T t = math::binary_apply<OP, T>(x[v], y[v]);                                 \
      if (math::BinaryCanDomainError<OP>::value)                                         \
        bad |= (t != t) && (x[v] == x[v]) && (y[v] == y[v]) && (v < nvalid);             \
      x[v] = t;                                                                          \
    }                                                                                    \
    break;
#define MCM_T(OP)                                                                        \
  case MCM_TAPE_TERNARY + OP:                                                            \
    _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = math::ternary_apply<OP, T>(x[v], y[v], z[v]); \
    break;
    switch (op) {
      MCM_U(MCM_U_PLUS) MCM_U(MCM_U_NEG) MCM_U(MCM_U_EXP) MCM_U(MCM_U_EXP2) MCM_U(MCM_U_EXP10)
      MCM_U(MCM_U_EXPM1) MCM_U(MCM_U_LOG) MCM_U(MCM_U_LOG10) MCM_U(MCM_U_LOG2) MCM_U(MCM_U_LOG1P)
      MCM_U(MCM_U_SIN) MCM_U(MCM_U_COS) MCM_U(MCM_U_TAN) MCM_U(MCM_U_SIND) MCM_U(MCM_U_COSD)
      MCM_U(MCM_U_TAND) MCM_U(MCM_U_SINH) MCM_U(MCM_U_COSH) MCM_U(MCM_U_TANH) MCM_U(MCM_U_ASIN)
      MCM_U(MCM_U_ACOS) MCM_U(MCM_U_ATAN) MCM_U(MCM_U_ASIND) MCM_U(MCM_U_ACOSD) MCM_U(MCM_U_ATAND)
      MCM_U(MCM_U_ASINH) MCM_U(MCM_U_ACOSH) MCM_U(MCM_U_ATANH) MCM_U(MCM_U_ZERO) MCM_U(MCM_U_SIGN)
      MCM_U(MCM_U_ABS) MCM_U(MCM_U_SQRT) MCM_U(MCM_U_RAD2DEG) MCM_U(MCM_U_DEG2RAD) MCM_U(MCM_U_FLOAT)
      MCM_U(MCM_U_ROUND) MCM_U(MCM_U_FLOOR) MCM_U(MCM_U_CEIL) MCM_U(MCM_U_TRUNC) MCM_U(MCM_U_INV)
      MCM_U(MCM_U_SQR) MCM_U(MCM_U_CUBE)
      MCM_B(MCM_B_ADD) MCM_B(MCM_B_SUB) MCM_B(MCM_B_MUL) MCM_B(MCM_B_DIV) MCM_B(MCM_B_POW)
      MCM_B(MCM_B_MAX) MCM_B(MCM_B_MIN) MCM_B(MCM_B_MOD) MCM_B(MCM_B_MOD1) MCM_B(MCM_B_ATAN2)
      MCM_B(MCM_B_ATAND2) MCM_B(MCM_B_ADD_SUM) MCM_B(MCM_B_HYPOT) MCM_B(MCM_B_REM) MCM_B(MCM_B_FDIV)
      MCM_T(MCM_T_CLAMP) MCM_T(MCM_T_MULADD) MCM_T(MCM_T_MUL_ADD)
      default: break;
    }
#undef MCM_U
#undef MCM_B
#undef MCM_T
    slot_put<T, V>(r, ins.dst, x);
  }
}

template <typename T, bool STATS>
__global__ void __launch_bounds__(kThreads)
k_chain(const __grid_constant__ ChainParams p, Mom* partials, unsigned int* ticket, double* stats_out,
        uint32_t* flags) {
  constexpr int V = Pack<T>::N;
  const size_t n = p.n;
  const size_t nvec = (n + V - 1) / V;
  const size_t stride = (size_t)gridDim.x * kThreads;
  bool bad = false;
  Mom m = mom_empty();

  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < nvec; j += stride) {
    const int nvalid = (j + 1) * V <= n ? V : (int)(n - j * V);
    T r[S][V];
#pragma unroll
    for (int k = 0; k < S; ++k)
#pragma unroll
      for (int v = 0; v < V; ++v) r[k][v] = T(0);
    exec_range<T, V>(r, p, 0, p.loop_begin, j, nvalid, bad);
    for (int rep = 0; rep < p.reps; ++rep) exec_range<T, V>(r, p, p.loop_begin, p.loop_end, j, nvalid, bad);
    exec_range<T, V>(r, p, p.loop_end, p.ntape, j, nvalid, bad);
    if constexpr (STATS) {
      T x[V];
      slot_get<T, V>(r, p, p.stats_slot, x);
      if (nvalid == V) {
        double d[V];
#pragma unroll
        for (int v = 0; v < V; ++v) d[v] = (double)x[v];
        mom_push_chunk<V>(m, d);
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v)
          if (v < nvalid) mom_push_one(m, (double)x[v]);
      }
    }
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
  if constexpr (STATS) {
    m = mom_block_reduce(m);
    mom_grid_finish(m, partials, ticket, stats_out);
  }
}

double* moments_result_slot();
int read_back(double* host_dst, const double* dev_src, size_t count);

static int validate(const ChainParams& p, int nconst, int nin, int nout) {
  if (p.ntape < 0 || p.ntape > MCM_TAPE_MAX_INS) return fail(MCM_ERR_INVALID, "mcm_chain: ntape %d out of range", p.ntape);
  if (p.loop_begin < 0 || p.loop_end < p.loop_begin || p.loop_end > p.ntape)
    return fail(MCM_ERR_INVALID, "mcm_chain: bad loop section [%d,%d) of %d", p.loop_begin, p.loop_end, p.ntape);
  if (p.reps < 0) return fail(MCM_ERR_INVALID, "mcm_chain: reps < 0");
  auto operand_ok = [&](int o) { return (o & MCM_TAPE_CONST) ? ((o & ~MCM_TAPE_CONST) < nconst) : (o < S); };
  for (int i = 0; i < p.ntape; ++i) {
    const mcm_tape_ins& t = p.tape[i];
    const int op = t.op;
    bool ok = true;
    if (op == MCM_TAPE_LOAD) ok = t.dst < S && t.a < nin;
    else if (op == MCM_TAPE_STORE) ok = t.dst < nout && operand_ok(t.a);
    else if (op == MCM_TAPE_MOV || op == MCM_TAPE_POWI) ok = t.dst < S && operand_ok(t.a);
    else if (op < MCM_TAPE_BINARY) ok = op < MCM_U_COUNT && t.dst < S && operand_ok(t.a);
    else if (op < MCM_TAPE_TERNARY)
      ok = (op - MCM_TAPE_BINARY) < MCM_B_COUNT && (op - MCM_TAPE_BINARY) != MCM_B_RATDIV && t.dst < S &&
           operand_ok(t.a) && operand_ok(t.b);
    else if (op < MCM_TAPE_LOAD)
      ok = (op - MCM_TAPE_TERNARY) < MCM_T_COUNT && t.dst < S && operand_ok(t.a) && operand_ok(t.b) && operand_ok(t.c);
    else ok = false;
    if (!ok) return fail(MCM_ERR_INVALID, "mcm_chain: invalid instruction %d (op 0x%02x)", i, op);
  }
  return MCM_OK;
}

template <typename T>
static int run_chain(const ChainParams& p, bool stats, double* stats_dev) {
  Ctx& c = ctx();
  constexpr int V = Pack<T>::N;
  static int occ[2] = {0, 0};
  if (occ[0] == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_chain<T, false>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_chain<T, true>, kThreads, 0);
    if (occ[0] < 1) occ[0] = 1;
    if (occ[1] < 1) occ[1] = 1;
  }
  const size_t nvec = (p.n + V - 1) / V;
  int grid = stream_grid(nvec, 1, occ[stats ? 1 : 0]);
  const size_t max_grid = c.scratch_doubles / 8;
  if ((size_t)grid > max_grid) grid = (int)max_grid;
  Mom* partials = reinterpret_cast<Mom*>(c.d_scratch);
  if (stats)
    k_chain<T, true><<<grid, kThreads, 0, c.stream>>>(p, partials, c.d_tickets, stats_dev, c.d_flags);
  else
    k_chain<T, false><<<grid, kThreads, 0, c.stream>>>(p, partials, c.d_tickets, stats_dev, c.d_flags);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_chain(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                         int reps, const double* consts, int nconst, const void* const* inputs, int nin,
                         void* const* outputs, int nout, size_t n, int stats_slot, double* stats_out_host) {
  MCM_NEED_INIT();
  if (!tape && ntape) return fail(MCM_ERR_INVALID, "mcm_chain: null tape");
  if (nconst < 0 || nconst > MCM_TAPE_MAX_CONST || nin < 0 || nin > MCM_TAPE_MAX_IN || nout < 0 ||
      nout > MCM_TAPE_MAX_OUT)
    return fail(MCM_ERR_INVALID, "mcm_chain: nconst/nin/nout out of range");
  if (stats_slot >= S) return fail(MCM_ERR_INVALID, "mcm_chain: stats_slot out of range");
  if (stats_slot >= 0 && !stats_out_host) return fail(MCM_ERR_INVALID, "mcm_chain: stats requested without output");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_chain: bad dtype %d", dtype);
  ChainParams p;
  memset(&p, 0, sizeof(p));
  if (ntape > MCM_TAPE_MAX_INS || ntape < 0) return fail(MCM_ERR_INVALID, "mcm_chain: ntape out of range");
  for (int i = 0; i < ntape; ++i) p.tape[i] = tape[i];
  for (int i = 0; i < nconst; ++i) p.consts[i] = consts[i];
  for (int i = 0; i < nin; ++i) {
    if (!inputs[i] && n) return fail(MCM_ERR_INVALID, "mcm_chain: null input %d", i);
    if (reinterpret_cast<uintptr_t>(inputs[i]) & 15u) return fail(MCM_ERR_INVALID, "mcm_chain: input %d not 16-byte aligned", i);
    p.in[i] = inputs[i];
  }
  for (int i = 0; i < nout; ++i) {
    p.out[i] = outputs ? outputs[i] : nullptr;
    if (reinterpret_cast<uintptr_t>(p.out[i]) & 15u) return fail(MCM_ERR_INVALID, "mcm_chain: output %d not 16-byte aligned", i);
  }
  p.n = n; p.ntape = ntape; p.loop_begin = loop_begin; p.loop_end = loop_end; p.reps = reps;
  p.stats_slot = stats_slot < 0 ? 0 : stats_slot;
  int rc = validate(p, nconst, nin, nout);
  if (rc) return rc;
  const bool stats = stats_slot >= 0;
  if (n == 0) {
    if (stats) { stats_out_host[0] = 0; stats_out_host[1] = NAN; stats_out_host[2] = 0;
                 stats_out_host[3] = INFINITY; stats_out_host[4] = -INFINITY; stats_out_host[5] = 0; }
    return MCM_OK;
  }
  double* slot = moments_result_slot();
  rc = (dtype == MCM_F64) ? run_chain<double>(p, stats, slot) : run_chain<float>(p, stats, slot);
  if (rc) return rc;
  if (stats) return read_back(stats_out_host, slot, 6);
  return MCM_OK;
}
