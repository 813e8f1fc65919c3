// chain.cu — fused evaluation of a whole chain of registered primitives.
//
// The reference evaluates `sin(a)*a^2 + exp(-a)` as five separate broadcasts,
// each allocating a Vector{T}(N) (src/register_primitive.jl:44-58,94-97).  At
// N = 1e8 that is 10 x 16-24 B/particle of HBM traffic for the canonical chain.
// Here the chain is a tape for a tiny register machine with 16 value slots per
// particle: one launch reads each input once, writes each requested output once
// and optionally folds the result into the one-pass moments of mcm_moments.cuh
// (pmean/pstd, src/particles.jl:302-323), so y never has to exist in HBM when
// only statistics are wanted.  A loop section [loop_begin, loop_end) x reps
// allows in-register time stepping (RK4 on the pendulum, examples/ode_sim.jl:20-25).
//
// Two execution engines behind the same mcm_chain entry point:
//  * REGISTERED chains (chain_registry.def): the analogue of `register_primitive`,
//    which `eval`s new methods at load time (src/register_primitive.jl:19-82) —
//    here a tape listed in the registry is compiled ahead of time into a
//    straight-line kernel by template recursion over the constexpr tape (slots
//    live in registers, no dispatch cost).
//  * any other tape: an interpreter whose slot file lives in shared memory
//    (one 128-bit LDS/STS per operand, one jump-table dispatch per instruction).
#include <string.h>
#include "mcm_chain_device.cuh"

namespace mcm {

// ---------------------------------------------------------------------------
// engine 1: registered (ahead-of-time specialised) chains
// ---------------------------------------------------------------------------
struct SIns { int op, dst, a, b, c; };

template <int IDX, typename T, int V>
__device__ __forceinline__ void sget(const T (&r)[S][V], const ChainParams& p, T (&x)[V]) {
  if constexpr ((IDX & MCM_TAPE_CONST) != 0) {
    const T c = (T)p.consts[IDX & (MCM_TAPE_MAX_CONST - 1)];
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = c;
  } else {
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = r[IDX][v];
  }
}

// W packs per thread (pack indices j[0..W)): W independent dependency chains per lane-pair for the FP64 pipe
template <int W> struct PackIdx { size_t j[W]; int nvalid[W]; };

template <class TP, int PC, int END, typename T, int V, int W, bool EXACT>
__device__ __forceinline__ void exec_static(T (&r)[S][V], const ChainParams& p, const PackIdx<W>& ix, int nvalid, bool& bad,
                                            unsigned& susp) {
  constexpr int PV = V / W;
  if constexpr (PC < END) {
    constexpr SIns I = TP::ins[PC];
    T x[V], y[V], z[V];
    if constexpr (I.op == MCM_TAPE_LOAD) {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        T t[PV];
        load_input<T, PV>(p.in[I.a], ix.j[w], ix.nvalid[w], t);
#pragma unroll
        for (int v = 0; v < PV; ++v) r[I.dst][w * PV + v] = t[v];
      }
    } else if constexpr (I.op == MCM_TAPE_STORE) {
      sget<I.a, T, V>(r, p, x);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        T t[PV];
#pragma unroll
        for (int v = 0; v < PV; ++v) t[v] = x[w * PV + v];
        if (ix.nvalid[w] > 0) store_output<T, PV>(p.out[I.dst], ix.j[w], ix.nvalid[w], t);
      }
    } else if constexpr (I.op == MCM_TAPE_MOV) {
      sget<I.a, T, V>(r, p, x);
#pragma unroll
      for (int v = 0; v < V; ++v) r[I.dst][v] = x[v];
    } else {
      sget<I.a, T, V>(r, p, x);
      if constexpr (I.op >= MCM_TAPE_BINARY && I.op < MCM_TAPE_LOAD) sget<I.b, T, V>(r, p, y);
      if constexpr (I.op >= MCM_TAPE_TERNARY && I.op < MCM_TAPE_LOAD) sget<I.c, T, V>(r, p, z);
      apply_op<I.op, T, V, EXACT>(x, y, z, (int)(signed char)I.b, nvalid, bad, susp);
#pragma unroll
      for (int v = 0; v < V; ++v) r[I.dst][v] = x[v];
    }
    exec_static<TP, PC + 1, END, T, V, W, EXACT>(r, p, ix, nvalid, bad, susp);
  }
}

#ifndef MCM_CHAIN_W
#define MCM_CHAIN_W 1   // packs per thread in the registered-chain kernel
#endif
#ifndef MCM_CHAIN_PREFETCH
#define MCM_CHAIN_PREFETCH 0   // 0 none (default: measured no effect), 1 prefetch.global.L1, 2 prefetch.global.L2 of the next inputs
#endif

// The register budget (64 at 4 CTAs/SM) leaves no room to hold the next iteration's inputs, and the pc samples put
// 18 % of all warp stalls on the first use of the loaded pack: ask the cache for it one iteration ahead instead.
template <class TP, int PC, typename T>
__device__ __forceinline__ void prefetch_static(const ChainParams& p, size_t jn) {
  if constexpr (PC < TP::n) {
    constexpr SIns I = TP::ins[PC];
    if constexpr (I.op == MCM_TAPE_LOAD) {
      const char* a = reinterpret_cast<const char*>(p.in[I.a]) + jn * sizeof(Pack<T>);
#if MCM_CHAIN_PREFETCH == 1
      asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
#elif MCM_CHAIN_PREFETCH == 2
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
#endif
    }
    prefetch_static<TP, PC + 1, T>(p, jn);
  }
}

// Cold path of the domain check: the same pack once more with the per-lane test (stores repeat the same values).
template <class TP, typename T, int V, int W>
__device__ __noinline__ bool replay_exact(const ChainParams& p, size_t j, size_t stride) {
  constexpr int PV = V / W;
  const size_t n = p.n, nvec = (n + PV - 1) / PV;
  PackIdx<W> ix;
  int nvalid = 0;
  bool bad = false;
#pragma unroll
  for (int w = 0; w < W; ++w) {
    ix.j[w] = j + (size_t)w * stride;
    ix.nvalid[w] = ix.j[w] >= nvec ? 0 : ((ix.j[w] + 1) * PV <= n ? PV : (int)(n - ix.j[w] * PV));
    nvalid += ix.nvalid[w];
  }
  T r[S][V];
#pragma unroll
  for (int k = 0; k < S; ++k)
#pragma unroll
    for (int v = 0; v < V; ++v) r[k][v] = T(0);
  unsigned unused = 0;
  exec_static<TP, 0, TP::loop_begin, T, V, W, true>(r, p, ix, nvalid, bad, unused);
  if constexpr (TP::loop_end > TP::loop_begin) {
    for (int rep = 0; rep < p.reps; ++rep) exec_static<TP, TP::loop_begin, TP::loop_end, T, V, W, true>(r, p, ix, nvalid, bad, unused);
  }
  exec_static<TP, TP::loop_end, TP::n, T, V, W, true>(r, p, ix, nvalid, bad, unused);
  return bad;
}

template <class TP, typename T, bool STATS>
__global__ void __launch_bounds__(kThreads, MCM_CHAIN_MIN_CTAS)
k_chain_static(const __grid_constant__ ChainParams p, Mom* partials, unsigned int* ticket, double* stats_out,
               uint32_t* flags, const __grid_constant__ MomExchange px) {
  constexpr int W = MCM_CHAIN_W;
  constexpr int PV = Pack<T>::N;
  constexpr int V = W * PV;
  const size_t n = p.n;
  const size_t nvec = (n + PV - 1) / PV;
  const size_t stride = (size_t)gridDim.x * kThreads;
  bool bad = false;
  const int ragged_n = (int)(n % PV);
  const size_t ragged_pack = ragged_n ? nvec - 1 : ~(size_t)0;
  MomStream acc = ms_empty();
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < nvec; j += W * stride) {
    PackIdx<W> ix;
    int nvalid = 0;   // valid lanes form a prefix of the V lanes of this thread only when W == 1; used for the domain flag
    if constexpr (W == 1) {
      ix.j[0] = j;                                       // only the very last pack can be ragged
      ix.nvalid[0] = nvalid = (j == ragged_pack) ? ragged_n : PV;
    } else {
#pragma unroll
      for (int w = 0; w < W; ++w) {
        ix.j[w] = j + (size_t)w * stride;
        ix.nvalid[w] = ix.j[w] >= nvec ? 0 : ((ix.j[w] + 1) * PV <= n ? PV : (int)(n - ix.j[w] * PV));
        nvalid += ix.nvalid[w];
      }
    }
#if MCM_CHAIN_PREFETCH
    if (j + W * stride < nvec) {
#pragma unroll
      for (int w = 0; w < W; ++w)
        if (j + (W + w) * stride < nvec) prefetch_static<TP, 0, T>(p, j + (W + w) * stride);
    }
#endif
    T r[S][V];
#pragma unroll
    for (int k = 0; k < S; ++k)
#pragma unroll
      for (int v = 0; v < V; ++v) r[k][v] = T(0);
    unsigned susp = 0;
    exec_static<TP, 0, TP::loop_begin, T, V, W, false>(r, p, ix, nvalid, bad, susp);
    if constexpr (TP::loop_end > TP::loop_begin) {
      for (int rep = 0; rep < p.reps; ++rep) exec_static<TP, TP::loop_begin, TP::loop_end, T, V, W, false>(r, p, ix, nvalid, bad, susp);
    }
    exec_static<TP, TP::loop_end, TP::n, T, V, W, false>(r, p, ix, nvalid, bad, susp);
    if (nan_key_hit<T>(susp)) bad |= replay_exact<TP, T, V, W>(p, j, stride);   // rare: a checked primitive returned NaN
    if constexpr (STATS) {
      T x[V];
      // stats slot is a runtime choice among 8 registers: resolved by a uniform switch once per vector
      switch (p.stats_slot) {
#define MCM_SS(K) case K: _Pragma("unroll") for (int v = 0; v < V; ++v) x[v] = r[K][v]; break;
        MCM_SS(0) MCM_SS(1) MCM_SS(2) MCM_SS(3) MCM_SS(4) MCM_SS(5) MCM_SS(6) MCM_SS(7) MCM_SS(8) MCM_SS(9)
        MCM_SS(10) MCM_SS(11) MCM_SS(12) MCM_SS(13) MCM_SS(14) default: MCM_SS(15)
#undef MCM_SS
      }
#pragma unroll
      for (int w = 0; w < W; ++w) {
        T t[PV];
#pragma unroll
        for (int v = 0; v < PV; ++v) t[v] = x[w * PV + v];
        if (ix.nvalid[w] > 0) stats_push<T, PV>(acc, t, ix.nvalid[w]);
      }
    }
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
  if constexpr (STATS) {
    Mom m = mom_block_reduce(ms_finish(acc));
    mom_grid_finish(m, partials, ticket, stats_out, &px);
  }
}

// ---- the registry ------------------------------------------------------------
#define TU(op) (MCM_TAPE_UNARY + (op))
#define TB(op) (MCM_TAPE_BINARY + (op))
#define TT(op) (MCM_TAPE_TERNARY + (op))
#define KC(k) (MCM_TAPE_CONST | (k))
#define MCM_CHAIN_BEGIN(NAME, LB, LE) \
  struct NAME {                       \
    static constexpr int loop_begin_decl = (LB), loop_end_decl = (LE); \
    static constexpr SIns ins[] = {
#define MCM_CHAIN_END(NAME)                                                      \
    };                                                                           \
    static constexpr int n = (int)(sizeof(ins) / sizeof(ins[0]));                \
    static constexpr int loop_begin = loop_begin_decl < 0 ? n : loop_begin_decl; \
    static constexpr int loop_end = loop_end_decl < 0 ? n : loop_end_decl;       \
  };
#include "chain_registry.def"
#undef MCM_CHAIN_BEGIN
#undef MCM_CHAIN_END

template <class TP>
static bool tape_matches(const ChainParams& p) {
  if (p.ntape != TP::n || p.loop_begin != TP::loop_begin || p.loop_end != TP::loop_end) return false;
  for (int i = 0; i < TP::n; ++i) {
    const SIns& s = TP::ins[i];
    const mcm_tape_ins& t = p.tape[i];
    if (t.op != s.op || t.dst != s.dst || t.a != s.a) return false;
    const bool uses_b = (s.op >= MCM_TAPE_BINARY && s.op < MCM_TAPE_LOAD) || s.op == MCM_TAPE_POWI;
    const bool uses_c = (s.op >= MCM_TAPE_TERNARY && s.op < MCM_TAPE_LOAD);
    if (uses_b && t.b != (uint8_t)s.b) return false;
    if (uses_c && t.c != s.c) return false;
  }
  return true;
}

MomExchange mom_exchange_none();
template <class TP, typename T>
static int launch_static(const ChainParams& p, bool stats, double* stats_dev, const MomExchange& px) {
  Ctx& c = ctx();
  constexpr int V = Pack<T>::N;
  static int occ[2] = {0, 0};
  if (occ[0] == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_chain_static<TP, T, false>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_chain_static<TP, T, true>, kThreads, 0);
    if (occ[0] < 1) occ[0] = 1;
    if (occ[1] < 1) occ[1] = 1;
  }
  const size_t nvec = (p.n + V - 1) / V;
  int grid = stream_grid(nvec, 1, occ[stats ? 1 : 0]);
  const size_t max_grid = c.scratch_doubles / 8;
  if ((size_t)grid > max_grid) grid = (int)max_grid;
  Mom* partials = reinterpret_cast<Mom*>(c.d_scratch);
  if (stats)
    k_chain_static<TP, T, true><<<grid, kThreads, 0, c.stream>>>(p, partials, c.d_tickets, stats_dev, c.d_flags, px);
  else
    k_chain_static<TP, T, false><<<grid, kThreads, 0, c.stream>>>(p, partials, c.d_tickets, stats_dev, c.d_flags, px);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

// returns -1 if no registered chain matches
template <typename T>
static int try_registered(const ChainParams& p, bool stats, double* stats_dev, const MomExchange& px) {
#define MCM_CHAIN_USE(NAME) \
  if (tape_matches<NAME>(p)) return launch_static<NAME, T>(p, stats, stats_dev, px);
  MCM_REGISTERED_CHAINS(MCM_CHAIN_USE)
#undef MCM_CHAIN_USE
  return -1;
}

// ---------------------------------------------------------------------------
// engine 2: interpreter, slot file in shared memory
// ---------------------------------------------------------------------------
// smem layout: slot s of thread t is the Pack at sm[s * kThreads + t] -> every
// quarter-warp touches 128 contiguous bytes: conflict-free 128-bit LDS/STS.
template <typename T, int V>
__device__ __forceinline__ void iget(const Pack<T>* sm, const ChainParams& p, int idx, T (&x)[V]) {
  if (idx & MCM_TAPE_CONST) {
    const T c = (T)p.consts[idx & (MCM_TAPE_MAX_CONST - 1)];
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = c;
  } else {
    const Pack<T> t = sm[(idx & (S - 1)) * kThreads + threadIdx.x];
#pragma unroll
    for (int v = 0; v < V; ++v) x[v] = t.v[v];
  }
}
template <typename T, int V>
__device__ __forceinline__ void iput(Pack<T>* sm, int idx, const T (&x)[V]) {
  Pack<T> t;
#pragma unroll
  for (int v = 0; v < V; ++v) t.v[v] = x[v];
  sm[(idx & (S - 1)) * kThreads + threadIdx.x] = t;
}

template <typename T, int V>
__device__ __forceinline__ void exec_range(Pack<T>* sm, const ChainParams& p, int pc0, int pc1, size_t j,
                                           int nvalid, bool& bad) {
  unsigned unused_susp = 0;
  for (int pc = pc0; pc < pc1; ++pc) {
    const mcm_tape_ins ins = p.tape[pc];
    const int op = ins.op;
    T x[V], y[V], z[V];
    if (op == MCM_TAPE_LOAD) {
      load_input<T, V>(p.in[ins.a & (MCM_TAPE_MAX_IN - 1)], j, nvalid, x);
      iput<T, V>(sm, ins.dst, x);
      continue;
    }
    iget<T, V>(sm, p, ins.a, x);
    if (op == MCM_TAPE_STORE) {
      store_output<T, V>(p.out[ins.dst & (MCM_TAPE_MAX_OUT - 1)], j, nvalid, x);
      continue;
    }
    if (op >= MCM_TAPE_BINARY && op < MCM_TAPE_LOAD) iget<T, V>(sm, p, ins.b, y);
    if (op >= MCM_TAPE_TERNARY && op < MCM_TAPE_LOAD) iget<T, V>(sm, p, ins.c, z);
    const int imm = (int)(signed char)ins.b;
#define MCM_OPC(OPC) \
  case OPC: apply_op<OPC, T, V, true>(x, y, z, imm, nvalid, bad, unused_susp); break;
#define MCM_U(OP) MCM_OPC(MCM_TAPE_UNARY + OP)
#define MCM_B(OP) MCM_OPC(MCM_TAPE_BINARY + OP)
#define MCM_T(OP) MCM_OPC(MCM_TAPE_TERNARY + OP)
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
      MCM_U(MCM_U_ROUND_TIESAWAY) MCM_U(MCM_U_ROUND_TIESUP) MCM_U(MCM_U_ROUND_FROMZERO)
      MCM_U(MCM_U_REM2PI_NEAREST) MCM_U(MCM_U_REM2PI_TOZERO) MCM_U(MCM_U_REM2PI_DOWN) MCM_U(MCM_U_REM2PI_UP)
      MCM_B(MCM_B_ADD) MCM_B(MCM_B_SUB) MCM_B(MCM_B_MUL) MCM_B(MCM_B_DIV) MCM_B(MCM_B_POW)
      MCM_B(MCM_B_MAX) MCM_B(MCM_B_MIN) MCM_B(MCM_B_MOD) MCM_B(MCM_B_MOD1) MCM_B(MCM_B_ATAN2)
      MCM_B(MCM_B_ATAND2) MCM_B(MCM_B_ADD_SUM) MCM_B(MCM_B_HYPOT) MCM_B(MCM_B_REM) MCM_B(MCM_B_FDIV)
      MCM_B(MCM_B_REM_NEAREST) MCM_B(MCM_B_REM_DOWN) MCM_B(MCM_B_REM_UP) MCM_B(MCM_B_REM_FROMZERO)
      MCM_B(MCM_B_DIV_NEAREST) MCM_B(MCM_B_DIV_DOWN) MCM_B(MCM_B_DIV_UP) MCM_B(MCM_B_DIV_FROMZERO)
      MCM_T(MCM_T_CLAMP) MCM_T(MCM_T_MULADD) MCM_T(MCM_T_MUL_ADD)
      MCM_OPC(MCM_TAPE_POWI)
      default: break;  // MOV
    }
#undef MCM_U
#undef MCM_B
#undef MCM_T
#undef MCM_OPC
    iput<T, V>(sm, ins.dst, x);
  }
}

template <typename T, bool STATS>
__global__ void __launch_bounds__(kThreads)
k_chain(const __grid_constant__ ChainParams p, Mom* partials, unsigned int* ticket, double* stats_out,
        uint32_t* flags, const __grid_constant__ MomExchange px) {
  constexpr int V = Pack<T>::N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Pack<T>* sm = reinterpret_cast<Pack<T>*>(smem_raw);
  const size_t n = p.n;
  const size_t nvec = (n + V - 1) / V;
  const size_t stride = (size_t)gridDim.x * kThreads;
  bool bad = false;
  Mom m = mom_empty();
  {
    T zero[V];
#pragma unroll
    for (int v = 0; v < V; ++v) zero[v] = T(0);
#pragma unroll
    for (int k = 0; k < S; ++k) iput<T, V>(sm, k, zero);
  }
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < nvec; j += stride) {
    const int nvalid = (j + 1) * V <= n ? V : (int)(n - j * V);
    exec_range<T, V>(sm, p, 0, p.loop_begin, j, nvalid, bad);
    for (int rep = 0; rep < p.reps; ++rep) exec_range<T, V>(sm, p, p.loop_begin, p.loop_end, j, nvalid, bad);
    exec_range<T, V>(sm, p, p.loop_end, p.ntape, j, nvalid, bad);
    if constexpr (STATS) {
      T x[V];
      iget<T, V>(sm, p, p.stats_slot, x);
      stats_push<T, V>(m, x, nvalid);
    }
  }
  if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
  if constexpr (STATS) {
    m = mom_block_reduce(m);
    mom_grid_finish(m, partials, ticket, stats_out, &px);
  }
}

double* moments_result_slot();
MomExchange mom_exchange_none();
int mom_exchange_next(MomExchange* px);
int ring_acquire(double** dev_slot, unsigned long long* ticket);
int ring_publish(unsigned long long ticket);
int ring_wait(unsigned long long ticket, double* out6);
unsigned long long ring_last_ticket();
double* ring_host_slot(unsigned long long ticket);
int read_back(double* host_dst, const double* dev_src, size_t count);

int chain_validate(const ChainParams& p, int nconst, int nin, int nout) {
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

static int g_last_engine = 0;  // 1 = registered/static, 2 = interpreter, 3 = generated at run time (introspection for tests)
static int g_force_engine = 0; // 0 = automatic, 1 = registered (no generation), 2 = interpreter, 3 = always generate
int jit_try_launch(const ChainParams& p, int nconst, bool f64, bool stats, double* stats_dev, const MomExchange& px,
                   bool force);

template <typename T>
static int run_chain(const ChainParams& p, bool stats, double* stats_dev, bool allow_registered, const MomExchange& px) {
  if (allow_registered && g_force_engine != 1 && g_force_engine != 2) {
    // engine 3: a kernel generated from this tape (chain_jit.cu) — cached, or compiled now if the launch is large
    const int j = jit_try_launch(p, p.nconst, sizeof(T) == 8, stats, stats_dev, px, g_force_engine == 3);
    if (j > 0) { g_last_engine = 3; return MCM_OK; }
    if (j < 0) return -j;
  }
  int rc = (allow_registered && g_force_engine != 2) ? try_registered<T>(p, stats, stats_dev, px) : -1;
  if (rc >= 0) { g_last_engine = 1; return rc; }
  g_last_engine = 2;
  Ctx& c = ctx();
  constexpr int V = Pack<T>::N;
  constexpr size_t kSmem = (size_t)S * kThreads * sizeof(Pack<T>);  // 64 KiB: 16 slots x 256 threads x 16 B
  static int occ[2] = {0, 0};
  if (occ[0] == 0) {
    MCM_CUDA(cudaFuncSetAttribute(k_chain<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    MCM_CUDA(cudaFuncSetAttribute(k_chain<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_chain<T, false>, kThreads, kSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_chain<T, true>, kThreads, kSmem);
    if (occ[0] < 1) occ[0] = 1;
    if (occ[1] < 1) occ[1] = 1;
  }
  const size_t nvec = (p.n + V - 1) / V;
  int grid = stream_grid(nvec, 1, occ[stats ? 1 : 0]);
  const size_t max_grid = c.scratch_doubles / 8;
  if ((size_t)grid > max_grid) grid = (int)max_grid;
  Mom* partials = reinterpret_cast<Mom*>(c.d_scratch);
  if (stats)
    k_chain<T, true><<<grid, kThreads, kSmem, c.stream>>>(p, partials, c.d_tickets, stats_dev, c.d_flags, px);
  else
    k_chain<T, false><<<grid, kThreads, kSmem, c.stream>>>(p, partials, c.d_tickets, stats_dev, c.d_flags, px);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

// 1 if the last mcm_chain call ran a registered (ahead-of-time specialised) kernel, 2 if interpreted
extern "C" int mcm_chain_last_engine(void) { return g_last_engine; }

static int g_force_interpreter = 0;
extern "C" int mcm_chain_force_interpreter(int on) { g_force_interpreter = on; return MCM_OK; }
// 0 = automatic (generated kernel if cached or the launch is large, else registered, else interpreter), 1 = registered
// chains + interpreter only, 2 = interpreter only, 3 = always generate (compiles on first use whatever the size)
extern "C" int mcm_chain_force_engine(int engine) {
  if (engine < 0 || engine > 3) return fail(MCM_ERR_INVALID, "mcm_chain_force_engine: 0..3");
  g_force_engine = engine;
  return MCM_OK;
}

// Validate + launch; statistics (if stats_slot >= 0) land in the 6 doubles at stats_dev.  No sync.
static int chain_launch(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                        int reps, const double* consts, int nconst, const void* const* inputs, int nin,
                        void* const* outputs, int nout, size_t n, int stats_slot, double* stats_dev,
                        bool sharded = false) {
  if (!tape && ntape) return fail(MCM_ERR_INVALID, "mcm_chain: null tape");
  if (nconst < 0 || nconst > MCM_TAPE_MAX_CONST || nin < 0 || nin > MCM_TAPE_MAX_IN || nout < 0 ||
      nout > MCM_TAPE_MAX_OUT)
    return fail(MCM_ERR_INVALID, "mcm_chain: nconst/nin/nout out of range");
  if (stats_slot >= S) return fail(MCM_ERR_INVALID, "mcm_chain: stats_slot out of range");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_chain: bad dtype %d", dtype);
  if (ntape > MCM_TAPE_MAX_INS || ntape < 0) return fail(MCM_ERR_INVALID, "mcm_chain: ntape out of range");
  ChainParams p;
  memset(&p, 0, sizeof(p));
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
  p.nconst = nconst;
  int rc = chain_validate(p, nconst, nin, nout);
  if (rc) return rc;
  const bool stats = stats_slot >= 0;
  MomExchange px = mom_exchange_none();
  if (sharded && stats) {             // collective: every rank takes the next epoch, an empty shard included
    rc = mom_exchange_next(&px);
    if (rc) return rc;
  }
  if (n == 0 && !(sharded && stats && px.world > 1)) return MCM_OK;
  const bool allow = !g_force_interpreter;
  return (dtype == MCM_F64) ? run_chain<double>(p, stats, stats_dev, allow, px) : run_chain<float>(p, stats, stats_dev, allow, px);
}

extern "C" int mcm_comm_info(int* nranks, int* rank);
static void empty_stats(double* s) {
  s[0] = 0; s[1] = NAN; s[2] = 0; s[3] = INFINITY; s[4] = -INFINITY; s[5] = 0;
}

// Statistics travel through the result ring (reduce.cu): the last CTA writes them into mapped pinned host memory;
// with stats_out_host == NULL the call returns at once and the caller collects them later by ticket.
static int chain_entry(bool sharded, int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                       int reps, const double* consts, int nconst, const void* const* inputs, int nin,
                       void* const* outputs, int nout, size_t n, int stats_slot, double* stats_out_host) {
  if (stats_slot < 0)
    return chain_launch(dtype, tape, ntape, loop_begin, loop_end, reps, consts, nconst, inputs, nin, outputs, nout, n,
                        stats_slot, nullptr, false);
  double* slot = nullptr;
  unsigned long long ticket = 0;
  int rc = ring_acquire(&slot, &ticket);
  if (rc) return rc;
  int world = 1, rank = 0;
  if (sharded) mcm_comm_info(&world, &rank);
  if (n == 0 && world == 1) {
    empty_stats(ring_host_slot(ticket));
  } else {
    rc = chain_launch(dtype, tape, ntape, loop_begin, loop_end, reps, consts, nconst, inputs, nin, outputs, nout, n,
                      stats_slot, slot, sharded);
    if (rc) return rc;
  }
  rc = ring_publish(ticket);
  if (rc) return rc;
  if (stats_out_host) return ring_wait(ticket, stats_out_host);
  return MCM_OK;
}

extern "C" int mcm_chain(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                         int reps, const double* consts, int nconst, const void* const* inputs, int nin,
                         void* const* outputs, int nout, size_t n, int stats_slot, double* stats_out_host) {
  MCM_NEED_INIT();
  return chain_entry(false, dtype, tape, ntape, loop_begin, loop_end, reps, consts, nconst, inputs, nin, outputs, nout, n,
                     stats_slot, stats_out_host);
}

// The same chain on a particle-sharded variable (one process per GPU): the statistics are those of the GLOBAL
// particle set, merged across ranks inside the kernel's epilogue over peer memory (mcm_moments.cuh: mom_exchange).
// Collective when stats_slot >= 0: every rank of the communicator must call it, in the same order.
extern "C" int mcm_chain_sharded(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                                 int reps, const double* consts, int nconst, const void* const* inputs, int nin,
                                 void* const* outputs, int nout, size_t n, int stats_slot, double* stats_out_host) {
  MCM_NEED_INIT();
  return chain_entry(true, dtype, tape, ntape, loop_begin, loop_end, reps, consts, nconst, inputs, nin, outputs, nout, n,
                     stats_slot, stats_out_host);
}

// Fetch the statistics of the most recent mcm_chain launched with stats_out_host == NULL (waits for that launch only).
extern "C" int mcm_chain_stats_fetch(double* stats_out_host) {
  MCM_NEED_INIT();
  if (!stats_out_host) return fail(MCM_ERR_INVALID, "mcm_chain_stats_fetch: null pointer");
  return ring_wait(ring_last_ticket(), stats_out_host);
}

// Ticket of the most recent statistics-producing mcm_chain / mcm_chain_sharded call, and a wait for a given ticket
// (one of the last 8): launch k+1 can be queued before the statistics of launch k are collected.
extern "C" uint64_t mcm_chain_ticket(void) { return ring_last_ticket(); }
extern "C" int mcm_chain_stats_wait(uint64_t ticket, double* stats_out_host) {
  MCM_NEED_INIT();
  if (!stats_out_host) return fail(MCM_ERR_INVALID, "mcm_chain_stats_wait: null pointer");
  return ring_wait(ticket, stats_out_host);
}

// Host-buffer entry point: the whole chain with inputs/outputs in HOST memory (what a Julia caller holding
// ordinary Vector{T}s would pass).  The particle range is cut into chunks; chunk c+1 is copied host->device on
// a copy stream while chunk c computes, results stream back on a third stream; per-chunk moments are merged
// on the host in chunk order.  Pinned host memory (mcm_host_alloc) makes the copies truly asynchronous.
extern "C" int mcm_moments_merge(const double* a, const double* b, double* out);

extern "C" int mcm_chain_host(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                              int reps, const double* consts, int nconst, const void* const* host_inputs, int nin,
                              void* const* host_outputs, int nout, size_t n, int stats_slot,
                              double* stats_out_host) {
  MCM_NEED_INIT();
  if (nin < 1 || nin > MCM_TAPE_MAX_IN || nout < 0 || nout > MCM_TAPE_MAX_OUT)
    return fail(MCM_ERR_INVALID, "mcm_chain_host: nin/nout out of range");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_chain_host: bad dtype %d", dtype);
  if (stats_slot >= 0 && !stats_out_host) return fail(MCM_ERR_INVALID, "mcm_chain_host: stats requested without output");
  for (int i = 0; i < nin; ++i)
    if (!host_inputs || (!host_inputs[i] && n)) return fail(MCM_ERR_INVALID, "mcm_chain_host: null input %d", i);
  if (n == 0) { if (stats_slot >= 0) empty_stats(stats_out_host); return MCM_OK; }
  Ctx& c = ctx();
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  const size_t chunk = (size_t)1 << 23;  // 8 Mi particles per chunk (64 MiB of Float64 per stream buffer)
  const size_t nchunks = (n + chunk - 1) / chunk;
  static cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  if (!s_h2d) {
    MCM_CUDA(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    MCM_CUDA(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
  }
  const size_t cbytes = (n < chunk ? n : chunk) * esz;
  void* dbuf_in[2][MCM_TAPE_MAX_IN];
  void* dbuf_out[2][MCM_TAPE_MAX_OUT];
  for (int b = 0; b < 2; ++b) {
    for (int i = 0; i < nin; ++i) MCM_CUDA(cudaMallocAsync(&dbuf_in[b][i], cbytes, c.stream));
    for (int i = 0; i < nout; ++i) {
      dbuf_out[b][i] = nullptr;
      if (host_outputs && host_outputs[i]) MCM_CUDA(cudaMallocAsync(&dbuf_out[b][i], cbytes, c.stream));
    }
  }
  double* d_stats = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_stats, nchunks * 6 * sizeof(double), c.stream));
  cudaEvent_t ev_in[2], ev_comp[2], ev_out[2], ev_alloc;
  for (int b = 0; b < 2; ++b) {
    MCM_CUDA(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
    MCM_CUDA(cudaEventCreateWithFlags(&ev_comp[b], cudaEventDisableTiming));
    MCM_CUDA(cudaEventCreateWithFlags(&ev_out[b], cudaEventDisableTiming));
  }
  MCM_CUDA(cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming));
  MCM_CUDA(cudaEventRecord(ev_alloc, c.stream));
  MCM_CUDA(cudaStreamWaitEvent(s_h2d, ev_alloc, 0));
  MCM_CUDA(cudaStreamWaitEvent(s_d2h, ev_alloc, 0));
  int rc = MCM_OK;
  for (size_t k = 0; k < nchunks && rc == MCM_OK; ++k) {
    const int b = (int)(k & 1);
    const size_t off = k * chunk;
    const size_t len = (n - off < chunk) ? n - off : chunk;
    // the input buffer b is free once chunk k-2 has computed; the output buffer once its D2H finished
    if (k >= 2) {
      MCM_CUDA(cudaStreamWaitEvent(s_h2d, ev_comp[b], 0));
      MCM_CUDA(cudaStreamWaitEvent(c.stream, ev_out[b], 0));
    }
    for (int i = 0; i < nin; ++i)
      MCM_CUDA(cudaMemcpyAsync(dbuf_in[b][i], static_cast<const char*>(host_inputs[i]) + off * esz, len * esz,
                               cudaMemcpyHostToDevice, s_h2d));
    MCM_CUDA(cudaEventRecord(ev_in[b], s_h2d));
    MCM_CUDA(cudaStreamWaitEvent(c.stream, ev_in[b], 0));
    rc = chain_launch(dtype, tape, ntape, loop_begin, loop_end, reps, consts, nconst, dbuf_in[b], nin, dbuf_out[b],
                      nout, len, stats_slot, d_stats + 6 * k);
    if (rc) break;
    MCM_CUDA(cudaEventRecord(ev_comp[b], c.stream));
    bool any_out = false;
    for (int i = 0; i < nout; ++i)
      if (dbuf_out[b][i]) {
        if (!any_out) { MCM_CUDA(cudaStreamWaitEvent(s_d2h, ev_comp[b], 0)); any_out = true; }
        MCM_CUDA(cudaMemcpyAsync(static_cast<char*>(host_outputs[i]) + off * esz, dbuf_out[b][i], len * esz,
                                 cudaMemcpyDeviceToHost, s_d2h));
      }
    MCM_CUDA(cudaEventRecord(ev_out[b], s_d2h));
  }
  if (rc == MCM_OK && stats_slot >= 0) {
    double* h = new double[nchunks * 6];
    MCM_CUDA(cudaMemcpyAsync(h, d_stats, nchunks * 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    MCM_CUDA(cudaStreamSynchronize(c.stream));
    double acc[6];
    memcpy(acc, h, sizeof(acc));
    for (size_t k = 1; k < nchunks; ++k) {
      double t[6];
      mcm_moments_merge(acc, h + 6 * k, t);
      memcpy(acc, t, sizeof(t));
    }
    memcpy(stats_out_host, acc, sizeof(acc));
    delete[] h;
  }
  MCM_CUDA(cudaStreamSynchronize(s_d2h));
  MCM_CUDA(cudaStreamSynchronize(s_h2d));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  for (int b = 0; b < 2; ++b) {
    for (int i = 0; i < nin; ++i) MCM_CUDA(cudaFreeAsync(dbuf_in[b][i], c.stream));
    for (int i = 0; i < nout; ++i)
      if (dbuf_out[b][i]) MCM_CUDA(cudaFreeAsync(dbuf_out[b][i], c.stream));
    cudaEventDestroy(ev_in[b]); cudaEventDestroy(ev_comp[b]); cudaEventDestroy(ev_out[b]);
  }
  cudaEventDestroy(ev_alloc);
  MCM_CUDA(cudaFreeAsync(d_stats, c.stream));
  return rc;
}
