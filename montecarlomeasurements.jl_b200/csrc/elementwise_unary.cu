// elementwise_unary.cu — f(p) for the 35 register_primitive_single entries
// (src/particles.jl:185-190, bodies at src/register_primitive.jl:94-97 and
// src/sleefpirates.jl:7-16,26-35) plus round/floor/ceil/trunc/inv/sqr/cube
// (src/particles.jl:346-348,388-391), fill and convert (types.jl:57-63,
// particles.jl:333).
#include "mcm_map.cuh"
#include "mcm_math.cuh"
#include "mcm_bulk.cuh"
#include "mcm_chain_device.cuh"   // fast:: branch-free forms (sin / cos per warp, exp, inline sqrt / division) and their windows

namespace mcm {

// ---- Float64 primitives with a branch-free fast form ------------------------------------------------------------
// The generic map kernel evaluates math::unary_apply per value: for sin / cos / exp / sqrt / inv that is a range
// branch, a library fallback call and (sqrt, inv) the compiler's slow-path region per value, which keeps ptxas from
// interleaving the values of a thread (sin: 0.73 of the HBM rate).  Here a thread evaluates its NV values on the
// fast forms the generated chain kernels use (mcm_chain_device.cuh) — one basic block — and, if any argument left
// its window, redoes them with the exact per-value function: identical results (same arithmetic inside the windows).
template <int OP> struct FastUnary { static constexpr bool has = false; };
template <> struct FastUnary<MCM_U_SIN> {
  static constexpr bool has = true;
  template <int NV> static __device__ __forceinline__ void apply(const double (&x)[NV], double (&y)[NV], bool& ok) {
    fast::sincos_fast_v<NV>(x, y, 0, ok);
#pragma unroll
    for (int v = 0; v < NV; ++v) ok = ok && fast::abs_le(x[v], 0x410921f7u);      // |x| < 205887 = 2^17 pi/2
  }
};
template <> struct FastUnary<MCM_U_COS> {
  static constexpr bool has = true;
  template <int NV> static __device__ __forceinline__ void apply(const double (&x)[NV], double (&y)[NV], bool& ok) {
    fast::sincos_fast_v<NV>(x, y, 1, ok);
#pragma unroll
    for (int v = 0; v < NV; ++v) ok = ok && fast::abs_le(x[v], 0x410921f7u);
  }
};
// (exp stays on the generic map kernel: its range branch is cheap and 0.2645 ms beat every fast-form variant, 0.277-0.29)
template <> struct FastUnary<MCM_U_SQRT> {
  static constexpr bool has = true;
  template <int NV> static __device__ __forceinline__ void apply(const double (&x)[NV], double (&y)[NV], bool& ok) {
#pragma unroll
    for (int v = 0; v < NV; ++v) { y[v] = fast::sqrt_fast(x[v]); ok = ok && fast::pos_in(x[v], fast::kSqrtLo, fast::kSqrtHi); }
  }
};
template <> struct FastUnary<MCM_U_INV> {
  static constexpr bool has = true;
  template <int NV> static __device__ __forceinline__ void apply(const double (&x)[NV], double (&y)[NV], bool& ok) {
#pragma unroll
    for (int v = 0; v < NV; ++v) { y[v] = fast::div_fast(1.0, x[v]); ok = ok && fast::abs_in(x[v], fast::kDivLo, fast::kDivHi); }
  }
};

// The packs of the NEXT iteration are requested before this iteration's arithmetic starts: with 46-80 registers per
// thread only 3-4 CTAs are resident, and one iteration's loads alone (UNROLL x 16 B per thread) do not keep enough
// bytes in flight to cover the HBM latency.  Two register buffers used in ping-pong, the loop unrolled twice: copying
// "next" into "current" at the end of an iteration would wait for the load it was meant to hide (ncu: 53 % of the
// stall samples sat on that copy).
template <int UNROLL>
__device__ __forceinline__ void map_fast_load(Pack<double> (&b)[UNROLL], const double* __restrict__ in, size_t i, size_t stride, size_t nvec) {
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const size_t j = i + (size_t)u * stride;
    b[u].v[0] = 1.0; b[u].v[1] = 1.0;
    if (j < nvec) b[u] = ld_stream(reinterpret_cast<const Pack<double>*>(in) + j);
  }
}
template <int OP, int UNROLL>
__device__ __forceinline__ void map_fast_step(const Pack<double> (&b)[UNROLL], double* __restrict__ out, size_t i, size_t stride, size_t nvec, bool& bad) {
  using P = Pack<double>;
  constexpr int V = P::N, NV = UNROLL * V;
  double x[NV], y[NV];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
#pragma unroll
    for (int e = 0; e < V; ++e) x[u * V + e] = b[u].v[e];
  bool ok = true;
  FastUnary<OP>::template apply<NV>(x, y, ok);
  if (!ok) {                                  // rare: the exact per-value function (library fallbacks, DomainError test)
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      y[v] = math::unary_apply<OP, double>(x[v]);
      if constexpr (math::UnaryCanDomainError<OP>::value) bad |= is_nan_fast(y[v]) && !is_nan_fast(x[v]);
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const size_t j = i + (size_t)u * stride;
    if (j < nvec) {
      P r;
#pragma unroll
      for (int e = 0; e < V; ++e) r.v[e] = y[u * V + e];
      st_stream(reinterpret_cast<P*>(out) + j, r);
    }
  }
}
template <int OP, int UNROLL, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_map_fast(const double* __restrict__ in, double* __restrict__ out, size_t n, uint32_t* __restrict__ flags) {
  using P = Pack<double>;
  constexpr int V = P::N;
  const size_t nvec = n / V;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t step = stride * UNROLL;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  bool bad = false;
  P b0[UNROLL], b1[UNROLL];
  map_fast_load<UNROLL>(b0, in, tid, stride, nvec);
  for (size_t i = tid; i < nvec; i += 2 * step) {
    map_fast_load<UNROLL>(b1, in, i + step, stride, nvec);
    map_fast_step<OP, UNROLL>(b0, out, i, stride, nvec, bad);
    map_fast_load<UNROLL>(b0, in, i + 2 * step, stride, nvec);
    map_fast_step<OP, UNROLL>(b1, out, i + step, stride, nvec, bad);     // past the end: nothing is stored
  }
  for (size_t i = nvec * V + tid; i < n; i += stride) {
    const double y = math::unary_apply<OP, double>(in[i]);
    if constexpr (math::UnaryCanDomainError<OP>::value) bad |= is_nan_fast(y) && !is_nan_fast(in[i]);
    out[i] = y;
  }
  if constexpr (math::UnaryCanDomainError<OP>::value) {
    if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
  }
}

template <int OP>
struct FUnary {
  template <typename T>
  __device__ __forceinline__ T operator()(T x) const { return math::unary_apply<OP, T>(x); }
};

// cheap exact-tier ops stream with 4 requests in flight, transcendental ones with 2
template <int OP>
constexpr int unary_unroll() {
  return (OP == MCM_U_PLUS || OP == MCM_U_NEG || OP == MCM_U_ZERO || OP == MCM_U_SIGN ||
          OP == MCM_U_ABS || OP == MCM_U_RAD2DEG || OP == MCM_U_DEG2RAD || OP == MCM_U_FLOAT ||
          OP == MCM_U_ROUND || OP == MCM_U_FLOOR || OP == MCM_U_CEIL || OP == MCM_U_TRUNC ||
          OP == MCM_U_SQR || OP == MCM_U_CUBE || OP == MCM_U_ROUND_TIESAWAY || OP == MCM_U_ROUND_TIESUP ||
          OP == MCM_U_ROUND_FROMZERO)
             ? 4
             : 2;
}

// The same step with the input staged through shared memory by bulk asynchronous copies (cp.async.bulk, one elected
// thread, an mbarrier per stage): the bytes in flight per SM are MF_STAGES x 8 KiB per resident CTA whatever the
// register budget, which the register prefetch above cannot reach at 80 registers per thread (ptxas sinks the early
// loads back to their first use).  Tile t of a CTA is tile blockIdx.x + t gridDim.x of the input; within a tile the
// pack of thread tid, unroll u is pack u * kThreads + tid (conflict-free 128-bit LDS, coalesced 128-bit stores).
constexpr int MF_STAGES = 4;
template <int OP, int UNROLL, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
k_map_fast_bulk(const double* __restrict__ in, double* __restrict__ out, size_t n, uint32_t* __restrict__ flags) {
  using P = Pack<double>;
  constexpr int V = P::N;
  constexpr int EPT = kThreads * UNROLL * V;                 // doubles per tile
  constexpr unsigned TILE_BYTES = EPT * 8;
  extern __shared__ __align__(128) unsigned char mf_tiles[];
  __shared__ __align__(8) unsigned long long full_bar[MF_STAGES];
  const int tid = threadIdx.x;
  const size_t ntiles = n / EPT;
  const size_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < MF_STAGES; ++s) mbar_init(&full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < MF_STAGES && (size_t)s < my_tiles; ++s) {
      const size_t tile = blockIdx.x + (size_t)s * gridDim.x;
      mbar_expect_tx(&full_bar[s], TILE_BYTES);
      bulk_g2s(mf_tiles + (size_t)s * TILE_BYTES, in + tile * EPT, TILE_BYTES, &full_bar[s]);
    }
  }
  bool bad = false;
  for (size_t i = 0; i < my_tiles; ++i) {
    const int s = (int)(i % MF_STAGES);
    mbar_wait(&full_bar[s], (unsigned)((i / MF_STAGES) & 1));
    const P* tile = reinterpret_cast<const P*>(mf_tiles + (size_t)s * TILE_BYTES);
    P b[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) b[u] = tile[u * kThreads + tid];
    __syncthreads();                           // every thread has read stage s: it may be refilled
    if (tid == 0 && i + MF_STAGES < my_tiles) {
      const size_t tile_next = blockIdx.x + (i + MF_STAGES) * gridDim.x;
      mbar_expect_tx(&full_bar[s], TILE_BYTES);
      bulk_g2s(mf_tiles + (size_t)s * TILE_BYTES, in + tile_next * EPT, TILE_BYTES, &full_bar[s]);
    }
    const size_t base = (blockIdx.x + i * gridDim.x) * (size_t)(EPT / V);      // first pack of this tile
    map_fast_step<OP, UNROLL>(b, out, base + tid, kThreads, (size_t)-1, bad);
  }
  // remainder (< one tile) per value
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t j = ntiles * EPT + (size_t)blockIdx.x * kThreads + tid; j < n; j += stride) {
    const double y = math::unary_apply<OP, double>(in[j]);
    if constexpr (math::UnaryCanDomainError<OP>::value) bad |= is_nan_fast(y) && !is_nan_fast(in[j]);
    out[j] = y;
  }
  if constexpr (math::UnaryCanDomainError<OP>::value) {
    if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
  }
}

template <int OP, typename T>
static int run_unary(const void* in, void* out, size_t n) {
  if constexpr (sizeof(T) == 8 && FastUnary<OP>::has) {
    if (n >= 4096 && aligned16(in) && aligned16(out)) {
      Ctx& c = ctx();
      // measured at N = 1e8 on a B200 (scripts/ab_unary.py, profiles/r2_ab_unary.txt): sin / cos need the shared-memory
      // ring (0.338 -> 0.265 ms = 0.94 of the measured HBM rate; every register variant stayed at 0.33-0.34), sqrt / inv
      // are best with the register ping-pong at 4 CTAs per SM (0.264 / 0.261 ms)
      if constexpr (OP == MCM_U_SIN || OP == MCM_U_COS) {
        constexpr int smem = MF_STAGES * kThreads * 1 * 16;
        static int occ = 0;
        if (occ == 0) {
          cudaFuncSetAttribute(k_map_fast_bulk<OP, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_map_fast_bulk<OP, 1, 4>, kThreads, smem);
          if (occ < 1) occ = 1;
        }
        k_map_fast_bulk<OP, 1, 4><<<c.sm_count * occ, kThreads, smem, c.stream>>>((const double*)in, (double*)out, n, c.d_flags);
      } else {
        static int occ = 0;
        if (occ == 0) {
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_map_fast<OP, 2, 4>, kThreads, 0);
          if (occ < 1) occ = 1;
        }
        k_map_fast<OP, 2, 4><<<stream_grid((n + 1) / 2, 2, occ), kThreads, 0, c.stream>>>((const double*)in, (double*)out, n, c.d_flags);
      }
      MCM_CHECK_LAUNCH();
      return MCM_OK;
    }
  }
  Operands<T, 1> ops;
  ops.p[0] = static_cast<const T*>(in);
  ops.s[0] = T(0);
  return launch_map<T, 1, unary_unroll<OP>(), math::UnaryCanDomainError<OP>::value>(
      ops, static_cast<T*>(out), n, FUnary<OP>());
}

template <typename T>
static int dispatch_unary(int op, const void* in, void* out, size_t n) {
  switch (op) {
#define MCM_CASE(OP) \
  case OP: return run_unary<OP, T>(in, out, n);
    MCM_CASE(MCM_U_PLUS) MCM_CASE(MCM_U_NEG) MCM_CASE(MCM_U_EXP) MCM_CASE(MCM_U_EXP2)
    MCM_CASE(MCM_U_EXP10) MCM_CASE(MCM_U_EXPM1) MCM_CASE(MCM_U_LOG) MCM_CASE(MCM_U_LOG10)
    MCM_CASE(MCM_U_LOG2) MCM_CASE(MCM_U_LOG1P) MCM_CASE(MCM_U_SIN) MCM_CASE(MCM_U_COS)
    MCM_CASE(MCM_U_TAN) MCM_CASE(MCM_U_SIND) MCM_CASE(MCM_U_COSD) MCM_CASE(MCM_U_TAND)
    MCM_CASE(MCM_U_SINH) MCM_CASE(MCM_U_COSH) MCM_CASE(MCM_U_TANH) MCM_CASE(MCM_U_ASIN)
    MCM_CASE(MCM_U_ACOS) MCM_CASE(MCM_U_ATAN) MCM_CASE(MCM_U_ASIND) MCM_CASE(MCM_U_ACOSD)
    MCM_CASE(MCM_U_ATAND) MCM_CASE(MCM_U_ASINH) MCM_CASE(MCM_U_ACOSH) MCM_CASE(MCM_U_ATANH)
    MCM_CASE(MCM_U_ZERO) MCM_CASE(MCM_U_SIGN) MCM_CASE(MCM_U_ABS) MCM_CASE(MCM_U_SQRT)
    MCM_CASE(MCM_U_RAD2DEG) MCM_CASE(MCM_U_DEG2RAD) MCM_CASE(MCM_U_FLOAT) MCM_CASE(MCM_U_ROUND)
    MCM_CASE(MCM_U_FLOOR) MCM_CASE(MCM_U_CEIL) MCM_CASE(MCM_U_TRUNC) MCM_CASE(MCM_U_INV)
    MCM_CASE(MCM_U_SQR) MCM_CASE(MCM_U_CUBE)
    MCM_CASE(MCM_U_ROUND_TIESAWAY) MCM_CASE(MCM_U_ROUND_TIESUP) MCM_CASE(MCM_U_ROUND_FROMZERO)
    MCM_CASE(MCM_U_REM2PI_NEAREST) MCM_CASE(MCM_U_REM2PI_TOZERO) MCM_CASE(MCM_U_REM2PI_DOWN) MCM_CASE(MCM_U_REM2PI_UP)
#undef MCM_CASE
    default: return fail(MCM_ERR_INVALID, "mcm_unary: unknown op %d", op);
  }
}

struct FPowi {
  int k;
  template <typename T>
  __device__ __forceinline__ T operator()(T x) const { return math::m_powi(x, k); }
};

template <typename T>
struct FFill {
  T v;
  __device__ __forceinline__ T operator()(T) const { return v; }
};

// convert between element types (one pack in, one pack out would need two pack
// widths; a scalar grid-stride loop with 8-byte/4-byte accesses is still coalesced)
template <typename S, typename D>
__global__ void __launch_bounds__(kThreads) k_convert(const S* __restrict__ src, D* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) dst[i] = (D)src[i];
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_unary(int op, int dtype, const void* in, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_unary: null pointer");
  if (dtype == MCM_F64) return dispatch_unary<double>(op, in, out, n);
  if (dtype == MCM_F32) return dispatch_unary<float>(op, in, out, n);
  return fail(MCM_ERR_INVALID, "mcm_unary: bad dtype %d", dtype);
}

extern "C" int mcm_powi(int dtype, const void* in, int k, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_powi: null pointer");
  FPowi f{k};
  if (dtype == MCM_F64) {
    Operands<double, 1> ops; ops.p[0] = (const double*)in; ops.s[0] = 0;
    return launch_map<double, 1, 2, false>(ops, (double*)out, n, f);
  }
  if (dtype == MCM_F32) {
    Operands<float, 1> ops; ops.p[0] = (const float*)in; ops.s[0] = 0;
    return launch_map<float, 1, 2, false>(ops, (float*)out, n, f);
  }
  return fail(MCM_ERR_INVALID, "mcm_powi: bad dtype %d", dtype);
}

extern "C" int mcm_fill(int dtype, void* out, size_t n, double value) {
  MCM_NEED_INIT();
  if (n && !out) return fail(MCM_ERR_INVALID, "mcm_fill: null pointer");
  if (dtype == MCM_F64) {
    Operands<double, 1> ops; ops.p[0] = nullptr; ops.s[0] = value;
    return launch_map<double, 1, 4, false>(ops, (double*)out, n, FFill<double>{value});
  }
  if (dtype == MCM_F32) {
    Operands<float, 1> ops; ops.p[0] = nullptr; ops.s[0] = (float)value;
    return launch_map<float, 1, 4, false>(ops, (float*)out, n, FFill<float>{(float)value});
  }
  return fail(MCM_ERR_INVALID, "mcm_fill: bad dtype %d", dtype);
}

extern "C" int mcm_convert(int src_dtype, const void* src, int dst_dtype, void* dst, size_t n) {
  MCM_NEED_INIT();
  if (n == 0) return MCM_OK;
  if (!src || !dst) return fail(MCM_ERR_INVALID, "mcm_convert: null pointer");
  Ctx& c = ctx();
  const int grid = stream_grid(n, 4, 8);
  if (src_dtype == MCM_F32 && dst_dtype == MCM_F64)
    k_convert<float, double><<<grid, kThreads, 0, c.stream>>>((const float*)src, (double*)dst, n);
  else if (src_dtype == MCM_F64 && dst_dtype == MCM_F32)
    k_convert<double, float><<<grid, kThreads, 0, c.stream>>>((const double*)src, (float*)dst, n);
  else if (src_dtype == dst_dtype && (src_dtype == MCM_F32 || src_dtype == MCM_F64)) {
    MCM_CUDA(cudaMemcpyAsync(dst, src, n * (src_dtype == MCM_F64 ? 8 : 4), cudaMemcpyDeviceToDevice, c.stream));
    return MCM_OK;
  } else
    return fail(MCM_ERR_INVALID, "mcm_convert: bad dtypes %d -> %d", src_dtype, dst_dtype);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
