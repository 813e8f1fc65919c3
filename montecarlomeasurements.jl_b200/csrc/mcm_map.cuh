// mcm_map.cuh — the one elementwise kernel shape used by every registered
// primitive: 128-bit coalesced streaming loads/stores, grid-stride over a
// persistent grid sized in multiples of the SM count, UNROLL independent
// 16-byte requests in flight per thread per operand.
//
// Replaces the broadcast bodies `f.(p.particles, ...)` of
// src/register_primitive.jl:44-58,94-97.
#pragma once
#include "mcm_common.cuh"

namespace mcm {

template <typename T, int NIN>
struct Operands {
  const T* p[NIN];  // nullptr -> use the scalar
  T s[NIN];
};

template <class F, typename T>
__device__ __forceinline__ T call_f(const F& f, const T (&a)[1]) { return f(a[0]); }
template <class F, typename T>
__device__ __forceinline__ T call_f(const F& f, const T (&a)[2]) { return f(a[0], a[1]); }
template <class F, typename T>
__device__ __forceinline__ T call_f(const F& f, const T (&a)[3]) { return f(a[0], a[1], a[2]); }

template <typename T, int NIN>
__device__ __forceinline__ bool any_nan(const T (&a)[NIN]) {
  bool r = false;
#pragma unroll
  for (int k = 0; k < NIN; ++k) r |= is_nan_fast(a[k]);
  return r;
}

// VECTOR=true : all pointers 16-byte aligned, packs of 16/sizeof(T) elements.
// VECTOR=false: scalar fallback for unaligned views.
template <typename T, int NIN, int UNROLL, bool FLAG, bool VECTOR, class F>
__global__ void __launch_bounds__(kThreads)
k_map(Operands<T, NIN> ops, T* __restrict__ out, size_t n, F f, uint32_t* __restrict__ flags) {
  using P = Pack<T>;
  constexpr int V = VECTOR ? P::N : 1;
  const size_t nvec = VECTOR ? n / V : 0;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  bool bad = false;

  if constexpr (VECTOR) {
    for (size_t i = tid; i < nvec; i += stride * UNROLL) {
      P x[NIN][UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t j = i + (size_t)u * stride;
        if (j < nvec) {
#pragma unroll
          for (int k = 0; k < NIN; ++k)
            if (ops.p[k] != nullptr) x[k][u] = ld_stream(reinterpret_cast<const P*>(ops.p[k]) + j);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t j = i + (size_t)u * stride;
        if (j < nvec) {
          P r;
#pragma unroll
          for (int e = 0; e < V; ++e) {
            T a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = (ops.p[k] != nullptr) ? x[k][u].v[e] : ops.s[k];
            const T y = call_f(f, a);
            if constexpr (FLAG) bad |= is_nan_fast(y) && !any_nan(a);
            r.v[e] = y;
          }
          st_stream(reinterpret_cast<P*>(out) + j, r);
        }
      }
    }
  }
  // tail (VECTOR) or everything (scalar path)
  for (size_t i = nvec * V + tid; i < n; i += stride) {
    T a[NIN];
#pragma unroll
    for (int k = 0; k < NIN; ++k) a[k] = (ops.p[k] != nullptr) ? ops.p[k][i] : ops.s[k];
    const T y = call_f(f, a);
    if constexpr (FLAG) bad |= is_nan_fast(y) && !any_nan(a);
    out[i] = y;
  }
  if constexpr (FLAG) {
    if (bad) atomicOr(flags, (uint32_t)MCM_FLAG_DOMAIN);
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T, int NIN, int UNROLL, bool FLAG, class F>
int launch_map(const Operands<T, NIN>& ops, T* out, size_t n, F f) {
  if (n == 0) return MCM_OK;
  Ctx& c = ctx();
  bool vec = aligned16(out);
  for (int k = 0; k < NIN; ++k) vec = vec && aligned16(ops.p[k]);
  constexpr int V = Pack<T>::N;
  if (vec) {
    static int occ = 0;  // resident CTAs per SM for this instantiation
    if (occ == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_map<T, NIN, UNROLL, FLAG, true, F>, kThreads, 0);
      if (occ < 1) occ = 1;
    }
    const int grid = stream_grid((n + V - 1) / V, UNROLL, occ);
    k_map<T, NIN, UNROLL, FLAG, true, F><<<grid, kThreads, 0, c.stream>>>(ops, out, n, f, c.d_flags);
  } else {
    const int grid = stream_grid(n, 1, 8);
    k_map<T, NIN, 1, FLAG, false, F><<<grid, kThreads, 0, c.stream>>>(ops, out, n, f, c.d_flags);
  }
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

}  // namespace mcm
