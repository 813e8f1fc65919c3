// mcm_moments.cuh — single-pass, numerically robust moments {n, mean, M2, min, max, sum}.
//
// pmean / pvar / pstd / pextrema in the reference are Statistics.mean (pairwise
// sum / N) and a two-pass corrected variance over p.particles
// (src/particles.jl:302-323).  Here ONE pass over HBM produces all of them:
// each thread folds the 16-byte packs it loads into chunk statistics
// (chunk mean, chunk sum of squared deviations) and merges chunks with the
// Chan/Golub/LeVeque update, so there is no catastrophic cancellation for any
// data offset (unlike sum-of-squares) at one division per chunk.  Lanes, warps
// and CTAs are then merged in a fixed order (shuffle tree -> shared memory ->
// last-CTA-done ticket), which makes results reproducible launch to launch.
#pragma once
#include "mcm_common.cuh"

namespace mcm {

// The mean is carried as the unevaluated sum K + dm: K is a fixed shift taken from the data (so that
// deviations v - K are formed exactly for data with a large common offset, e.g. 1e9 +- 1), dm the mean of
// the shifted values.  All second-moment arithmetic happens on O(sigma)-sized numbers.
struct Mom {
  double n, K, dm, m2, mn, mx, sum;
};

__host__ __device__ __forceinline__ Mom mom_empty() {
  Mom m;
  m.n = 0.0; m.K = 0.0; m.dm = 0.0; m.m2 = 0.0;
  m.mn = __builtin_huge_val(); m.mx = -__builtin_huge_val();
  m.sum = 0.0;
  return m;
}

// NaN-propagating min/max (Base.minimum/maximum propagate NaN)
__host__ __device__ __forceinline__ double nmin(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (b < a ? b : a));
}
__host__ __device__ __forceinline__ double nmax(double a, double b) {
  return (a != a) ? a : ((b != b) ? b : (b > a ? b : a));
}

__host__ __device__ __forceinline__ Mom mom_merge(const Mom& a, const Mom& b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  Mom r;
  r.n = a.n + b.n;
  const double d = (b.K - a.K) + (b.dm - a.dm);  // difference of the two means, shifts subtracted first
  const double fb = b.n / r.n;
  r.K = a.K;
  r.dm = a.dm + d * fb;
  r.m2 = (a.m2 + b.m2) + (d * d) * (a.n * fb);
  r.mn = nmin(a.mn, b.mn);
  r.mx = nmax(a.mx, b.mx);
  r.sum = a.sum + b.sum;
  return r;
}

#ifdef __CUDACC__

// Fold a chunk of C values (already in registers) into the thread's running moments.
__device__ __forceinline__ double median3(double a, double b, double c) {
  return fmax(fmin(a, b), fmin(fmax(a, b), c));
}

template <int C>
__device__ __forceinline__ void mom_push_chunk(Mom& m, const double (&v)[C]) {
  if (m.n == 0.0) {
    // the thread's shift: a value from the data; median of three guards against a lone outlier
    if constexpr (C >= 3) m.K = median3(v[0], v[1], v[2]);
    else m.K = v[0];
    if (!(fabs(m.K) <= 1.7976931348623157e308)) m.K = 0.0;  // Inf/NaN shift would poison everything
  }
  double s = 0.0, sd = 0.0, lo = v[0], hi = v[0];
  bool nan = false;
  double dv[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
    s += v[i];
    dv[i] = v[i] - m.K;
    sd += dv[i];
    lo = fmin(lo, v[i]);
    hi = fmax(hi, v[i]);
    nan |= (v[i] != v[i]);
  }
  const double cm = sd * (1.0 / C);  // chunk mean in shifted coordinates
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const double d = dv[i] - cm;
    q = fma(d, d, q);
  }
  if (nan) lo = hi = s;  // s is NaN
  const double nn = m.n + (double)C;
  const double fb = (double)C / nn;
  const double d = cm - m.dm;
  m.dm = (m.n == 0.0) ? cm : m.dm + d * fb;
  m.m2 = (m.m2 + q) + (d * d) * (m.n * fb);
  m.n = nn;
  m.sum += s;
  m.mn = nmin(m.mn, lo);
  m.mx = nmax(m.mx, hi);
}

__device__ __forceinline__ void mom_push_one(Mom& m, double x) {
  const double v[1] = {x};
  mom_push_chunk<1>(m, v);
}

// Accumulator for kernels that see only a pack per loop iteration (the fused chains): mom_push_chunk would pay a
// full pairwise merge, with its division, for every two values.  Here the hot step is a shifted power sum
// (d = v - K; s1 += d; s2 += d*d), folded into running (n, mean, m2) every MS_FOLD values by the same merge formula;
// at every fold the shift K moves onto the running mean, so from the second fold on the in-chunk cancellation
// s2 - s1^2/c loses log2(1 + (mean_chunk - K)^2 / var_chunk) bits of at most MS_FOLD values.
constexpr int MS_FOLD = 64;
struct MomStream {
  double n, K, m2, mn, mx, s1, s2;
  int c;             // values in the open chunk; -1 until the first value fixed K
  unsigned nanw;     // max over values of (high word << 1): > 0xffe00000 iff a NaN was pushed
};
// The running sum is not accumulated separately (one FP64 add per value on a pipe-bound loop): after the last fold
// K is the mean of everything pushed, so sum = n * K to 1 ulp.
__device__ __forceinline__ MomStream ms_empty() {
  MomStream a;
  a.n = 0.0; a.K = 0.0; a.m2 = 0.0; a.mn = __builtin_huge_val(); a.mx = -__builtin_huge_val();
  a.s1 = 0.0; a.s2 = 0.0; a.c = -1; a.nanw = 0u;
  return a;
}
__device__ __forceinline__ void ms_fold(MomStream& a) {
  const double cc = (double)a.c;
  const double cm = a.s1 / cc;                 // chunk mean relative to K
  double q = fma(-a.s1, cm, a.s2);
  q = q < 0.0 ? 0.0 : q;
  const double nn = a.n + cc;
  const double f = cc / nn;
  a.m2 = (a.m2 + q) + (cm * cm) * (a.n * f);
  a.K += cm * f;                               // K is the running mean from here on
  a.n = nn; a.s1 = 0.0; a.s2 = 0.0; a.c = 0;
}
// the first value a thread sees fixes its shift; kept out of line so that the steady-state loop carries one
// predictable branch instead of a compare and four selects per pack
static __device__ __noinline__ double ms_first_shift(double v0) { return (fabs(v0) <= 1.7976931348623157e308) ? v0 : 0.0; }
template <int C>
__device__ __forceinline__ void ms_push(MomStream& a, const double (&v)[C]) {
  if (a.c < 0) { a.K = ms_first_shift(v[0]); a.c = 0; }
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const double d = v[i] - a.K;
    a.s1 += d;
    a.s2 = fma(d, d, a.s2);
    a.nanw = max(a.nanw, (unsigned)__double2hiint(v[i]) << 1);
  }
  // extrema: one compare orders a pair, then one compare each against the running minimum / maximum (3 FP64
  // compares per pair instead of 4).  A NaN compares false and is caught by nanw, like fmin/fmax would skip it.
#pragma unroll
  for (int i = 0; i + 1 < C; i += 2) {
    const bool lt = v[i] < v[i + 1];
    const double lo = lt ? v[i] : v[i + 1], hi = lt ? v[i + 1] : v[i];
    a.mn = lo < a.mn ? lo : a.mn;
    a.mx = hi > a.mx ? hi : a.mx;
  }
  if constexpr (C & 1) {
    a.mn = v[C - 1] < a.mn ? v[C - 1] : a.mn;
    a.mx = v[C - 1] > a.mx ? v[C - 1] : a.mx;
  }
  a.c += C;
  if (a.c >= MS_FOLD) ms_fold(a);
}
__device__ __forceinline__ Mom ms_finish(MomStream& a) {
  if (a.c > 0) ms_fold(a);
  Mom m;
  m.n = a.n; m.K = a.K; m.dm = 0.0; m.m2 = a.m2; m.mn = a.mn; m.mx = a.mx; m.sum = a.n * a.K;
  if (a.nanw > 0xffe00000u) m.mn = m.mx = m.sum = __builtin_nan("");   // a NaN was pushed
  return m;
}

__device__ __forceinline__ Mom mom_shfl_down(const Mom& a, int d) {
  Mom b;
  b.n = shfl_down_d(a.n, d); b.K = shfl_down_d(a.K, d); b.dm = shfl_down_d(a.dm, d); b.m2 = shfl_down_d(a.m2, d);
  b.mn = shfl_down_d(a.mn, d); b.mx = shfl_down_d(a.mx, d); b.sum = shfl_down_d(a.sum, d);
  return b;
}

// Fixed-order CTA reduction; result valid in thread 0.
__device__ __forceinline__ Mom mom_block_reduce(Mom m) {
  __shared__ Mom s_w[kThreads / 32];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = mom_merge(m, mom_shfl_down(m, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_w[warp] = m;
  __syncthreads();
  if (warp == 0) {
    m = (lane < kThreads / 32) ? s_w[lane] : mom_empty();
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) m = mom_merge(m, mom_shfl_down(m, d));
  }
  __syncthreads();
  return m;
}

// ---- cross-rank merge inside the kernel (only the reduction partials cross NVLink) -------------
// Every rank owns a small peer-visible window (CUDA IPC, csrc/comm.cu): MX_SLOTS x MX_MAXRANKS records of 8 doubles
// {n, K, dm, m2, min, max, sum, epoch}.  The last CTA of a reducing kernel stores its rank's partial straight into
// the window of EVERY rank (P2P stores over NVLink / NVSwitch), then waits until the records of all ranks for this
// epoch have arrived in its own window and merges them in rank order — all ranks get bit-identical results with no
// host round trip, no separate collective launch and no data-path synchronisation with the CPU.  (First version:
// kernel -> D2H -> H2D -> ncclAllGather(48 B) -> D2H -> host merge: 0.8 ms per step at 8 ranks for a 48-byte exchange.)
// Slots are indexed by epoch % MX_SLOTS: a rank can run at most one exchange ahead of the slowest peer, because its
// own merge needs that peer's record of the same epoch.
constexpr int MX_MAXRANKS = 16;
constexpr int MX_SLOTS = 4;
constexpr int MX_REC = 8;
struct MomExchange {
  double* peer[MX_MAXRANKS];     // window base on every rank (peer[rank] is the local one)
  int world, rank;
  unsigned long long epoch;      // sequence number of this exchange (>= 1, the same on all ranks)
  uint32_t* flags;               // sticky flags (a peer that never arrives raises MCM_FLAG_COMM)
};
__device__ __forceinline__ unsigned long long mx_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// called by all threads of ONE CTA; m valid in thread 0; returns the merged moments in thread 0
__device__ __forceinline__ Mom mom_exchange(Mom m, const MomExchange& px) {
  __shared__ double s_rec[MX_REC];
  if (threadIdx.x == 0) {
    s_rec[0] = m.n; s_rec[1] = m.K; s_rec[2] = m.dm; s_rec[3] = m.m2; s_rec[4] = m.mn; s_rec[5] = m.mx; s_rec[6] = m.sum;
  }
  __syncthreads();
  const size_t slot = (size_t)(px.epoch % MX_SLOTS) * MX_MAXRANKS * MX_REC;
  if ((int)threadIdx.x < px.world) {
    // my record into the window of rank t
    volatile double* dst = px.peer[threadIdx.x] + slot + (size_t)px.rank * MX_REC;
#pragma unroll
    for (int k = 0; k < 7; ++k) dst[k] = s_rec[k];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(dst + 7) = px.epoch;
    // the record of rank t in my window
    const volatile double* src = px.peer[px.rank] + slot + (size_t)threadIdx.x * MX_REC;
    const volatile unsigned long long* flag = reinterpret_cast<const volatile unsigned long long*>(src + 7);
    const unsigned long long t0 = mx_timer_ns();
    while (*flag != px.epoch) {
      if (mx_timer_ns() - t0 > 20000000000ull) { atomicOr(px.flags, 4u); break; }   // 20 s: a rank is missing (MCM_FLAG_COMM)
      __nanosleep(200);
    }
    __threadfence_system();
  }
  __syncthreads();
  Mom r = mom_empty();
  if (threadIdx.x == 0) {
    for (int t = 0; t < px.world; ++t) {
      const volatile double* src = px.peer[px.rank] + slot + (size_t)t * MX_REC;
      Mom b;
      b.n = src[0]; b.K = src[1]; b.dm = src[2]; b.m2 = src[3]; b.mn = src[4]; b.mx = src[5]; b.sum = src[6];
      r = mom_merge(r, b);
    }
  }
  return r;
}

// Grid-level finish: every CTA deposits its partial; the last CTA to arrive
// merges all partials in index order and writes out[0..5].  `ticket` must be 0
// on entry and is reset to 0 on exit (so launches can be queued back to back).
// px (optional): merge with the other ranks' results before writing (mom_exchange).
__device__ __forceinline__ void mom_grid_finish(Mom block_m, Mom* partials, unsigned int* ticket,
                                                double* out, const MomExchange* px = nullptr) {
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_m;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  Mom m = mom_empty();
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += kThreads) {
    const volatile Mom* p = partials + i;
    Mom b;
    b.n = p->n; b.K = p->K; b.dm = p->dm; b.m2 = p->m2; b.mn = p->mn; b.mx = p->mx; b.sum = p->sum;
    m = mom_merge(m, b);
  }
  m = mom_block_reduce(m);
  if (px != nullptr && px->world > 1) m = mom_exchange(m, *px);
  if (threadIdx.x == 0) {
    out[0] = m.n; out[1] = m.K + m.dm; out[2] = m.m2; out[3] = m.mn; out[4] = m.mx; out[5] = m.sum;
    __threadfence_system();     // `out` may be mapped host memory (result ring)
    *ticket = 0u;
  }
}

#endif  // __CUDACC__
}  // namespace mcm
