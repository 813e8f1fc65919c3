// select_batch.cu — pquantile / pmedian of M uncertain variables in one call
// (src/particles.jl:316-323 applied per element of a Vector{Particles}: the ribbon / C5 case,
// 1024 variables x 1e6 particles).
//
// Stage 1 (n >= 131072), BRACKETING — one read of the data at streaming speed, no histograms:
//   k_bq_sample  one CTA per segment sorts a stratified sample (64 runs x 128 keys) in shared memory
//                and publishes, for every group of requested ranks, two pivots lo <= hi taken
//                4.5 standard deviations of the sampling error either side of the rank;
//   k_bq_stream  grid (chunks, M): every key is compared with the pivots: per-thread register
//                counters of keys below lo, keys inside [lo, hi] (a few % of the data) are appended
//                to the bracket's candidate list (warp-aggregated atomics).
//   The wanted order statistic of the segment is order statistic (rank - below) of its bracket; a
//   rank that falls outside its bracket (probability ~1e-5, or adversarial data) is flagged and
//   re-done by the general path, so results are always exact.
// Stage 2, RADIX SELECT on the (small, L2-resident) candidate lists — or directly on the data when
// n is small.  The single-variable path (select.cu) narrows the ranks on the host between passes;
// with M segments that is M x passes round trips.  Here the narrowing state lives on the device:
//   k_bs_hist    grid (chunks, M): 2048-bin histograms of the next 11 key bits for the keys of
//                segment m that match one of its live prefixes (shared-memory histograms with
//                warp-aggregated atomics, flushed with integer global atomics => exact and
//                order-independent);
//   k_bs_pick    one CTA per segment: block scan of each live histogram, every requested rank
//                moves to its bin; bins with <= 2048 keys are final; histograms are re-zeroed;
//   (repeated a FIXED number of times — segments whose ranks are all final skip the reads)
//   k_bs_compact grid (chunks, M): keys matching a final prefix are appended to the segment's
//                candidate lists;
//   k_bs_final   grid (groups, M): one-CTA bitonic sort of <= 2048 candidates, ranks read off.
// One device->host copy at the end returns M x nr order statistics; the type-7 interpolation is
// done on the host (mcm_quantile.cuh).  A segment that cannot be resolved in the fixed pass count
// (more than 2048 identical keys around a rank) is flagged and re-done by the single path, so
// results are always the exact order statistics.
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_keys.cuh"
#include "mcm_quantile.cuh"

namespace mcm {

constexpr int BS_BITS = 11;
constexpr int BS_BINS = 1 << BS_BITS;
constexpr int BS_MAXG = 8;       // live prefixes / ranks per segment
constexpr int BS_THRESH = 2048;  // candidates per final group
constexpr int BS_PASSES = 6;     // 66 >= 64 key bits: every group ends final or fully resolved; 3 are enough for 32-bit keys

static int g_bq_fallbacks = 0;  // segments re-done by the general path (diagnostics / tests)

struct BSeg {
  unsigned long long prefix[BS_MAXG];
  uint32_t count[BS_MAXG];
  int rem[BS_MAXG];             // low key bits of group g not fixed yet
  uint32_t rank_group[BS_MAXG];
  uint32_t rank_local[BS_MAXG];  // rank inside its group
  int G;
  int nr;            // requested ranks of this segment
  uint32_t active;   // groups still being narrowed
  uint32_t resolved; // groups whose key is fully determined (rem == 0): value = prefix
  uint32_t flags;    // 1 = NaN seen, 2 = unresolved after BS_PASSES
};

// Segment descriptors for the radix stage.  Uniform mode: nseg == nullptr, every segment has n keys and
// the same ranks.  Bracket mode: segment sidx = m*NB + b holds the candidates of bracket b of variable m:
// nseg[sidx] keys (must be <= cap), its ranks are the parent's ranks of that bracket minus below[sidx];
// all candidates share the key bits above `hb` = highest bit in which the two pivots differ.
struct BInit {
  int nb;                        // brackets per variable (1 in uniform mode)
  int nr[4];                     // ranks per bracket
  uint32_t rank[4][BS_MAXG];     // global ranks of bracket b
  uint32_t cap[4];
};
__global__ void k_bs_init(BSeg* __restrict__ seg, int nsegs, int key_bits, uint32_t n_uniform, BInit bi,
                          const uint32_t* __restrict__ nseg, const uint32_t* __restrict__ nseg_global,
                          const uint32_t* __restrict__ below, const unsigned long long* __restrict__ pivots,
                          uint32_t* __restrict__ parent_flags) {
  const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx >= nsegs) return;
  const int b = sidx % bi.nb, m = sidx / bi.nb;
  BSeg s;
  memset(&s, 0, sizeof(s));
  s.G = 1;
  s.nr = bi.nr[b];
  uint32_t n = n_uniform, sub = 0;
  int rem = key_bits;
  unsigned long long prefix = 0;
  if (nseg) {
    n = nseg_global ? nseg_global[sidx] : nseg[sidx];   // sharded: counts and ranks are global, lists are local
    sub = below[sidx];
    const unsigned long long lo = pivots[2 * sidx], hi = pivots[2 * sidx + 1];
    const unsigned long long diff = lo ^ hi;
    rem = diff ? 64 - __clzll((long long)diff) : 0;
    if (rem > key_bits) rem = key_bits;
    prefix = rem >= 64 ? 0ull : (lo >> rem);
    if (nseg[sidx] > bi.cap[b]) { atomicOr(parent_flags + m, 2u); n = 0; }
  }
  s.count[0] = n;
  s.rem[0] = rem;
  s.prefix[0] = prefix;
  for (int r = 0; r < bi.nr[b]; ++r) {
    const uint32_t want = bi.rank[b][r];
    if (want < sub || want - sub >= n) {         // the rank is not inside its bracket: general path
      if (parent_flags) atomicOr(parent_flags + m, 2u);
      s.flags |= 2u;
    }
    s.rank_group[r] = 0;
    s.rank_local[r] = want - sub;
  }
  if (s.flags & 2u) n = 0;
  if (rem == 0 && n > (uint32_t)BS_THRESH) s.resolved = 1u;
  else s.active = (n > (uint32_t)BS_THRESH) ? 1u : 0u;
  seg[sidx] = s;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_bs_hist(const T* const* __restrict__ vars, size_t n_uniform, const uint32_t* __restrict__ nseg, BSeg* __restrict__ segs,
          uint32_t* __restrict__ hist) {
  using K = typename KeyOf<T>::type;
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int UNROLL = 4;
  extern __shared__ uint32_t sh[];  // [gmax][BS_BINS], indexed by group id
  const int m = blockIdx.y;
  BSeg* sg = segs + m;
  const uint32_t active = sg->active;
  if (active == 0) return;
  const size_t n = nseg ? (size_t)nseg[m] : n_uniform;
  if ((size_t)blockIdx.x * kThreads >= n && blockIdx.x > 0) return;   // more chunks than this segment needs
  const int G = sg->G;
  unsigned long long pre[BS_MAXG];
  int rem = 0;
#pragma unroll
  for (int g = 0; g < BS_MAXG; ++g) {
    const bool on = g < G && ((active >> g) & 1u);
    pre[g] = on ? sg->prefix[g] : ~0ull;  // live groups share `rem`; an impossible prefix disables the rest
    if (on) rem = sg->rem[g];
  }
  const int dbits = rem < BS_BITS ? rem : BS_BITS;
  const int shift = rem - dbits;
  const uint32_t dmask = (1u << dbits) - 1u;
  const int nb_total = G * BS_BINS;
  const bool first_pass = rem >= (int)(8 * sizeof(K));   // nothing fixed yet: every key belongs to group 0
  for (int i = threadIdx.x; i < nb_total; i += kThreads) sh[i] = 0;
  __syncthreads();
  const T* __restrict__ in = vars[m];
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool nan = false;
  // a prefix of ~0 can only be matched by rem == 0 keys of all ones, and live groups have rem > 0
  auto account = [&](bool valid, T x) {
    uint32_t code = 0xffffffffu - lane;
    if (valid) {
      const K key = to_key(x, nan);
      const unsigned long long kp = first_pass ? 0ull : ((unsigned long long)key >> rem);
#pragma unroll
      for (int g = 0; g < BS_MAXG; ++g)
        if (kp == pre[g]) code = (uint32_t)g * BS_BINS + (uint32_t)(((unsigned long long)key >> shift) & dmask);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, code);
    if (code < (uint32_t)nb_total && lane == __ffs(peers) - 1) atomicAdd(&sh[code], (uint32_t)__popc(peers));
  };
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const size_t nvec = vec ? n / V : 0;
  const P* pin = reinterpret_cast<const P*>(in);
  const size_t rows = (nvec + stride * UNROLL - 1) / (stride * UNROLL);
  for (size_t r = 0; r < rows; ++r) {
    P x[UNROLL];
    bool ok[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t j = (r * UNROLL + u) * stride + tid;
      ok[u] = j < nvec;
      if (ok[u]) x[u] = ld_stream(pin + j);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int e = 0; e < V; ++e) account(ok[u], ok[u] ? x[u].v[e] : T(0));
  }
  const size_t tail0 = nvec * V;
  const size_t trips = (n - tail0 + stride - 1) / stride;
  for (size_t r = 0; r < trips; ++r) {
    const size_t i = tail0 + r * stride + tid;
    account(i < n, i < n ? in[i] : T(0));
  }
  __syncthreads();
  uint32_t* gh = hist + (size_t)m * BS_MAXG * BS_BINS;
  for (int i = threadIdx.x; i < nb_total; i += kThreads) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(gh + i, v);
  }
  if (nan) atomicOr(&sg->flags, 1u);
}

// one CTA (256 threads) per segment: thread t owns bins 8t .. 8t+7
__global__ void __launch_bounds__(256)
k_bs_pick(BSeg* __restrict__ segs, uint32_t* __restrict__ hist, int last_pass) {
  const int m = blockIdx.x;
  BSeg* sg = segs + m;
  __shared__ BSeg s;
  __shared__ uint32_t warp_sum[8];
  __shared__ uint32_t nbin[BS_MAXG], nlocal[BS_MAXG], ncount[BS_MAXG];
  if (threadIdx.x == 0) s = *sg;
  __syncthreads();
  if (s.active == 0) return;
  const int nr = s.nr;
  uint32_t* gh = hist + (size_t)m * BS_MAXG * BS_BINS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int g = 0; g < s.G; ++g) {
    if (!((s.active >> g) & 1u)) continue;
    uint32_t b[8], tot = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      b[e] = gh[g * BS_BINS + threadIdx.x * 8 + e];
      gh[g * BS_BINS + threadIdx.x * 8 + e] = 0;  // clean for the next pass
      tot += b[e];
    }
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += warp_sum[w];
    const uint32_t lo = base + inc - tot;  // keys in bins before mine
    for (int r = 0; r < nr; ++r) {
      if (s.rank_group[r] != (uint32_t)g) continue;
      const uint32_t want = s.rank_local[r];
      if (want >= lo && want < lo + tot) {
        uint32_t cum = lo;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (want >= cum && want < cum + b[e]) { nbin[r] = threadIdx.x * 8 + e; nlocal[r] = want - cum; ncount[r] = b[e]; }
          cum += b[e];
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    BSeg o;
    memset(&o, 0, sizeof(o));
    o.flags = s.flags;
    o.nr = s.nr;
    int src_group[BS_MAXG];
    uint32_t src_bin[BS_MAXG];
    for (int r = 0; r < nr; ++r) {
      const int og = (int)s.rank_group[r];
      const bool live = (s.active >> og) & 1u;
      const uint32_t bin = live ? nbin[r] : 0xffffffffu;
      int ng = -1;
      for (int q = 0; q < o.G; ++q)
        if (src_group[q] == og && src_bin[q] == bin) ng = q;
      if (ng < 0) {
        ng = o.G++;
        src_group[ng] = og; src_bin[ng] = bin;
        if (live) {
          const int rem = s.rem[og];
          const int dbits = rem < BS_BITS ? rem : BS_BITS;
          o.prefix[ng] = (s.prefix[og] << dbits) | (unsigned long long)bin;
          o.rem[ng] = rem - dbits;
          o.count[ng] = ncount[r];
          if (o.rem[ng] == 0 && o.count[ng] > (uint32_t)BS_THRESH) o.resolved |= 1u << ng;
          else if (o.count[ng] > (uint32_t)BS_THRESH) o.active |= 1u << ng;
        } else {
          o.prefix[ng] = s.prefix[og]; o.rem[ng] = s.rem[og]; o.count[ng] = s.count[og];
          if ((s.resolved >> og) & 1u) o.resolved |= 1u << ng;
        }
      }
      o.rank_group[r] = (uint32_t)ng;
      o.rank_local[r] = live ? nlocal[r] : s.rank_local[r];
    }
    if (last_pass && o.active) { o.flags |= 2u; o.active = 0; }
    *sg = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_bs_compact(const T* const* __restrict__ vars, size_t n_uniform, const uint32_t* __restrict__ nseg,
             const BSeg* __restrict__ segs, unsigned long long* __restrict__ cand, uint32_t* __restrict__ cnt) {
  using K = typename KeyOf<T>::type;
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int UNROLL = 4;
  const int m = blockIdx.y;
  const BSeg* sg = segs + m;
  if (sg->flags & 2u) return;
  const size_t n = nseg ? (size_t)nseg[m] : n_uniform;
  if ((size_t)blockIdx.x * kThreads >= n) return;
  const int G = sg->G;
  unsigned long long pre[BS_MAXG];
  int rem[BS_MAXG];
  bool any = false;
#pragma unroll
  for (int g = 0; g < BS_MAXG; ++g) {
    const bool on = g < G && !((sg->resolved >> g) & 1u);
    pre[g] = on ? sg->prefix[g] : ~0ull;
    rem[g] = on ? sg->rem[g] : 0;   // (key >> 0) == ~0 only for NaN keys, which are never candidates of a live rank
    any |= on;
  }
  if (!any) return;
  const T* __restrict__ in = vars[m];
  unsigned long long* mc = cand + (size_t)m * BS_MAXG * BS_THRESH;
  uint32_t* mcnt = cnt + (size_t)m * BS_MAXG;
  bool nan = false;
  auto consider = [&](T x) {
    const K key = to_key(x, nan);
#pragma unroll
    for (int g = 0; g < BS_MAXG; ++g) {
      if (rem[g] < 64 ? (((unsigned long long)key >> rem[g]) == pre[g]) : (pre[g] == 0ull)) {
        if (g < G && !((sg->resolved >> g) & 1u)) {
          const uint32_t pos = atomicAdd(mcnt + g, 1u);
          if (pos < (uint32_t)BS_THRESH) mc[(size_t)g * BS_THRESH + pos] = (unsigned long long)key;
        }
        break;
      }
    }
  };
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const size_t nvec = vec ? n / V : 0;
  const P* pin = reinterpret_cast<const P*>(in);
  for (size_t i = tid; i < nvec; i += stride * UNROLL) {
    P x[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t j = i + (size_t)u * stride;
      if (j < nvec) x[u] = ld_stream(pin + j);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t j = i + (size_t)u * stride;
      if (j < nvec) {
#pragma unroll
        for (int e = 0; e < V; ++e) consider(x[u].v[e]);
      }
    }
  }
  for (size_t i = nvec * V + tid; i < n; i += stride) consider(in[i]);
}

template <typename T>
__global__ void __launch_bounds__(1024)
k_bs_final(BSeg* __restrict__ segs, const unsigned long long* __restrict__ cand, const uint32_t* __restrict__ cnt,
           double* __restrict__ out) {
  using K = typename KeyOf<T>::type;
  __shared__ unsigned long long s[BS_THRESH];
  const int g = blockIdx.x, m = blockIdx.y;
  BSeg* sg = segs + m;
  if (g >= sg->G || (sg->flags & 2u)) return;
  const int nr = sg->nr;
  double* mo = out + (size_t)m * BS_MAXG;
  if ((sg->resolved >> g) & 1u) {
    if (threadIdx.x < nr && sg->rank_group[threadIdx.x] == (uint32_t)g) mo[threadIdx.x] = (double)from_key((K)sg->prefix[g]);
    return;
  }
  const uint32_t have = cnt[(size_t)m * BS_MAXG + g];
  if (have != sg->count[g] || have > (uint32_t)BS_THRESH) {  // cannot happen unless the data changed under us
    if (threadIdx.x == 0) atomicOr(&sg->flags, 2u);
    return;
  }
  const unsigned long long* mc = cand + ((size_t)m * BS_MAXG + g) * BS_THRESH;
  for (int i = threadIdx.x; i < BS_THRESH; i += 1024) s[i] = i < (int)have ? mc[i] : ~0ull;
  __syncthreads();
  // sort only as many elements as needed (next power of two >= have)
  int len = 2;
  while (len < (int)have) len <<= 1;
  for (int k = 2; k <= len; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < len; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = s[i], b = s[l];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { s[i] = b; s[l] = a; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x < nr && sg->rank_group[threadIdx.x] == (uint32_t)g)
    mo[threadIdx.x] = (double)from_key((K)s[sg->rank_local[threadIdx.x]]);
}

// ---- stage 1: bracketing ------------------------------------------------------------------------
constexpr int BQ_SAMPLE = 8192;     // keys sampled per segment
constexpr int BQ_RUN = 128;         // contiguous keys per sampled run
constexpr int BQ_MAXB = 4;          // brackets per sweep
constexpr size_t BQ_MIN_N = 131072; // below this the radix stage runs on the data directly

struct BqParams {
  int nb;
  int a[BQ_MAXB], b[BQ_MAXB];  // sample positions of the pivots: a < 0 -> -inf, b >= BQ_SAMPLE -> +inf
  uint32_t cap[BQ_MAXB];        // capacity of each candidate list
};

template <typename T>
__global__ void __launch_bounds__(1024)
k_bq_sample(const T* const* __restrict__ vars, size_t n, BqParams bp, unsigned long long* __restrict__ pivots,
            unsigned long long* __restrict__ sub_out, int world) {
  using K = typename KeyOf<T>::type;
  extern __shared__ unsigned long long sk[];  // BQ_SAMPLE keys
  const int m = blockIdx.x;
  const T* __restrict__ in = vars[m];
  constexpr int RUNS = BQ_SAMPLE / BQ_RUN;
  bool nan = false;
  for (int i = threadIdx.x; i < BQ_SAMPLE; i += 1024) {
    const int run = i / BQ_RUN, off = i % BQ_RUN;
    const size_t start = (size_t)(((unsigned __int128)run * (n - BQ_RUN)) / (RUNS - 1));  // evenly spaced, both ends included
    sk[i] = (unsigned long long)to_key(in[start + off], nan);
  }
  __syncthreads();
  for (int k = 2; k <= BQ_SAMPLE; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < BQ_SAMPLE; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = sk[i], y = sk[l];
          const bool up = ((i & k) == 0);
          if ((x > y) == up) { sk[i] = y; sk[l] = x; }
        }
      }
      __syncthreads();
    }
  if (sub_out) {
    // sharded: this rank contributes every world-th key of its sorted sample (a stratified subsample); the union
    // over the ranks is the 8192-key sample every rank derives the SAME pivots from (k_bq_pivots)
    const int per = BQ_SAMPLE / world;
    for (int i = threadIdx.x; i < per; i += 1024) sub_out[(size_t)m * per + i] = sk[i * world + world / 2];
    return;
  }
  if (threadIdx.x < bp.nb) {
    const int b = threadIdx.x;
    const unsigned long long kmax = sizeof(K) == 8 ? ~0ull : 0xffffffffull;
    // 64-bit keys: the pivots are widened to boundaries of the HIGH key word, so the streaming pass classifies
    // with 32-bit compares only (the bracket grows by < 2^32 key units, i.e. by ~1e-6 of the value)
    unsigned long long lo = bp.a[b] < 0 ? 0ull : sk[bp.a[b]];
    unsigned long long hi = bp.b[b] >= BQ_SAMPLE ? kmax : sk[bp.b[b]];
    if (sizeof(K) == 8) { lo &= 0xffffffff00000000ull; hi |= 0x00000000ffffffffull; }
    pivots[2 * ((size_t)m * bp.nb + b)] = lo;
    pivots[2 * ((size_t)m * bp.nb + b) + 1] = hi;
  }
}

// sharded: pivots from the union of the per-rank subsamples ([rank][M][per] keys, identical on every rank)
__global__ void __launch_bounds__(1024)
k_bq_pivots(const unsigned long long* __restrict__ uni, int world, int M, BqParams bp, int key64,
            unsigned long long* __restrict__ pivots) {
  extern __shared__ unsigned long long sk[];
  const int m = blockIdx.x;
  const int per = BQ_SAMPLE / world;
  // world need not divide BQ_SAMPLE: the < world missing keys are padded with the largest key, which moves every
  // sample position by < world places — far inside the 4.5 sigma bracket margins (exactness never depends on pivots)
  const int real = per * world;
  for (int e = threadIdx.x; e < BQ_SAMPLE; e += 1024)
    sk[e] = e < real ? uni[((size_t)(e / per) * M + m) * per + e % per] : (key64 ? ~0ull : 0xffffffffull);
  __syncthreads();
  for (int k = 2; k <= BQ_SAMPLE; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < BQ_SAMPLE; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long x = sk[i], y = sk[l];
          const bool up = ((i & k) == 0);
          if ((x > y) == up) { sk[i] = y; sk[l] = x; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x < bp.nb) {
    const int b = threadIdx.x;
    unsigned long long lo = bp.a[b] < 0 ? 0ull : sk[bp.a[b]];
    unsigned long long hi = bp.b[b] >= BQ_SAMPLE ? (key64 ? ~0ull : 0xffffffffull) : sk[bp.b[b]];
    if (key64) { lo &= 0xffffffff00000000ull; hi |= 0x00000000ffffffffull; }
    pivots[2 * ((size_t)m * bp.nb + b)] = lo;
    pivots[2 * ((size_t)m * bp.nb + b) + 1] = hi;
  }
}

// The classification only needs the high word of each 64-bit key, and ptxas narrows a 128-bit load whose low
// words are dead into two strided 4-byte loads that fetch every sector twice (ncu: 14.6 GB of DRAM reads for
// 8.19 GB of data).  Folding the low words into a live value (one LOP3 per pack) keeps the load whole.
template <typename T>
__device__ __forceinline__ uint32_t low_words(const Pack<T>& p) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&p);
  return sizeof(T) == 8 ? (w[0] | w[2]) : 0u;
}

// Candidates are staged PER WARP in shared memory — as 32-bit element INDICES — and flushed by the warp
// with one global atomic per flush; the flush gathers the values (L1/L2 hits) into the bracket's list.
// There is no CTA barrier and no shared-memory atomic anywhere in the loop: the per-lane counts of all brackets
// are scanned at once (16-bit fields of a 64-bit word, 5 shuffle steps) and the warp's fill counts live in
// registers.  History (C5, 8.19 GB of keys): one global atomic per warp-step 7.1 ms -> CTA staging of values
// 5.1 ms (97 instructions per key, 3 barriers per iteration, no load/compute overlap) -> CTA staging of
// indices + prefetch + one barrier 3.6 ms (flush every 4th iteration stalls all 8 warps) -> this version.
constexpr int BQ_WARP_KEYS = 32 * 4 * 2;                   // keys per warp iteration (both dtypes)

template <typename T, int NB>
__global__ void __launch_bounds__(kThreads)
k_bq_stream(const T* const* __restrict__ vars, size_t n, BqParams bp, const unsigned long long* __restrict__ pivots,
            T* __restrict__ cand, size_t cand_stride, uint32_t* __restrict__ cnt, uint32_t* __restrict__ below,
            uint32_t* __restrict__ flags, int flush_at) {
  using K = typename KeyOf<T>::type;
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int UNROLL = BQ_WARP_KEYS / (32 * V);          // 4 (f64) / 2 (f32)
  extern __shared__ __align__(16) unsigned char bq_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int capw = flush_at + BQ_WARP_KEYS;                // staging entries per (warp, bracket)
  // candidates are staged BY VALUE: staging their indices and re-reading the keys at flush time cost 47 % of the
  // kernel's DRAM sectors (ncu: 8-byte re-reads are fetched as 64-byte pairs of sectors long after the line left L2)
  T* wstage = reinterpret_cast<T*>(bq_smem) + (size_t)warp * NB * capw;
  __shared__ uint32_t sh[NB][kThreads / 32];
  const int m = blockIdx.y;
  const T* __restrict__ in = vars[m];
  uint32_t lo[NB], span[NB];    // pivots as high key words (the whole key for Float32): in bracket <=> k - lo <= span
  uint32_t nbelow[NB], wcnt[NB];
  T* cb[NB];
  size_t coff = 0;
  constexpr int PSH = sizeof(K) == 8 ? 32 : 0;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    lo[b] = (uint32_t)(pivots[2 * ((size_t)m * NB + b)] >> PSH);
    span[b] = (uint32_t)(pivots[2 * ((size_t)m * NB + b) + 1] >> PSH) - lo[b];
    nbelow[b] = 0;
    wcnt[b] = 0;
    cb[b] = cand + (size_t)m * cand_stride + coff;
    coff += bp.cap[b];
  }
  uint32_t* mcnt = cnt + (size_t)m * NB;
  // high key word of x: order-preserving on the sign/exponent/top-mantissa bits.  NaNs are the keys above
  // +Inf or below -Inf: tracked with a running min/max instead of a per-key test.
  uint32_t kmin = 0xffffffffu, kmax = 0u, lowor = 0u;
  uint32_t inmask[NB];
  auto classify = [&](T x, int bit) {
    uint32_t w;
    if constexpr (sizeof(T) == 8) w = (uint32_t)__double2hiint(x); else w = __float_as_uint(x);
    const uint32_t k = w ^ ((uint32_t)((int32_t)w >> 31) | 0x80000000u);
    kmin = min(kmin, k); kmax = max(kmax, k);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      nbelow[b] += (k < lo[b]) ? 1u : 0u;
      inmask[b] |= ((k - lo[b]) <= span[b]) ? (1u << bit) : 0u;
    }
  };
  // warp flush of bracket b: one global atomic, gather of the staged elements, coalesced stores
  auto warp_flush = [&](int b) {
    const uint32_t c = wcnt[b];
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(mcnt + b, c);
    base = __shfl_sync(0xffffffffu, base, 0);
    __syncwarp();
    const T* st = wstage + b * capw;
    for (uint32_t i = lane; i < c; i += 32)
      if (base + i < bp.cap[b]) cb[b][base + i] = st[i];
    __syncwarp();
    wcnt[b] = 0;
  };
  // Append the candidates flagged in inmask[] (bit (u*V + e) = vals[u].v[e]).  Warp-uniform call.
  auto append = [&](const P (&vals)[UNROLL]) {
    unsigned long long packed = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) packed |= (unsigned long long)__popc(inmask[b]) << (16 * b);
    if (!__any_sync(0xffffffffu, packed != 0)) return;
    unsigned long long incl = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned long long excl = incl - packed;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      uint32_t mk = inmask[b];
      if (mk) {
        T* p = wstage + b * capw + wcnt[b] + ((uint32_t)(excl >> (16 * b)) & 0xffffu);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int e = 0; e < V; ++e)
            if (mk & (1u << (u * V + e))) *p++ = vals[u].v[e];     // static register names: no local memory
      }
      wcnt[b] += (uint32_t)(total >> (16 * b)) & 0xffffu;
      if (wcnt[b] > (uint32_t)flush_at) warp_flush(b);
    }
  };
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const size_t nvec = vec ? n / V : 0;
  const P* pin = reinterpret_cast<const P*>(in);
  const size_t full_rows = nvec / (stride * UNROLL);   // rows in which every thread of the grid has UNROLL packs
  P cur[UNROLL], nxt[UNROLL];
  if (full_rows) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) cur[u] = ld_stream(pin + (size_t)u * stride + tid);
  }
  for (size_t r = 0; r < full_rows; ++r) {
    if (r + 1 < full_rows) {     // next row's loads are in flight while this row is classified
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) nxt[u] = ld_stream(pin + ((r + 1) * UNROLL + u) * stride + tid);
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) inmask[b] = 0;
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      lowor |= low_words(cur[u]);
#pragma unroll
      for (int e = 0; e < V; ++e) classify(cur[u].v[e], u * V + e);
    }
    append(cur);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) cur[u] = nxt[u];
  }
  // remaining packs: fewer than UNROLL per thread, uniform trip count, bounds-checked
  for (size_t j0 = full_rows * UNROLL * stride; j0 < nvec; j0 += stride) {
    const size_t j = j0 + tid;
#pragma unroll
    for (int b = 0; b < NB; ++b) inmask[b] = 0;
    if (j < nvec) {
      cur[0] = ld_stream(pin + j);
      lowor |= low_words(cur[0]);
#pragma unroll
      for (int e = 0; e < V; ++e) classify(cur[0].v[e], e);
    }
    append(cur);
  }
  const size_t tail0 = nvec * V;
  const size_t trips = (n - tail0 + stride - 1) / stride;
  for (size_t r = 0; r < trips; ++r) {
    const size_t i = tail0 + r * stride + tid;
#pragma unroll
    for (int b = 0; b < NB; ++b) inmask[b] = 0;
    if (i < n) { cur[0].v[0] = in[i]; classify(cur[0].v[0], 0); }
    append(cur);
  }
#pragma unroll
  for (int b = 0; b < NB; ++b)
    if (wcnt[b]) warp_flush(b);
  // CTA totals of the below-counters, one atomic per bracket per CTA
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    uint32_t v = nbelow[b];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) sh[b][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NB) {
    uint32_t t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh[threadIdx.x][w];
    if (t) atomicAdd(below + (size_t)m * NB + threadIdx.x, t);
  }
  // NaN keys lie above +Inf / below -Inf
  constexpr uint32_t kPosInf = sizeof(T) == 8 ? 0xfff00000u : 0xff800000u;
  constexpr uint32_t kNegInf = sizeof(T) == 8 ? 0x000fffffu : 0x007fffffu;
  if (kmax > kPosInf || kmin < kNegInf) atomicOr(flags + m, 1u);
  if (lowor == 0x5a5a5a5au && kmax == 0u) atomicOr(flags + m, 0u);   // keeps `lowor` (hence the full loads) alive
}

// sharded merge of the final candidates: rank r's keys of group (sidx, g) go to offset sum_{q<r} cnt[q] of a zeroed
// global-size list; an all-reduce (sum) of the lists then equals their concatenation
__global__ void __launch_bounds__(256)
k_bs_place(const unsigned long long* __restrict__ cand_local, const uint32_t* __restrict__ cnt_all, int world, int rank,
           int ngroups, unsigned long long* __restrict__ cand_global, uint32_t* __restrict__ cnt_total) {
  const int gidx = blockIdx.x;   // sidx * BS_MAXG + g
  uint32_t off = 0, tot = 0;
  for (int r = 0; r < world; ++r) {
    const uint32_t c = cnt_all[(size_t)r * ngroups + gidx];
    if (r < rank) off += c;
    tot += c;
  }
  const uint32_t mine = cnt_all[(size_t)rank * ngroups + gidx];
  for (uint32_t i = threadIdx.x; i < mine && i < (uint32_t)BS_THRESH; i += 256)
    if (off + i < (uint32_t)BS_THRESH) cand_global[(size_t)gidx * BS_THRESH + off + i] = cand_local[(size_t)gidx * BS_THRESH + i];
  if (threadIdx.x == 0) cnt_total[gidx] = tot;
}
__global__ void k_split_flags(const uint32_t* __restrict__ flags, int M, uint32_t* __restrict__ out2) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) { out2[m] = flags[m] & 1u; out2[M + m] = (flags[m] >> 1) & 1u; }
}

__global__ void k_bs_flags(const BSeg* __restrict__ segs, int nsegs, int nb, uint32_t* __restrict__ flags) {
  const int sidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (sidx < nsegs && segs[sidx].flags) atomicOr(flags + sidx / nb, segs[sidx].flags);
}

template <typename T>
int order_statistics(const T* in, size_t n, const unsigned long long* ranks, int nr, double* vals_host, int* has_nan);

// vals_host[m*nr + r] = sorted(vars[m])[ranks[r]]; nr <= BS_MAXG
int comm_world(int* nranks, int* rank);
int comm_allreduce_sum_u32(uint32_t* d, size_t count);
int comm_allreduce_sum_u64(unsigned long long* d, size_t count);
int comm_allgather_bytes(const void* send, void* recv, size_t bytes_per_rank);

struct ShardInfo {
  int world, rank;
  size_t n_global;
  bool full;                   // skip the bracketing stage: radix select over the whole (sharded) data
  std::vector<int>* flagged;   // out: variables whose ranks escaped their bracket (same list on every rank)
};

// sharded radix select without brackets: the "candidate list" of a variable is its whole local shard
__global__ void k_fill_full(uint32_t* __restrict__ qcnt, uint32_t* __restrict__ qcnt_g, unsigned long long* __restrict__ piv,
                            int S, uint32_t n_local, uint32_t n_global, int key64) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S) return;
  qcnt[i] = n_local;
  qcnt_g[i] = n_global;
  piv[2 * i] = 0ull;
  piv[2 * i + 1] = key64 ? ~0ull : 0xffffffffull;
}

// n = number of LOCAL keys per variable.  sh != nullptr: the variables are sharded over sh->world ranks (every rank
// calls with its shard), `ranks` are ranks in the GLOBAL order, every rank returns the same values.
template <typename T>
static int order_statistics_batch(const T* const* vars_host, int M, size_t n, const unsigned long long* ranks, int nr,
                                  double* vals_host, int* has_nan, const ShardInfo* sh = nullptr) {
  using K = typename KeyOf<T>::type;
  Ctx& c = ctx();
  *has_nan = 0;
  const size_t n_plan = sh ? sh->n_global : n;     // the order statistics refer to this many keys
  const bool bracket = sh ? !sh->full : n >= BQ_MIN_N;
  // ---- host plan: brackets in sample space (data independent) --------------------------------
  BqParams bp;
  BInit bi;
  memset(&bp, 0, sizeof(bp));
  memset(&bi, 0, sizeof(bi));
  int slot_bracket[BS_MAXG], slot_index[BS_MAXG];   // where rank r of the caller lives
  if (bracket) {
    struct Iv { int a, b; std::vector<int> ranks; };
    std::vector<Iv> iv;
    std::vector<int> order(nr);
    for (int r = 0; r < nr; ++r) order[r] = r;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return ranks[x] < ranks[y]; });
    for (int oi = 0; oi < nr; ++oi) {
      const int r = order[oi];
      const double p = ((double)ranks[r] + 0.5) / (double)n_plan;
      const double pos = p * BQ_SAMPLE;
      const double w = 4.5 * sqrt(BQ_SAMPLE * p * (1.0 - p)) + 3.0;
      int a = (int)floor(pos - w), b = (int)ceil(pos + w);
      if (a < 0) a = -1;
      if (b >= BQ_SAMPLE) b = BQ_SAMPLE;
      if (!iv.empty() && a <= iv.back().b + 1) { iv.back().b = std::max(iv.back().b, b); iv.back().ranks.push_back(r); }
      else iv.push_back({a, b, {r}});
    }
    while ((int)iv.size() > BQ_MAXB) {  // cannot happen for <= 4 probabilities; merge the closest pair otherwise
      size_t best = 0;
      for (size_t i = 1; i + 1 < iv.size(); ++i)
        if (iv[i + 1].a - iv[i].b < iv[best + 1].a - iv[best].b) best = i;
      iv[best].b = iv[best + 1].b;
      iv[best].ranks.insert(iv[best].ranks.end(), iv[best + 1].ranks.begin(), iv[best + 1].ranks.end());
      iv.erase(iv.begin() + best + 1);
    }
    bp.nb = bi.nb = (int)iv.size();
    for (int b = 0; b < bp.nb; ++b) {
      bp.a[b] = iv[b].a; bp.b[b] = iv[b].b;
      const double width = (double)(std::min(iv[b].b, BQ_SAMPLE) - std::max(iv[b].a, 0) + 2);
      // sharded: one rank may hold ALL candidates of a bracket (e.g. a globally sorted array), so the bound is
      // the global candidate count, clipped to the local size
      double cap = (width + 6.0 * sqrt(width) + 8.0) * (double)n_plan / BQ_SAMPLE + 1024.0;
      if (cap > (double)n) cap = (double)n;
      bp.cap[b] = bi.cap[b] = ((uint32_t)cap + 3u) & ~3u;   // 16-byte aligned lists
      bi.nr[b] = (int)iv[b].ranks.size();
      for (int i = 0; i < bi.nr[b]; ++i) {
        bi.rank[b][i] = (uint32_t)ranks[iv[b].ranks[i]];
        slot_bracket[iv[b].ranks[i]] = b; slot_index[iv[b].ranks[i]] = i;
      }
    }
  } else {
    bi.nb = 1; bi.nr[0] = nr; bi.cap[0] = (uint32_t)n;
    for (int r = 0; r < nr; ++r) { bi.rank[0][r] = (uint32_t)ranks[r]; slot_bracket[r] = 0; slot_index[r] = r; }
  }
  const int NB = bi.nb;
  const int S = M * NB;  // radix-stage segments
  size_t cand_stride = 0;
  uint32_t max_cap = 0;
  for (int b = 0; b < NB; ++b) { cand_stride += bi.cap[b]; max_cap = std::max(max_cap, bi.cap[b]); }

  // ---- device workspace -----------------------------------------------------------------------
  const size_t tab_bytes = (size_t)(M + S) * sizeof(void*);
  const size_t seg_bytes = (size_t)S * sizeof(BSeg);
  const size_t hist_bytes = (size_t)S * BS_MAXG * BS_BINS * sizeof(uint32_t);
  const size_t cnt_bytes = (size_t)S * BS_MAXG * sizeof(uint32_t);
  const size_t out_bytes = (size_t)S * BS_MAXG * sizeof(double);
  const size_t piv_bytes = (size_t)S * 2 * sizeof(unsigned long long);
  const int world = sh ? sh->world : 1;
  const size_t per = BQ_SAMPLE / (size_t)world;                       // subsample keys per rank per variable
  const size_t shard_bytes = sh ? ((size_t)S * 4 + 2 * (size_t)M * 4 + (size_t)world * cnt_bytes + cnt_bytes +
                                   (size_t)M * per * 8 + (size_t)world * M * per * 8 + 256) : 0;
  const size_t zero_bytes = hist_bytes + cnt_bytes + 3 * (size_t)S * 4 + (size_t)M * 4 + 64;  // hist, cnt, qcnt, below, pad, flags
  unsigned char* d_all = nullptr;
  const size_t total = tab_bytes + seg_bytes + zero_bytes + out_bytes + piv_bytes + shard_bytes + 1024;
  MCM_CUDA(cudaMallocAsync((void**)&d_all, total, c.stream));
  size_t off = 0;
  auto take = [&](size_t bytes) { unsigned char* p = d_all + off; off += (bytes + 15) & ~(size_t)15; return p; };
  unsigned char* d_zero0 = d_all;
  uint32_t* d_hist = (uint32_t*)take(hist_bytes);
  uint32_t* d_cnt = (uint32_t*)take(cnt_bytes);
  uint32_t* d_qcnt = (uint32_t*)take((size_t)S * 4);    // bracket list lengths
  uint32_t* d_below = (uint32_t*)take((size_t)S * 4);
  uint32_t* d_flags = (uint32_t*)take((size_t)M * 4);
  const size_t zero_span = off;
  const T** d_tab = (const T**)take(tab_bytes);          // [M] data pointers, then [S] list pointers
  BSeg* d_seg = (BSeg*)take(seg_bytes);
  double* d_out = (double*)take(out_bytes);
  unsigned long long* d_piv = (unsigned long long*)take(piv_bytes);
  uint32_t *d_qcnt_g = nullptr, *d_flags2 = nullptr, *d_cnt_all = nullptr, *d_cnt_tot = nullptr;
  unsigned long long *d_sub = nullptr, *d_uni = nullptr;
  if (sh) {
    d_qcnt_g = (uint32_t*)take((size_t)S * 4);
    d_flags2 = (uint32_t*)take(2 * (size_t)M * 4);
    d_cnt_all = (uint32_t*)take((size_t)world * cnt_bytes);
    d_cnt_tot = (uint32_t*)take(cnt_bytes);
    d_sub = (unsigned long long*)take((size_t)M * per * 8);
    d_uni = (unsigned long long*)take((size_t)world * M * per * 8);
  }
  MCM_CUDA(cudaMemsetAsync(d_zero0, 0, zero_span, c.stream));
  T* d_qcand = nullptr;
  if (bracket) MCM_CUDA(cudaMallocAsync((void**)&d_qcand, (size_t)M * cand_stride * sizeof(T), c.stream));
  unsigned long long* d_cand = nullptr;
  const size_t cand_keys = (size_t)S * BS_MAXG * BS_THRESH;
  MCM_CUDA(cudaMallocAsync((void**)&d_cand, cand_keys * sizeof(unsigned long long) * (sh ? 2 : 1), c.stream));
  unsigned long long* d_cand_g = sh ? d_cand + cand_keys : nullptr;   // merged lists (sharded)
  if (sh) MCM_CUDA(cudaMemsetAsync(d_cand_g, 0, cand_keys * sizeof(unsigned long long), c.stream));
  std::vector<const T*> htab((size_t)M + S);
  for (int m = 0; m < M; ++m) htab[m] = vars_host[m];
  for (int m = 0; m < M; ++m) {
    size_t o = 0;
    for (int b = 0; b < NB; ++b) {
      htab[(size_t)M + (size_t)m * NB + b] = bracket ? d_qcand + (size_t)m * cand_stride + o : vars_host[m];
      o += bi.cap[b];
    }
  }
  MCM_CUDA(cudaMemcpyAsync(d_tab, htab.data(), tab_bytes, cudaMemcpyHostToDevice, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));  // htab is a local
  static bool attr = false;
  if (!attr) {
    MCM_CUDA(cudaFuncSetAttribute(k_bs_hist<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_MAXG * BS_BINS * 4));
    MCM_CUDA(cudaFuncSetAttribute(k_bs_hist<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, BS_MAXG * BS_BINS * 4));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_sample<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, BQ_SAMPLE * 8));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_sample<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, BQ_SAMPLE * 8));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_pivots, cudaFuncAttributeMaxDynamicSharedMemorySize, BQ_SAMPLE * 8));
    constexpr int kStreamSmem = (kThreads / 32) * 4 * (256 + BQ_WARP_KEYS) * 8;
    MCM_CUDA(cudaFuncSetAttribute(k_bq_stream<double, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_stream<double, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_stream<double, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_stream<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_stream<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    MCM_CUDA(cudaFuncSetAttribute(k_bq_stream<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSmem));
    attr = true;
  }
  auto chunks = [&](size_t keys, int segs) {
    size_t want = (keys + 8191) / 8192;
    size_t cap = ((size_t)c.sm_count * 8 + segs - 1) / segs;
    if (cap < 1) cap = 1;
    return (unsigned)std::max<size_t>(1, std::min(want, cap));
  };
  // ---- stage 1 ------------------------------------------------------------------------------------
  if (bracket) {
    if (!sh) {
      k_bq_sample<T><<<M, 1024, BQ_SAMPLE * 8, c.stream>>>(d_tab, n, bp, d_piv, nullptr, 1);
      MCM_CHECK_LAUNCH();
    } else {
      k_bq_sample<T><<<M, 1024, BQ_SAMPLE * 8, c.stream>>>(d_tab, n, bp, d_piv, d_sub, world);
      MCM_CHECK_LAUNCH();
      if (int rc = comm_allgather_bytes(d_sub, d_uni, (size_t)M * per * 8)) return rc;
      k_bq_pivots<<<M, 1024, BQ_SAMPLE * 8, c.stream>>>(d_uni, world, M, bp, sizeof(K) == 8 ? 1 : 0, d_piv);
      MCM_CHECK_LAUNCH();
    }
    const dim3 g1(chunks(n, M), (unsigned)M);
    // few variables: many warps flush into the same three counters -> flush rarely; many variables: small staging
    const int flush_at = M >= 64 ? 128 : 256;
    const size_t smem1 = (size_t)(kThreads / 32) * NB * (flush_at + BQ_WARP_KEYS) * sizeof(T);
    switch (NB) {
      case 1: k_bq_stream<T, 1><<<g1, kThreads, smem1, c.stream>>>(d_tab, n, bp, d_piv, d_qcand, cand_stride, d_qcnt, d_below, d_flags, flush_at); break;
      case 2: k_bq_stream<T, 2><<<g1, kThreads, smem1, c.stream>>>(d_tab, n, bp, d_piv, d_qcand, cand_stride, d_qcnt, d_below, d_flags, flush_at); break;
      case 3: k_bq_stream<T, 3><<<g1, kThreads, smem1, c.stream>>>(d_tab, n, bp, d_piv, d_qcand, cand_stride, d_qcnt, d_below, d_flags, flush_at); break;
      default: k_bq_stream<T, 4><<<g1, kThreads, smem1, c.stream>>>(d_tab, n, bp, d_piv, d_qcand, cand_stride, d_qcnt, d_below, d_flags, flush_at); break;
    }
    MCM_CHECK_LAUNCH();
    if (sh) {   // global below-counts and candidate counts (the lists themselves stay where they are)
      MCM_CUDA(cudaMemcpyAsync(d_qcnt_g, d_qcnt, (size_t)S * 4, cudaMemcpyDeviceToDevice, c.stream));
      if (int rc = comm_allreduce_sum_u32(d_qcnt_g, S)) return rc;
      if (int rc = comm_allreduce_sum_u32(d_below, S)) return rc;
    }
  }
  if (sh && !bracket) {
    k_fill_full<<<(S + 127) / 128, 128, 0, c.stream>>>(d_qcnt, d_qcnt_g, d_piv, S, (uint32_t)n, (uint32_t)sh->n_global, sizeof(K) == 8 ? 1 : 0);
    MCM_CHECK_LAUNCH();
  }
  // ---- stage 2: radix select on S segments -------------------------------------------------------
  const T* const* seg_tab = d_tab + M;
  const uint32_t* d_nseg = (bracket || sh) ? d_qcnt : nullptr;
  const size_t seg_keys = bracket ? max_cap : n;
  k_bs_init<<<(S + 127) / 128, 128, 0, c.stream>>>(d_seg, S, 8 * (int)sizeof(K), (uint32_t)n, bi, d_nseg, d_qcnt_g, d_below, d_piv, d_flags);
  MCM_CHECK_LAUNCH();
  const dim3 grid(chunks(seg_keys, S), (unsigned)S);
  const int passes = sizeof(K) == 8 ? BS_PASSES : 3;
  int gmax = 1;
  for (int b = 0; b < NB; ++b) gmax = std::max(gmax, std::min(BS_MAXG, bi.nr[b]));
  for (int p = 0; p < passes; ++p) {
    const size_t smem = (size_t)(p == 0 ? 1 : gmax) * BS_BINS * 4;
    k_bs_hist<T><<<grid, kThreads, smem, c.stream>>>(seg_tab, n, d_nseg, d_seg, d_hist);
    MCM_CHECK_LAUNCH();
    if (sh) {   // every rank narrows on the GLOBAL histograms, so the narrowing state stays identical everywhere
      if (int rc = comm_allreduce_sum_u32(d_hist, (size_t)S * BS_MAXG * BS_BINS)) return rc;
    }
    k_bs_pick<<<S, 256, 0, c.stream>>>(d_seg, d_hist, p == passes - 1);
    MCM_CHECK_LAUNCH();
  }
  k_bs_compact<T><<<grid, kThreads, 0, c.stream>>>(seg_tab, n, d_nseg, d_seg, d_cand, d_cnt);
  MCM_CHECK_LAUNCH();
  const unsigned long long* fin_cand = d_cand;
  const uint32_t* fin_cnt = d_cnt;
  if (sh) {   // the <= 2048 finalists of every group are the only keys that ever cross NVLink
    const int ngroups = S * BS_MAXG;
    if (int rc = comm_allgather_bytes(d_cnt, d_cnt_all, cnt_bytes)) return rc;
    k_bs_place<<<ngroups, 256, 0, c.stream>>>(d_cand, d_cnt_all, world, sh->rank, ngroups, d_cand_g, d_cnt_tot);
    MCM_CHECK_LAUNCH();
    if (int rc = comm_allreduce_sum_u64(d_cand_g, cand_keys)) return rc;
    fin_cand = d_cand_g;
    fin_cnt = d_cnt_tot;
  }
  k_bs_final<T><<<dim3(gmax, S), 1024, 0, c.stream>>>(d_seg, fin_cand, fin_cnt, d_out);
  MCM_CHECK_LAUNCH();
  k_bs_flags<<<(S + 127) / 128, 128, 0, c.stream>>>(d_seg, S, NB, d_flags);
  MCM_CHECK_LAUNCH();
  if (sh) {   // a NaN or an escaped rank seen by any rank is seen by all
    k_split_flags<<<(M + 127) / 128, 128, 0, c.stream>>>(d_flags, M, d_flags2);
    MCM_CHECK_LAUNCH();
    if (int rc = comm_allreduce_sum_u32(d_flags2, 2 * (size_t)M)) return rc;
  }
  std::vector<double> hout((size_t)S * BS_MAXG);
  std::vector<uint32_t> hflags(M), hflags2(sh ? 2 * (size_t)M : 0);
  MCM_CUDA(cudaMemcpyAsync(hout.data(), d_out, out_bytes, cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaMemcpyAsync(hflags.data(), d_flags, (size_t)M * 4, cudaMemcpyDeviceToHost, c.stream));
  if (sh) MCM_CUDA(cudaMemcpyAsync(hflags2.data(), d_flags2, 2 * (size_t)M * 4, cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  if (sh)
    for (int m = 0; m < M; ++m) hflags[m] = (hflags2[m] ? 1u : 0u) | (hflags2[(size_t)M + m] ? 2u : 0u);
  MCM_CUDA(cudaFreeAsync(d_cand, c.stream));
  if (d_qcand) MCM_CUDA(cudaFreeAsync(d_qcand, c.stream));
  MCM_CUDA(cudaFreeAsync(d_all, c.stream));
  int rc = MCM_OK;
  for (int m = 0; m < M && rc == MCM_OK; ++m) {
    if (hflags[m] & 1u) *has_nan = 1;
    if ((hflags[m] & 2u) && sh) {
      if (sh->full || !sh->flagged) rc = fail(MCM_ERR_CUDA, "sharded order statistics: variable %d did not resolve", m);
      else sh->flagged->push_back(m);   // heavy ties or an escaped rank: the caller re-runs these without brackets
    } else if (hflags[m] & 2u) {  // a rank outside its bracket: the general path resolves any input
      int hn = 0;
      rc = order_statistics<T>(vars_host[m], n, ranks, nr, vals_host + (size_t)m * nr, &hn);
      *has_nan |= hn;
      g_bq_fallbacks++;
    } else {
      for (int r = 0; r < nr; ++r)
        vals_host[(size_t)m * nr + r] = hout[((size_t)m * NB + slot_bracket[r]) * BS_MAXG + slot_index[r]];
    }
  }
  return rc;
}

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_quantile_batch_fallbacks(void) { return g_bq_fallbacks; }

extern "C" int mcm_order_statistics(int dtype, const void* in, size_t n, const uint64_t* ranks_host, int nr, double* out_host) {
  MCM_NEED_INIT();
  if (!in || !ranks_host || !out_host || nr < 1) return fail(MCM_ERR_INVALID, "mcm_order_statistics: bad arguments");
  if (n == 0) return fail(MCM_ERR_INVALID, "mcm_order_statistics: empty particles");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_order_statistics: bad dtype %d", dtype);
  if (n >= 0xFFFF0000ull) return fail(MCM_ERR_UNSUPPORTED, "mcm_order_statistics: n must be < 2^32");
  if (reinterpret_cast<uintptr_t>(in) % (dtype == MCM_F64 ? 8 : 4)) return fail(MCM_ERR_INVALID, "mcm_order_statistics: misaligned input");
  for (int r = 0; r < nr; ++r)
    if (ranks_host[r] >= n) return fail(MCM_ERR_INVALID, "mcm_order_statistics: rank %llu out of range", (unsigned long long)ranks_host[r]);
  int any_nan = 0;
  for (int r0 = 0; r0 < nr; r0 += BS_MAXG) {
    const int cnt = std::min(BS_MAXG, nr - r0);
    unsigned long long rk[BS_MAXG];
    for (int i = 0; i < cnt; ++i) rk[i] = ranks_host[r0 + i];
    int hn = 0;
    const int rc = dtype == MCM_F64 ? order_statistics_batch<double>((const double* const*)&in, 1, n, rk, cnt, out_host + r0, &hn)
                                    : order_statistics_batch<float>((const float* const*)&in, 1, n, rk, cnt, out_host + r0, &hn);
    if (rc) return rc;
    any_nan |= hn;
  }
  if (any_nan) return fail(MCM_ERR_INVALID, "mcm_order_statistics: order is undefined in presence of NaNs");
  return MCM_OK;
}

extern "C" int mcm_quantile_batch(int dtype, const void* const* vars, int M, size_t n, const double* q_host, int nq,
                                  double* out_host) {
  MCM_NEED_INIT();
  if (!vars || !q_host || !out_host || nq < 1 || M < 1) return fail(MCM_ERR_INVALID, "mcm_quantile_batch: bad arguments");
  if (n == 0) return fail(MCM_ERR_INVALID, "mcm_quantile_batch: empty particles");
  if (M > 16383) return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_batch: at most 16383 variables per call");
  if (n >= 0xFFFF0000ull) return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_batch: n must be < 2^32");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_quantile_batch: bad dtype %d", dtype);
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  for (int m = 0; m < M; ++m)
    if (!vars[m] || (reinterpret_cast<uintptr_t>(vars[m]) % esz) != 0)
      return fail(MCM_ERR_INVALID, "mcm_quantile_batch: variable %d is null or misaligned", m);
  for (int i = 0; i < nq; ++i)
    if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0))
      return fail(MCM_ERR_INVALID, "mcm_quantile_batch: input probability out of [0,1] range");
  // 4 probabilities (8 ranks) per sweep
  std::vector<double> vals((size_t)M * BS_MAXG);
  int any_nan = 0;
  for (int q0 = 0; q0 < nq; q0 += BS_MAXG / 2) {
    const int cq = std::min(BS_MAXG / 2, nq - q0);
    unsigned long long ranks[BS_MAXG];
    double gam[BS_MAXG / 2];
    for (int i = 0; i < cq; ++i) quantile_rank_pair(n, q_host[q0 + i], ranks[2 * i], ranks[2 * i + 1], gam[i]);
    const int nr = 2 * cq;
    int hn = 0;
    const int rc = dtype == MCM_F64
                       ? order_statistics_batch<double>((const double* const*)vars, M, n, ranks, nr, vals.data(), &hn)
                       : order_statistics_batch<float>((const float* const*)vars, M, n, ranks, nr, vals.data(), &hn);
    if (rc) return rc;
    any_nan |= hn;
    for (int m = 0; m < M; ++m)
      for (int i = 0; i < cq; ++i)
        out_host[(size_t)m * nq + q0 + i] =
            quantile_combine(dtype, vals[(size_t)m * nr + 2 * i], vals[(size_t)m * nr + 2 * i + 1], gam[i]);
  }
  if (any_nan) return fail(MCM_ERR_INVALID, "mcm_quantile_batch: quantiles are undefined in presence of NaNs");
  return MCM_OK;
}

// pquantile over a particle index sharded across the ranks of the library communicator: every
// rank passes its shard of each variable; all ranks return the same quantiles of the GLOBAL particle set.
// Communication: 8192 sample keys per variable (all-gather), the below/candidate counts and the per-pass
// histograms (all-reduce), and the <= 2048 finalists per rank group — no other key leaves its GPU.
extern "C" int mcm_quantile_sharded(int dtype, const void* const* vars, int M, size_t n_local, const double* q_host, int nq,
                                    double* out_host) {
  MCM_NEED_INIT();
  int world = 1, rank = 0;
  comm_world(&world, &rank);
  if (world == 1) return mcm_quantile_batch(dtype, vars, M, n_local, q_host, nq, out_host);
  if (!vars || !q_host || !out_host || nq < 1 || M < 1) return fail(MCM_ERR_INVALID, "mcm_quantile_sharded: bad arguments");
  if (M > 16383) return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_sharded: at most 16383 variables per call");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_quantile_sharded: bad dtype %d", dtype);
  if (world > 64) return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_sharded: at most 64 ranks");
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  for (int m = 0; m < M; ++m)
    if (!vars[m] || (reinterpret_cast<uintptr_t>(vars[m]) % esz) != 0)
      return fail(MCM_ERR_INVALID, "mcm_quantile_sharded: variable %d is null or misaligned", m);
  for (int i = 0; i < nq; ++i)
    if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0))
      return fail(MCM_ERR_INVALID, "mcm_quantile_sharded: input probability out of [0,1] range");
  Ctx& c = ctx();
  // shard sizes of all ranks
  unsigned long long* d_n = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_n, (size_t)(world + 1) * 8, c.stream));
  const unsigned long long nl = n_local;
  MCM_CUDA(cudaMemcpyAsync(d_n + world, &nl, 8, cudaMemcpyHostToDevice, c.stream));
  if (int rc = comm_allgather_bytes(d_n + world, d_n, 8)) return rc;
  std::vector<unsigned long long> hn(world);
  MCM_CUDA(cudaMemcpyAsync(hn.data(), d_n, (size_t)world * 8, cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d_n, c.stream));
  size_t n_global = 0;
  for (int r = 0; r < world; ++r) n_global += hn[r];
  for (int r = 0; r < world; ++r) {
    if (hn[r] < (unsigned long long)BQ_SAMPLE) return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_sharded: every shard needs at least %d particles (rank %d has %llu)", BQ_SAMPLE, r, hn[r]);
    const double ideal = (double)n_global / world;
    if (fabs((double)hn[r] - ideal) > 0.01 * ideal + 2.0)
      return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_sharded: shards must be balanced to 1 %% (rank %d has %llu of %zu)", r, hn[r], n_global);
  }
  if (n_global >= 0xFFFF0000ull) return fail(MCM_ERR_UNSUPPORTED, "mcm_quantile_sharded: global N must be < 2^32");
  std::vector<int> flagged;
  ShardInfo sh{world, rank, n_global, false, &flagged};
  std::vector<double> vals((size_t)M * BS_MAXG);
  int any_nan = 0;
  for (int q0 = 0; q0 < nq; q0 += BS_MAXG / 2) {
    const int cq = std::min(BS_MAXG / 2, nq - q0);
    unsigned long long ranks[BS_MAXG];
    double gam[BS_MAXG / 2];
    for (int i = 0; i < cq; ++i) quantile_rank_pair(n_global, q_host[q0 + i], ranks[2 * i], ranks[2 * i + 1], gam[i]);
    const int nr = 2 * cq;
    int h = 0;
    const int rc = dtype == MCM_F64
                       ? order_statistics_batch<double>((const double* const*)vars, M, n_local, ranks, nr, vals.data(), &h, &sh)
                       : order_statistics_batch<float>((const float* const*)vars, M, n_local, ranks, nr, vals.data(), &h, &sh);
    if (rc) return rc;
    any_nan |= h;
    if (!flagged.empty()) {
      // variables with more ties around a rank than a bracket can hold (or, with probability ~1e-5, an escaped
      // rank): radix select over their whole shards with all-reduced histograms — 64 key bits in 6 passes
      std::vector<const void*> sub(flagged.size());
      for (size_t i = 0; i < flagged.size(); ++i) sub[i] = vars[flagged[i]];
      std::vector<double> v2(flagged.size() * BS_MAXG);
      ShardInfo full{world, rank, n_global, true, nullptr};
      int h2 = 0;
      const int rc2 = dtype == MCM_F64
                          ? order_statistics_batch<double>((const double* const*)sub.data(), (int)sub.size(), n_local, ranks, nr, v2.data(), &h2, &full)
                          : order_statistics_batch<float>((const float* const*)sub.data(), (int)sub.size(), n_local, ranks, nr, v2.data(), &h2, &full);
      if (rc2) return rc2;
      any_nan |= h2;
      for (size_t i = 0; i < flagged.size(); ++i)
        for (int r = 0; r < nr; ++r) vals[(size_t)flagged[i] * nr + r] = v2[i * nr + r];
      g_bq_fallbacks += (int)flagged.size();
      flagged.clear();
    }
    for (int m = 0; m < M; ++m)
      for (int i = 0; i < cq; ++i)
        out_host[(size_t)m * nq + q0 + i] =
            quantile_combine(dtype, vals[(size_t)m * nr + 2 * i], vals[(size_t)m * nr + 2 * i + 1], gam[i]);
  }
  if (any_nan) return fail(MCM_ERR_INVALID, "mcm_quantile_sharded: quantiles are undefined in presence of NaNs");
  return MCM_OK;
}
