// select.cu — order statistics by MSD radix SELECT (no full sort) for pquantile / pmedian.
//
// Statistics.quantile(p.particles, q) (src/particles.jl:321) needs at most two order statistics
// per probability.  Sorting costs 8 passes x 24 B/key; selecting costs one 8 B/key read per pass
// and stops as soon as the candidate set of every requested rank is small:
//   pass: histogram (2048 bins) of the next 11 key bits over the keys that match one of up to 8
//         active prefixes (shared-memory histograms, warp-aggregated atomics, per-CTA partials
//         summed in fixed order) -> the host narrows every rank to one bin;
//   end:  keys matching the final prefixes are compacted (<= 4096 per prefix), sorted by a one-CTA
//         bitonic network in shared memory, and the requested ranks are read off.
// Continuous data resolves in 3 histogram passes + 1 compaction pass (32 B/key in total).
// Results are exact order statistics (bit-identical to sort + index).
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_keys.cuh"

namespace mcm {

constexpr int SEL_BITS = 11;
constexpr int SEL_BINS = 1 << SEL_BITS;
constexpr int SEL_MAXG = 8;          // active prefixes per launch
constexpr int SEL_THRESH = 4096;     // candidate-set size at which histogramming stops
constexpr int SEL_IPT = 8;

struct SelParams {
  unsigned long long prefix[SEL_MAXG];  // value of the already-fixed high bits (key >> rem)
  int G;
  int rem;     // number of low bits not fixed yet
  int dbits;   // digit width of this pass
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_sel_hist(const T* __restrict__ in, size_t n, SelParams p, uint32_t* __restrict__ part, uint32_t* __restrict__ nan_flag) {
  using K = typename KeyOf<T>::type;
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int UNROLL = 4;
  extern __shared__ uint32_t sh[];  // [G][SEL_BINS]
  const int nbins_total = p.G * SEL_BINS;
  for (int i = threadIdx.x; i < nbins_total; i += kThreads) sh[i] = 0;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int shift = p.rem - p.dbits;
  const uint32_t dmask = (1u << p.dbits) - 1u;
  const bool first_pass = p.rem >= (int)(8 * sizeof(K));
  bool nan = false;

  auto account = [&](bool valid, T x) {
    uint32_t code = 0xffffffffu - lane;  // unique "no match" codes
    if (valid) {
      const K key = to_key(x, nan);
      const unsigned long long kp = first_pass ? 0ull : ((unsigned long long)key >> p.rem);
#pragma unroll
      for (int g = 0; g < SEL_MAXG; ++g)
        if (g < p.G && kp == p.prefix[g]) code = (uint32_t)g * SEL_BINS + (uint32_t)(((unsigned long long)key >> shift) & dmask);
    }
    // warp-aggregated shared atomics: one atomic per distinct (prefix, digit) in the warp
    const uint32_t peers = __match_any_sync(0xffffffffu, code);
    if (code < (uint32_t)nbins_total && lane == __ffs(peers) - 1) atomicAdd(&sh[code], (uint32_t)__popc(peers));
  };

  const bool vec = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
  const size_t nvec = vec ? n / V : 0;
  const P* pin = reinterpret_cast<const P*>(in);
  const size_t rows = (nvec + stride * UNROLL - 1) / (stride * UNROLL);  // every lane runs the same trip count
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
  uint32_t* mine = part + (size_t)blockIdx.x * nbins_total;
  for (int i = threadIdx.x; i < nbins_total; i += kThreads) mine[i] = sh[i];
  if (nan) atomicOr(nan_flag, 1u);
}

__global__ void __launch_bounds__(256) k_sel_sum(const uint32_t* __restrict__ part, int ncta, int nbins_total, uint32_t* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nbins_total) return;
  uint32_t s = 0;
  for (int c = 0; c < ncta; ++c) s += part[(size_t)c * nbins_total + i];
  out[i] = s;
}

struct CompactParams {
  unsigned long long prefix[2 * SEL_MAXG];
  int rem[2 * SEL_MAXG];
  uint32_t offset[2 * SEL_MAXG];  // start of each group's region in the candidate buffer
  int G;
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
k_sel_compact(const T* __restrict__ in, size_t n, CompactParams p, unsigned long long* __restrict__ cand, uint32_t* __restrict__ counters) {
  using K = typename KeyOf<T>::type;
  using P = Pack<T>;
  constexpr int V = P::N;
  constexpr int UNROLL = 4;
  const size_t stride = (size_t)gridDim.x * kThreads;
  const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
  bool nan = false;
  auto consider = [&](T x) {
    const K key = to_key(x, nan);
    for (int g = 0; g < p.G; ++g) {
      if (((unsigned long long)key >> p.rem[g]) == p.prefix[g]) {
        const uint32_t pos = atomicAdd(&counters[g], 1u);
        cand[p.offset[g] + pos] = (unsigned long long)key;
        break;
      }
    }
  };
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

// one CTA per group: bitonic sort of <= SEL_THRESH keys in shared memory, then read the ranks
struct PickParams {
  uint32_t offset[2 * SEL_MAXG], count[2 * SEL_MAXG];
  uint32_t rank_begin[2 * SEL_MAXG + 1];   // ranks of group g: rank_local[rank_begin[g] .. rank_begin[g+1])
  uint32_t rank_local[64];
  uint32_t rank_out[64];                   // where to put the result
};

template <typename T>
__global__ void __launch_bounds__(1024) k_sel_pick(const unsigned long long* __restrict__ cand, PickParams p, double* __restrict__ out) {
  using K = typename KeyOf<T>::type;
  __shared__ unsigned long long s[SEL_THRESH];
  const int g = blockIdx.x;
  const uint32_t cnt = p.count[g];
  for (int i = threadIdx.x; i < SEL_THRESH; i += 1024) s[i] = i < (int)cnt ? cand[p.offset[g] + i] : ~0ull;
  __syncthreads();
  for (int k = 2; k <= SEL_THRESH; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < SEL_THRESH; i += 1024) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = s[i], b = s[l];
          const bool up = ((i & k) == 0);
          if ((a > b) == up) { s[i] = b; s[l] = a; }
        }
      }
      __syncthreads();
    }
  for (uint32_t r = p.rank_begin[g] + threadIdx.x; r < p.rank_begin[g + 1]; r += 1024)
    out[p.rank_out[r]] = (double)from_key((K)s[p.rank_local[r]]);
}

int read_back(double* host_dst, const double* dev_src, size_t count);

struct SelGroup {
  unsigned long long prefix = 0;
  int rem = 0;
  unsigned long long count = 0;
  std::vector<std::pair<unsigned long long, int>> ranks;  // (rank within the group, output slot)
};

// Exact order statistics: vals_host[i] = sorted(in)[ranks[i]] (0-based ranks).  *has_nan reports NaNs.
template <typename T>
int order_statistics(const T* in, size_t n, const unsigned long long* ranks, int nr, double* vals_host, int* has_nan) {
  using K = typename KeyOf<T>::type;
  constexpr int KB = 8 * (int)sizeof(K);
  Ctx& c = ctx();
  if (nr > 64) return fail(MCM_ERR_INVALID, "order_statistics: at most 64 ranks per call");
  if (n >= 0xFFFF0000ull) return fail(MCM_ERR_UNSUPPORTED, "order_statistics: n must be < 2^32");
  // 4 CTAs/SM: the G = 1 passes use 8 KiB of shared memory, the widest (G = 8) 64 KiB
  const int ncta = std::max(1, std::min(c.sm_count * 4, (int)((n + kThreads * SEL_IPT - 1) / (kThreads * SEL_IPT))));
  uint32_t *d_part = nullptr, *d_hist = nullptr, *d_flag = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_part, (size_t)ncta * SEL_MAXG * SEL_BINS * sizeof(uint32_t), c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_hist, (size_t)(SEL_MAXG * SEL_BINS + 32) * sizeof(uint32_t), c.stream));
  d_flag = d_hist + SEL_MAXG * SEL_BINS;
  MCM_CUDA(cudaMemsetAsync(d_flag, 0, 32 * sizeof(uint32_t), c.stream));
  static bool attr = false;
  if (!attr) {
    MCM_CUDA(cudaFuncSetAttribute(k_sel_hist<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_MAXG * SEL_BINS * 4));
    MCM_CUDA(cudaFuncSetAttribute(k_sel_hist<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SEL_MAXG * SEL_BINS * 4));
    attr = true;
  }
  std::vector<SelGroup> done, active;
  {
    SelGroup g0;
    g0.prefix = 0; g0.rem = KB; g0.count = n;
    for (int i = 0; i < nr; ++i) {
      if (ranks[i] >= n) return fail(MCM_ERR_INVALID, "order_statistics: rank out of range");
      g0.ranks.push_back({ranks[i], i});
    }
    active.push_back(g0);
  }
  std::vector<uint32_t> h((size_t)SEL_MAXG * SEL_BINS + 32);
  bool first = true;
  *has_nan = 0;
  while (!active.empty()) {
    std::vector<SelGroup> next;
    for (size_t b0 = 0; b0 < active.size(); b0 += SEL_MAXG) {
      const int G = (int)std::min<size_t>(SEL_MAXG, active.size() - b0);
      SelParams p;
      memset(&p, 0, sizeof(p));
      p.G = G; p.rem = active[b0].rem; p.dbits = std::min(SEL_BITS, p.rem);
      for (int g = 0; g < G; ++g) p.prefix[g] = active[b0 + g].prefix;
      const int nbt = G * SEL_BINS;
      k_sel_hist<T><<<ncta, kThreads, (size_t)nbt * 4, c.stream>>>(in, n, p, d_part, d_flag);
      MCM_CHECK_LAUNCH();
      k_sel_sum<<<(nbt + 255) / 256, 256, 0, c.stream>>>(d_part, ncta, nbt, d_hist);
      MCM_CHECK_LAUNCH();
      MCM_CUDA(cudaMemcpyAsync(c.h_pinned, d_hist, ((size_t)SEL_MAXG * SEL_BINS + 32) * 4, cudaMemcpyDeviceToHost, c.stream));
      MCM_CUDA(cudaStreamSynchronize(c.stream));
      memcpy(h.data(), c.h_pinned, h.size() * 4);
      if (first) { *has_nan = h[(size_t)SEL_MAXG * SEL_BINS] != 0; first = false; }
      for (int g = 0; g < G; ++g) {
        SelGroup& grp = active[b0 + g];
        std::sort(grp.ranks.begin(), grp.ranks.end());
        const uint32_t* hg = h.data() + (size_t)g * SEL_BINS;
        unsigned long long cum = 0;
        size_t ri = 0;
        const int nb = 1 << p.dbits;
        for (int b = 0; b < nb && ri < grp.ranks.size(); ++b) {
          const unsigned long long hi = cum + hg[b];
          if (grp.ranks[ri].first < hi) {
            SelGroup ng;
            ng.prefix = (grp.prefix << p.dbits) | (unsigned long long)b;
            ng.rem = grp.rem - p.dbits;
            ng.count = hg[b];
            while (ri < grp.ranks.size() && grp.ranks[ri].first < hi) {
              ng.ranks.push_back({grp.ranks[ri].first - cum, grp.ranks[ri].second});
              ++ri;
            }
            if (ng.rem == 0 || ng.count <= SEL_THRESH) done.push_back(ng); else next.push_back(ng);
          }
          cum = hi;
        }
        if (ri < grp.ranks.size()) return fail(MCM_ERR_CUDA, "order_statistics: histogram does not cover the requested rank");
      }
    }
    active.swap(next);
  }
  // groups whose key is fully resolved need no candidates
  std::vector<SelGroup> need;
  for (auto& g : done) {
    if (g.rem == 0) {
      // every key in the group is identical: the value is the prefix itself
      K key = (K)g.prefix;
      double v;
      {
        // decode on the host (same transform as from_key)
        if (sizeof(K) == 8) {
          unsigned long long k = (unsigned long long)key;
          if (k == ~0ull) v = NAN;
          else { unsigned long long u = k ^ ((k >> 63) ? 0x8000000000000000ull : ~0ull); memcpy(&v, &u, 8); }
        } else {
          uint32_t k = (uint32_t)key;
          float f;
          if (k == ~0u) f = NAN;
          else { uint32_t u = k ^ ((k >> 31) ? 0x80000000u : ~0u); memcpy(&f, &u, 4); }
          v = (double)f;
        }
      }
      for (auto& r : g.ranks) vals_host[r.second] = v;
    } else {
      need.push_back(g);
    }
  }
  int rc = MCM_OK;
  for (size_t b0 = 0; b0 < need.size() && rc == MCM_OK; b0 += 2 * SEL_MAXG) {
    const int G = (int)std::min<size_t>(2 * SEL_MAXG, need.size() - b0);
    CompactParams cp;
    PickParams pp;
    memset(&cp, 0, sizeof(cp));
    memset(&pp, 0, sizeof(pp));
    cp.G = G;
    uint32_t off = 0, nrk = 0;
    for (int g = 0; g < G; ++g) {
      const SelGroup& grp = need[b0 + g];
      cp.prefix[g] = grp.prefix; cp.rem[g] = grp.rem; cp.offset[g] = off;
      pp.offset[g] = off; pp.count[g] = (uint32_t)grp.count; pp.rank_begin[g] = nrk;
      for (auto& r : grp.ranks) {
        pp.rank_local[nrk] = (uint32_t)r.first;
        pp.rank_out[nrk] = (uint32_t)r.second;
        ++nrk;
      }
      off += (uint32_t)grp.count;
    }
    pp.rank_begin[G] = nrk;
    unsigned long long* d_cand = nullptr;
    uint32_t* d_cnt = nullptr;
    double* d_out = nullptr;
    MCM_CUDA(cudaMallocAsync((void**)&d_cand, (size_t)std::max<uint32_t>(off, 1) * 8, c.stream));
    MCM_CUDA(cudaMallocAsync((void**)&d_cnt, 2 * SEL_MAXG * 4, c.stream));
    MCM_CUDA(cudaMallocAsync((void**)&d_out, 64 * 8, c.stream));
    MCM_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * SEL_MAXG * 4, c.stream));
    k_sel_compact<T><<<ncta, kThreads, 0, c.stream>>>(in, n, cp, d_cand, d_cnt);
    MCM_CHECK_LAUNCH();
    k_sel_pick<T><<<G, 1024, 0, c.stream>>>(d_cand, pp, d_out);
    MCM_CHECK_LAUNCH();
    double tmp[64];
    rc = read_back(tmp, d_out, 64);
    if (rc == MCM_OK)
      for (uint32_t r = 0; r < nrk; ++r) vals_host[pp.rank_out[r]] = tmp[pp.rank_out[r]];
    MCM_CUDA(cudaFreeAsync(d_cand, c.stream));
    MCM_CUDA(cudaFreeAsync(d_cnt, c.stream));
    MCM_CUDA(cudaFreeAsync(d_out, c.stream));
  }
  MCM_CUDA(cudaFreeAsync(d_part, c.stream));
  MCM_CUDA(cudaFreeAsync(d_hist, c.stream));
  return rc;
}

template int order_statistics<double>(const double*, size_t, const unsigned long long*, int, double*, int*);
template int order_statistics<float>(const float*, size_t, const unsigned long long*, int, double*, int*);

}  // namespace mcm
