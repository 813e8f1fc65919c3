// sort.cu — hand-written stable LSD radix sort (8-bit digits) for particle
// vectors, sortperm, and pquantile/pmedian on top of it.
//
// Reference: pquantile(p,q) = Statistics.quantile(p.particles, q), pmedian
// (src/particles.jl:316-323); sort is also what wasserstein (src/distances.jl:7-13)
// needs.  Order is Julia's `isless` on floats: -0.0 < +0.0, NaN last.
//
// Per pass: k_hist (per-CTA digit histogram: every thread owns a byte column of a [256 digits][256 threads] table,
// thread d sums row d with dp4a — no atomics), an exclusive scan over the [digit][CTA] table, and k_scatter, which
// ranks keys stably without warp votes or atomics (own-column byte counters per warp, turned in place into lane
// prefixes so that a key finds its rank inside the warp with one 32-bit load), stages the tile in digit order in shared memory and writes runs of equal digits to consecutive
// addresses.  (Earlier versions ranked with __match_any_sync / ballots and shared atomics: see the history notes at
// the kernels.)
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_keys.cuh"
#include "mcm_quantile.cuh"

namespace mcm {

constexpr int RB = 8;
constexpr int NBINS = 1 << RB;
constexpr int ST = 256;          // threads per CTA (== NBINS, one thread per digit in the prefix step)
constexpr int IPT = 16;          // keys per thread
constexpr int TILE = ST * IPT;   // keys per CTA
constexpr int SEG = 32 * IPT;    // keys per warp segment

// keys + the bitwise OR and AND of all of them (orand[0] |= key, orand[1] &= key): a digit whose bits are the same in
// every key (OR == AND there) needs no pass — typical particle data shares sign and most of the exponent
template <typename T, typename K>
__global__ void __launch_bounds__(256) k_to_keys(const T* __restrict__ in, K* __restrict__ keys, size_t n, uint32_t* flags,
                                                 unsigned long long* __restrict__ orand) {
  const size_t stride = (size_t)gridDim.x * 256;
  bool nan = false;
  unsigned long long o = 0ull, a = ~0ull;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    const K k = to_key(in[i], nan);
    keys[i] = k;
    o |= (unsigned long long)k; a &= (unsigned long long)k;
  }
  if (nan) atomicOr(flags, (uint32_t)MCM_FLAG_NAN_QUANTILE);
  if (orand) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      o |= __shfl_xor_sync(0xffffffffu, o, d);
      a &= __shfl_xor_sync(0xffffffffu, a, d);
    }
    if ((threadIdx.x & 31) == 0) { atomicOr(&orand[0], o); atomicAnd(&orand[1], a); }
  }
}
template <typename T, typename K>
__global__ void __launch_bounds__(256) k_from_keys(const K* __restrict__ keys, T* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) out[i] = from_key(keys[i]);
}
__global__ void __launch_bounds__(256) k_iota(uint32_t* v, size_t n) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) v[i] = (uint32_t)i;
}
__global__ void __launch_bounds__(256) k_u32_to_i64(const uint32_t* __restrict__ v, int64_t* __restrict__ o, size_t n) {
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) o[i] = (int64_t)v[i];
}

// 16-byte asynchronous copies global -> shared (LDGSTS): both kernels of a pass are persistent and fetch the NEXT tile's
// keys while they work on the current one.  Before round 2's last session every CTA loaded its tile, waited out the HBM
// latency and only then started: ncu showed 22-26 % of the stall samples on the first use of the keys and half of the
// HBM rate (k_hist read at 3.4 TB/s).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Lanes of the warp that hold the same 8-bit digit, by 8 ballots (VOTE + LOP3 each).  __match_any_sync gives the same
// mask in one instruction but is far slower here: the histogram kernel built on it ran at 0.53 keys/clk/SM
// (0.675 ms per 1e8 keys).  Invalid lanes get an empty mask.
__device__ __forceinline__ uint32_t digit_peers(uint32_t d, bool valid) {
  uint32_t mask = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t vote = __ballot_sync(0xffffffffu, bit);
    mask &= bit ? vote : ~vote;
  }
  return valid ? mask : 0u;
}

// ---- per-CTA digit histogram ---------------------------------------------------
// Counting WITHOUT atomics: shared-memory atomics cost 2 cycles per lane, and with near-uniform digits every lane is
// its own leader, so the atomic unit alone needs 64 cycles per 32 keys (0.675 ms per 1e8 keys, measured).  Here every
// thread owns one byte column of a [256 digits][256 threads] table (a thread sees at most IPT = 16 keys, so a byte
// is enough), increments it with plain loads/stores, and thread d then sums row d with dp4a.  Rows are padded to
// 17 x 16 bytes: the row sums are 128-bit loads (rows 17 slots apart: the eight threads of a quarter-warp hit eight
// different 16-byte slots), the column updates of a warp spread over all banks (bank = 4 d + lane / 4 mod 32).
constexpr int HROW = ST + 16;   // bytes per digit row
__device__ __forceinline__ void hist_fold_orand(unsigned long long o, unsigned long long a, bool nan, uint32_t* flags,
                                                unsigned long long* orand) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    o |= __shfl_xor_sync(0xffffffffu, o, d);
    a &= __shfl_xor_sync(0xffffffffu, a, d);
  }
  if (__any_sync(0xffffffffu, nan) && (threadIdx.x & 31) == 0) atomicOr(flags, (uint32_t)MCM_FLAG_NAN_QUANTILE);
  if ((threadIdx.x & 31) == 0) {
    const unsigned long long co = *reinterpret_cast<volatile unsigned long long*>(&orand[0]);
    const unsigned long long ca = *reinterpret_cast<volatile unsigned long long*>(&orand[1]);
    if (o & ~co) atomicOr(&orand[0], o);
    if (~a & ca) atomicAnd(&orand[1], a);
  }
}
// XIN: `keys` holds the raw bits of the floating-point input; the order-preserving key is formed in registers (the
// first pass of mcm_sort reads the particles directly: no separate key-transform pass), NaNs raise the sticky flag,
// and the OR / AND of all keys are folded into orand[0] / orand[1] (atomics only while a warp still changes them).
template <typename K> struct FloatOf;
template <> struct FloatOf<uint64_t> { using type = double; };
template <> struct FloatOf<uint32_t> { using type = float; };
template <typename K> __device__ __forceinline__ K key_of_bits(K raw, bool& nan) {
  typename FloatOf<K>::type x;
  memcpy(&x, &raw, sizeof(K));
  return to_key(x, nan);
}
template <typename K> __device__ __forceinline__ K bits_of_key(K k) {
  const typename FloatOf<K>::type x = from_key(k);
  K raw;
  memcpy(&raw, &x, sizeof(K));
  return raw;
}
constexpr int HIST_TABLE = NBINS * HROW;
template <typename K> constexpr int hist_smem() { return HIST_TABLE + TILE * (int)sizeof(K); }
template <typename K>
__device__ __forceinline__ void hist_prefetch(unsigned char* kbuf, const K* __restrict__ keys, uint32_t tile) {
  const unsigned char* src = reinterpret_cast<const unsigned char*>(keys + (size_t)tile * TILE);
#pragma unroll
  for (int q = 0; q < TILE * (int)sizeof(K) / 16 / ST; ++q) cp_async16(kbuf + (q * ST + threadIdx.x) * 16, src + (size_t)(q * ST + threadIdx.x) * 16);
  cp_async_commit();
}
template <typename K, bool XIN>
__global__ void __launch_bounds__(ST, 2)
k_hist(const K* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist, uint32_t nblocks,
       uint32_t* __restrict__ flags, unsigned long long* __restrict__ orand) {
  extern __shared__ __align__(16) unsigned char hraw[];
  unsigned char* hcnt = hraw;                       // [NBINS][HROW]
  unsigned char* kbuf = hraw + HIST_TABLE;          // the tile's keys, fetched one tile ahead
  const K* kb = reinterpret_cast<const K*>(kbuf);
  const uint32_t nfull = n / TILE;                  // tiles below nfull are full
  uint32_t tile = blockIdx.x;
  if (tile < nfull) hist_prefetch<K>(kbuf, keys, tile);
  bool nan = false;                                 // XIN: folded over all tiles of this CTA, reduced once at the end
  unsigned long long o = 0ull, a = ~0ull;
  for (; tile < nblocks; tile += gridDim.x) {
    {
      uint4* z = reinterpret_cast<uint4*>(hcnt);
#pragma unroll
      for (int i = 0; i < HIST_TABLE / 16 / ST; ++i) z[i * ST + threadIdx.x] = make_uint4(0, 0, 0, 0);
      static_assert(HIST_TABLE % (16 * ST) == 0, "table zeroed in whole rounds");
    }
    const uint32_t base = tile * TILE;
    const bool full = tile < nfull;
    K k[IPT];
    if (full) {
      cp_async_wait_all();
      __syncthreads();                               // every thread's copies have landed; the table is zeroed
#pragma unroll
      for (int i = 0; i < IPT; ++i) k[i] = kb[i * ST + threadIdx.x];
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const uint32_t idx = base + i * ST + threadIdx.x;
        k[i] = idx < n ? keys[idx] : (K)0;
      }
    }
    __syncthreads();                                 // the key buffer has been read: the next tile may overwrite it
    if (tile + gridDim.x < nfull) hist_prefetch<K>(kbuf, keys, tile + gridDim.x);
    if constexpr (XIN) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        if (full || base + i * ST + threadIdx.x < n) { k[i] = key_of_bits<K>(k[i], nan); o |= (unsigned long long)k[i]; a &= (unsigned long long)k[i]; }
      }
    }
    if (full) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const uint32_t d = (uint32_t)((k[i] >> shift) & (NBINS - 1));
        hcnt[d * HROW + threadIdx.x] += 1;
      }
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        if (base + i * ST + threadIdx.x < n) {
          const uint32_t d = (uint32_t)((k[i] >> shift) & (NBINS - 1));
          hcnt[d * HROW + threadIdx.x] += 1;
        }
      }
    }
    __syncthreads();
    const uint4* row = reinterpret_cast<const uint4*>(hcnt + threadIdx.x * HROW);
    uint32_t tot = 0, tot2 = 0;
#pragma unroll 8
    for (int w = 0; w < ST / 16; ++w) {
      const uint4 r = row[w];
      tot = __dp4a(r.x, 0x01010101u, tot); tot2 = __dp4a(r.y, 0x01010101u, tot2);
      tot = __dp4a(r.z, 0x01010101u, tot); tot2 = __dp4a(r.w, 0x01010101u, tot2);
    }
    hist[(size_t)threadIdx.x * nblocks + tile] = tot + tot2;
    __syncthreads();                                 // the row sums are done before the table is zeroed again
  }
  if constexpr (XIN) hist_fold_orand(o, a, nan, flags, orand);
}

// ---- exclusive scan (3 kernels; chunk = 2048 per CTA) ---------------------------
constexpr int SCAN_IPT = 8;
constexpr int SCAN_CHUNK = 256 * SCAN_IPT;

template <typename U>
__device__ __forceinline__ U block_exclusive_scan(U v, U* total) {
  // v: this thread's value; returns exclusive prefix within the CTA of 256 threads
  __shared__ U wsum[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  U inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const U t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  U woff = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < w) woff += wsum[i];
    tot += wsum[i];
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

template <typename U>
__global__ void __launch_bounds__(256) k_scan_reduce(const U* __restrict__ a, size_t n, U* __restrict__ sums) {
  const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_IPT;
  U s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; ++i)
    if (base + i < n) s += a[base + i];
  U tot;
  block_exclusive_scan<U>(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}
template <typename U>
__global__ void __launch_bounds__(256) k_scan_sums(U* sums, size_t nb, U* grand_total) {
  __shared__ U carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (size_t b0 = 0; b0 < nb; b0 += 256) {
    const size_t i = b0 + threadIdx.x;
    const U v = i < nb ? sums[i] : (U)0;
    U tot;
    const U ex = block_exclusive_scan<U>(v, &tot);
    const U carry = carry_s;
    if (i < nb) sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && grand_total) *grand_total = carry_s;
}
// INCLUSIVE=false: exclusive scan in place; true: inclusive
template <typename U, bool INCLUSIVE>
__global__ void __launch_bounds__(256) k_scan_final(U* a, size_t n, const U* __restrict__ sums) {
  const size_t base = (size_t)blockIdx.x * SCAN_CHUNK + (size_t)threadIdx.x * SCAN_IPT;
  U v[SCAN_IPT];
  U s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_IPT; ++i) {
    v[i] = (base + i < n) ? a[base + i] : (U)0;
    s += v[i];
  }
  U tot;
  U run = block_exclusive_scan<U>(s, &tot) + sums[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_IPT; ++i) {
    if (INCLUSIVE) run += v[i];
    if (base + i < n) a[base + i] = run;
    if (!INCLUSIVE) run += v[i];
  }
}

template <typename U, bool INCLUSIVE>
int device_scan(U* a, size_t n, U* grand_total_dev) {
  Ctx& c = ctx();
  const size_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
  U* sums = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&sums, nb * sizeof(U), c.stream));
  k_scan_reduce<U><<<(unsigned)nb, 256, 0, c.stream>>>(a, n, sums);
  MCM_CHECK_LAUNCH();
  k_scan_sums<U><<<1, 256, 0, c.stream>>>(sums, nb, grand_total_dev);
  MCM_CHECK_LAUNCH();
  k_scan_final<U, INCLUSIVE><<<(unsigned)nb, 256, 0, c.stream>>>(a, n, sums);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(sums, c.stream));
  return MCM_OK;
}
template int device_scan<uint32_t, false>(uint32_t*, size_t, uint32_t*);
template int device_scan<unsigned long long, true>(unsigned long long*, size_t, unsigned long long*);

// ---- stable scatter ---------------------------------------------------------------
// No warp votes, no atomics.  Thread t owns the 16 CONSECUTIVE keys [16t, 16t+16) of the tile.
//   1. every thread counts its keys in its own byte column of a per-warp table cnt[warp][digit][lane]; the value
//      read before each increment is the key's rank among the thread's own earlier keys of that digit;
//   2. thread d turns every row (warp, d) in place into lane prefixes (one 32-bit load per key then gives the number
//      of equal-digit keys held by lower lanes of the warp), the row totals into offsets over the warps and the tile
//      totals into an exclusive scan over the digits;
//   3. keys (and values) are placed into a shared-memory copy of the tile IN DIGIT ORDER (the copy reuses the
//      memory of the count table) and written out from there, so consecutive threads store runs of equal digits to
//      consecutive addresses.
// History at 1e8 64-bit keys per pass: __match_any ranking + 32 scattered stores per warp instruction 1.74 ms ->
// ballot ranking + digit-ordered staging 1.19 ms (113 instructions per key, vote-bound) -> this version.
// Count table: one 32-byte row of lane counters per (warp, digit), no padding (71 KiB of shared memory per CTA).  The
// per-key accesses (random digit per lane) spread over the banks by themselves; the per-digit pass over whole rows
// uses conflict-free 128-bit accesses (see 2a).  ncu (round 2): the kernel is latency-bound, not wavefront-bound
// (shared pipe 47 %, issue 38 %), so it runs 2 CTAs per SM with ~120 registers: the 16 keys of a thread in flight at once.
constexpr int CROW = 32;

template <typename K, bool HAS_VAL> struct ScatterSmem {
  static constexpr size_t KEYS = (size_t)TILE * sizeof(K);
  static constexpr size_t VALS = HAS_VAL ? (size_t)TILE * sizeof(uint32_t) : 0;
  static constexpr size_t CNT = (size_t)(ST / 32) * NBINS * CROW;
  static constexpr size_t UNION = (KEYS + VALS) > CNT ? (KEYS + VALS) : CNT;    // the staged tile reuses the count table
  static constexpr size_t WOFF = (size_t)(ST / 32) * NBINS * sizeof(uint16_t);     // offsets inside a tile: < 4096
  static constexpr size_t TABLES = UNION + WOFF + NBINS * sizeof(uint32_t);
  // the next tile's keys, fetched while this one is processed: one row of IPT keys per thread, padded by 16 bytes so
  // that the threads' 128-bit reads of their own rows are conflict-free (row stride = an odd number of 16-byte slots)
  static constexpr size_t KROW = IPT * sizeof(K) + 16;
  static constexpr size_t BYTES = TABLES + (size_t)ST * KROW;
};
// FULL: the tile holds TILE keys — no per-key bounds test, so the 16 keys of a thread are independent straight-line
// code (with the tests every key sat in its own BSSY/BSYNC region and its shared-memory loads could not be batched:
// ncu showed 52 % of the stall samples on short_scoreboard).  Only the last tile of a pass runs FULL = false.
// XIN: kin holds raw floating-point bits (keys formed in registers) and the values are the global indices themselves;
// XOUT: the keys leave as floating-point bits and the values as Int64 (the last pass of mcm_sort / mcm_sortperm).
// chunk g (16 bytes) of the tile goes to row g / CPR, slot g % CPR: coalesced global reads, conflict-free shared writes
template <typename K, bool HAS_VAL>
__device__ __forceinline__ void scatter_prefetch(unsigned char* kbuf, const K* __restrict__ kin, uint32_t tile) {
  constexpr int CPR = IPT * (int)sizeof(K) / 16;      // chunks per row
  constexpr int KROW = (int)ScatterSmem<K, HAS_VAL>::KROW;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(kin + (size_t)tile * TILE);
#pragma unroll
  for (int q = 0; q < CPR; ++q) {
    const int g = q * ST + threadIdx.x;
    cp_async16(kbuf + (g / CPR) * KROW + (g % CPR) * 16, src + (size_t)g * 16);
  }
  cp_async_commit();
}
template <typename K, bool HAS_VAL, bool FULL, bool XIN, bool XOUT>
__device__ __forceinline__ void scatter_tile(unsigned char* sraw, uint32_t tile, bool prefetch_next, uint32_t next_tile,
                                             const K* __restrict__ kin, K* __restrict__ kout,
                                             const uint32_t* __restrict__ vin, uint32_t* __restrict__ vout,
                                             int64_t* __restrict__ vout64, uint32_t n,
                                             int shift, const uint32_t* __restrict__ hist_scanned, uint32_t nblocks) {
  using SM = ScatterSmem<K, HAS_VAL>;
  unsigned char* cnt = sraw;                                             // [warp][digit][CROW]
  K* skeys = reinterpret_cast<K*>(sraw);                                 // (after the counting phase)
  uint32_t* svals = reinterpret_cast<uint32_t*>(sraw + SM::KEYS);
  uint16_t (*woff)[NBINS] = reinterpret_cast<uint16_t (*)[NBINS]>(sraw + SM::UNION);
  uint32_t* gdel = reinterpret_cast<uint32_t*>(sraw + SM::UNION + SM::WOFF);   // global position minus tile slot
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  {
    uint4* z = reinterpret_cast<uint4*>(cnt);
    for (int i = threadIdx.x; i < (int)(SM::CNT / 16); i += ST) z[i] = make_uint4(0, 0, 0, 0);
  }
  unsigned char* kbuf = sraw + SM::TABLES;
  const uint32_t tile0 = tile * TILE;
  const uint32_t mine = tile0 + threadIdx.x * IPT;                      // first key of this thread
  const uint32_t ntile = min((uint32_t)TILE, n - tile0);
  K key[IPT];
  uint32_t val[IPT];
  if constexpr (FULL) {
    cp_async_wait_all();
    __syncthreads();                                                    // every thread's copies of this tile have landed
  }
  if (FULL || mine + IPT <= n) {                                        // 128 contiguous bytes per thread
    constexpr int KV = 16 / (int)sizeof(K);
    const uint4* p = FULL ? reinterpret_cast<const uint4*>(kbuf + (size_t)threadIdx.x * SM::KROW)
                          : reinterpret_cast<const uint4*>(kin + mine);
#pragma unroll
    for (int q = 0; q < IPT / KV; ++q) {
      const uint4 t = p[q];
      const K* tk = reinterpret_cast<const K*>(&t);
#pragma unroll
      for (int e = 0; e < KV; ++e) key[q * KV + e] = tk[e];
    }
    if constexpr (HAS_VAL && !XIN) {
      const uint4* pv = reinterpret_cast<const uint4*>(vin + mine);
#pragma unroll
      for (int q = 0; q < IPT / 4; ++q) {
        const uint4 t = pv[q];
        val[4 * q] = t.x; val[4 * q + 1] = t.y; val[4 * q + 2] = t.z; val[4 * q + 3] = t.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      key[j] = mine + j < n ? kin[mine + j] : (K)0;
      if constexpr (HAS_VAL && !XIN) val[j] = mine + j < n ? vin[mine + j] : 0u;
    }
  }
  if constexpr (XIN) {
    bool nan_unused = false;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
      key[j] = key_of_bits<K>(key[j], nan_unused);
      if constexpr (HAS_VAL) val[j] = mine + j;
    }
  }
  __syncthreads();                                                      // table zeroed; the key buffer has been read
  if (prefetch_next) scatter_prefetch<K, HAS_VAL>(kbuf, kin, next_tile);
  // 1. own-column counting; own[j] = rank among this thread's earlier keys with the same digit
  unsigned char* wcnt = cnt + (size_t)w * NBINS * CROW;
  unsigned long long own64 = 0;   // 16 x 4-bit ranks packed (a thread has at most 16 keys per digit: rank <= 15)
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if (FULL || mine + j < n) {
      const uint32_t d = (uint32_t)((key[j] >> shift) & (NBINS - 1));
      unsigned char* cell = wcnt + d * CROW + lane;
      const unsigned char c = *cell;
      *cell = (unsigned char)(c + 1);
      own64 |= (unsigned long long)c << (4 * j);
    }
  }
  __syncthreads();
  // 2a. per digit d (thread d), per warp: the 32 lane counters of row (warp, d) become, IN PLACE, what a key needs to
  //     find the equal-digit keys held by lower lanes with ONE 32-bit load: byte k >= 1 of word q = counters of lanes
  //     4q .. 4q+k-1 (<= 48: 6 bits); byte 0 (always 0 as an in-word prefix) and bit 15 = the 9-bit total of the words
  //     before q (<= 448).  (Until round 2 every key summed the eight words of its row with masked dp4a: 8 random-row
  //     loads per key were 28 of the kernel's ~58 shared-memory wavefronts per warp instruction.)
  //     Row totals -> offsets over the warps; tile totals -> exclusive scan over the digits, folded into woff.
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
    uint32_t wrun[ST / 32];
    // a row is two 128-bit words; threads 4..7 of every eight take the halves in the other order, so the eight
    // threads of a quarter-warp (rows 32 bytes apart) cover eight different 16-byte slots: no bank conflicts
    const bool swp = (d >> 2) & 1;
#pragma unroll
    for (int ww = 0; ww < ST / 32; ++ww) {
      uint4* row = reinterpret_cast<uint4*>(cnt + ((size_t)ww * NBINS + d) * CROW);
      const uint4 f = row[swp ? 1 : 0], g = row[swp ? 0 : 1];
      uint32_t wv[8];
      wv[0] = swp ? g.x : f.x; wv[1] = swp ? g.y : f.y; wv[2] = swp ? g.z : f.z; wv[3] = swp ? g.w : f.w;
      wv[4] = swp ? f.x : g.x; wv[5] = swp ? f.y : g.y; wv[6] = swp ? f.z : g.z; wv[7] = swp ? f.w : g.w;
      uint32_t wp = 0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t c = wv[q];
        wv[q] = ((c * 0x01010100u) & 0x3f3f3f00u) | (wp & 0xffu) | ((wp & 0x100u) << 7);
        wp = __dp4a(c, 0x01010101u, wp);
      }
      const uint4 lo = make_uint4(wv[0], wv[1], wv[2], wv[3]), hi = make_uint4(wv[4], wv[5], wv[6], wv[7]);
      row[swp ? 1 : 0] = swp ? hi : lo;
      row[swp ? 0 : 1] = swp ? lo : hi;
      wrun[ww] = run;
      run += wp;
    }
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan<uint32_t>(run, &tot);
#pragma unroll
    for (int ww = 0; ww < ST / 32; ++ww) woff[ww][d] = (uint16_t)(ex + wrun[ww]);   // first tile slot of (warp, digit)
    gdel[d] = hist_scanned[(size_t)d * nblocks + tile] - ex;
  }
  __syncthreads();
  // 2b. tile slot = first slot of (warp, digit) + equal-digit keys of lower lanes + rank among the thread's own keys
  uint32_t slot[IPT];
  const int lq = lane >> 2;
  const uint32_t ish = 8u * (uint32_t)(lane & 3), imask = (lane & 3) ? 0x3fu : 0u;
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    slot[j] = 0;
    if (FULL || mine + j < n) {
      const uint32_t d = (uint32_t)((key[j] >> shift) & (NBINS - 1));
      const uint32_t wv = reinterpret_cast<const uint32_t*>(wcnt + d * CROW)[lq];
      const uint32_t before = (wv & 0xffu) + ((wv >> 7) & 0x100u) + ((wv >> ish) & imask);
      slot[j] = (uint32_t)woff[w][d] + before + (uint32_t)((own64 >> (4 * j)) & 15ull);
    }
  }
  __syncthreads();   // the count table is dead: its memory becomes the staged tile
#pragma unroll
  for (int j = 0; j < IPT; ++j) {
    if (FULL || mine + j < n) {
      skeys[slot[j]] = key[j];
      if constexpr (HAS_VAL) svals[slot[j]] = val[j];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < IPT; ++r) {
    const uint32_t sl = r * ST + threadIdx.x;
    if (FULL || sl < ntile) {
      const K k = skeys[sl];
      const uint32_t d = (uint32_t)((k >> shift) & (NBINS - 1));
      const uint32_t pos = gdel[d] + sl;
      kout[pos] = XOUT ? bits_of_key<K>(k) : k;
      if constexpr (HAS_VAL) {
        if constexpr (XOUT) vout64[pos] = (int64_t)svals[sl];
        else vout[pos] = svals[sl];
      }
    }
  }
  __syncthreads();                                                      // the staged tile and gdel are dead: next tile
}
#ifndef MCM_SCATTER_MIN_CTAS
#define MCM_SCATTER_MIN_CTAS 2
#endif
template <typename K, bool HAS_VAL, bool XIN, bool XOUT>
__global__ void __launch_bounds__(ST, MCM_SCATTER_MIN_CTAS)
k_scatter(const K* __restrict__ kin, K* __restrict__ kout, const uint32_t* __restrict__ vin,
          uint32_t* __restrict__ vout, int64_t* __restrict__ vout64, uint32_t n, int shift,
          const uint32_t* __restrict__ hist_scanned, uint32_t nblocks) {
  extern __shared__ __align__(16) unsigned char sraw[];
  const uint32_t nfull = n / TILE;                  // tiles below nfull are full; only they are fetched ahead
  uint32_t tile = blockIdx.x;
  if (tile < nfull) scatter_prefetch<K, HAS_VAL>(sraw + ScatterSmem<K, HAS_VAL>::TABLES, kin, tile);
  for (; tile < nblocks; tile += gridDim.x) {
    const uint32_t next = tile + gridDim.x;
    if (tile < nfull)
      scatter_tile<K, HAS_VAL, true, XIN, XOUT>(sraw, tile, next < nfull, next, kin, kout, vin, vout, vout64, n, shift, hist_scanned, nblocks);
    else
      scatter_tile<K, HAS_VAL, false, XIN, XOUT>(sraw, tile, false, next, kin, kout, vin, vout, vout64, n, shift, hist_scanned, nblocks);
  }
}
template <typename K, bool HAS_VAL, bool XIN, bool XOUT>
static int scatter_launch(const K* kin, K* kout, const uint32_t* vin, uint32_t* vout, int64_t* vout64, uint32_t n, int shift,
                          const uint32_t* hist_scanned, uint32_t nblocks) {
  static bool attr = false;
  if (!attr) {
    MCM_CUDA(cudaFuncSetAttribute(k_scatter<K, HAS_VAL, XIN, XOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ScatterSmem<K, HAS_VAL>::BYTES));
    attr = true;
  }
  const uint32_t grid = std::min<uint32_t>(nblocks, (uint32_t)(MCM_SCATTER_MIN_CTAS * ctx().sm_count));   // persistent: tiles strided over the CTAs
  k_scatter<K, HAS_VAL, XIN, XOUT><<<grid, ST, ScatterSmem<K, HAS_VAL>::BYTES, ctx().stream>>>(kin, kout, vin, vout, vout64, n, shift, hist_scanned, nblocks);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}
template <typename K, bool XIN>
static int hist_launch(const K* keys, uint32_t n, int shift, uint32_t* hist, uint32_t nblocks, unsigned long long* orand) {
  static bool attr = false;
  if (!attr) {
    MCM_CUDA(cudaFuncSetAttribute(k_hist<K, XIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, hist_smem<K>()));
    attr = true;
  }
  const uint32_t grid = std::min<uint32_t>(nblocks, (uint32_t)(2 * ctx().sm_count));
  k_hist<K, XIN><<<grid, ST, hist_smem<K>(), ctx().stream>>>(keys, n, shift, hist, nblocks, ctx().d_flags, orand);
  MCM_CHECK_LAUNCH();
  return MCM_OK;
}

// Sort n keys held in bufA (ping-pong with bufB); the result is back in bufA (vals likewise).  `varying`: bits that
// differ between some two keys (OR ^ AND from k_to_keys; all ones if unknown) — passes over digits that are the same in
// every key are skipped (a stable pass on a constant digit is the identity).
template <typename K, bool HAS_VAL>
int radix_sort_keys_masked(K* bufA, K* bufB, uint32_t* valA, uint32_t* valB, size_t n, K varying) {
  Ctx& c = ctx();
  if (n >= 0xFFFF0000ull) return fail(MCM_ERR_UNSUPPORTED, "sort: n must be < 2^32");
  const uint32_t nblocks = (uint32_t)((n + TILE - 1) / TILE);
  uint32_t* hist = nullptr;
  const size_t hlen = (size_t)NBINS * nblocks;
  MCM_CUDA(cudaMallocAsync((void**)&hist, hlen * sizeof(uint32_t), c.stream));
  K* src = bufA; K* dst = bufB;
  uint32_t* vs = valA; uint32_t* vd = valB;
  constexpr int passes = (int)sizeof(K);
  for (int p = 0; p < passes; ++p) {
    const int shift = p * RB;
    if (((varying >> shift) & (K)(NBINS - 1)) == 0) continue;
    int rc = hist_launch<K, false>(src, (uint32_t)n, shift, hist, nblocks, nullptr);
    if (rc) return rc;
    rc = device_scan<uint32_t, false>(hist, hlen, nullptr);
    if (rc) return rc;
    rc = scatter_launch<K, HAS_VAL, false, false>(src, dst, vs, vd, nullptr, (uint32_t)n, shift, hist, nblocks);
    if (rc) return rc;
    K* t = src; src = dst; dst = t;
    uint32_t* tv = vs; vs = vd; vd = tv;
  }
  if (src != bufA) {          // an odd number of passes ran: bring the result home
    MCM_CUDA(cudaMemcpyAsync(bufA, src, n * sizeof(K), cudaMemcpyDeviceToDevice, c.stream));
    if (HAS_VAL) MCM_CUDA(cudaMemcpyAsync(valA, vs, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
  }
  MCM_CUDA(cudaFreeAsync(hist, c.stream));
  return MCM_OK;
}
template <typename K, bool HAS_VAL>
int radix_sort_keys(K* bufA, K* bufB, uint32_t* valA, uint32_t* valB, size_t n) {
  return radix_sort_keys_masked<K, HAS_VAL>(bufA, bufB, valA, valB, n, (K)~(K)0);
}
template int radix_sort_keys<uint64_t, true>(uint64_t*, uint64_t*, uint32_t*, uint32_t*, size_t);

template <typename T, typename K>
static int sort_impl_unfused(const T* in, T* out, int64_t* perm_out, size_t n) {
  Ctx& c = ctx();
  if (n == 0) return MCM_OK;
  K* A = nullptr; K* B = nullptr;
  uint32_t* vA = nullptr; uint32_t* vB = nullptr;
  const bool own_A = (out == nullptr);
  if (own_A) MCM_CUDA(cudaMallocAsync((void**)&A, n * sizeof(K), c.stream));
  else A = reinterpret_cast<K*>(out);
  MCM_CUDA(cudaMallocAsync((void**)&B, n * sizeof(K), c.stream));
  const int grid = stream_grid(n, 4, 8);
  unsigned long long* d_orand = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_orand, 2 * sizeof(unsigned long long), c.stream));
  const unsigned long long init[2] = {0ull, ~0ull};
  MCM_CUDA(cudaMemcpyAsync(d_orand, init, sizeof(init), cudaMemcpyHostToDevice, c.stream));
  k_to_keys<T, K><<<grid, 256, 0, c.stream>>>(in, A, n, c.d_flags, d_orand);
  MCM_CHECK_LAUNCH();
  unsigned long long h_orand[2] = {~0ull, 0ull};
  MCM_CUDA(cudaMemcpyAsync(h_orand, d_orand, sizeof(h_orand), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));      // one 16-byte read-back per sort decides which passes run
  MCM_CUDA(cudaFreeAsync(d_orand, c.stream));
  const K varying = (K)(h_orand[0] ^ h_orand[1]);
  int rc;
  if (perm_out) {
    MCM_CUDA(cudaMallocAsync((void**)&vA, n * sizeof(uint32_t), c.stream));
    MCM_CUDA(cudaMallocAsync((void**)&vB, n * sizeof(uint32_t), c.stream));
    k_iota<<<grid, 256, 0, c.stream>>>(vA, n);
    MCM_CHECK_LAUNCH();
    rc = radix_sort_keys_masked<K, true>(A, B, vA, vB, n, varying);
    if (rc) return rc;
    k_u32_to_i64<<<grid, 256, 0, c.stream>>>(vA, perm_out, n);
    MCM_CHECK_LAUNCH();
    MCM_CUDA(cudaFreeAsync(vA, c.stream));
    MCM_CUDA(cudaFreeAsync(vB, c.stream));
  } else {
    rc = radix_sort_keys_masked<K, false>(A, B, nullptr, nullptr, n, varying);
    if (rc) return rc;
  }
  if (!own_A) {
    k_from_keys<T, K><<<grid, 256, 0, c.stream>>>(A, out, n);  // in place: same element size
    MCM_CHECK_LAUNCH();
  } else {
    MCM_CUDA(cudaFreeAsync(A, c.stream));
  }
  MCM_CUDA(cudaFreeAsync(B, c.stream));
  return MCM_OK;
}

// mcm_sort / mcm_sortperm without separate transform passes: the first pass reads the particles themselves (its
// histogram kernel also returns the OR / AND of all keys, which decide the remaining passes), the last pass writes
// floating-point values and Int64 indices.  Against sort_impl_unfused this saves the key-transform round trip
// (k_to_keys + k_from_keys, 0.57 ms per 1e8 Float64) and, for sortperm, k_iota + k_u32_to_i64.
template <typename T, typename K, bool HAS_VAL>
static int sort_fused_passes(const K* in_bits, K* A, K* B, uint32_t* vA, uint32_t* vB, int64_t* perm_out, size_t n, K varying,
                             uint32_t* hist, uint32_t nblocks) {
  int plist[sizeof(K)], m = 0;
  for (int p = 0; p < (int)sizeof(K); ++p)
    if (((varying >> (p * RB)) & (K)(NBINS - 1)) != 0) plist[m++] = p;
  const size_t hlen = (size_t)NBINS * nblocks;
  const K* src = in_bits;
  const uint32_t* vsrc = nullptr;
  for (int i = 1; i <= m; ++i) {
    const int shift = plist[i - 1] * RB;
    K* dst = ((m - i) % 2 == 0) ? A : B;
    uint32_t* vdst = HAS_VAL ? (((m - i) % 2 == 0) ? vA : vB) : nullptr;
    int rc = MCM_OK;
    if (i > 1) rc = hist_launch<K, false>(src, (uint32_t)n, shift, hist, nblocks, nullptr);   // pass 1: done by the caller
    if (rc) return rc;
    rc = device_scan<uint32_t, false>(hist, hlen, nullptr);
    if (rc) return rc;
    const bool first = i == 1, last = i == m;
    if (first && last) rc = scatter_launch<K, HAS_VAL, true, true>(src, dst, vsrc, vdst, perm_out, (uint32_t)n, shift, hist, nblocks);
    else if (first) rc = scatter_launch<K, HAS_VAL, true, false>(src, dst, vsrc, vdst, perm_out, (uint32_t)n, shift, hist, nblocks);
    else if (last) rc = scatter_launch<K, HAS_VAL, false, true>(src, dst, vsrc, vdst, perm_out, (uint32_t)n, shift, hist, nblocks);
    else rc = scatter_launch<K, HAS_VAL, false, false>(src, dst, vsrc, vdst, perm_out, (uint32_t)n, shift, hist, nblocks);
    if (rc) return rc;
    src = dst; vsrc = vdst;
  }
  return MCM_OK;
}

static int g_sort_fused = 1;
extern "C" int mcm_sort_use_fused(int on) { g_sort_fused = on; return MCM_OK; }

template <typename T, typename K>
static int sort_impl(const T* in, T* out, int64_t* perm_out, size_t n) {
  Ctx& c = ctx();
  if (n == 0) return MCM_OK;
  // the fused form reads `in` while it writes `out` (no in-place calls), and needs 16-byte aligned rows for its vector loads
  if (!g_sort_fused || (const void*)in == (const void*)out || n >= 0xFFFF0000ull ||
      (reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) % 16 != 0)
    return sort_impl_unfused<T, K>(in, out, perm_out, n);
  const uint32_t nblocks = (uint32_t)((n + TILE - 1) / TILE);
  uint32_t* hist = nullptr;
  unsigned long long* d_orand = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&hist, (size_t)NBINS * nblocks * sizeof(uint32_t), c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_orand, 2 * sizeof(unsigned long long), c.stream));
  const unsigned long long init[2] = {0ull, ~0ull};
  MCM_CUDA(cudaMemcpyAsync(d_orand, init, sizeof(init), cudaMemcpyHostToDevice, c.stream));
  const K* in_bits = reinterpret_cast<const K*>(in);
  int rc = hist_launch<K, true>(in_bits, (uint32_t)n, 0, hist, nblocks, d_orand);   // digit 0 + OR / AND + NaN flag
  if (rc) { cudaFreeAsync(hist, c.stream); cudaFreeAsync(d_orand, c.stream); return rc; }
  unsigned long long h_orand[2] = {~0ull, 0ull};
  MCM_CUDA(cudaMemcpyAsync(h_orand, d_orand, sizeof(h_orand), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));      // one 16-byte read-back per sort decides which passes run
  MCM_CUDA(cudaFreeAsync(d_orand, c.stream));
  const K varying = (K)(h_orand[0] ^ h_orand[1]);
  if ((varying & (K)(NBINS - 1)) == 0) {          // the lowest digit is the same in every key: its histogram was for nothing
    MCM_CUDA(cudaFreeAsync(hist, c.stream));
    return sort_impl_unfused<T, K>(in, out, perm_out, n);
  }
  K* A = nullptr; K* B = nullptr;
  uint32_t* vA = nullptr; uint32_t* vB = nullptr;
  const bool own_A = (out == nullptr);
  if (own_A) MCM_CUDA(cudaMallocAsync((void**)&A, n * sizeof(K), c.stream));
  else A = reinterpret_cast<K*>(out);
  MCM_CUDA(cudaMallocAsync((void**)&B, n * sizeof(K), c.stream));
  if (perm_out) {
    MCM_CUDA(cudaMallocAsync((void**)&vA, n * sizeof(uint32_t), c.stream));
    MCM_CUDA(cudaMallocAsync((void**)&vB, n * sizeof(uint32_t), c.stream));
    rc = sort_fused_passes<T, K, true>(in_bits, A, B, vA, vB, perm_out, n, varying, hist, nblocks);
    MCM_CUDA(cudaFreeAsync(vA, c.stream));
    MCM_CUDA(cudaFreeAsync(vB, c.stream));
  } else {
    rc = sort_fused_passes<T, K, false>(in_bits, A, B, nullptr, nullptr, nullptr, n, varying, hist, nblocks);
  }
  if (own_A) MCM_CUDA(cudaFreeAsync(A, c.stream));
  MCM_CUDA(cudaFreeAsync(B, c.stream));
  MCM_CUDA(cudaFreeAsync(hist, c.stream));
  return rc;
}

template <typename T>
__global__ void k_pick(const T* __restrict__ sorted, const uint64_t* __restrict__ idx, int cnt, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cnt) out[i] = (double)sorted[idx[i]];
}

int read_back(double* host_dst, const double* dev_src, size_t count);

}  // namespace mcm

using namespace mcm;

extern "C" int mcm_sort(int dtype, const void* in, void* out, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !out)) return fail(MCM_ERR_INVALID, "mcm_sort: null pointer");
  if (dtype == MCM_F64) return sort_impl<double, uint64_t>((const double*)in, (double*)out, nullptr, n);
  if (dtype == MCM_F32) return sort_impl<float, uint32_t>((const float*)in, (float*)out, nullptr, n);
  return fail(MCM_ERR_INVALID, "mcm_sort: bad dtype %d", dtype);
}

extern "C" int mcm_sortperm(int dtype, const void* in, int64_t* perm_out, void* sorted_out_or_null, size_t n) {
  MCM_NEED_INIT();
  if (n && (!in || !perm_out)) return fail(MCM_ERR_INVALID, "mcm_sortperm: null pointer");
  if (dtype == MCM_F64) return sort_impl<double, uint64_t>((const double*)in, (double*)sorted_out_or_null, perm_out, n);
  if (dtype == MCM_F32) return sort_impl<float, uint32_t>((const float*)in, (float*)sorted_out_or_null, perm_out, n);
  return fail(MCM_ERR_INVALID, "mcm_sortperm: bad dtype %d", dtype);
}

static int g_use_select = 1;
extern "C" int mcm_quantile_use_select(int on) { g_use_select = on; return MCM_OK; }

namespace mcm {
template <typename T>
int order_statistics(const T* in, size_t n, const unsigned long long* ranks, int nr, double* vals_host, int* has_nan);
}

// order statistics through a full sort (A/B reference for the select path)
static int order_statistics_sort(int dtype, const void* in, size_t n, const unsigned long long* ranks, int nr,
                                 double* vals, int* has_nan) {
  Ctx& c = ctx();
  const size_t esz = dtype == MCM_F64 ? 8 : 4;
  void* sorted = nullptr;
  MCM_CUDA(cudaMallocAsync(&sorted, n * esz, c.stream));
  int rc = mcm_sort(dtype, in, sorted, n);
  if (rc) return rc;
  const int cnt = nr + 1;
  uint64_t* h_idx = static_cast<uint64_t*>(c.h_pinned);
  for (int i = 0; i < nr; ++i) h_idx[i] = ranks[i];
  h_idx[nr] = n - 1;
  uint64_t* d_idx = nullptr;
  double* d_val = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_idx, cnt * sizeof(uint64_t), c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_val, cnt * sizeof(double), c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_idx, h_idx, cnt * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
  if (dtype == MCM_F64) k_pick<double><<<(cnt + 127) / 128, 128, 0, c.stream>>>((const double*)sorted, d_idx, cnt, d_val);
  else k_pick<float><<<(cnt + 127) / 128, 128, 0, c.stream>>>((const float*)sorted, d_idx, cnt, d_val);
  MCM_CHECK_LAUNCH();
  std::vector<double> v(cnt);
  rc = read_back(v.data(), d_val, cnt);
  MCM_CUDA(cudaFreeAsync(d_idx, c.stream));
  MCM_CUDA(cudaFreeAsync(d_val, c.stream));
  MCM_CUDA(cudaFreeAsync(sorted, c.stream));
  if (rc) return rc;
  for (int i = 0; i < nr; ++i) vals[i] = v[i];
  *has_nan = v[nr] != v[nr];
  return MCM_OK;
}

// Statistics.quantile(v, p) with the default alpha = beta = 1 (type 7):
//   aleph = n*p + (1-p); j = clamp(trunc(aleph), 1, n-1); g = clamp(aleph-j, 0, 1)
//   a = v[j], b = v[j+1] (1-based, sorted);  a + g*(b-a)  (or (1-g)a + g b if not finite)
extern "C" int mcm_quantile(int dtype, const void* in, size_t n, const double* q_host, int nq, double* out_host) {
  MCM_NEED_INIT();
  if (!in || !q_host || !out_host || nq < 1) return fail(MCM_ERR_INVALID, "mcm_quantile: null pointer");
  if (n == 0) return fail(MCM_ERR_INVALID, "mcm_quantile: empty particles");
  if (nq > 4096) return fail(MCM_ERR_INVALID, "mcm_quantile: at most 4096 probabilities per call");
  for (int i = 0; i < nq; ++i)
    if (!(q_host[i] >= 0.0 && q_host[i] <= 1.0))
      return fail(MCM_ERR_INVALID, "mcm_quantile: input probability out of [0,1] range");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_quantile: bad dtype %d", dtype);
  // large inputs, few probabilities: sampled brackets + one streaming read (select_batch.cu)
  if (g_use_select == 1 && n >= 131072 && nq <= 8 && (reinterpret_cast<uintptr_t>(in) % (dtype == MCM_F64 ? 8 : 4)) == 0)
    return mcm_quantile_batch(dtype, &in, 1, n, q_host, nq, out_host);
  std::vector<double> gam(nq), vals(2 * (size_t)nq);
  std::vector<unsigned long long> ranks(2 * (size_t)nq);
  for (int i = 0; i < nq; ++i) quantile_rank_pair(n, q_host[i], ranks[2 * i], ranks[2 * i + 1], gam[i]);
  int has_nan = 0, rc = MCM_OK;
  const int nr = 2 * nq;
  if (g_use_select && (reinterpret_cast<uintptr_t>(in) % (dtype == MCM_F64 ? 8 : 4)) == 0) {
    for (int r0 = 0; r0 < nr && rc == MCM_OK; r0 += 64) {
      const int cnt = nr - r0 < 64 ? nr - r0 : 64;
      int hn = 0;
      rc = (dtype == MCM_F64)
               ? order_statistics<double>((const double*)in, n, ranks.data() + r0, cnt, vals.data() + r0, &hn)
               : order_statistics<float>((const float*)in, n, ranks.data() + r0, cnt, vals.data() + r0, &hn);
      has_nan |= hn;
    }
  } else {
    for (int r0 = 0; r0 < nr && rc == MCM_OK; r0 += 1024) {
      const int cnt = nr - r0 < 1024 ? nr - r0 : 1024;
      int hn = 0;
      rc = order_statistics_sort(dtype, in, n, ranks.data() + r0, cnt, vals.data() + r0, &hn);
      has_nan |= hn;
    }
  }
  if (rc) return rc;
  if (has_nan) return fail(MCM_ERR_INVALID, "mcm_quantile: quantiles are undefined in presence of NaNs");
  for (int i = 0; i < nq; ++i) out_host[i] = quantile_combine(dtype, vals[2 * i], vals[2 * i + 1], gam[i]);
  return MCM_OK;
}
