// gram_tc.cu — pcov on the 5th-generation tensor cores (tcgen05 + TMEM) for Float64 particles.
//
// tcgen05 has no f64 kind, so the Float64 Gram is computed by an Ozaki-style error-free
// splitting onto the int8 tensor path, which accumulates EXACTLY in int32:
//   1. every centred row (variable) x_i - mu_i is scaled by a power of two 2^-(e_i+1) to |y| < 1/2
//      and split into S = 6 signed 8-bit digits  y = sum_s d_s 2^(-8(s+1)),  d_s in [-128, 127]
//      (k_rowstats, k_slice; all steps exact in Float64; 48 bits are kept);
//   2. sum_k y_i y_j = sum_{s,t} 2^(-8(s+t+2)) D_s^i . D_t^j  with exact integer dot products.
//      Pairs with s+t = d share the weight, so they accumulate into ONE int32 TMEM accumulator
//      per diagonal d = 0..5 (tcgen05.mma kind::i8, M = 128, N = 64, K = 32 per instruction): 21 MMAs
//      per k-block; pairs with s + t > 5 fall below 2^-56 relative to |y_i||y_j| and are dropped;
//   3. every 16384 particles (int32 headroom: 6 pairs x 128 x 128 x 16384 < 2^31) the accumulators are
//      read back with tcgen05.ld, weighted in Float64 and added to Float64 registers.
// Error: the only inexact step is dropping digits beyond 48 bits / pairs beyond the 6th diagonal, a
// zero-mean perturbation of at most 2^-47 max_i max_j per product (max_i = max_k |x_i - mu_i|).  Relative
// to the sigma_i sigma_j N scale of the result the worst case is (max/sigma)^2 2^-46 and the expected
// deviation is ~1/sqrt(N) of that.  gram_tensor_impl measures max/sigma per row and declines (the caller
// falls back to the FP64 DMMA kernel) above 10; Gaussian-like data at N = 1e5 has max/sigma ~ 4.5
// (measured deviation from the DMMA result 1.3e-13 of sigma_i sigma_j); heavy-tailed rows take DMMA.
// Centring happens before the split, so the result is the same centred Gram the DMMA kernel
// (gram.cu) computes, to ~1e-14 relative to sigma_i sigma_j N (tests compare the two).
//
// Kernel organisation (one CTA per 128 x 64 output tile, 256 threads, 1 CTA/SM):
//   all threads : cp.async the 7+7 digit tiles of the next k-block into a 4-stage shared-memory ring
//                 (canonical K-major no-swizzle UMMA layout: 8-row x 16-byte core matrices)
//   thread 0    : issues the 28 tcgen05.mma of a k-block, tcgen05.commit -> mbarrier frees the stage
//   all threads : flush / epilogue through tcgen05.ld (warp w owns TMEM lanes 32*(w%4)..+31)
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_tcgen05.cuh"
#include "mcm_moments.cuh"

namespace mcm {

// Digits per value: 6 by default; 7 or 8 for heavy-tailed rows (the truncation error is (max/sigma)^2 2^(2 - 8 S)
// relative to sigma_i sigma_j N in the worst case: 6 digits hold 1e-12 up to max/sigma = 10, 7 up to 160, 8 up to 2560 —
// and max/sigma <= sqrt(N) always, so 8 digits cover every row up to N = 6.5e6 particles).  More digits = more MMAs
// per k-block (21 / 28 / 36 digit pairs), still far ahead of the FP64 DMMA kernel (C4: 12.7 / ~17 / ~22 ms against 69).
constexpr int OZ_S_MIN = 6, OZ_S_MAX = 8;
constexpr int OZ_TM = 128, OZ_TN = 64, OZ_KB = 32;   // tile rows, tile cols, k per block (== UMMA K for i8)
#ifndef MCM_OZ_STAGES
#define MCM_OZ_STAGES 4
#endif
constexpr int OZ_STAGES = MCM_OZ_STAGES;
constexpr int OZ_THREADS = 256;
constexpr int OZ_A_SLICE = OZ_TM * OZ_KB;            // 4096 B
constexpr int OZ_B_SLICE = OZ_TN * OZ_KB;            // 2048 B
template <int S> constexpr int oz_stage_bytes() { return S * (OZ_A_SLICE + OZ_B_SLICE); }   // 36864 / 43008 / 49152 B
// k-blocks between flushes: int32 headroom S pairs x 128 x 128 x K < 2^31 -> K <= 16384 for S <= 7, 8192 for S = 8
template <int S> constexpr int oz_chunk_kb() { return (S <= 7 ? 16384 : 8192) / OZ_KB; }
constexpr int OZ_DBITS = 8;                          // bits per digit
constexpr double OZ_MAX_RATIO = 10.0;                // max/sigma that 6 digits hold at 1e-12; x16 per extra digit
constexpr int OZ_TMEM_COLS = 512;

// ---- one-pass row statistics ----------------------------------------------------------------------------
// mean, exponent of max |x - mean| and max/sigma of every variable from ONE read (the one-pass moments of
// mcm_moments.cuh: {n, mean, M2, min, max}); replaces a column-mean pass followed by a centred max/sum-of-squares
// pass (0.46 ms each at C4).  max |x - mean| = max(max - mean, mean - min).
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_varstats(const T* const* __restrict__ vars, size_t n, size_t slice, Mom* __restrict__ partial) {
  using P = Pack<T>;
  constexpr int V = P::N;
  const T* x = vars[blockIdx.x];
  const size_t b = (size_t)blockIdx.y * slice;          // slice is a multiple of V and vars are 16-byte aligned
  size_t e = b + slice;
  if (e > n) e = n;
  // shifted power sums folded every 16 values (MomStream) instead of a pairwise merge per pack, four packs in flight
  MomStream acc = ms_empty();
  const size_t nvec = (e - b) / V;
  const P* px = reinterpret_cast<const P*>(x + b);
  size_t i = threadIdx.x;
  for (; i + 3 * (size_t)kThreads < nvec; i += 4 * (size_t)kThreads) {
    P t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) t[u] = ld_stream(px + i + (size_t)u * kThreads);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      double d[V];
#pragma unroll
      for (int v = 0; v < V; ++v) d[v] = (double)t[u].v[v];
      ms_push<V>(acc, d);
    }
  }
  for (; i < nvec; i += kThreads) {
    const P t = ld_stream(px + i);
    double d[V];
#pragma unroll
    for (int v = 0; v < V; ++v) d[v] = (double)t.v[v];
    ms_push<V>(acc, d);
  }
  for (size_t q = b + nvec * V + threadIdx.x; q < e; q += kThreads) {
    const double d[1] = {(double)x[q]};
    ms_push<1>(acc, d);
  }
  Mom m = mom_block_reduce(ms_finish(acc));
  if (threadIdx.x == 0) partial[(size_t)blockIdx.x * gridDim.y + blockIdx.y] = m;
}
__global__ void k_varstats_finish(const Mom* __restrict__ partial, int M, int nsplit, double* __restrict__ means,
                                  int* __restrict__ ex, double* __restrict__ ratio) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  Mom a = mom_empty();
  for (int s = 0; s < nsplit; ++s) a = mom_merge(a, partial[(size_t)j * nsplit + s]);
  const double mu = a.K + a.dm;
  means[j] = mu;
  const double m = fmax(a.mx - mu, mu - a.mn);
  int e = -2000;
  double r = 0.0;
  if (m > 0.0 && m < 1.7e308) {
    const double mant = frexp(m, &e);
    if (mant > 0.99) ++e;                 // the max is formed from rounded differences: a little more headroom than k_rowstats
    r = m / sqrt(a.m2 / a.n);
  } else if (!(m == 0.0)) {
    e = 2000; r = 1e300;                  // Inf / NaN in the row: force the fallback
  }
  ex[j] = e;
  ratio[j] = r;
}
template <typename T>
int varstats_device(const T* const* d_vars, int M, size_t n, double* d_means, int* d_ex, double* d_ratio) {
  Ctx& c = ctx();
  constexpr size_t V = Pack<T>::N;
  int nsplit = (4 * c.sm_count + M - 1) / M;
  const size_t max_split = (n + 65535) / 65536;
  if ((size_t)nsplit > max_split) nsplit = (int)max_split;
  if (nsplit < 1) nsplit = 1;
  size_t slice = (n + nsplit - 1) / nsplit;
  slice = (slice + V - 1) / V * V;
  nsplit = (int)((n + slice - 1) / slice);
  Mom* partial = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&partial, (size_t)M * nsplit * sizeof(Mom), c.stream));
  k_varstats<T><<<dim3(M, nsplit), kThreads, 0, c.stream>>>(d_vars, n, slice, partial);
  MCM_CHECK_LAUNCH();
  k_varstats_finish<<<(M + 127) / 128, 128, 0, c.stream>>>(partial, M, nsplit, d_means, d_ex, d_ratio);
  MCM_CHECK_LAUNCH();
  MCM_CUDA(cudaFreeAsync(partial, c.stream));
  return MCM_OK;
}
template int varstats_device<double>(const double* const*, int, size_t, double*, int*, double*);
template int varstats_device<float>(const float* const*, int, size_t, double*, int*, double*);

// digits in TILE-READY order: slices[s][kb][rg][c][r8][16] (kb = k / 32, rg = row / 8, c = (k % 32) / 16,
// r8 = row % 8): the 8-row x 16-byte core matrices of the canonical K-major no-swizzle UMMA layout are
// contiguous, so the 128 (or 64) rows of one k-block of one digit plane are ONE contiguous 4096 (2048) byte
// run that a single cp.async.bulk drops into shared memory already in MMA layout.
// One thread = 16 consecutive k of one row (reads 128 contiguous bytes of Float64, writes 16 bytes per digit).
template <typename T, int OZ_S>
__global__ void __launch_bounds__(256)
k_slice(const T* const* __restrict__ vars, const double* __restrict__ means, const int* __restrict__ ex, int M,
        size_t n, size_t kpad, size_t mpad, int8_t* __restrict__ slices) {
  // work item w -> (kb, rg, c, r8) with r8 fastest: the 16 lanes of a half-warp write one 256-byte core block
  const size_t nkb = kpad / OZ_KB, nrg = mpad / 8;
  const size_t total = nkb * nrg * 16;
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t w = (size_t)blockIdx.x * 256 + threadIdx.x; w < total; w += stride) {
    // work order: within a group of OZ_KBT k-blocks the k-block index runs fastest, then the row group — a CTA's 16 blocks
    // are 16 consecutive k-blocks of the same 8 rows, i.e. 4 KiB CONTIGUOUS per row (the rows are separate arrays, ~1 MB
    // apart: with the row group running fastest every row contributed 256 bytes per CTA and the reads hopped DRAM pages)
    const size_t r8 = w & 7, c = (w >> 3) & 1, blk = w >> 4;
    constexpr size_t OZ_KBT = 16;
    const size_t per_group = nrg * OZ_KBT;
    const size_t kbt = blk / per_group, rem = blk % per_group;
    size_t rg, kb;
    if ((kbt + 1) * OZ_KBT <= nkb) { rg = rem / OZ_KBT; kb = kbt * OZ_KBT + rem % OZ_KBT; }
    else { const size_t left = nkb - kbt * OZ_KBT; rg = rem / left; kb = kbt * OZ_KBT + rem % left; }   // ragged last group
    const size_t row = rg * 8 + r8;
    int4 pl[OZ_S];
    if (row < (size_t)M) {
      const T* x = vars[row];
      const double mu = means[row];
      const int e = ex[row];
      // I = rint(y * 2^48), |I| < 2^47 (|y| < 127.4/256); balanced base-256 digits from the low end:
      // every digit lands in [-128, 127] and the carries make sum_s d_s 256^(5-s) == I exactly.
      const double sc = (e > -1500 && e < 1500) ? scalbn(1.0, -(e + 1) + OZ_DBITS * OZ_S) : 0.0;
      const size_t k0 = kb * OZ_KB + c * 16;
      // 16 consecutive particles of this row: 128-bit loads when the whole run is inside the data (the lanes of a
      // warp read 32 different 128-byte runs, so wider requests halve the LSU wavefronts per byte)
      T xv[16];
      if (k0 + 16 <= n && (reinterpret_cast<uintptr_t>(x + k0) & 15u) == 0) {
        using P = Pack<T>;
        constexpr int V = P::N;
#pragma unroll
        for (int q = 0; q < 16 / V; ++q) {
          const P t = *reinterpret_cast<const P*>(x + k0 + q * V);
#pragma unroll
          for (int v = 0; v < V; ++v) xv[q * V + v] = t.v[v];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = (k0 + j < n) ? x[k0 + j] : (T)mu;
      }
      unsigned long long Kd[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const size_t k = k0 + j;
        const long long I = (k < n) ? __double2ll_rn(((double)xv[j] - mu) * sc) : 0ll;
        Kd[j] = balanced_digit_bytes<OZ_S>(I);          // all digits in one addition (mcm_tcgen05.cuh)
      }
      pack_digit_planes<OZ_S>(Kd, pl);
    } else {
#pragma unroll
      for (int s = 0; s < OZ_S; ++s) pl[s] = make_int4(0, 0, 0, 0);
    }
    const size_t off = (((kb * nrg + rg) * 2 + c) * 8 + r8) * 16;
#pragma unroll
    for (int s = 0; s < OZ_S; ++s)
      *reinterpret_cast<int4*>(slices + (size_t)s * (mpad * kpad) + off) = pl[s];
  }
}

// tile t -> (bi over 128-row blocks, bj over 64-col blocks) with bj >= 2*bi (upper part incl. the diagonal)
__device__ __forceinline__ void oz_tile_decode(int t, int mb, int nbk, int& bi, int& bj) {
  int i = 0, rem = t;
  while (rem >= nbk - 2 * i) { rem -= nbk - 2 * i; ++i; }
  bi = i; bj = 2 * i + rem;
}

// Warp-specialised: warp 0 lane 0 = TMA producer (14 bulk copies per k-block), warp 1 lane 0 = MMA issuer
// (28 tcgen05.mma per k-block), all 8 warps = TMEM flush + epilogue.  K range [kb_begin, kb_end) of this CTA
// (split-K when there are fewer tiles than SMs); `out` is C (ksplit == 1) or the split workspace.
template <int OZ_S>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1)
k_gram_i8(const int8_t* __restrict__ slices, const int* __restrict__ ex, int M, size_t mpad, size_t kpad,
          int ksplit, int kb_per_split, double scale, double* __restrict__ C, double* __restrict__ ws, int a_in_tmem) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_full[OZ_STAGES];
  __shared__ __align__(8) unsigned long long bar_free[OZ_STAGES];
  __shared__ __align__(8) unsigned long long bar_acc;
  __shared__ uint32_t tmem_base_sh;

  constexpr int OZ_STAGE_BYTES = oz_stage_bytes<OZ_S>();
  constexpr int OZ_CHUNK_KB = oz_chunk_kb<OZ_S>();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = (int)(mpad / OZ_TM), nbk = (int)(mpad / OZ_TN);
  int bi, bj;
  oz_tile_decode(blockIdx.x, mb, nbk, bi, bj);
  const size_t plane = mpad * kpad;                 // bytes per digit plane
  const size_t nrg = mpad / 8;
  const int nkb_total = (int)(kpad / OZ_KB);
  const int kb_begin = blockIdx.y * kb_per_split;
  const int kb_end = min(nkb_total, kb_begin + kb_per_split);
  const int nkb = max(0, kb_end - kb_begin);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_sh)), "r"(OZ_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
#pragma unroll
    // bar_free counts the MMA commits of BOTH CTAs of the pair: the A planes of a stage are multicast into both
#ifdef MCM_EXP_NO_MC
    for (int s = 0; s < OZ_STAGES; ++s) { mbar_init_tc(&bar_full[s], 1); mbar_init_tc(&bar_free[s], 1); }
#else
    for (int s = 0; s < OZ_STAGES; ++s) { mbar_init_tc(&bar_full[s], 1); mbar_init_tc(&bar_free[s], 2); }
#endif
    mbar_init_tc(&bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before any multicast copy or commit targets them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t crank = cluster_ctarank();
  const uint32_t tmem_base = tmem_base_sh;
  const uint32_t idesc = umma_idesc_i8(OZ_TM, OZ_TN);
  const uint32_t smem0 = smem_addr(smem);

  double facc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) facc[j] = 0.0;

  int chunk = 0;
  for (int c0 = 0; c0 < nkb; c0 += OZ_CHUNK_KB, ++chunk) {
    const int c1 = min(nkb, c0 + OZ_CHUNK_KB);
    if (warp == 0) {
      // ---- producer: digit planes of k-block kb -> stage (it % STAGES); `it` counts k-blocks of this CTA.
      // One elected thread owns the loop; the next stage's free-barrier is tested while this stage's copies are issued.
      if (elect_one()) {
        const int8_t* srcA0 = slices + ((size_t)bi * (OZ_TM / 8)) * 256;
        const int8_t* srcB0 = slices + ((size_t)bj * (OZ_TN / 8)) * 256;
        const size_t kb_stride = nrg * 256;                 // bytes between consecutive k-blocks of one plane
        bool ready = true;
        if (c0 >= OZ_STAGES) ready = mbar_try_tc(&bar_free[c0 % OZ_STAGES], (unsigned)((c0 / OZ_STAGES - 1) & 1));
        for (int it = c0; it < c1; ++it) {
          const int st = it % OZ_STAGES;
          if (!ready) mbar_wait_tc(&bar_free[st], (unsigned)((it / OZ_STAGES - 1) & 1));
          // test the NEXT stage now; the answer is needed only after this stage's copies have been issued
          ready = (it + 1 >= OZ_STAGES && it + 1 < c1)
                      ? mbar_try_tc(&bar_free[(it + 1) % OZ_STAGES], (unsigned)(((it + 1) / OZ_STAGES - 1) & 1))
                      : true;
#ifdef MCM_EXP_HALF_B
          mbar_expect_tx_tc(&bar_full[st], OZ_STAGE_BYTES - 3 * OZ_B_SLICE);   // EXPERIMENT: timing sensitivity to operand bytes
#else
          mbar_expect_tx_tc(&bar_full[st], OZ_STAGE_BYTES);
#endif
          const uint32_t sbase = smem0 + (uint32_t)st * OZ_STAGE_BYTES;
#ifdef MCM_EXP_SAME_KB
          const size_t kb = (size_t)kb_begin;   // EXPERIMENT: every k-block re-reads the same (L2-hot) planes
#else
          const size_t kb = (size_t)(kb_begin + it);
#endif
          const int8_t* srcA = srcA0 + kb * kb_stride;
          const int8_t* srcB = srcB0 + kb * kb_stride;
          // the pair (tiles 2p, 2p+1) has the same row block bi: each CTA fetches every other A plane and
          // multicasts it to both, so an A plane crosses the L2 -> SM fabric once per pair
#pragma unroll
          for (int s = 0; s < OZ_S; ++s) {
#ifdef MCM_EXP_NO_MC
            bulk_g2s_tc(sbase + s * OZ_A_SLICE, srcA + (size_t)s * plane, OZ_A_SLICE, &bar_full[st]);
#else
            if ((uint32_t)(s & 1) == crank)
              bulk_g2s_mc(sbase + s * OZ_A_SLICE, srcA + (size_t)s * plane, OZ_A_SLICE, &bar_full[st], (uint16_t)0x3);
#endif
#ifdef MCM_EXP_HALF_B
            if (s < 3)
#endif
            bulk_g2s_tc(sbase + OZ_S * OZ_A_SLICE + s * OZ_B_SLICE, srcB + (size_t)s * plane, OZ_B_SLICE, &bar_full[st]);
          }
        }
      }
    } else if (warp == 1) {
      // ---- MMA issuer.  tcgen05.mma issue is nearly synchronous with execution (the queue holds only a couple of
      // these 128-cycle MMAs), so every cycle the issuing thread spends between k-blocks is a cycle the tensor pipe
      // idles.  Measured on an isolated replica of this pipeline (scripts/micro/pipe.cu): wait-then-issue per k-block
      // = 1009 cycles, against 696 for the MMAs alone.  Hence: ONE elected thread owns the loop (no per-k-block
      // elect / __syncwarp), the NEXT stage's barrier is TESTED before the current k-block's MMAs are issued (its
      // latency hides behind the blocking issue), no tcgen05.fence inside the loop (TMA -> MMA needs none), and the
      // shared-memory descriptors advance by a constant per stage: 740 cycles per k-block on the replica.
      if (elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");   // TMEM was read by the flush before this chunk
        const uint64_t dA00 = umma_desc_kmajor(smem0, 128, 256);
        const uint64_t dB00 = umma_desc_kmajor(smem0 + OZ_S * OZ_A_SLICE, 128, 256);
        if (c0 < c1) mbar_wait_tc(&bar_full[c0 % OZ_STAGES], (unsigned)((c0 / OZ_STAGES) & 1));
        for (int it = c0; it < c1; ++it) {
          const int st = it % OZ_STAGES;
          const int nst = (it + 1) % OZ_STAGES;
          const unsigned npar = (unsigned)(((it + 1) / OZ_STAGES) & 1);
          const bool nready = (it + 1 < c1) ? mbar_try_tc(&bar_full[nst], npar) : true;   // consumed after the MMAs
          const uint64_t dA0 = dA00 + (uint64_t)(((uint32_t)st * OZ_STAGE_BYTES) >> 4);
          const uint64_t dB0 = dB00 + (uint64_t)(((uint32_t)st * OZ_STAGE_BYTES) >> 4);
          // pair (s, t) feeds accumulator d = s + t; the first pair of every accumulator is (0, d), so only the
          // s == 0 row of the first k-block of a chunk overwrites
          const uint32_t acc0 = (it == c0) ? 0u : 1u;
          // WIDE-N issue: the accumulators of diagonals d, d+1, ... are adjacent TMEM column blocks and the B planes
          // t, t+1, ... are adjacent 64-row blocks of one K-major tile, so for a fixed A plane s ONE MMA with
          // N = 64 (6 - s) (cut at 256) updates all its diagonals: 8 MMAs per k-block instead of 21, and the
          // 4 KiB A plane is read from shared memory once per MMA instead of once per diagonal
          // (microbenchmark: M128 N64 SS costs 48 cycles, smem-bound; N >= 128 runs at the 32-cycles-per-64-columns floor)
          uint32_t abuf = 0;
          if (a_in_tmem) {
            abuf = tmem_base + (uint32_t)(OZ_S * OZ_TN) + (uint32_t)(it & 1) * (OZ_S * 8);
#pragma unroll
            for (int s = 0; s < OZ_S; ++s) tmem_cp_128x256b(abuf + s * 8, dA0 + (uint64_t)((s * OZ_A_SLICE) >> 4));
          }
          // issue order: plane 0 first (it alone may overwrite), the LONGEST MMAs last, so that the tensor queue still
          // holds ~250 cycles of work while this thread commits, loops and builds the next descriptors
          auto issue = [&](int s, int c0n, int nn) {
            const uint32_t idn = umma_idesc_i8(OZ_TM, nn);
            const uint32_t dcol = tmem_base + (uint32_t)(s * OZ_TN + c0n);
            const uint64_t db = dB0 + (uint64_t)(((c0n / OZ_TN) * OZ_B_SLICE) >> 4);   // B planes start at t = c0n/64
            if (a_in_tmem) umma_i8_ts(dcol, abuf + s * 8, db, idn, s == 0 ? acc0 : 1u);
            else umma_i8(dcol, dA0 + (uint64_t)((s * OZ_A_SLICE) >> 4), db, idn, s == 0 ? acc0 : 1u);
          };
          static_assert(OZ_TN == 64, "accumulators are 64 columns wide");
          if constexpr (OZ_S == 6) {
            issue(0, 0, 256); issue(0, 256, 128);     // diagonals 0..3, 4..5
            issue(5, 0, 64);                          // 5
            issue(4, 0, 128);                         // 4..5
            issue(3, 0, 192);                         // 3..5
            issue(1, 256, 64);                        // plane 1 x plane 4 -> diagonal 5
            issue(2, 0, 256);                         // 2..5
            issue(1, 0, 256);                         // 1..4
          } else {
            // plane s meets B planes 0 .. S-1-s: 64 (S - s) accumulator columns starting at diagonal s, cut at 256 per
            // MMA; plane 0 first (it alone may overwrite), then the short ones, the long ones last
#pragma unroll
            for (int c0n = 0; c0n < 64 * OZ_S; c0n += 256) issue(0, c0n, (64 * OZ_S - c0n) < 256 ? (64 * OZ_S - c0n) : 256);
#pragma unroll
            for (int sp = OZ_S - 1; sp >= 1; --sp) {
#pragma unroll
              for (int c0n = (64 * (OZ_S - sp) - 1) / 256 * 256; c0n >= 0; c0n -= 256)
                issue(sp, c0n, (64 * (OZ_S - sp) - c0n) < 256 ? (64 * (OZ_S - sp) - c0n) : 256);
            }
          }
#ifdef MCM_EXP_NO_MC
          umma_commit(&bar_free[st]);
#else
          umma_commit_mc(&bar_free[st], (uint16_t)0x3);   // stage reusable in BOTH CTAs once these MMAs have read it
#endif
          if (it == c1 - 1) umma_commit(&bar_acc);        // chunk accumulated
          if (!nready) mbar_wait_tc(&bar_full[nst], npar);
        }
      }
    }
    __syncwarp();
    // ---- flush: int32 accumulators -> Float64 registers (all warps)
    mbar_wait_backoff(&bar_acc, (unsigned)(chunk & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int d = 0; d < OZ_S; ++d) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(d * OZ_TN + (warp >> 2) * 32);
      tmem_ld32(taddr, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const double w = scalbn(1.0, -OZ_DBITS * (d + 2));
#pragma unroll
      for (int j = 0; j < 32; ++j) facc[j] = fma((double)(int)v[j], w, facc[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // TMEM drained before the next chunk overwrites it
  }

  // ---- epilogue: TMEM lane = tile row, 32 columns per thread
  {
    const int r = bi * OZ_TM + (warp & 3) * 32 + lane;
    const int cbase = bj * OZ_TN + (warp >> 2) * 32;
    if (ksplit == 1) {
      if (r < M) {
        const int er = ex[r];
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const int c = cbase + j;
          if (c < M) {
            const int ec = ex[c];
            double v = 0.0;
            if (er > -1500 && er < 1500 && ec > -1500 && ec < 1500) v = scalbn(facc[j], er + ec + 2) * scale;
            C[(size_t)r * M + c] = v;
            C[(size_t)c * M + r] = v;
          }
        }
      }
    } else {
      // raw partial sums (still in scaled units) for the fixed-order split-K reduction
      double* w = ws + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(OZ_TM * OZ_TN) +
                  (size_t)((warp & 3) * 32 + lane) * OZ_TN + (warp >> 2) * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 2) *reinterpret_cast<double2*>(w + j) = make_double2(facc[j], facc[j + 1]);
    }
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(OZ_TMEM_COLS));
  cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
}

__global__ void __launch_bounds__(256)
k_gram_i8_reduce(const double* __restrict__ ws, const int* __restrict__ ex, int M, int mb, int nbk, int ntiles,
                 int ksplit, double scale, double* __restrict__ C) {
  int bi, bj;
  oz_tile_decode(blockIdx.x, mb, nbk, bi, bj);
  for (int e = threadIdx.x; e < OZ_TM * OZ_TN; e += 256) {
    double sum = 0.0;
    for (int sp = 0; sp < ksplit; ++sp) sum += ws[((size_t)sp * ntiles + blockIdx.x) * (size_t)(OZ_TM * OZ_TN) + e];
    const int r = bi * OZ_TM + e / OZ_TN, c = bj * OZ_TN + e % OZ_TN;
    if (r < M && c < M) {
      const int er = ex[r], ec = ex[c];
      double v = 0.0;
      if (er > -1500 && er < 1500 && ec > -1500 && ec < 1500) v = scalbn(sum, er + ec + 2) * scale;
      C[(size_t)r * M + c] = v;
      C[(size_t)c * M + r] = v;
    }
  }
}

int g_a_in_tmem = 0;   // 0 (default): A planes read from shared memory by wide-N MMAs; mcm_gram_use_tensor(2): staged in TMEM (A/B runs)

int g_oz_last_digits = 0;   // digits per value of the last tensor-core Gram (introspection for tests / bench)

// slicing + MMA kernel + split-K reduction for S digits per value; ex = per-row exponents (device, mpad entries)
template <typename T, int OZ_S>
static int gram_tensor_run(const T* const* d_vars, const double* d_means, int M, size_t n, double scale, double* C, const int* ex) {
  Ctx& c = ctx();
  constexpr int OZ_STAGE_BYTES = oz_stage_bytes<OZ_S>();
  const size_t mpad = ((size_t)M + OZ_TM - 1) / OZ_TM * OZ_TM;
  const size_t kpad = (n + OZ_KB - 1) / OZ_KB * OZ_KB;
  int8_t* slices = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&slices, (size_t)OZ_S * mpad * kpad, c.stream));
  {
    const size_t total = (kpad / OZ_KB) * (mpad / 8) * 16;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)c.sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_slice<T, OZ_S><<<(unsigned)blocks, 256, 0, c.stream>>>(d_vars, d_means, ex, M, n, kpad, mpad, slices);
    MCM_CHECK_LAUNCH();
  }
  const int mb = (int)(mpad / OZ_TM), nbk = (int)(mpad / OZ_TN);
  int ntiles = 0;
  for (int i = 0; i < mb; ++i) ntiles += nbk - 2 * i;
  constexpr size_t kSmem = (size_t)OZ_STAGES * OZ_STAGE_BYTES;
  static bool attr = false;
  if (!attr) {
    MCM_CUDA(cudaFuncSetAttribute(k_gram_i8<OZ_S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    attr = true;
  }
  // split-K when the tiles alone cannot fill the GPU
  const int nkb_total = (int)(kpad / OZ_KB);
  int ksplit = 1;
  const int max_split = std::max(1, nkb_total / 256);   // keep >= 8192 particles per split
  if (ntiles < c.sm_count) {
    ksplit = (2 * c.sm_count + ntiles - 1) / ntiles;
    if (ksplit > max_split) ksplit = max_split;
  } else {
    // wave quantisation: 1056 tiles on 148 SMs are 7.1 waves = 8 rounds (89 % efficient); cutting K in 3 gives
    // 21.4 -> 22 rounds of a third the length (97 %).  Pick the split (<= 4) with the best efficiency, if it
    // gains more than the ~1 % the extra workspace pass costs.
    double best = 0.0;
    for (int k = 1; k <= std::min(4, max_split); ++k) {
      const double waves = (double)ntiles * k / c.sm_count;
      const double eff = waves / ceil(waves) - (k > 1 ? 0.01 : 0.0);
      if (eff > best + 1e-9) { best = eff; ksplit = k; }
    }
  }
  int kb_per_split = (nkb_total + ksplit - 1) / ksplit;
  ksplit = (nkb_total + kb_per_split - 1) / kb_per_split;
  double* ws = nullptr;
  if (ksplit > 1) MCM_CUDA(cudaMallocAsync((void**)&ws, (size_t)ksplit * ntiles * OZ_TM * OZ_TN * sizeof(double), c.stream));
  k_gram_i8<OZ_S><<<dim3(ntiles, ksplit), OZ_THREADS, kSmem, c.stream>>>(slices, ex, M, mpad, kpad, ksplit, kb_per_split, scale, C, ws,
                                                                        OZ_S == 6 ? g_a_in_tmem : 0);
  MCM_CHECK_LAUNCH();
  if (ksplit > 1) {
    k_gram_i8_reduce<<<ntiles, 256, 0, c.stream>>>(ws, ex, M, mb, nbk, ntiles, ksplit, scale, C);
    MCM_CHECK_LAUNCH();
    MCM_CUDA(cudaFreeAsync(ws, c.stream));
  }
  MCM_CUDA(cudaFreeAsync(slices, c.stream));
  g_oz_last_digits = OZ_S;
  return MCM_OK;
}

// pre_ex / pre_ratio: row statistics already on device (varstats_device), or nullptr to compute them here about
// the given means (k_rowstats: the centre was supplied by the caller, e.g. the global means of a sharded pcov)
template <typename T>
int gram_tensor_impl(const T* const* d_vars, const double* d_means, int M, size_t n, double scale, double* C,
                     const int* pre_ex, const double* pre_ratio) {
  Ctx& c = ctx();
  const size_t mpad = ((size_t)M + OZ_TM - 1) / OZ_TM * OZ_TM;
  int* ex = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&ex, mpad * sizeof(int), c.stream));
  if (mpad > (size_t)M) MCM_CUDA(cudaMemsetAsync(ex + M, 0, (mpad - M) * sizeof(int), c.stream));
  double* d_ratio = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_ratio, (size_t)M * sizeof(double), c.stream));
  if (pre_ex) {
    MCM_CUDA(cudaMemcpyAsync(ex, pre_ex, (size_t)M * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
    MCM_CUDA(cudaMemcpyAsync(d_ratio, pre_ratio, (size_t)M * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
  } else {
    k_rowstats<T><<<M, 256, 0, c.stream>>>(d_vars, d_means, n, ex, d_ratio);
    MCM_CHECK_LAUNCH();
  }
  // accuracy guard: (max/sigma)^2 2^(2 - 8 S) must stay below ~1e-12 -> the digit count follows the heaviest row
  std::vector<double> h_ratio(M);
  MCM_CUDA(cudaMemcpyAsync(h_ratio.data(), d_ratio, (size_t)M * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d_ratio, c.stream));
  double worst = 0.0;
  for (int i = 0; i < M; ++i) worst = std::max(worst, h_ratio[i]);
  int rc;
  if (worst <= OZ_MAX_RATIO) rc = gram_tensor_run<T, 6>(d_vars, d_means, M, n, scale, C, ex);
  else if (worst <= 16.0 * OZ_MAX_RATIO) rc = gram_tensor_run<T, 7>(d_vars, d_means, M, n, scale, C, ex);
  else if (worst <= 256.0 * OZ_MAX_RATIO) rc = gram_tensor_run<T, 8>(d_vars, d_means, M, n, scale, C, ex);
  else rc = -1;      // Inf / NaN rows or max/sigma > 2560 (needs N > 6.5e6 and one dominant outlier): the caller runs the DMMA kernel
  MCM_CUDA(cudaFreeAsync(ex, c.stream));
  return rc;
}

template int gram_tensor_impl<double>(const double* const*, const double*, int, size_t, double, double*, const int*, const double*);
template int gram_tensor_impl<float>(const float* const*, const double*, int, size_t, double, double*, const int*, const double*);

}  // namespace mcm
