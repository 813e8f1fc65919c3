// pgemm_tc.cu — constant matrix x vector of uncertain variables on tcgen05 + TMEM (Float64 in, Float64 out).
//
//   y_j[i] = bias_j + sum_k A[j,k] (x_k[i] - shift_k)        (sigmapoints.jl:71, particles.jl:688-707)
//
// Same Ozaki-style error-free split as the Gram kernel (gram_tc.cu), with 7 digits because the result is a
// per-particle value, not an average:
//   x'_k[i]  = (x_k[i] - shift_k) 2^-(e_k+1)        |x'| < 1/2,  e_k from max_i |x_k[i] - shift_k|  (k_rowstats)
//   a''_jk   = A[j,k] 2^(e_k+1) 2^-(f_j+1)          |a''| < 1/2, f_j from max_k |A[j,k]| 2^(e_k+1)  (k_pz_rows)
//   both cut into 7 balanced base-256 digits (56 bits); digit-plane products are exact int32 dot products on
//   tcgen05.mma kind::i8; the 28 plane pairs with s + t <= 6 accumulate into 7 TMEM accumulators (one per
//   weight 256^-(s+t+2)); y = bias + 2^(f_j+1) sum_d 256^-(d+2) acc_d.
// Dropped terms are below 2^-55 |x'|max |a''|max per product: the deviation from the FP64 DMMA kernel
// (pgemm.cu) is ~1e-15 of the row scale sum_k |A[j,k]| max|x_k - shift_k| (tests compare both with the
// long-double oracle).  Inputs whose max/sigma exceeds 40 for some variable, or K > 16384, are declined and run
// on the DMMA kernel.
//
// Tile: 128 particles (TMEM lanes, so the epilogue writes 256 contiguous bytes of one output per warp store) x
// 64 outputs; k-blocks of 32 variables.  PERSISTENT kernel, one CTA per SM: four CTAs share a particle tile (rotating parts) and
// split its output tiles, so the 148 CTAs keep 37 digit tiles (7 planes x 128 x K bytes) hot in L2 while they
// sweep all output tiles; the TMA producer runs ahead across tile boundaries.
//   warp 0 lane 0 : producer, 14 cp.async.bulk per k-block (digit planes are stored tile-ready) -> 4-stage ring
//   warp 1 lane 0 : 28 tcgen05.mma (M128 N64 K32) per k-block, tcgen05.commit -> mbarriers
//   warps 2..5    : tcgen05.ld of the 7 accumulators, exact int64 recombination, 3 conversions, then release
//                   TMEM (the next tile's MMAs start) and store 256 contiguous bytes per warp per output
// Output tiles whose rows of A end early (triangular factors) stop their k loop early.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "mcm_common.cuh"
#include "mcm_tcgen05.cuh"

namespace mcm {

constexpr int PZ_S = 7;                               // digits per value
constexpr int PZ_TM = 128, PZ_TN = 64, PZ_KB = 32;    // particles per tile, outputs per tile, variables per k-block
// MCM_PZ_BULK_STORE (experiment): the epilogue stages the tile's results in shared memory and hands them to the bulk-copy
// engine (cp.async.bulk shared -> global, 1 KiB per output column) instead of 32 global stores per thread; the staging
// tile (128 x 64 results) takes the place of the fourth operand stage.
// Measured at C5 (1024 x 1024 triangular / dense, 1e6 particles): 13.09 / 19.5 ms against 12.59 / 19.0 ms for direct
// stores with a 4-stage ring — the stores are not what limits the kernel (the per-tile overhead is ~9 k cycles whether
// the k loop is short or long: accumulator drain + read-back, serial with the MMAs), so it is compiled out.
#ifndef MCM_PZ_BULK_STORE
#define MCM_PZ_BULK_STORE 0
#endif
// MCM_PZ_BATCH_LD (experiment): the three tcgen05.ld of a diagonal group issued together, one wait per group instead of
// one per load.  Measured equal (12.63 / 19.2 ms against 12.65 / 19.2 ms): the read-back is not latency-bound on the waits.
// What the measurements leave as the bound of the MMA phase is the SM's shared-memory bandwidth: per k-block the MMAs
// read 10 x 4 KiB of particle planes + 56 KiB of coefficient planes and the bulk copies write 42 KiB — ~140 KiB against
// 128 B/clk = ~1.1 k cycles, above the 896 cycles the tensor pipe needs (traced: ~970 per k-block).  Halving the
// coefficient-side traffic needs cta_group::2 (each CTA of a pair holds half of the N operand).
#ifndef MCM_PZ_BATCH_LD
#define MCM_PZ_BATCH_LD 0
#endif
#ifndef MCM_PZ_STAGES
#define MCM_PZ_STAGES (MCM_PZ_BULK_STORE ? 3 : 4)
#endif
constexpr int PZ_STAGES = MCM_PZ_STAGES;
constexpr int PZ_THREADS = 320;     // producer warp, MMA warp, 8 epilogue warps
constexpr int PZ_X_SLICE = PZ_TM * PZ_KB;             // 4096 B
constexpr int PZ_A_SLICE = PZ_TN * PZ_KB;             // 2048 B
constexpr int PZ_STAGE_BYTES = PZ_S * (PZ_X_SLICE + PZ_A_SLICE);   // 43008 B
constexpr int PZ_TMEM_COLS = 512;
constexpr int PZ_MAX_K = 16384;                       // int32 headroom: 7 pairs x 128 x 128 x K < 2^31
constexpr double PZ_MAX_RATIO = 40.0;

// digits of the particle operand in TILE-READY order: planes[s][kb][rg][c][r8][16]  (kb = k / 32, rg = particle / 8,
// c = (k % 32) / 16, r8 = particle % 8) — the canonical K-major no-swizzle UMMA layout, so the 128 particles of one
// k-block of one plane are ONE contiguous 4096-byte run.  One thread = 16 consecutive variables of one particle:
// it reads 16 arrays at the same particle index (8 consecutive particles per 8 lanes = 64-byte segments) and
// writes 16 bytes per plane (fully coalesced).
template <typename T>
__global__ void __launch_bounds__(256)
k_pslice(const T* const* __restrict__ xs, const double* __restrict__ shift, const double* __restrict__ scale56, int K,
         size_t n, size_t npad, size_t kpad, int8_t* __restrict__ planes) {
  const size_t nrg = npad / 8;
  const size_t total = (kpad / PZ_KB) * nrg * 16;
  // a CTA takes PS_RUN consecutive chunks of 256 work items: 512 consecutive particles of the same 32 variables, i.e.
  // 4 KiB contiguous per variable array (one chunk alone is a 1 KiB run per array: the reads hopped DRAM pages)
  constexpr size_t PS_RUN = 4;
  const size_t nchunks = (total + 255) / 256;
  for (size_t ch = (size_t)blockIdx.x * PS_RUN; ch < nchunks; ch += (size_t)gridDim.x * PS_RUN)
  for (size_t ci = 0; ci < PS_RUN; ++ci) {
    const size_t w = (ch + ci) * 256 + threadIdx.x;
    if (w >= total) break;
    const size_t r8 = w & 7, c = (w >> 3) & 1, blk = w >> 4;
    const size_t rg = blk % nrg, kb = blk / nrg;
    const size_t row = rg * 8 + r8;
    const int k0 = (int)(kb * PZ_KB + c * 16);
    double xv[16];                   // all 16 loads issued before the first digit is extracted
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = k0 + j;
      xv[j] = (k < K && row < n) ? ((double)__ldcs(xs[k] + row) - shift[k]) * scale56[k] : 0.0;
    }
    unsigned long long Kd[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) Kd[j] = balanced_digit_bytes<PZ_S>(__double2ll_rn(xv[j]));
    int4 pl[PZ_S];
    pack_digit_planes<PZ_S>(Kd, pl);
    const size_t off = w * 16;
#pragma unroll
    for (int s = 0; s < PZ_S; ++s) *reinterpret_cast<int4*>(planes + (size_t)s * (npad * kpad) + off) = pl[s];
  }
}

// ---- coefficient operand (device side) -------------------------------------------------------
// per-variable powers: sc56[k] = 2^(-(e_k+1)+56) for the particle digits, apow[k] = 2^(e_k+1) folded into A
__global__ void k_pz_scales(const int* __restrict__ ex, int K, size_t kpad, double* __restrict__ sc56, double* __restrict__ apow) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kpad) return;
  const bool live = k < (size_t)K && ex[k] != -2000;
  sc56[k] = live ? scalbn(1.0, -(ex[k] + 1) + 8 * PZ_S) : 0.0;
  apow[k] = live ? scalbn(1.0, ex[k] + 1) : 0.0;
}
// one CTA per output row j: f_j from max_k |A[j,k]| apow[k]; rowscale = 2^(f_j+1), cscale = 2^(-(f_j+1)+56);
// kend[j/64] = columns needed by the output tile (multiple of 32); flag bit 0 = out-of-range scale
__global__ void __launch_bounds__(256)
k_pz_rows(const double* __restrict__ A, int Mi, const double* __restrict__ apow, double* __restrict__ rowscale,
          double* __restrict__ cscale, int* __restrict__ kend, uint32_t* __restrict__ flag) {
  const int j = blockIdx.x;
  double amax = 0.0;
  int last = 0;
  for (int k = threadIdx.x; k < Mi; k += 256) {
    const double v = fabs(A[(size_t)j * Mi + k] * apow[k]);
    if (v != 0.0) last = k + 1;
    amax = fmax(amax, v);    // NaN-free maximum is enough: a NaN product is caught through amax below
    if (!(v <= 1.7e308)) amax = INFINITY;
  }
  __shared__ double shm[8];
  __shared__ int shl[8];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    amax = fmax(amax, __shfl_down_sync(0xffffffffu, amax, d));
    last = max(last, __shfl_down_sync(0xffffffffu, last, d));
  }
  if ((threadIdx.x & 31) == 0) { shm[threadIdx.x >> 5] = amax; shl[threadIdx.x >> 5] = last; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { amax = fmax(amax, shm[w]); last = max(last, shl[w]); }
    double rs = 0.0, cs = 0.0;
    if (amax > 0.0) {
      int f;
      const double mant = frexp(amax, &f);
      if (mant > 0.995) ++f;
      if (!(amax <= 1.7e308) || f < -900 || f > 900) atomicOr(flag, 1u);
      else { rs = scalbn(1.0, f + 1); cs = scalbn(1.0, -(f + 1) + 8 * PZ_S); }
    }
    rowscale[j] = rs;
    cscale[j] = cs;
    atomicMax(&kend[j / PZ_TN], (last + PZ_KB - 1) / PZ_KB * PZ_KB);
  }
}
// digits of a''_jk in tile-ready order aplanes[t][kb][jg][c][r8][16] (same layout as the particle operand)
__global__ void __launch_bounds__(256)
k_pz_coef(const double* __restrict__ A, int Mo, int Mi, const double* __restrict__ apow, const double* __restrict__ cscale,
          size_t mopad, size_t kpad, int8_t* __restrict__ planes) {
  const size_t njg = mopad / 8;
  const size_t total = (kpad / PZ_KB) * njg * 16;
  const size_t w = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (w >= total) return;
  const size_t r8 = w & 7, c = (w >> 3) & 1, blk = w >> 4;
  const size_t jg = blk % njg, kb = blk / njg;
  const size_t j = jg * 8 + r8;
  const int k0 = (int)(kb * PZ_KB + c * 16);
  const double cs = j < (size_t)Mo ? cscale[j] : 0.0;
  unsigned long long Kd[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int k = k0 + q;
    long long I = 0;
    if (k < Mi && cs != 0.0) I = __double2ll_rn(A[j * (size_t)Mi + k] * apow[k] * cs);
    Kd[q] = balanced_digit_bytes<PZ_S>(I);
  }
  int4 pl[PZ_S];
  pack_digit_planes<PZ_S>(Kd, pl);
#pragma unroll
  for (int s = 0; s < PZ_S; ++s) *reinterpret_cast<int4*>(planes + (size_t)s * (mopad * kpad) + w * 16) = pl[s];
}

struct PzSchedule {
  int nptiles, njt, share, slots;   // share = CTAs per particle tile, slots = particle tiles in flight
};

// tile sequence of CTA c: particle tiles slot, slot + slots, ...; for each, output tiles njt-1-part, -share, ... where the
// part rotates with the particle tile: for a triangular A the k loops shrink with jt, so a fixed part gave one CTA of every
// pair 144 k-blocks per particle tile and the other 128 (share = 2, njt = 16) — 11 % more work on half of the CTAs
#define PZ_TILE_LOOP(sch)                                                                      \
  for (int ptile = (int)blockIdx.x / (sch).share; ptile < (sch).nptiles; ptile += (sch).slots) \
    for (int jt = (sch).njt - 1 - (((int)blockIdx.x % (sch).share + ptile / (sch).slots) % (sch).share); jt >= 0; jt -= (sch).share)

// warp roles: 0 = TMA producer, 1 = MMA issuer, 2..5 = epilogue (TMEM lane group = warp % 4)
constexpr int PZ_EPI_WARPS = 8;     // two per TMEM lane group: 32 of the 64 output columns each

template <typename T>
__global__ void __launch_bounds__(PZ_THREADS, 1)
k_pgemm_i8(const int8_t* __restrict__ xplanes, const int8_t* __restrict__ aplanes, const int* __restrict__ kend,
           const double* __restrict__ rowscale, const double* __restrict__ bias, T* const* __restrict__ outs, int Mo,
           size_t n, size_t npad, size_t kpad, size_t mopad, PzSchedule sch, int bulk_ok) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_full[PZ_STAGES];
  __shared__ __align__(8) unsigned long long bar_free[PZ_STAGES];
  __shared__ __align__(8) unsigned long long bar_acc;     // accumulators of a tile complete
  __shared__ __align__(8) unsigned long long bar_drained; // accumulators read back: TMEM reusable
  __shared__ uint32_t tmem_base_sh;
  __shared__ T* s_out[2][PZ_TN];          // per-tile output pointers / row scales / biases (double-buffered by tile parity)
  __shared__ __align__(16) double s_rs[2][PZ_TN], s_bi[2][PZ_TN];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t xplane = npad * kpad, aplane = mopad * kpad;
  const size_t nrg = npad / 8, njg = mopad / 8;
  if ((int)blockIdx.x / sch.share >= sch.nptiles) return;   // more CTAs than work (uniform per CTA)

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_sh)), "r"(PZ_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < PZ_STAGES; ++s) { mbar_init_tc(&bar_full[s], 1); mbar_init_tc(&bar_free[s], 1); }
    mbar_init_tc(&bar_acc, 1);
    mbar_init_tc(&bar_drained, PZ_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_sh;
  const uint32_t idesc = umma_idesc_i8(PZ_TM, PZ_TN);
  const uint32_t smem0 = smem_addr(smem);

  if (warp == 0) {
    // ---- producer: runs ahead over the whole tile sequence; `it` counts k-blocks of this CTA.  One elected thread
    // owns the loop; the next stage's free-barrier is tested while this stage's copies are issued.
    if (elect_one()) {
      int it = 0;
      bool ready = true;
      PZ_TILE_LOOP(sch) {
        const int nkb = kend[jt] / PZ_KB;
        const int8_t* srcX0 = xplanes + ((size_t)ptile * (PZ_TM / 8)) * 256;
        const int8_t* srcA0 = aplanes + ((size_t)jt * (PZ_TN / 8)) * 256;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = it % PZ_STAGES;
          if (it >= PZ_STAGES && !ready) mbar_wait_tc(&bar_free[st], (unsigned)((it / PZ_STAGES - 1) & 1));
          ready = (it + 1 >= PZ_STAGES) ? mbar_try_tc(&bar_free[(it + 1) % PZ_STAGES], (unsigned)(((it + 1) / PZ_STAGES - 1) & 1)) : true;
          mbar_expect_tx_tc(&bar_full[st], PZ_STAGE_BYTES);
          const uint32_t sbase = smem0 + (uint32_t)st * PZ_STAGE_BYTES;
          const int8_t* srcX = srcX0 + (size_t)kb * nrg * 256;
          const int8_t* srcA = srcA0 + (size_t)kb * njg * 256;
#pragma unroll
          for (int s = 0; s < PZ_S; ++s) {
            bulk_g2s_tc(sbase + s * PZ_X_SLICE, srcX + (size_t)s * xplane, PZ_X_SLICE, &bar_full[st]);
            bulk_g2s_tc(sbase + PZ_S * PZ_X_SLICE + s * PZ_A_SLICE, srcA + (size_t)s * aplane, PZ_A_SLICE, &bar_full[st]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer (see gram_tc.cu for the measurements behind this shape): one elected thread owns the loop, the
    // next stage's barrier is tested before the current k-block's MMAs are issued, WIDE-N MMAs (for a fixed particle
    // plane s one MMA covers the adjacent accumulators of diagonals s..6 and the adjacent coefficient planes 0..6-s:
    // 10 MMAs per k-block instead of 28), plane 0 first (it alone may overwrite), the longest MMAs last.
    if (elect_one()) {
      const uint64_t dX00 = umma_desc_kmajor(smem0, 128, 256);
      const uint64_t dA00 = umma_desc_kmajor(smem0 + PZ_S * PZ_X_SLICE, 128, 256);
      int it = 0;
      int tile = 0;
      PZ_TILE_LOOP(sch) {
        const int nkb = kend[jt] / PZ_KB;
        if (tile > 0) {   // the previous tile's accumulators must have been read back
          mbar_wait_tc(&bar_drained, (unsigned)((tile - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (nkb > 0) mbar_wait_tc(&bar_full[it % PZ_STAGES], (unsigned)((it / PZ_STAGES) & 1));
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = it % PZ_STAGES;
          const int nst = (it + 1) % PZ_STAGES;
          const unsigned npar = (unsigned)(((it + 1) / PZ_STAGES) & 1);
          const bool nready = (kb + 1 < nkb) ? mbar_try_tc(&bar_full[nst], npar) : true;   // consumed after the MMAs
          const uint64_t dX0 = dX00 + (uint64_t)(((uint32_t)st * PZ_STAGE_BYTES) >> 4);
          const uint64_t dA0 = dA00 + (uint64_t)(((uint32_t)st * PZ_STAGE_BYTES) >> 4);
          const uint32_t acc0 = kb == 0 ? 0u : 1u;
          auto issue = [&](int s, int c0n, int nn) {
            umma_i8(tmem_base + (uint32_t)(s * PZ_TN + c0n), dX0 + (uint64_t)((s * PZ_X_SLICE) >> 4),
                    dA0 + (uint64_t)(((c0n / PZ_TN) * PZ_A_SLICE) >> 4), umma_idesc_i8(PZ_TM, nn), s == 0 ? acc0 : 1u);
          };
          static_assert(PZ_S == 7 && PZ_TN == 64, "issue schedule is written for 7 digits and 64-column accumulators");
          issue(0, 0, 256); issue(0, 256, 192);     // diagonals 0..3, 4..6
          issue(6, 0, 64);                          // 6
          issue(5, 0, 128);                         // 5..6
          issue(4, 0, 192);                         // 4..6
          issue(2, 256, 64);                        // plane 2 x coefficient plane 4 -> 6
          issue(1, 256, 128);                       // plane 1 x coefficient planes 4..5 -> 5..6
          issue(3, 0, 256);                         // 3..6
          issue(2, 0, 256);                         // 2..5
          issue(1, 0, 256);                         // 1..4
          umma_commit(&bar_free[st]);
          if (kb == nkb - 1) umma_commit(&bar_acc);
          if (!nready) mbar_wait_tc(&bar_full[nst], npar);
        }
        if (nkb == 0) umma_commit(&bar_acc);   // nothing to accumulate: release the epilogue at once
        ++tile;
      }
    }
    __syncwarp();
  } else if (warp < 2 + PZ_EPI_WARPS) {
    // ---- epilogue warps: TMEM lane = particle; warps (2 + lg) and (6 + lg) share lane group lg and take 32 output
    // columns each.  The tile's MMAs cannot be followed by the next tile's until all 448 accumulator columns have been
    // read (no TMEM left for a second set), so this read-back is on the critical path: 8 warps halve it (per-tile
    // overhead measured at 12-14 k cycles with 4 warps and a sleeping wait).
    const int lg = warp & 3;             // TMEM lane group this warp may access (warp id modulo 4)
    const int h = (warp - 2) >> 2;       // which half of the 64 output columns
    const int et = tid - 64;             // index among the 256 epilogue threads
    int tile = 0;
    PZ_TILE_LOOP(sch) {
      const int nkb = kend[jt] / PZ_KB;
      // Output pointers, row scales and biases of the tile's 64 columns go to shared memory BEFORE the wait for the
      // accumulators.  While the tensor cores stream operands from shared memory every ordinary load / store of the
      // SM crawls (traced: ~130 cycles per LDS, ~250 per STG; constant-bank loads are no better): the 32 stores per
      // thread of a tile take ~20 k cycles, which hides behind the 31 k-cycle MMA phase of a dense tile but not
      // behind a triangular one's.  Global loads inside the store loop made it 29 k.
      {
        const int buf = tile & 1;
        if (et < PZ_TN) {
          const int jj = jt * PZ_TN + et;
          s_out[buf][et] = jj < Mo ? outs[jj] : nullptr;
          s_rs[buf][et] = jj < Mo ? rowscale[jj] : 0.0;
          s_bi[buf][et] = (jj < Mo && bias) ? bias[jj] : 0.0;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps only
      }
      mbar_wait_tc(&bar_acc, (unsigned)(tile & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // TMEM -> exact int64 recombination of the diagonals (0,1,2), (3,4,5), (6) (|acc_d| < 2^31, 16 extra bits)
      // -> 3 conversions per output
      double f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = 0.0;
      if (nkb > 0) {
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          long long comb[32];
          const int dn = g < 2 ? 3 : 1;
#if MCM_PZ_BATCH_LD
          // the group's loads are issued together and waited for once (7 serial load + wait pairs per tile before)
          uint32_t v[3][32];
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (q < dn) tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((g * 3 + q) * PZ_TN + h * 32), v[q]);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (q < dn) {
#pragma unroll
              for (int j = 0; j < 32; ++j) comb[j] = (q == 0) ? (long long)(int)v[q][j] : comb[j] * 256 + (long long)(int)v[q][j];
            }
#else
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (q < dn) {
              uint32_t v[32];
              const int d = g * 3 + q;
              tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(d * PZ_TN + h * 32), v);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j) comb[j] = (q == 0) ? (long long)(int)v[j] : comb[j] * 256 + (long long)(int)v[j];
            }
          }
#endif
          const double w = scalbn(1.0, -8 * (g * 3 + dn - 1 + 2));   // weight of the group's last diagonal
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fma((double)comb[j], w, f[j]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&bar_drained)) : "memory");
      const size_t i = (size_t)ptile * PZ_TM + lg * 32 + lane;
      const int buf = tile & 1;
#if MCM_PZ_BULK_STORE
      // Results -> staging tile [column][particle] in shared memory (a warp writes 256 contiguous bytes per column) ->
      // one bulk copy of the column's 128 results per output (issued by lane 0 of each warp for its 8 columns).  The
      // copies drain on the bulk-copy engine while these warps wait for the next tile's accumulators; the staging tile is
      // reused once the previous tile's copies have READ it (wait_group.read).  A ragged last particle tile stores directly.
      T* stage_t = reinterpret_cast<T*>(smem + (size_t)PZ_STAGES * PZ_STAGE_BYTES);
      const bool full_tile = bulk_ok && (size_t)ptile * PZ_TM + PZ_TM <= n;
      if (full_tile) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const double2* rs2 = reinterpret_cast<const double2*>(&s_rs[buf][h * 32]);
        const double2* bi2 = reinterpret_cast<const double2*>(&s_bi[buf][h * 32]);
        T* dst = stage_t + (size_t)(h * 32) * PZ_TM + lg * 32 + lane;
#pragma unroll
        for (int j2 = 0; j2 < 16; ++j2) {
          const double2 r = rs2[j2], b = bi2[j2];
          dst[(size_t)(2 * j2) * PZ_TM] = (T)fma(f[2 * j2], r.x, b.x);
          dst[(size_t)(2 * j2 + 1) * PZ_TM] = (T)fma(f[2 * j2 + 1], r.y, b.y);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (lane == 0) {
          const int ew = warp - 2;                       // 0..7: columns 8 ew .. 8 ew + 7
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int col = ew * 8 + q;
            T* o = s_out[buf][col];
            if (o)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           ::"l"(o + (size_t)ptile * PZ_TM), "r"(smem_addr(stage_t + (size_t)col * PZ_TM)), "r"((unsigned)(PZ_TM * sizeof(T)))
                           : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else
#endif
      if (i < n) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          T* o = s_out[buf][h * 32 + j];
          if (o) __stcs(o + i, (T)fma(f[j], s_rs[buf][h * 32 + j], s_bi[buf][h * 32 + j]));
        }
      }
      ++tile;
    }
#if MCM_PZ_BULK_STORE
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staging tile must outlive its copies
#endif
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(PZ_TMEM_COLS));
}

// returns MCM_OK, an error, or -1 = declined (caller runs the DMMA kernel).
// A_dev != nullptr: the coefficient matrix is already on the device (row-major Mo x Mi, e.g. the factor product of
// mcm_transform_moments) — no host copy of it is read, checked or uploaded; a non-finite entry is caught by k_pz_rows.
template <typename T>
int pgemm_tensor_impl(const double* A_host, int Mo, int Mi, const T* const* xs_host, const double* shift_host,
                      const double* bias_host, T* const* outs_host, size_t n, const double* A_dev) {
  Ctx& c = ctx();
  const size_t kpad = ((size_t)Mi + PZ_KB - 1) / PZ_KB * PZ_KB;
  const size_t npad = (n + PZ_TM - 1) / PZ_TM * PZ_TM;
  const size_t mopad = ((size_t)Mo + PZ_TN - 1) / PZ_TN * PZ_TN;
  if (kpad > (size_t)PZ_MAX_K) return -1;
  const int njt = (int)(mopad / PZ_TN);
  const size_t nptiles = npad / PZ_TM;
  if (nptiles > 0x7fffffffull) return -1;
  if (!A_dev)
    for (size_t q = 0; q < (size_t)Mo * Mi; ++q)
      if (!std::isfinite(A_host[q])) return -1;

  // ---- one upload: input/output pointer tables, shift, bias, A
  auto al = [](size_t b) { return (b + 15) / 16 * 16; };
  const size_t o_xs = 0, o_outs = al((size_t)Mi * sizeof(void*)), o_shift = o_outs + al((size_t)Mo * sizeof(void*));
  const size_t o_bias = o_shift + kpad * sizeof(double), o_A = o_bias + mopad * sizeof(double);
  const size_t up_bytes = o_A + (A_dev ? 0 : (size_t)Mo * Mi * sizeof(double));
  // staged in pinned memory (slot 2; a pageable upload of the 8 MB of A at C5 cost ~2 ms and synchronised)
  std::vector<unsigned char> up_fallback;
  unsigned char* up = static_cast<unsigned char*>(pinned_scratch(2, up_bytes));
  if (!up) { up_fallback.resize(up_bytes); up = up_fallback.data(); }
  memset(up, 0, o_A);
  memcpy(up + o_xs, xs_host, (size_t)Mi * sizeof(void*));
  memcpy(up + o_outs, outs_host, (size_t)Mo * sizeof(void*));
  if (shift_host) memcpy(up + o_shift, shift_host, (size_t)Mi * sizeof(double));
  if (bias_host) memcpy(up + o_bias, bias_host, (size_t)Mo * sizeof(double));
  if (!A_dev) memcpy(up + o_A, A_host, (size_t)Mo * Mi * sizeof(double));
  // ---- device workspace: ex, ratio, sc56, apow, rowscale, cscale, kend, flag, coefficient planes
  const size_t w_ex = 0, w_ratio = al((size_t)Mi * sizeof(int)), w_sc = w_ratio + al((size_t)Mi * sizeof(double));
  const size_t w_apow = w_sc + kpad * sizeof(double), w_rows = w_apow + kpad * sizeof(double);
  const size_t w_cs = w_rows + mopad * sizeof(double), w_kend = w_cs + mopad * sizeof(double);
  const size_t w_flag = w_kend + al((size_t)njt * sizeof(int)), w_planes = w_flag + 16;
  const size_t ws_bytes = w_planes + (size_t)PZ_S * mopad * kpad;
  unsigned char *d_up = nullptr, *d_ws = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_up, up_bytes, c.stream));
  MCM_CUDA(cudaMallocAsync((void**)&d_ws, ws_bytes, c.stream));
  MCM_CUDA(cudaMemcpyAsync(d_up, up, up_bytes, cudaMemcpyHostToDevice, c.stream));
  MCM_CUDA(cudaMemsetAsync(d_ws + w_rows, 0, w_planes - w_rows, c.stream));   // rowscale, cscale, kend, flag
  const T* const* d_xs = reinterpret_cast<const T* const*>(d_up + o_xs);
  const double* d_shift = reinterpret_cast<const double*>(d_up + o_shift);
  const double* d_A = A_dev ? A_dev : reinterpret_cast<const double*>(d_up + o_A);
  int* d_ex = reinterpret_cast<int*>(d_ws + w_ex);
  double* d_ratio = reinterpret_cast<double*>(d_ws + w_ratio);
  double* d_sc = reinterpret_cast<double*>(d_ws + w_sc);
  double* d_apow = reinterpret_cast<double*>(d_ws + w_apow);
  double* d_rows = reinterpret_cast<double*>(d_ws + w_rows);
  double* d_cs = reinterpret_cast<double*>(d_ws + w_cs);
  int* d_kend = reinterpret_cast<int*>(d_ws + w_kend);
  uint32_t* d_flag = reinterpret_cast<uint32_t*>(d_ws + w_flag);
  int8_t* d_aplanes = reinterpret_cast<int8_t*>(d_ws + w_planes);
  auto release = [&]() { cudaFreeAsync(d_up, c.stream); cudaFreeAsync(d_ws, c.stream); };

  StatsHandoff& ho = stats_handoff();
  if (ho.valid && ho.M == Mi && ho.n == n && ho.first_var == (const void*)xs_host[0]) {
    // statistics about exactly these shifts (the column means) came with the Gram: no third pass over the particles
    MCM_CUDA(cudaMemcpyAsync(d_ex, ho.d_ex, (size_t)Mi * sizeof(int), cudaMemcpyDeviceToDevice, c.stream));
    MCM_CUDA(cudaMemcpyAsync(d_ratio, ho.d_ratio, (size_t)Mi * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
  } else {
    k_rowstats<T><<<Mi, 256, 0, c.stream>>>(d_xs, d_shift, n, d_ex, d_ratio);
  }
  MCM_CHECK_LAUNCH();
  std::vector<int> ex(Mi);
  std::vector<double> ratio(Mi);
  MCM_CUDA(cudaMemcpyAsync(ex.data(), d_ex, (size_t)Mi * sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaMemcpyAsync(ratio.data(), d_ratio, (size_t)Mi * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));   // also: `up` has been consumed
  bool ok = true;
  for (int k = 0; k < Mi; ++k) ok &= (ratio[k] <= PZ_MAX_RATIO) && (ex[k] == -2000 || (ex[k] > -900 && ex[k] < 900));
  if (!ok) { release(); return -1; }

  k_pz_scales<<<(unsigned)((kpad + 255) / 256), 256, 0, c.stream>>>(d_ex, Mi, kpad, d_sc, d_apow);
  MCM_CHECK_LAUNCH();
  k_pz_rows<<<Mo, 256, 0, c.stream>>>(d_A, Mi, d_apow, d_rows, d_cs, d_kend, d_flag);
  MCM_CHECK_LAUNCH();
  {
    const size_t total = (kpad / PZ_KB) * (mopad / 8) * 16;
    k_pz_coef<<<(unsigned)((total + 255) / 256), 256, 0, c.stream>>>(d_A, Mo, Mi, d_apow, d_cs, mopad, kpad, d_aplanes);
    MCM_CHECK_LAUNCH();
  }
  int8_t* d_xplanes = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d_xplanes, (size_t)PZ_S * npad * kpad, c.stream));
  {
    const size_t total = (kpad / PZ_KB) * (npad / 8) * 16;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)c.sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_pslice<T><<<(unsigned)blocks, 256, 0, c.stream>>>(d_xs, d_shift, d_sc, Mi, n, npad, kpad, d_xplanes);
    MCM_CHECK_LAUNCH();
  }
  PzSchedule sch;
  sch.nptiles = (int)nptiles;
  sch.njt = njt;
  {
    // CTAs per particle tile.  With the rotating part (PZ_TILE_LOOP) the CTAs of a group stay on the same particle tile,
    // so its digit planes are read from DRAM once and served from L2 to the others.  Measured at C5 (ncu, triangular /
    // dense A): share 1: 12.1 / 22.8 ms with 39.7 / 107 GB of DRAM reads; share 2: 10.7 / 18.7 ms, 10.1 / 19.2 GB;
    // share 4: 10.6 / 18.1 ms, 7.4 / 7.9 GB = the 7.2 GB of digit planes.  (With a fixed part share 2 read 37 GB: the
    // unevenly loaded CTAs of a pair drifted apart.)  MCM_PGEMM_SHARE overrides for experiments.
    const char* e = getenv("MCM_PGEMM_SHARE");
    int want = e ? atoi(e) : 4;
    while (want > 1 && (njt < want || c.sm_count % want != 0)) want >>= 1;
    sch.share = want;
  }
  int bulk_ok = 1;      // bulk copies need 16-byte aligned destinations
  for (int j = 0; j < Mo; ++j) bulk_ok &= (reinterpret_cast<uintptr_t>(outs_host[j]) % 16 == 0) ? 1 : 0;
  int ncta = c.sm_count;
  if ((size_t)ncta > nptiles * sch.share) ncta = (int)(nptiles * sch.share);
  ncta = std::max(sch.share, ncta / sch.share * sch.share);
  sch.slots = ncta / sch.share;
  constexpr size_t kSmem = (size_t)PZ_STAGES * PZ_STAGE_BYTES + (MCM_PZ_BULK_STORE ? (size_t)PZ_TM * PZ_TN * sizeof(T) : 0);
  static bool attr = false;
  if (!attr) {
    constexpr int kSmemMax = PZ_STAGES * PZ_STAGE_BYTES + (MCM_PZ_BULK_STORE ? PZ_TM * PZ_TN * 8 : 0);
    static_assert(kSmemMax <= 227 * 1024, "operand ring + staging tile exceed the shared memory of an SM");
    MCM_CUDA(cudaFuncSetAttribute(k_pgemm_i8<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    MCM_CUDA(cudaFuncSetAttribute(k_pgemm_i8<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    attr = true;
  }
  k_pgemm_i8<T><<<ncta, PZ_THREADS, kSmem, c.stream>>>(
      d_xplanes, d_aplanes, d_kend, d_rows, bias_host ? reinterpret_cast<const double*>(d_up + o_bias) : nullptr,
      reinterpret_cast<T* const*>(d_up + o_outs), Mo, n, npad, kpad, mopad, sch, bulk_ok);
  MCM_CHECK_LAUNCH();
  uint32_t flag = 0;
  MCM_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(flag), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d_xplanes, c.stream));
  release();
  return flag ? -1 : MCM_OK;   // out-of-range coefficient scale: the caller recomputes on the DMMA kernel
}

template int pgemm_tensor_impl<double>(const double*, int, int, const double* const*, const double*, const double*, double* const*, size_t, const double*);
template int pgemm_tensor_impl<float>(const double*, int, int, const float* const*, const double*, const double*, float* const*, size_t, const double*);

}  // namespace mcm
