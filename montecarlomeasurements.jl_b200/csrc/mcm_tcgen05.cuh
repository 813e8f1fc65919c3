// mcm_tcgen05.cuh — thin inline-PTX wrappers for the 5th-generation tensor-core path (sm_100a):
// shared-memory matrix descriptors, tcgen05.mma kind::i8, commit -> mbarrier, tcgen05.ld, TMA bulk copies.
// Shared by gram_tc.cu (pcov) and pgemm_tc.cu (matrix x particle vector).
#pragma once
#include <math.h>
#include "mcm_common.cuh"
#include "mcm_digits.cuh"

namespace mcm {

// ---- tcgen05 helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle ("interleave"): 8-row x 16-byte core matrices; LBO = byte distance between the two
// core matrices along K, SBO = byte distance between consecutive 8-row groups.  version = 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor, kind::i8: D = s32, A = B = signed int8, both K-major, M x N
__device__ __forceinline__ uint32_t umma_idesc_i8(int Mt, int Nt) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Nt >> 3) << 17) | ((uint32_t)(Mt >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init_tc(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait_tc(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE_%=;\n"
      "bra TC_WAIT_%=;\n"
      "TC_DONE_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity)
      : "memory");
}
// Wait used by warps that have nothing to do for a long time (the epilogue warps during a tile's MMA phase): a
// bare try_wait spin from 6 idle warps hammers the mbarrier in shared memory while UMMA operand fetches and TMA
// writes need the same port; back off between polls.
__device__ __forceinline__ bool mbar_try_tc(unsigned long long* bar, unsigned parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_backoff(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_tc(bar, parity)) __nanosleep(512);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void cp_async16_tc(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx_tc(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s_tc(uint32_t dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}


// A operand from TMEM ("TS" form).  The SS form reads the 128 x 32-byte A tile from shared memory for EVERY
// MMA (32 wavefronts of 128 B, plus 16 for a 64-row B tile, against a 32-cycle MMA: ncu shows l1tex at 80 %
// and the tensor pipe at 53 %); staging each A plane once per k-block with tcgen05.cp and reusing it from
// TMEM leaves only the B reads on the shared-memory datapath.
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One elected lane of a converged warp.  Inside `if (elect_one())` ptxas knows a single thread is active and
// emits the uniform-datapath instructions (UTCIMMA, UBLKCP, UTCBAR) straight; under `if (lane == 0)` it wraps
// EVERY one of them in an ELECT / BRA.U.ANY serialisation loop (~9 issue slots and an R2UR stall per MMA —
// at 32-cycle MMAs the single issuing thread then paces the tensor pipe, measured 49 %).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// L2 residency hints for bulk copies: operand tiles that are re-read by later output tiles should outlive the
// streaming traffic (result stores, other CTAs' tiles) in L2.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, unsigned bytes, unsigned long long* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
               : "memory");
}

// ---- thread-block clusters: TMA multicast + multicast commit -----------------------------------------
// A CTA pair that needs the same operand tile loads half of it each and multicasts: the data lands at the
// same shared-memory offset in every CTA of `mask` and completes bytes on the mbarrier at the same offset.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, unsigned bytes, unsigned long long* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_addr(bar)), "h"(mask)
               : "memory");
}
// tcgen05.commit that arrives on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(unsigned long long* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_addr(bar)), "h"(mask)
               : "memory");
}

// ---- operand preparation ------------------------------------------------------------------
// Per row: e[i] = exponent with max_k |x_i[k] - mu_i| * 256/255 < 2^e[i] (so that the leading 8-bit digit
// cannot overflow), and ratio[i] = max / sigma (accuracy guard).  e = -2000 marks an all-zero row.
template <typename T>
__global__ void __launch_bounds__(256)
k_rowstats(const T* const* __restrict__ vars, const double* __restrict__ means, size_t n, int* __restrict__ ex,
           double* __restrict__ ratio) {
  const T* x = vars[blockIdx.x];
  const double mu = means[blockIdx.x];
  double m = 0.0, q = 0.0;
  for (size_t i = threadIdx.x; i < n; i += 256) {
    const double d = (double)x[i] - mu;
    m = fmax(m, fabs(d));
    q = fma(d, d, q);
  }
  __shared__ double shm[8], shq[8];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    m = fmax(m, __shfl_down_sync(0xffffffffu, m, d));
    q += __shfl_down_sync(0xffffffffu, q, d);
  }
  if ((threadIdx.x & 31) == 0) { shm[threadIdx.x >> 5] = m; shq[threadIdx.x >> 5] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { m = fmax(m, shm[w]); q += shq[w]; }
    int e = -2000;
    double r = 0.0;
    if (m > 0.0 && m < 1.7e308) {
      double mant = frexp(m, &e);           // m = mant * 2^e, mant in [0.5, 1)
      if (mant > 0.995) ++e;                // keep 256 * y below 127.5
      r = m / sqrt(q / (double)n);
    } else if (m > 0.0) {
      e = 2000; r = 1e300;                  // Inf/NaN in the row: force the fallback
    }
    ex[blockIdx.x] = e;
    ratio[blockIdx.x] = r;
  }
}



// ---- balanced base-256 digits of a 64-bit integer in one addition ------------------------------------------------
// The Ozaki slicers need the S signed digits d_b in [-128, 127] with I = sum_b d_b 256^b.  Peeling them off one at a
// time (d = ((I + 128) & 255) - 128; I = (I - d) >> 8) costs ~10 integer instructions per digit in 64-bit arithmetic:
// ncu showed k_pslice at 59 instructions per value, ALU pipe 56 %, half the HBM rate.  Since
//     I + sum_b 128 * 256^b = sum_b (d_b + 128) 256^b     with every (d_b + 128) in [0, 255],
// the UNSIGNED bytes of I + C (C = S bytes of 0x80) are d_b + 128, and flipping bit 7 of each byte turns them into the
// int8 digits: two 64-bit operations for all digits.  The top byte wraps exactly like the (int8_t) cast of the loop's
// remainder did, so the planes are bit-identical to the loop's for every I.
// (balanced_digit_bytes lives in mcm_digits.cuh: plain C++, also compiled by tests/test_digit_split.py)
// plane s (s = 0 most significant) of 16 values -> 16 bytes: byte S-1-s of every K[j], gathered with byte permutes
template <int S>
__device__ __forceinline__ void pack_digit_planes(const unsigned long long (&K)[16], int4 (&out)[S]) {
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const int b = S - 1 - s;
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = b < 4 ? (uint32_t)K[4 * q + i] : (uint32_t)(K[4 * q + i] >> 32);
      const uint32_t sel = (uint32_t)(b & 3) | ((4u + (uint32_t)(b & 3)) << 4);
      w[q] = __byte_perm(__byte_perm(v[0], v[1], sel), __byte_perm(v[2], v[3], sel), 0x5410);
    }
    out[s] = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
  }
}

}  // namespace mcm
