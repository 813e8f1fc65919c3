// Microbenchmark: the producer / MMA-issuer pipeline of gram_tc.cu in isolation (no flush, no epilogue).
// MODE bits: 1 = commit per k-block to bar_free, 2 = MMA waits bar_full (real dependency on TMA), 4 = producer waits bar_free
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_tcgen05.cuh"
using namespace mcm;
namespace mcm { Ctx& ctx() { static Ctx c; return c; } int fail(int code, const char*, ...) { return code; } }
constexpr int STAGE = 6 * 4096 + 6 * 2048;
__device__ int g_timeout = 0;
__device__ long long g_tfree[4096], g_tfull[4096], g_t0, g_pend[4096], g_marr[4096], g_mend[4096];
// bounded wait: a microbenchmark must never hang the GPU
__device__ __forceinline__ void wait_b(unsigned long long* bar, unsigned parity) {
  for (int i = 0; i < (1 << 20); ++i) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  atomicExch(&g_timeout, 1);
}
template <int MODE, int STAGES>
__global__ void __launch_bounds__(256, 1) k_pipe(const unsigned char* __restrict__ src, long long* out, int rounds) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar_full[STAGES], bar_free[STAGES], bar_done;
  __shared__ uint32_t tmem_base_sh;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < STAGE * STAGES / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_sh)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init_tc(&bar_full[s], 1); mbar_init_tc(&bar_free[s], 1); }
    mbar_init_tc(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_sh, s0 = smem_addr(smem);
  const long long t0 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) g_t0 = t0;
  long long wfree = 0, wfull = 0;
  if (warp == 0) {
    for (int it = 0; it < rounds; ++it) {
      const int st = it % STAGES;
      { const long long w0 = clock64(); if ((MODE & 4) && it >= STAGES) wait_b(&bar_free[st], (unsigned)((it / STAGES - 1) & 1)); wfree += clock64() - w0; if (blockIdx.x == 0 && threadIdx.x == 0 && it < 4096) g_tfree[it] = clock64(); }
      if (!(MODE & 4) && it >= STAGES) wait_b(&bar_full[st], (unsigned)((it / STAGES - 1) & 1));   // self-paced ring
      if (elect_one()) {
        mbar_expect_tx_tc(&bar_full[st], STAGE);
        const uint32_t dst = s0 + STAGE * st;
        const unsigned char* g = src + (size_t)((blockIdx.x * 7 + it) % 64) * STAGE;
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          bulk_g2s_tc(dst + s * 4096, g + s * 4096, 4096, &bar_full[st]);
          bulk_g2s_tc(dst + 6 * 4096 + s * 2048, g + 6 * 4096 + s * 2048, 2048, &bar_full[st]);
        }
      }
      __syncwarp();
      if (blockIdx.x == 0 && threadIdx.x == 0 && it < 4096) g_pend[it] = clock64();
    }
    if (!(MODE & 2)) {   // nobody else waits for the data: drain the ring before the CTA may exit
      for (int st = 0; st < STAGES && st < rounds; ++st) {
        const int last = ((rounds - 1 - st) / STAGES) * STAGES + st;
        wait_b(&bar_full[st], (unsigned)((last / STAGES) & 1));
      }
    }
  } else if (warp == 1 && (MODE & 16)) {
    // optimised issuer: one elected thread owns the whole loop; the NEXT stage's barrier is tested before the current
    // k-block's MMAs are issued (the test's latency hides behind the blocking MMA issue); descriptors precomputed
    if (elect_one()) {
      uint64_t dA[STAGES], dB[STAGES];
#pragma unroll
      for (int s = 0; s < STAGES; ++s) { dA[s] = umma_desc_kmajor(s0 + STAGE * s, 128, 256); dB[s] = umma_desc_kmajor(s0 + STAGE * s + 6 * 4096, 128, 256); }
      wait_b(&bar_full[0], 0);
      for (int it0 = 0; it0 < rounds; it0 += STAGES) {
#pragma unroll
        for (int st = 0; st < STAGES; ++st) {
          const int it = it0 + st;
          if (it < rounds) {
            const int nst = (st + 1) % STAGES;
            const unsigned npar = (unsigned)(((it + 1) / STAGES) & 1);
            const bool nready = (it + 1 < rounds) ? mbar_try_tc(&bar_full[nst], npar) : true;   // issued now, consumed after the MMAs
#pragma unroll
            for (int s = 0; s < 6; ++s) {
              const int ncols = 64 * (6 - s);
#pragma unroll
              for (int c0 = 0; c0 < ncols; c0 += 256) {
                const int nn = ncols - c0 < 256 ? ncols - c0 : 256;
                umma_i8(tb + s * 64 + c0, dA[st] + (uint64_t)((s * 4096) >> 4), dB[st] + (uint64_t)(((c0 / 64) * 2048) >> 4), umma_idesc_i8(128, nn), 1u);
              }
            }
            umma_commit(&bar_free[st]);
            if (it == rounds - 1) umma_commit(&bar_done);
            if (!nready) wait_b(&bar_full[nst], npar);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    for (int it = 0; it < rounds; ++it) {
      const int st = it % STAGES;
      if (blockIdx.x == 0 && threadIdx.x == 32 && it < 4096) g_marr[it] = clock64();
      if (MODE & 2) { const long long w0 = clock64(); wait_b(&bar_full[st], (unsigned)((it / STAGES) & 1)); wfull += clock64() - w0; if (blockIdx.x == 0 && threadIdx.x == 32 && it < 4096) g_tfull[it] = clock64(); if (MODE & 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      if (elect_one()) {
        const uint32_t sb = s0 + STAGE * st;
        const uint64_t dA0 = umma_desc_kmajor(sb, 128, 256), dB0 = umma_desc_kmajor(sb + 6 * 4096, 128, 256);
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          const int ncols = 64 * (6 - s);
#pragma unroll
          for (int c0 = 0; c0 < ncols; c0 += 256) {
            const int nn = ncols - c0 < 256 ? ncols - c0 : 256;
            umma_i8(tb + s * 64 + c0, dA0 + (uint64_t)((s * 4096) >> 4), dB0 + (uint64_t)(((c0 / 64) * 2048) >> 4), umma_idesc_i8(128, nn), 1u);
          }
        }
        if (MODE & 1) umma_commit(&bar_free[st]);
        if (it == rounds - 1) umma_commit(&bar_done);
      }
      __syncwarp();
      if (blockIdx.x == 0 && threadIdx.x == 32 && it < 4096) g_mend[it] = clock64();
    }
  }
  __syncthreads();
  wait_b(&bar_done, 0);
  if (threadIdx.x == 0) out[3 * blockIdx.x] = clock64() - t0;
  if (threadIdx.x == 0) out[3 * blockIdx.x + 1] = wfree;
  if (threadIdx.x == 32) out[3 * blockIdx.x + 2] = wfull;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}
template <int MODE, int STAGES>
void run(const char* name, const unsigned char* src) {
  const int grid = 148, rounds = 600;
  long long* d; cudaMalloc(&d, grid * 24);
  const int smem = STAGE * STAGES;
  cudaFuncSetAttribute(k_pipe<MODE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_pipe<MODE, STAGES><<<grid, 256, smem>>>(src, d, rounds);
  k_pipe<MODE, STAGES><<<grid, 256, smem>>>(src, d, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[444]; cudaMemcpy(h, d, grid * 24, cudaMemcpyDeviceToHost);
  long long mx = 0; double wf = 0, wu = 0; for (int i = 0; i < grid; ++i) { mx = h[3 * i] > mx ? h[3 * i] : mx; wf += h[3 * i + 1]; wu += h[3 * i + 2]; }
  int to = 0; cudaMemcpyFromSymbol(&to, g_timeout, sizeof(int));
  printf("%-64s %7.1f cycles per k-block; producer waits free %6.1f, MMA waits full %6.1f per k-block (%s%s)\n", name, (double)mx / rounds,
         wf / grid / rounds, wu / grid / rounds, cudaGetErrorString(e), to ? ", WAIT TIMEOUT" : "");
  to = 0; cudaMemcpyToSymbol(g_timeout, &to, sizeof(int));
  fflush(stdout);
  cudaFree(d);
}
void dump(int n0, int n1) {
  static long long a[4096], b[4096]; long long t0;
  cudaMemcpyFromSymbol(a, g_tfree, sizeof(a)); cudaMemcpyFromSymbol(b, g_tfull, sizeof(b)); cudaMemcpyFromSymbol(&t0, g_t0, 8);
  static long long pe[4096], ma[4096], me[4096];
  cudaMemcpyFromSymbol(pe, g_pend, sizeof(pe)); cudaMemcpyFromSymbol(ma, g_marr, sizeof(ma)); cudaMemcpyFromSymbol(me, g_mend, sizeof(me));
  for (int i = n0; i < n1; ++i)
    printf("  it %3d: producer released %7lld issued %7lld | MMA arrives %7lld passes %7lld issued+committed %7lld\n", i, a[i] - t0, pe[i] - t0,
           ma[i] - t0, b[i] - t0, me[i] - t0);
}
int main() {
  unsigned char* src; cudaMalloc(&src, (size_t)64 * STAGE); cudaMemset(src, 1, (size_t)64 * STAGE);
  run<0, 4>("independent TMA ring and MMAs (no barriers between them)", src);
  run<1, 4>("+ commit per k-block", src);
  run<5, 4>("+ producer paced by bar_free, MMA free-running", src);
  run<7, 4>("+ MMA waits bar_full (NO tcgen05.fence): full pipeline, 4 stages", src);
  run<23, 4>("full pipeline, optimised issuer (prefetched barrier test, one thread), 4 stages", src);
  run<23, 2>("full pipeline, optimised issuer, 2 stages", src);
  run<7, 5>("full pipeline, 5 stages", src);
  run<7, 6>("full pipeline, 6 stages", src);
  run<7, 2>("full pipeline, 2 stages", src);
  return 0;
}
