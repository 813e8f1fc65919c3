// Microbenchmark: do TMA bulk copies into shared memory slow down tcgen05.mma (and vice versa)?
// warp 0: back-to-back "k-blocks" of int8 MMAs (21 narrow N=64 or the 8 wide-N of gram_tc.cu) on fixed operands;
// warp 1: a 4-deep ring of cp.async.bulk copies (12 per round: 6 x 4 KiB + 6 x 2 KiB = 36 KiB) from an L2-hot buffer.
// Reports cycles per MMA k-block and bytes/cycle of TMA, alone and together.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_tcgen05.cuh"
using namespace mcm;
namespace mcm { Ctx& ctx() { static Ctx c; return c; } int fail(int code, const char*, ...) { return code; } }

constexpr int STAGE = 6 * 4096 + 6 * 2048;   // 36864
template <bool WIDE>
__global__ void __launch_bounds__(128, 1) k_both(const unsigned char* __restrict__ src, long long* out, int mma_rounds, int tma_rounds) {
  extern __shared__ __align__(1024) unsigned char smem[];   // [0,36864) MMA operands, then 4 TMA stages
  __shared__ __align__(8) unsigned long long bar_mma, bar_tma[4];
  __shared__ uint32_t tmem_base_sh;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < STAGE / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_sh)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) {
    mbar_init_tc(&bar_mma, 1);
    for (int s = 0; s < 4; ++s) mbar_init_tc(&bar_tma[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_sh, s0 = smem_addr(smem);
  if (warp == 0 && mma_rounds > 0) {
    long long t0 = clock64();
    if (elect_one()) {
      const uint64_t dA0 = umma_desc_kmajor(s0, 128, 256), dB0 = umma_desc_kmajor(s0 + 6 * 4096, 128, 256);
      for (int r = 0; r < mma_rounds; ++r) {
        if (WIDE) {
#pragma unroll
          for (int s = 0; s < 6; ++s) {
            const int ncols = 64 * (6 - s);
#pragma unroll
            for (int c0 = 0; c0 < ncols; c0 += 256) {
              const int nn = ncols - c0 < 256 ? ncols - c0 : 256;
              umma_i8(tb + s * 64 + c0, dA0 + (uint64_t)((s * 4096) >> 4), dB0 + (uint64_t)(((c0 / 64) * 2048) >> 4), umma_idesc_i8(128, nn), 1u);
            }
          }
        } else {
#pragma unroll
          for (int s = 0; s < 6; ++s)
#pragma unroll
            for (int t = 0; t < 6; ++t)
              if (s + t < 6) umma_i8(tb + (s + t) * 64, dA0 + (uint64_t)((s * 4096) >> 4), dB0 + (uint64_t)((t * 2048) >> 4), umma_idesc_i8(128, 64), 1u);
        }
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait_tc(&bar_mma, 0);
    if (threadIdx.x == 0) out[2 * blockIdx.x] = clock64() - t0;
  }
  if (warp == 1 && tma_rounds > 0) {
    long long t0 = clock64();
    for (int r = 0; r < tma_rounds; ++r) {
      const int st = r & 3;
      if (r >= 4) mbar_wait_tc(&bar_tma[st], (unsigned)(((r >> 2) - 1) & 1));
      if (elect_one()) {
        mbar_expect_tx_tc(&bar_tma[st], STAGE);
        const uint32_t dst = s0 + STAGE * (1 + st);
        const unsigned char* g = src + (size_t)((blockIdx.x * 7 + r) % 64) * STAGE;
#pragma unroll
        for (int s = 0; s < 6; ++s) {
          bulk_g2s_tc(dst + s * 4096, g + s * 4096, 4096, &bar_tma[st]);
          bulk_g2s_tc(dst + 6 * 4096 + s * 2048, g + 6 * 4096 + s * 2048, 2048, &bar_tma[st]);
        }
      }
      __syncwarp();
    }
    for (int st = 0; st < 4; ++st) {
      const int last = ((tma_rounds - 1 - st) / 4) * 4 + st;   // last round that used stage st
      if (tma_rounds > st) mbar_wait_tc(&bar_tma[st], (unsigned)((last >> 2) & 1));
    }
    if (threadIdx.x == 32) out[2 * blockIdx.x + 1] = clock64() - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

template <bool WIDE>
void run(const char* name, const unsigned char* src, int mr, int tr) {
  const int grid = 148;
  long long* d; cudaMalloc(&d, grid * 16); cudaMemset(d, 0, grid * 16);
  const int smem = STAGE * 5;
  cudaFuncSetAttribute(k_both<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_both<WIDE><<<grid, 128, smem>>>(src, d, mr, tr);
  k_both<WIDE><<<grid, 128, smem>>>(src, d, mr, tr);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[296]; cudaMemcpy(h, d, grid * 16, cudaMemcpyDeviceToHost);
  long long m0 = 0, m1 = 0; for (int i = 0; i < grid; ++i) { if (h[2 * i] > m0) m0 = h[2 * i]; if (h[2 * i + 1] > m1) m1 = h[2 * i + 1]; }
  printf("%-34s", name);
  if (mr) printf("  MMA: %7.1f cycles per k-block", (double)m0 / mr);
  if (tr) printf("  TMA: %6.1f B/cycle/SM (%7.1f cycles per 36 KiB round)", (double)STAGE * tr / m1, (double)m1 / tr);
  printf("  (%s)\n", cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  unsigned char* src; cudaMalloc(&src, (size_t)64 * STAGE); cudaMemset(src, 1, (size_t)64 * STAGE);
  run<false>("narrow MMAs alone", src, 2000, 0);
  run<true>("wide MMAs alone", src, 2000, 0);
  run<false>("TMA alone", src, 0, 2000);
  run<false>("narrow MMAs + TMA", src, 2000, 2000);
  run<true>("wide MMAs + TMA", src, 2000, 2000);
  run<true>("wide MMAs + TMA (TMA 2x rounds)", src, 2000, 4000);
  return 0;
}
