// Microbenchmark: cycles per tcgen05.mma kind::i8 (M128, N in {64,128,256}, K32), SS and TS forms, issued back to back
// by one elected thread on fixed shared-memory operands.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../montecarlomeasurements.jl_b200/csrc/mcm_tcgen05.cuh"
using namespace mcm;
namespace mcm { Ctx& ctx() { static Ctx c; return c; } int fail(int code, const char*, ...) { return code; } }

template <int N, bool TS, int NACC>
__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base_sh;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_sh)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x == 0) { mbar_init_tc(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_sh;
  const uint32_t idesc = umma_idesc_i8(128, N);
  const uint32_t s0 = smem_addr(smem);
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      const uint64_t dA = umma_desc_kmajor(s0, 128, 256);
      const uint64_t dB = umma_desc_kmajor(s0 + 8192, 128, 256);
      if (TS) tmem_cp_128x256b(tb + 480, dA);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          if (TS) umma_i8_ts(tb + a * N, tb + 480, dB, idesc, 1u);
          else umma_i8(tb + a * N, dA, dB, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait_tc(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512));
}

template <int N, bool TS, int NACC>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, grid * 8);
  const int iters = 4000;
  cudaFuncSetAttribute(k_rate<N, TS, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k_rate<N, TS, NACC><<<grid, 128, 65536>>>(d, iters);
  k_rate<N, TS, NACC><<<grid, 128, 65536>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-44s grid %3d: %.1f cycles per MMA (%s)\n", name, grid, (double)mx / (iters * NACC), cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int grid : {1, 148}) {
    run<64, false, 1>("i8 M128 N64  K32 SS, 1 accumulator", grid);
    run<64, false, 6>("i8 M128 N64  K32 SS, 6 accumulators", grid);
    run<64, true, 1>("i8 M128 N64  K32 TS, 1 accumulator", grid);
    run<64, true, 6>("i8 M128 N64  K32 TS, 6 accumulators", grid);
    run<128, false, 1>("i8 M128 N128 K32 SS, 1 accumulator", grid);
    run<128, true, 3>("i8 M128 N128 K32 TS, 3 accumulators", grid);
    run<256, false, 1>("i8 M128 N256 K32 SS, 1 accumulator", grid);
    run<256, true, 1>("i8 M128 N256 K32 TS, 1 accumulator", grid);
  }
  return 0;
}
