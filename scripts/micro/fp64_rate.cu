// Microbenchmark: FP64 instruction issue rate and dependent latency on one SM sub-partition, as a function of the
// number of resident warps and of the independent chains per thread.  One CTA per SM, WARPS*32 threads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_rate fp64_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { OP_DFMA = 0, OP_DMUL = 1, OP_DADD = 2, OP_MIX = 3, OP_DFMA_IMAD = 4, OP_DSETP = 5 };

template <int OP, int ILP>
__global__ void k_rate(double* out, long long* cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
  int acc = threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < ILP; ++i) {
        if (OP == OP_DFMA) x[i] = __fma_rn(x[i], a, b);
        if (OP == OP_DMUL) x[i] = __dmul_rn(x[i], a);
        if (OP == OP_DADD) x[i] = __dadd_rn(x[i], b);
        if (OP == OP_MIX) { x[i] = (u & 1) ? __dmul_rn(x[i], a) : __dadd_rn(x[i], b); }
        if (OP == OP_DFMA_IMAD) { x[i] = __fma_rn(x[i], a, b); acc = acc * 3 + it; }
        if (OP == OP_DSETP) { x[i] = x[i] > b ? __dadd_rn(x[i], a) : x[i]; }
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP, int ILP>
static void run(const char* name, int warps, double* out, long long* cyc) {
  const int iters = 2000;
  k_rate<OP, ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  k_rate<OP, ILP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double ninst = (double)iters * 8 * ILP * (OP == OP_DFMA_IMAD || OP == OP_DSETP ? 2 : 1);   // per warp
  const double per_smsp = ninst * (warps / 4.0);
  printf("%-10s warps/SM %2d ILP %d: %8lld cycles, %.2f cycles per warp-instr per SMSP, dependent step %.1f cycles  %s\n", name, warps, ILP,
         c, c / per_smsp, (double)c / (iters * 8), e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 16, 32}) {
    run<OP_DFMA, 1>("DFMA", warps, out, cyc);
    run<OP_DFMA, 2>("DFMA", warps, out, cyc);
    run<OP_DFMA, 4>("DFMA", warps, out, cyc);
    run<OP_DFMA, 8>("DFMA", warps, out, cyc);
  }
  for (int warps : {4, 32}) {
    run<OP_DMUL, 1>("DMUL", warps, out, cyc);
    run<OP_DMUL, 4>("DMUL", warps, out, cyc);
    run<OP_DADD, 1>("DADD", warps, out, cyc);
    run<OP_DADD, 4>("DADD", warps, out, cyc);
    run<OP_MIX, 4>("DMUL/DADD", warps, out, cyc);
    run<OP_DFMA_IMAD, 1>("DFMA+IMAD", warps, out, cyc);
    run<OP_DFMA_IMAD, 4>("DFMA+IMAD", warps, out, cyc);
    run<OP_DSETP, 1>("DSETP+DADD", warps, out, cyc);
    run<OP_DSETP, 4>("DSETP+DADD", warps, out, cyc);
  }
  return 0;
}
