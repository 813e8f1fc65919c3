// autocor.cu — StatsBase.autocor(p.particles, 1:maxlag) (demean = true), the kernel behind `ess`
// (src/sampling.jl:26-38, from MCMCDiagnostics.jl):
//   ac[k] = sum_{i=1}^{N-k} (x_i - m)(x_{i+k} - m) / sum_i (x_i - m)^2,   k = 1..maxlag (maxlag <= 256).
// One pass: each CTA stages a tile of 1024 centred values plus a halo of `maxlag` in shared memory; warp w
// accumulates lags w, w + 8, ... over the tile; per-CTA partial sums are combined in CTA order by a second
// kernel (fixed order => bit-reproducible).
#include <math.h>
#include <vector>
#include "mcm_common.cuh"

namespace mcm {

constexpr int AC_TILE = 1024;
constexpr int AC_MAXLAG = 256;

template <typename T>
__global__ void __launch_bounds__(256)
k_autocor(const T* __restrict__ x, size_t n, double mean, int maxlag, double* __restrict__ part) {
  __shared__ double s[AC_TILE + AC_MAXLAG];
  const size_t i0 = (size_t)blockIdx.x * AC_TILE;
  for (int e = threadIdx.x; e < AC_TILE + maxlag; e += 256) {
    const size_t i = i0 + e;
    s[e] = i < n ? (double)x[i] - mean : 0.0;      // zero padding: products beyond the end contribute nothing
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = warp; k <= maxlag; k += 8) {          // lag 0 = sum of squares
    double acc = 0.0;
    for (int e = lane; e < AC_TILE; e += 32) acc = fma(s[e], s[e + k], acc);
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
    if (lane == 0) part[(size_t)blockIdx.x * (AC_MAXLAG + 1) + k] = acc;
  }
}
__global__ void k_autocor_finish(const double* __restrict__ part, int nblk, int maxlag, double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > maxlag) return;
  double t = 0.0;
  for (int b = 0; b < nblk; ++b) t += part[(size_t)b * (AC_MAXLAG + 1) + k];
  out[k] = t;
}

int moments_device(int dtype, const void* in, size_t n, double* d_out6);
double* moments_result_slot();
int read_back(double* host_dst, const double* dev_src, size_t count);

}  // namespace mcm

using namespace mcm;

// out_host[k-1] = autocorrelation at lag k, k = 1..maxlag
extern "C" int mcm_autocor(int dtype, const void* in, size_t n, int maxlag, double* out_host) {
  MCM_NEED_INIT();
  if (!in || !out_host || maxlag < 1) return fail(MCM_ERR_INVALID, "mcm_autocor: bad arguments");
  if (maxlag > AC_MAXLAG) return fail(MCM_ERR_UNSUPPORTED, "mcm_autocor: at most %d lags", AC_MAXLAG);
  if ((size_t)maxlag >= n) return fail(MCM_ERR_INVALID, "mcm_autocor: lags must be smaller than the number of particles");
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_autocor: bad dtype %d", dtype);
  Ctx& c = ctx();
  double m6[6];
  int rc = moments_device(dtype, in, n, moments_result_slot());
  if (rc) return rc;
  rc = read_back(m6, moments_result_slot(), 6);
  if (rc) return rc;
  const size_t nblk = (n + AC_TILE - 1) / AC_TILE;
  if (nblk > 0x7fffffffull) return fail(MCM_ERR_UNSUPPORTED, "mcm_autocor: n too large");
  double* d = nullptr;
  MCM_CUDA(cudaMallocAsync((void**)&d, (nblk + 1) * (AC_MAXLAG + 1) * sizeof(double), c.stream));
  double* d_out = d + nblk * (AC_MAXLAG + 1);
  if (dtype == MCM_F64) k_autocor<double><<<(unsigned)nblk, 256, 0, c.stream>>>((const double*)in, n, m6[1], maxlag, d);
  else k_autocor<float><<<(unsigned)nblk, 256, 0, c.stream>>>((const float*)in, n, m6[1], maxlag, d);
  MCM_CHECK_LAUNCH();
  k_autocor_finish<<<(maxlag + 1 + 127) / 128, 128, 0, c.stream>>>(d, (int)nblk, maxlag, d_out);
  MCM_CHECK_LAUNCH();
  std::vector<double> h(maxlag + 1);
  MCM_CUDA(cudaMemcpyAsync(h.data(), d_out, (size_t)(maxlag + 1) * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  MCM_CUDA(cudaStreamSynchronize(c.stream));
  MCM_CUDA(cudaFreeAsync(d, c.stream));
  for (int k = 1; k <= maxlag; ++k) out_host[k - 1] = h[k] / h[0];
  return MCM_OK;
}
