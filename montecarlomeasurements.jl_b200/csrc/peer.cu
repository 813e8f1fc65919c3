// peer.cu — peer-memory windows for the one sharded operation whose reads leave the rank: `bootstrap`
// (src/resampling.jl:76-84) on particle-sharded variables.  Global uniform indices hit every shard,
// so (g-1)/g of the reads are remote.  Each rank exposes a window (plain cudaMalloc, CUDA IPC handle exchanged through
// the library's communicator); the gather kernel loads source elements straight from the owner's HBM over
// NVLink / NVSwitch (P2P loads, no staging, no request exchange) and writes its own output slice.  Ordering between
// "owners filled their windows" and "peers read them" is a stream-ordered collective on either side of the gather
// (mcm_comm_fence): a rank's contribution to the collective is ordered after its earlier work on the same stream.
#include <string.h>
#include <vector>
#include "mcm_common.cuh"

namespace mcm {

int comm_world(int* nranks, int* rank);
int comm_allreduce_sum_u32(uint32_t* d, size_t count);
int comm_allgather_bytes(const void* send, void* recv, size_t bytes_per_rank);

constexpr int PW_MAXRANKS = 16;
struct Window {
  void* local = nullptr;
  size_t bytes = 0;
  void* peers[PW_MAXRANKS] = {};
  int world = 1, rank = 0;
  bool live = false;
};
static std::vector<Window> g_windows;

struct PeerTable { const void* p[PW_MAXRANKS]; };

// Window layout: variables are interleaved in groups of PW_IL = 4 — element (group g, particle i, slot s) sits at
// ((g * per) + i) * 4 + s — so the four values a remote read needs for one particle share one 32-byte sector
// (Float64): NVLink moves sectors, and a one-variable-per-array layout would ship 32 bytes for every 8 useful ones.
constexpr int PW_IL = 4;
constexpr int PW_LAUNCH_VARS = 256;            // variables per launch (pointer table in the kernel parameters)
constexpr int PW_GROUPS_PER_BLOCK = 8;         // 8 independent 32-byte remote loads in flight per thread
struct VarTab { void* p[PW_LAUNCH_VARS]; };

template <typename T> struct Vec4 { T v[4]; };
__device__ __forceinline__ Vec4<double> ld_vec4(const double* p) {
  Vec4<double> r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3]) : "l"(p));
  return r;
}
__device__ __forceinline__ Vec4<float> ld_vec4(const float* p) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  return Vec4<float>{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void st_vec4(double* p, const Vec4<double>& r) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3]) : "memory");
}
__device__ __forceinline__ void st_vec4(float* p, const Vec4<float>& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}

// blockIdx.y = group (of this launch); src.p[4*y + s] may be null (padding slot)
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_window_fill(const __grid_constant__ VarTab src, T* __restrict__ win, size_t group0, size_t per, size_t n_local) {
  const int g = blockIdx.y;
  const T* s0 = static_cast<const T*>(src.p[4 * g + 0]);
  const T* s1 = static_cast<const T*>(src.p[4 * g + 1]);
  const T* s2 = static_cast<const T*>(src.p[4 * g + 2]);
  const T* s3 = static_cast<const T*>(src.p[4 * g + 3]);
  T* dst = win + (group0 + g) * per * PW_IL;
  const size_t stride = (size_t)gridDim.x * kThreads;
  for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n_local; i += stride) {
    Vec4<T> v;
    v.v[0] = s0 ? s0[i] : T(0); v.v[1] = s1 ? s1[i] : T(0); v.v[2] = s2 ? s2[i] : T(0); v.v[3] = s3 ? s3[i] : T(0);
    st_vec4(dst + i * PW_IL, v);
  }
}

// blockIdx.y = chunk of PW_GROUPS_PER_BLOCK groups; x (fastest) walks the outputs, so at any moment the owners serve
// few groups: their hot set (groups in flight x per x 32 B) stays in the owner's L2.
template <typename T>
__global__ void __launch_bounds__(kThreads)
k_gather_peer(const __grid_constant__ PeerTable tab, const __grid_constant__ VarTab outs, int ngroups, size_t group0,
              unsigned long long per, unsigned long long n_global, const long long* __restrict__ idx, size_t n_out,
              uint32_t* flags) {
  const int gfirst = blockIdx.y * PW_GROUPS_PER_BLOCK;
  const size_t stride = (size_t)gridDim.x * kThreads;
  bool oob = false;
  for (size_t j = (size_t)blockIdx.x * kThreads + threadIdx.x; j < n_out; j += stride) {
    const unsigned long long g = (unsigned long long)idx[j];
    if (g >= n_global) {                     // BoundsError: flag, and NaN outputs so the error cannot pass silently
      oob = true;
#pragma unroll
      for (int k = 0; k < PW_GROUPS_PER_BLOCK; ++k)
        if (gfirst + k < ngroups) {
#pragma unroll
          for (int s = 0; s < PW_IL; ++s) {
            T* o = static_cast<T*>(outs.p[(gfirst + k) * PW_IL + s]);
            if (o) o[j] = (T)__builtin_nan("");
          }
        }
      continue;
    }
    const unsigned owner = (unsigned)(g / per);
    const size_t i = (size_t)(g - (unsigned long long)owner * per);
    const T* src = static_cast<const T*>(tab.p[owner]);
    Vec4<T> v[PW_GROUPS_PER_BLOCK];
#pragma unroll
    for (int k = 0; k < PW_GROUPS_PER_BLOCK; ++k)
      if (gfirst + k < ngroups) v[k] = ld_vec4(src + ((group0 + gfirst + k) * per + i) * PW_IL);
#pragma unroll
    for (int k = 0; k < PW_GROUPS_PER_BLOCK; ++k)
      if (gfirst + k < ngroups) {
#pragma unroll
        for (int s = 0; s < PW_IL; ++s) {
          T* o = static_cast<T*>(outs.p[(gfirst + k) * PW_IL + s]);
          if (o) o[j] = v[k].v[s];
        }
      }
  }
  if (oob) atomicOr(flags, 0x80000000u);     // BoundsError bit, as mcm_gather
}

// non-collective teardown (process exit): close the peer mappings and free the local buffer without fences
int window_drop(int id) {
  if (id < 0 || id >= (int)g_windows.size() || !g_windows[id].live) return MCM_OK;
  Window& w = g_windows[id];
  for (int r = 0; r < w.world; ++r)
    if (r != w.rank && w.peers[r]) cudaIpcCloseMemHandle(w.peers[r]);
  cudaFree(w.local);
  w.live = false;
  return MCM_OK;
}

// for other translation units (sampling.cu: sharded resampling writes into index windows)
int window_lookup(int id, void** peers16, int* world, int* rank, size_t* bytes) {
  if (id < 0 || id >= (int)g_windows.size() || !g_windows[id].live) return fail(MCM_ERR_INVALID, "bad window %d", id);
  const Window& w = g_windows[id];
  for (int r = 0; r < PW_MAXRANKS; ++r) peers16[r] = w.peers[r];
  *world = w.world; *rank = w.rank; *bytes = w.bytes;
  return MCM_OK;
}

}  // namespace mcm

using namespace mcm;

// Stream-ordered fence across ranks: returns (asynchronously) once every rank has enqueued it; work enqueued after it
// on any rank runs after the work enqueued before it on every rank.  No-op without a communicator.
extern "C" int mcm_comm_fence(void) {
  MCM_NEED_INIT();
  int world, rank;
  comm_world(&world, &rank);
  if (world == 1) return MCM_OK;
  Ctx& c = ctx();
  uint32_t* d = reinterpret_cast<uint32_t*>(c.d_scratch + (c.scratch_doubles - 32));
  return comm_allreduce_sum_u32(d, 1);
}

// Collective: every rank allocates `bytes` of peer-visible memory and learns the address of every other rank's
// allocation.  *local_out = this rank's buffer.
extern "C" int mcm_window_create(size_t bytes, int* id_out, void** local_out) {
  MCM_NEED_INIT();
  if (!id_out || !local_out || bytes == 0) return fail(MCM_ERR_INVALID, "mcm_window_create: bad arguments");
  Ctx& c = ctx();
  Window w;
  comm_world(&w.world, &w.rank);
  if (w.world > PW_MAXRANKS) return fail(MCM_ERR_UNSUPPORTED, "mcm_window_create: at most %d ranks", PW_MAXRANKS);
  w.bytes = bytes;
  MCM_CUDA(cudaMalloc(&w.local, bytes));
  w.peers[w.rank] = w.local;
  if (w.world > 1) {
    cudaIpcMemHandle_t mine;
    MCM_CUDA(cudaIpcGetMemHandle(&mine, w.local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    unsigned char* d_h = nullptr;
    MCM_CUDA(cudaMallocAsync((void**)&d_h, 64 * (size_t)(w.world + 1), c.stream));
    MCM_CUDA(cudaMemcpyAsync(d_h, &mine, 64, cudaMemcpyHostToDevice, c.stream));
    int rc = comm_allgather_bytes(d_h, d_h + 64, 64);
    if (rc) return rc;
    std::vector<cudaIpcMemHandle_t> all(w.world);
    MCM_CUDA(cudaMemcpyAsync(all.data(), d_h + 64, 64 * (size_t)w.world, cudaMemcpyDeviceToHost, c.stream));
    MCM_CUDA(cudaStreamSynchronize(c.stream));
    MCM_CUDA(cudaFreeAsync(d_h, c.stream));
    for (int r = 0; r < w.world; ++r) {
      if (r == w.rank) continue;
      cudaError_t e = cudaIpcOpenMemHandle(&w.peers[r], all[r], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MCM_ERR_UNSUPPORTED, "mcm_window_create: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      }
    }
  }
  w.live = true;
  g_windows.push_back(w);
  *id_out = (int)g_windows.size() - 1;
  *local_out = w.local;
  return MCM_OK;
}

// Collective: all ranks must have finished reading (a fence is issued here) before the memory goes away.
extern "C" int mcm_window_free(int id) {
  MCM_NEED_INIT();
  if (id < 0 || id >= (int)g_windows.size() || !g_windows[id].live) return fail(MCM_ERR_INVALID, "mcm_window_free: bad window %d", id);
  Window& w = g_windows[id];
  int rc = mcm_comm_fence();
  if (rc) return rc;
  MCM_CUDA(cudaStreamSynchronize(ctx().stream));
  for (int r = 0; r < w.world; ++r)
    if (r != w.rank && w.peers[r]) cudaIpcCloseMemHandle(w.peers[r]);
  // second fence: nobody frees while a peer still has the mapping open
  rc = mcm_comm_fence();
  if (rc) return rc;
  MCM_CUDA(cudaStreamSynchronize(ctx().stream));
  MCM_CUDA(cudaFree(w.local));
  w.live = false;
  return MCM_OK;
}

template <typename T>
static int window_fill_t(const Window& w, const void* const* vars, int M, size_t n_local, size_t per) {
  Ctx& c = ctx();
  for (int v0 = 0; v0 < M; v0 += PW_LAUNCH_VARS) {
    VarTab tab{};
    const int nv = M - v0 < PW_LAUNCH_VARS ? M - v0 : PW_LAUNCH_VARS;
    for (int v = 0; v < nv; ++v) tab.p[v] = const_cast<void*>(vars[v0 + v]);
    const int ngroups = (nv + PW_IL - 1) / PW_IL;
    size_t gx = (n_local + kThreads - 1) / kThreads;
    const size_t cap = (size_t)c.sm_count * 8;
    if (gx > cap) gx = cap;
    if (gx == 0) gx = 1;
    k_window_fill<T><<<dim3((unsigned)gx, (unsigned)ngroups), kThreads, 0, c.stream>>>(tab, static_cast<T*>(w.local),
                                                                                   (size_t)v0 / PW_IL, per, n_local);
    MCM_CHECK_LAUNCH();
  }
  return MCM_OK;
}
template <typename T>
static int gather_peer_t(const Window& w, int M, uint64_t per, uint64_t n_global, const int64_t* idx, size_t n_out,
                         void* const* outs) {
  Ctx& c = ctx();
  PeerTable tab{};
  for (int r = 0; r < w.world; ++r) tab.p[r] = w.peers[r];
  for (int v0 = 0; v0 < M; v0 += PW_LAUNCH_VARS) {
    VarTab ot{};
    const int nv = M - v0 < PW_LAUNCH_VARS ? M - v0 : PW_LAUNCH_VARS;
    for (int v = 0; v < nv; ++v) ot.p[v] = outs[v0 + v];
    const int ngroups = (nv + PW_IL - 1) / PW_IL;
    size_t gx = (n_out + kThreads - 1) / kThreads;
    if (gx > 65535u * 16u) gx = 65535u * 16u;
    const int gy = (ngroups + PW_GROUPS_PER_BLOCK - 1) / PW_GROUPS_PER_BLOCK;
    k_gather_peer<T><<<dim3((unsigned)gx, (unsigned)gy), kThreads, 0, c.stream>>>(
        tab, ot, ngroups, (size_t)v0 / PW_IL, per, n_global, (const long long*)idx, n_out, c.d_flags);
    MCM_CHECK_LAUNCH();
  }
  return MCM_OK;
}

static size_t window_need(int M, uint64_t per, size_t es) { return (size_t)((M + PW_IL - 1) / PW_IL) * PW_IL * per * es; }

// Interleave this rank's slices vars[v][0..n_local) of M variables into its window (groups of four variables per
// 32-byte sector; mcm_window_bytes(M, per, dtype) gives the size to create).
extern "C" int mcm_window_fill(int dtype, int window, const void* const* vars, int M, size_t n_local, uint64_t per) {
  MCM_NEED_INIT();
  if (window < 0 || window >= (int)g_windows.size() || !g_windows[window].live)
    return fail(MCM_ERR_INVALID, "mcm_window_fill: bad window %d", window);
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_window_fill: bad dtype %d", dtype);
  if (M < 1 || !vars || n_local > per) return fail(MCM_ERR_INVALID, "mcm_window_fill: bad arguments");
  const Window& w = g_windows[window];
  if (window_need(M, per, dtype == MCM_F64 ? 8 : 4) > w.bytes) return fail(MCM_ERR_INVALID, "mcm_window_fill: window too small");
  for (int v = 0; v < M; ++v)
    if (!vars[v]) return fail(MCM_ERR_INVALID, "mcm_window_fill: vars[%d] is null", v);
  if (n_local == 0) return MCM_OK;
  return dtype == MCM_F64 ? window_fill_t<double>(w, vars, M, n_local, per) : window_fill_t<float>(w, vars, M, n_local, per);
}
extern "C" size_t mcm_window_bytes(int M, uint64_t per, int dtype) { return window_need(M, per, dtype == MCM_F64 ? 8 : 4); }

// outs[v][j] = variable v of global particle idx[j], read from the window of its owner idx[j] / per (filled by
// mcm_window_fill with the same M and per on every rank).  Bracket with mcm_comm_fence().
extern "C" int mcm_gather_peer(int dtype, int window, int M, uint64_t per, uint64_t n_global, const int64_t* idx,
                               size_t n_out, void* const* outs) {
  MCM_NEED_INIT();
  if (window < 0 || window >= (int)g_windows.size() || !g_windows[window].live)
    return fail(MCM_ERR_INVALID, "mcm_gather_peer: bad window %d", window);
  if (dtype != MCM_F64 && dtype != MCM_F32) return fail(MCM_ERR_INVALID, "mcm_gather_peer: bad dtype %d", dtype);
  if (M < 1 || !outs || (n_out && !idx) || per == 0) return fail(MCM_ERR_INVALID, "mcm_gather_peer: bad arguments");
  const Window& w = g_windows[window];
  if (window_need(M, per, dtype == MCM_F64 ? 8 : 4) > w.bytes) return fail(MCM_ERR_INVALID, "mcm_gather_peer: window too small");
  if ((n_global + per - 1) / per > (uint64_t)w.world) return fail(MCM_ERR_INVALID, "mcm_gather_peer: n_global exceeds world * per");
  for (int v = 0; v < M; ++v)
    if (!outs[v]) return fail(MCM_ERR_INVALID, "mcm_gather_peer: outs[%d] is null", v);
  if (n_out == 0) return MCM_OK;
  return dtype == MCM_F64 ? gather_peer_t<double>(w, M, per, n_global, idx, n_out, outs)
                          : gather_peer_t<float>(w, M, per, n_global, idx, n_out, outs);
}
