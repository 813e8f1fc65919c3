/*
 * mcm_b200.h — C ABI of the B200-native particle-arithmetic engine for
 * MonteCarloMeasurements.jl (reference: /root/reference, v1.5.1; all file:line
 * citations below are relative to it).
 *
 * The reference has NO FFI: its "operator interface" is Julia method dispatch on
 * `AbstractParticles{T,N}` (src/types.jl:4-28) whose method bodies are broadcasts
 * over the `particles::Vector{T}` field.  This header is what a `ccall`-based
 * device-resident subtype binds instead of those bodies (see INTEGRATION.md for
 * the Julia stub).  One entry point per reference method family; every function
 *   - takes plain pointers / sizes / scalars only (no torch, no C++ types),
 *   - returns an int status (MCM_OK == 0); `mcm_last_error()` gives the text,
 *   - is asynchronous and ordered on the library stream unless it returns a
 *     host scalar (reductions, quantiles, downloads synchronise).
 *
 * Device pointers are ordinary CUDA device addresses (from `mcm_alloc`, or any
 * other allocator in the same process, e.g. a torch tensor's data_ptr()).
 * The N samples of one uncertain scalar are ONE contiguous array of `T`
 * (SoA; src/types.jl:17-19).  A vector of M uncertain variables is M such
 * arrays (index-aligned: particle i of every variable is the same draw).
 */
#ifndef MCM_B200_H
#define MCM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCM_B200_VERSION 100 /* 0.1.0 */

/* ---- status codes ------------------------------------------------------- */
enum {
  MCM_OK = 0,
  MCM_ERR_INVALID = 1,     /* bad argument (ArgumentError in the reference)      */
  MCM_ERR_CUDA = 2,        /* CUDA runtime error                                  */
  MCM_ERR_NOMEM = 3,       /* allocation failed                                   */
  MCM_ERR_UNSUPPORTED = 4, /* op exists in the reference but not on device (MethodError) */
  MCM_ERR_NCCL = 5,        /* NCCL error / communicator missing                   */
  MCM_ERR_NOT_INIT = 6     /* mcm_init not called                                 */
};

/* element types of device-resident particles (src/types.jl:17, T) */
enum { MCM_F32 = 0, MCM_F64 = 1 };

/* sticky numerical flags (Base scalar functions THROW DomainError; kernels
 * cannot — they return NaN and raise a flag, surfaced by mcm_check_flags).   */
enum {
  MCM_FLAG_DOMAIN = 1u,     /* NaN produced from non-NaN input (sqrt(-1), log(-1), ...) */
  MCM_FLAG_NAN_QUANTILE = 2u, /* NaN met by quantile/sort (Statistics.quantile errors)    */
  MCM_FLAG_COMM = 4u,         /* a rank never arrived in an in-kernel cross-rank merge (20 s)  */
  MCM_FLAG_BOUNDS = 0x80000000u /* gather index outside [0, N) (BoundsError, resampling.jl:77,82); the output is NaN */
};

/* ---- unary primitives -----------------------------------------------------
 * The first 35 are exactly the list registered at src/particles.jl:185-190
 * through register_primitive_single (src/register_primitive.jl:89-100); for
 * Float64/Float32 particles exp/log/sin/cos/asin/acos/atan are the methods of
 * src/sleefpirates.jl:2-36.  The rest are the per-element bodies of
 * src/particles.jl:346-348 (round), :388-391 (literal_pow 2,3,-1), floor/ceil/
 * trunc (Base fallbacks through Real).                                        */
enum {
  MCM_U_PLUS = 0, MCM_U_NEG, MCM_U_EXP, MCM_U_EXP2, MCM_U_EXP10, MCM_U_EXPM1,
  MCM_U_LOG, MCM_U_LOG10, MCM_U_LOG2, MCM_U_LOG1P,
  MCM_U_SIN, MCM_U_COS, MCM_U_TAN, MCM_U_SIND, MCM_U_COSD, MCM_U_TAND,
  MCM_U_SINH, MCM_U_COSH, MCM_U_TANH,
  MCM_U_ASIN, MCM_U_ACOS, MCM_U_ATAN, MCM_U_ASIND, MCM_U_ACOSD, MCM_U_ATAND,
  MCM_U_ASINH, MCM_U_ACOSH, MCM_U_ATANH,
  MCM_U_ZERO, MCM_U_SIGN, MCM_U_ABS, MCM_U_SQRT, MCM_U_RAD2DEG, MCM_U_DEG2RAD,
  MCM_U_FLOAT, /* 34: last registered primitive */
  MCM_U_ROUND, /* round(x) RoundNearest (ties to even), particles.jl:346 */
  MCM_U_FLOOR, MCM_U_CEIL, MCM_U_TRUNC,
  MCM_U_INV,   /* literal_pow(^,x,Val(-1)) == inv(x) */
  MCM_U_SQR,   /* literal_pow(^,x,Val(2))  == x*x   (particles.jl:388-391) */
  MCM_U_CUBE,  /* literal_pow(^,x,Val(3))  == x*x*x */
  /* round(p, r::RoundingMode) (particles.jl:346): RoundNearest = MCM_U_ROUND, RoundToZero = MCM_U_TRUNC,
   * RoundUp = MCM_U_CEIL, RoundDown = MCM_U_FLOOR; the remaining modes of Base: */
  MCM_U_ROUND_TIESAWAY, MCM_U_ROUND_TIESUP, MCM_U_ROUND_FROMZERO,
  /* rem2pi(p, r::RoundingMode) (particles.jl:359); mod2pi(p) == rem2pi(p, RoundDown) (runtests.jl:476,480).
   * Correctly rounded remainders for every finite argument (Cody-Waite, Payne-Hanek beyond 2^17 pi/2). */
  MCM_U_REM2PI_NEAREST, MCM_U_REM2PI_TOZERO, MCM_U_REM2PI_DOWN, MCM_U_REM2PI_UP,
  MCM_U_COUNT
};

/* ---- binary primitives ------------------------------------------------------
 * register_primitive_multi list, src/particles.jl:183:
 *   + - * / // ^ max min mod mod1 atan atand add_sum hypot clamp
 * `//` (Rational) has no floating-point method in Base -> MCM_ERR_UNSUPPORTED.
 * clamp is ternary (mcm_ternary).  rem/div are src/particles.jl:354-358.      */
enum {
  MCM_B_ADD = 0, MCM_B_SUB, MCM_B_MUL, MCM_B_DIV, MCM_B_RATDIV, MCM_B_POW,
  MCM_B_MAX, MCM_B_MIN, MCM_B_MOD, MCM_B_MOD1, MCM_B_ATAN2, MCM_B_ATAND2,
  MCM_B_ADD_SUM, MCM_B_HYPOT,
  MCM_B_REM,    /* rem(x,y) == C fmod, exact  (particles.jl:354-355) */
  MCM_B_FDIV,   /* div(x,y) == round((x-rem(x,y))/y) (particles.jl:356-357) */
  /* rem(x, y, r) / div(x, y, r) with a RoundingMode (particles.jl:358,410-414; Base div.jl): RoundToZero is
   * MCM_B_REM / MCM_B_FDIV; Nearest = IEEE remainder, Down = mod(x,y), Up = mod(x,-y), FromZero = Up when the
   * signs agree else Down; div(x,y,r) = round((x - rem(x,y,r))/y).  All exact-tier (bit-exact). */
  MCM_B_REM_NEAREST, MCM_B_REM_DOWN, MCM_B_REM_UP, MCM_B_REM_FROMZERO,
  MCM_B_DIV_NEAREST, MCM_B_DIV_DOWN, MCM_B_DIV_UP, MCM_B_DIV_FROMZERO,
  MCM_B_COUNT
};

/* ---- ternary primitives */
enum {
  MCM_T_CLAMP = 0,  /* clamp(x,lo,hi), particles.jl:183            */
  MCM_T_MULADD,     /* muladd(x,y,z) — fused (particles.jl:416-422; H8: fusion allowed) */
  MCM_T_MUL_ADD,    /* x*y+z with two roundings (what `x*y+z` on particles gives) */
  MCM_T_COUNT
};

/* ---- reductions (src/particles.jl:302-323) */
enum {
  MCM_R_SUM = 0,
  MCM_R_MEAN,     /* pmean   = mean(p.particles)                 :321 */
  MCM_R_VAR,      /* pvar    = var(...; corrected=true)          :308-312 */
  MCM_R_VARP,     /* pvar(...; corrected=false)                        */
  MCM_R_STD,      /* pstd                                               */
  MCM_R_STDP,
  MCM_R_MIN,      /* pminimum */
  MCM_R_MAX,      /* pmaximum */
  MCM_R_COUNT
};

/* comparison kinds for mcm_count (the `@prob` macro, src/bymap.jl:128-136,
 * and the :montecarlo comparison mode, src/particles.jl:512-515) */
enum { MCM_CMP_LT = 0, MCM_CMP_LE, MCM_CMP_GT, MCM_CMP_GE, MCM_CMP_EQ, MCM_CMP_NE,
       /* predicates (threshold ignored): the all-reductions isinteger(p) / isfinite.(p) (particles.jl:581-585) */
       MCM_CMP_ISINTEGER, MCM_CMP_ISFINITE, MCM_CMP_ISNAN };
/* rounding modes of round(Int, p, r) / mcm_round_int (Base.RoundingMode) */
enum { MCM_RM_NEAREST = 0, MCM_RM_TOZERO, MCM_RM_UP, MCM_RM_DOWN, MCM_RM_TIESAWAY, MCM_RM_TIESUP, MCM_RM_FROMZERO };

/* ======================= lifecycle / memory ================================ */
int mcm_version(void);
/* bind the calling process to CUDA device `device`, create the library stream
 * and the stream-ordered pool.  Idempotent. */
int mcm_init(int device);
int mcm_finalize(void);
/* run all subsequent work on a caller-owned cudaStream_t (e.g. torch's); NULL restores the library stream */
int mcm_set_stream(void* cuda_stream);
void* mcm_get_stream(void);
int mcm_sync(void);
const char* mcm_last_error(void);
/* read (and optionally clear) the sticky numerical flags; synchronises */
int mcm_check_flags(uint32_t* bits, int reset);
/* clear only the sticky bits in `mask` (0x80000000 = BoundsError bit of the gathers); synchronises */
int mcm_clear_flags(uint32_t mask);
int mcm_device_info(int* sm_count, size_t* free_bytes, size_t* total_bytes);

/* storage of one `particles` vector (replaces `Vector{T}(undef,N)`; every
 * reference primitive allocates its result, register_primitive.jl:45,49,53,95) */
int mcm_alloc(void** dptr, size_t bytes);   /* stream-ordered (cudaMallocAsync pool) */
int mcm_free(void* dptr);                   /* stream-ordered; safe from a GC finalizer thread */
/* M vectors of bytes_each bytes in one call (the result of an operation on a vector of M uncertain variables) */
int mcm_alloc_many(void** dptrs, int M, size_t bytes_each);
int mcm_host_alloc(void** hptr, size_t bytes); /* pinned host staging */
int mcm_host_alloc_wc(void** hptr, size_t bytes); /* pinned, write-combined: host writes, device reads (no snoop on the DMA) */
int mcm_host_free(void* hptr);
/* Particles(v::Vector) (types.jl:55)  /  Array(p), Vector(p) (particles.jl:458-459) */
int mcm_upload(void* dst_dev, const void* src_host, size_t bytes);
int mcm_download(void* dst_host, const void* src_dev, size_t bytes); /* synchronises */
int mcm_copy(void* dst_dev, const void* src_dev, size_t bytes);

/* ======================= elementwise primitives ============================ */
/* PT{T,N}(n::Real) = fill(n,N)  (types.jl:57-63) */
int mcm_fill(int dtype, void* out, size_t n, double value);
/* convert(PT{T,N}, p::PT{S,N}) (particles.jl:333) */
int mcm_convert(int src_dtype, const void* src, int dst_dtype, void* dst, size_t n);
/* f(p)  — register_primitive.jl:94-97 */
int mcm_unary(int op, int dtype, const void* in, void* out, size_t n);
/* f(p1,p2) / f(p,a::Real) / f(a::Real,p) — register_primitive.jl:44-58.
 * A NULL pointer selects the scalar (`sa`/`sb`) for that operand. */
int mcm_binary(int op, int dtype, const void* a, double sa, const void* b, double sb,
               void* out, size_t n);
int mcm_ternary(int op, int dtype, const void* a, double sa, const void* b, double sb,
                const void* c, double sc, void* out, size_t n);
/* p^i::Integer (particles.jl:383-386); k in {-2..3} is bit-exact to literal_pow */
int mcm_powi(int dtype, const void* in, int k, void* out, size_t n);

/* ======================= fused chains ====================================== */
/* evalpoly(p, c) = c[0] + c[1]x + ... by Horner.  fused=1: muladd (FMA) per step
 * (what Base.evalpoly emits through particles.jl:416-422); fused=0: separate
 * rounding of * and + (what `c0 + x*(c1 + ...)` written with * and + gives). */
int mcm_polyval(int dtype, const void* in, const double* coeffs_host, int ncoef, int fused,
                void* out, size_t n);
/* alpha*p + beta with two roundings (bit-exact to the two reference broadcasts) */
int mcm_affine(int dtype, const void* in, double alpha, double beta, void* out, size_t n);
/* out = c0*x0 + c1*x1 + ... (left-to-right, unfused): RK4 stage / axpy updates */
int mcm_lincomb(int dtype, int nterms, const void* const* xs, const double* coeffs_host,
                void* out, size_t n);

/* Expression tape: a register machine with 16 value slots per particle; one launch evaluates the whole chain of registered
 * primitives (no temporaries in HBM — the reference materialises one Vector per
 * primitive).  Operand byte: 0..15 = slot, 0x40|k = constant k.               */
typedef struct {
  uint8_t op;   /* MCM_TAPE_* */
  uint8_t dst;  /* destination slot 0..15 (or output index for STORE) */
  uint8_t a, b, c; /* operands (or input index for LOAD; immediate for POWI in b as int8) */
  uint8_t pad[3];
} mcm_tape_ins;

enum {
  MCM_TAPE_UNARY = 0x00,   /* + MCM_U_*  : dst = f(a)            */
  MCM_TAPE_BINARY = 0x40,  /* + MCM_B_*  : dst = f(a,b)          */
  MCM_TAPE_TERNARY = 0x60, /* + MCM_T_*  : dst = f(a,b,c)        */
  MCM_TAPE_LOAD = 0x70,    /* dst = inputs[a][i]                 */
  MCM_TAPE_STORE = 0x71,   /* outputs[dst][i] = a                */
  MCM_TAPE_MOV = 0x72,     /* dst = a                            */
  MCM_TAPE_POWI = 0x73     /* dst = a ^ (int8)b                  */
};
#define MCM_TAPE_MAX_INS 96
#define MCM_TAPE_MAX_CONST 32
#define MCM_TAPE_MAX_IN 8
#define MCM_TAPE_MAX_OUT 4
#define MCM_TAPE_SLOTS 16
#define MCM_TAPE_CONST 0x40

/* Run `tape[0..loop_begin)` once, `tape[loop_begin..loop_end)` `reps` times and
 * `tape[loop_end..ntape)` once, per particle (reps>1: in-register time stepping,
 * e.g. RK4).  If stats_slot >= 0 the final value of that slot is also reduced
 * on the fly: stats_out_host[0..5] = {n, mean, M2, min, max, sum} (synchronises).
 * outputs may be NULL entries if only statistics are wanted.                  */
int mcm_chain(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
              int reps, const double* consts, int nconst,
              const void* const* inputs, int nin, void* const* outputs, int nout,
              size_t n, int stats_slot, double* stats_out_host);

/* Statistics of the most recent mcm_chain call made with stats_slot >= 0 and
 * stats_out_host == NULL (the launch is then fully asynchronous); waits for that launch only. */
int mcm_chain_stats_fetch(double* stats_out_host);
/* Deferred collection: mcm_chain_ticket() names the most recent statistics-producing launch; mcm_chain_stats_wait
 * collects the statistics of any of the last 8 tickets, waiting only for that launch (the kernel's last CTA writes
 * them into mapped pinned host memory, there is no copy to wait for) — step k's readback overlaps kernel k+1. */
uint64_t mcm_chain_ticket(void);
int mcm_chain_stats_wait(uint64_t ticket, double* stats_out_host);
/* mcm_chain on a particle-SHARDED variable (one process per GPU): same arguments; the statistics are
 * those of the GLOBAL particle set.  The last CTA stores this rank's partial into every rank's peer window (P2P
 * stores over NVLink) and merges all partials in rank order: no NCCL launch, no host round trip, bit-identical
 * results on all ranks.  COLLECTIVE when stats_slot >= 0 (same call order on every rank; needs mcm_comm_init). */
int mcm_chain_sharded(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                      int reps, const double* consts, int nconst,
                      const void* const* inputs, int nin, void* const* outputs, int nout,
                      size_t n, int stats_slot, double* stats_out_host);

/* Same as mcm_chain but inputs/outputs are HOST arrays (the `particles::Vector{T}` of an
 * ordinary reference Particles, types.jl:18): chunks are copied host->device, computed and
 * copied back on three overlapping streams.  Use mcm_host_alloc'ed (pinned) buffers for
 * full PCIe bandwidth.  Synchronises before returning. */
int mcm_chain_host(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                   int reps, const double* consts, int nconst,
                   const void* const* host_inputs, int nin, void* const* host_outputs, int nout,
                   size_t n, int stats_slot, double* stats_out_host);

/* Engine introspection: chains listed in csrc/chain_registry.def are compiled ahead of
 * time into straight-line kernels (the analogue of register_primitive's `eval`,
 * register_primitive.jl:19-82); every other tape runs on the interpreter kernel.
 * last_engine: 1 = registered kernel, 2 = interpreter.  force_interpreter(1) makes
 * registered tapes run on the interpreter too (parity tests compare the two). */
int mcm_chain_last_engine(void);
int mcm_chain_force_interpreter(int on);
/* Engine 3 — kernels GENERATED from the tape at run time (csrc/chain_jit.cu): the tape becomes straight-line CUDA
 * C++ (structural constants folded: `a^1.5` is the x*sqrt(x) kernel, integer powers unroll; branch-free fast paths
 * with a per-pack replay for out-of-range arguments), compiled for sm_100a with NVRTC and cached by tape — what
 * register_primitive's `eval` does for ANY function (register_primitive.jl:19-82), so user chains run as fast as the
 * built-in ones.  mcm_chain_last_engine() == 3 when it ran.  Results are bit-identical to the other engines.
 * mcm_chain_jit_mode: 0 = off, 1 (default) = generate when particles x instructions x reps >= 2^22 (compilation
 * costs ~1 s once per tape), 2 = always.  mcm_chain_force_engine: 0 = automatic, 1 = registered + interpreter only,
 * 2 = interpreter only, 3 = always generate.  mcm_chain_jit_compiles(): kernels compiled so far. */
int mcm_chain_jit_mode(int mode);
int mcm_chain_force_engine(int engine);
int mcm_chain_jit_compiles(void);
/* generation tunables for A/B measurements (part of the cache key): which = 0: resident CTAs per SM asked of the
 * compiler (default 3); which = 1: request the next pack's inputs one iteration ahead (default 0: measured no gain);
 * which = 2: 128-bit packs per thread, 1 / 2 / 4 or 0 = automatic (2 for short straight-line tapes, else 1);
 * which = 3: sqrt / division / inv as inline branch-free sequences guarded by the window predicate (default 1) */
int mcm_chain_jit_option(int which, int value);
/* Generate (do_compile = 0) or generate + compile the kernel of a tape WITHOUT launching it; needs no GPU.  src_out
 * (capacity src_cap) receives the generated source; *cubin_bytes the size of the sm_100a cubin; cubin_path, if not
 * NULL, a file to write it to (for cuobjdump -sass). */
int mcm_chain_jit_source(int dtype, const mcm_tape_ins* tape, int ntape, int loop_begin, int loop_end,
                         const double* consts, int nconst, int stats_slot, int do_compile, char* src_out,
                         size_t src_cap, size_t* cubin_bytes, const char* cubin_path);

/* ======================= statistics ======================================== */
/* one pass: out_host[0..5] = {n, mean, M2 = sum (x-mean)^2, min, max, sum} */
int mcm_moments(int dtype, const void* in, size_t n, double* out_host);
/* the same kernel without the host round trip: enqueue (asynchronous; the result stays in a device slot) and collect the
 * result of the LAST enqueue later */
int mcm_moments_enqueue(int dtype, const void* in, size_t n);
int mcm_moments_collect(double* out_host);
/* A/B switch for the moments kernel on large inputs: 1 (default) = bulk-async (TMA) shared-memory
 * pipeline, 0 = direct 128-bit global loads.  Results are identical up to merge order. */
int mcm_moments_use_bulk(int on);
/* pmean/pvar/pstd/pminimum/pmaximum/sum (particles.jl:302-323); N==1 -> var/std 0 */
int mcm_reduce(int kind, int dtype, const void* in, size_t n, double* out_host);
/* @prob p < a  -> count; (bymap.jl:128-136) */
int mcm_count(int cmp, int dtype, const void* in, size_t n, double threshold, uint64_t* count_host);
/* count of particles with a[i] cmp b[i] (cmp in MCM_CMP_LT..MCM_CMP_NE): Base.:(==)(p1, p2) = p1.particles ==
 * p2.particles (particles.jl:488) and the :montecarlo comparison mode (particles.jl:512-515) without a temporary */
int mcm_count2(int cmp, int dtype, const void* a, const void* b, size_t n, uint64_t* count_host);
/* round(Int, p[, r]) (particles.jl:347-348; runtests.jl:471-473): out[i] = Int64(round(in[i], r)).  A NaN, an
 * infinity or a value outside Int64 raises MCM_FLAG_DOMAIN (InexactError in Base) and stores INT64_MIN. */
int mcm_round_int(int dtype, const void* in, int mode, int64_t* out, size_t n);
/* merge two {n,mean,M2,min,max,sum} partials (Chan) — used across GPUs */
int mcm_moments_merge(const double* a, const double* b, double* out);

/* pcov(v::MvParticles; corrected) = cov(Matrix(v)) (particles.jl:476, 461-465).
 * vars: M device pointers (host array), each N elements.  cov_out: device, M x M
 * doubles, column-major == row-major (symmetric).  means_out: device M doubles or NULL.
 * scale: 1 -> divide by (N-1) [corrected], 0 -> by N, -1 -> raw centred Gram (for
 * multi-GPU partial Gram blocks). center_host: NULL -> centre on the local column
 * means; else M doubles to centre on (global means in the sharded case).      */
int mcm_gram(int dtype, const void* const* vars, int M, size_t n, int scale,
             const double* center_host, double* cov_out_dev, double* means_out_dev);
/* Gram engine for large problems (M >= 256, N >= 4096): 1 (default) = tcgen05/TMEM int8 tensor cores
 * with an Ozaki error-free split of the centred Float64 data (exact int32 accumulation, 48 bits kept;
 * rows with max/sigma > 10 make it decline and the FP64 kernel runs instead), 0 = always the FP64
 * tensor pipe (DMMA), 2 = tcgen05 with the A operand staged in TMEM (tcgen05.cp + TS-form MMAs) instead of
 * read from shared memory (A/B runs; bit-identical results).  All meet the 1e-12 tolerance; small problems always use DMMA.
 * mcm_gram_last_engine: 0 = DMMA, 1 = tcgen05 for the most recent mcm_gram call. */
int mcm_gram_use_tensor(int on);
int mcm_gram_last_engine(void);
/* digits per value (6, 7 or 8 signed 8-bit digits) the last tcgen05 Gram split its rows into: chosen from the
 * heaviest row's max/sigma so that heavy-tailed data (lognormal, t, exp(p)) stays on the tensor cores at 1e-12 */
int mcm_gram_last_digits(void);
/* column means of M variables into means_out_host[M] */
int mcm_colmeans(int dtype, const void* const* vars, int M, size_t n, double* means_out_host);

/* ======================= order statistics ================================== */
/* sort(p.particles) ; sortperm -> 0-based int64 permutation (stable) */
int mcm_sort(int dtype, const void* in, void* out, size_t n);
int mcm_sortperm(int dtype, const void* in, int64_t* perm_out, void* sorted_out_or_null, size_t n);
/* pquantile(p, q) = Statistics.quantile(p.particles, q) type-7 (particles.jl:321);
 * q_host: nq probabilities in [0,1]; out_host: nq values.  pmedian == q=0.5. */
int mcm_quantile(int dtype, const void* in, size_t n, const double* q_host, int nq, double* out_host);
/* A/B switch: 1 (default) = mcm_sort / mcm_sortperm read the particles in their first pass and write floating-point
 * values / Int64 indices in their last (no separate key-transform, iota and widening passes); 0 = separate passes. */
int mcm_sort_use_fused(int on);
/* A/B switch: 1 (default) = sampled brackets + one streaming read for n >= 131072 and <= 8 probabilities
 * (select_batch.cu), MSD radix select otherwise; 2 = MSD radix select always (select.cu, one 8 B/key read
 * per pass); 0 = full sort + index.  All return the exact order statistics. */
int mcm_quantile_use_select(int on);

/* pquantile of M variables of equal length in one call (particles.jl:316-323 per element of a
 * Vector{Particles}): out_host[m*nq + i] = quantile(vars[m], q[i]).  Sampled pivots bracket every rank,
 * one streaming read counts/collects, radix select on the candidates; one device->host copy in total
 * (select_batch.cu).  M <= 16383. */
int mcm_quantile_batch(int dtype, const void* const* vars, int M, size_t n, const double* q_host, int nq,
                       double* out_host);
/* exact order statistics: out_host[r] = sort(in)[ranks_host[r]] (0-based ranks) without sorting — what
 * Statistics.median (central one / two order statistics, particles.jl:319) and partialsort need */
int mcm_order_statistics(int dtype, const void* in, size_t n, const uint64_t* ranks_host, int nr, double* out_host);
/* diagnostics: number of variables the batched path has handed to the general path so far
 * (heavy ties around a rank, or a rank outside its sampled bracket) */
int mcm_quantile_batch_fallbacks(void);

/* ======================= constant matrix x particle vector ================= */
/* outs[j][i] = bias[j] + sum_k A[j*Mi + k] * (xs[k][i] - shift[k])  for j < Mo, k < Mi, i < n.
 * Replaces the `Matrix(b) * A'` GEMM of mul!(y, A, b::Vector{Particles}) (particles.jl:688-707),
 * `A * p` (_pgemv, :620-631) and the transform_moments application (sigmapoints.jl:71).
 * A_host: row-major Mo x Mi Float64 on the host; shift_host / bias_host may be NULL (zeros).
 * xs / outs: host arrays of device pointers; outputs must not alias inputs. */
int mcm_pgemm(int dtype, const double* A_host, int Mo, int Mi, const void* const* xs, const double* shift_host,
              const double* bias_host, void* const* outs, size_t n);
/* A/B switch: 1 (default) = tcgen05/TMEM int8 path (7-digit Ozaki split, exact int32 accumulation) for
 * Mo >= 64, Mi >= 128, n >= 8192 with automatic fallback; 0 = FP64 DMMA kernel always.
 * mcm_pgemm_last_engine: 0 = DMMA, 1 = tcgen05. */
int mcm_pgemm_use_tensor(int on);
int mcm_pgemm_last_engine(void);

/* ======================= M x M dense steps + transform_moments ============= */
/* In-place lower Cholesky factor of a symmetric positive definite M x M matrix in device memory (row-major,
 * lower triangle referenced, strict upper triangle zeroed) — `cholesky(C).L` of sigmapoints.jl:38,69.
 * *info_host = 0, or the 1-based column of the first non-positive pivot (then MCM_ERR_INVALID). */
int mcm_potrf(double* a_dev, int M, int* info_host);
/* X L = B solved in place in b_dev (nrows x M, row-major) for lower-triangular L (M x M): the `/` of
 * `cholesky(S).L / xl` (sigmapoints.jl:71). */
int mcm_trsm_right_lower(const double* l_dev, double* b_dev, int M, int nrows);
/* A/B switch for the dense M x M steps (mcm_potrf, mcm_trsm_right_lower, mcm_transform_moments): 2 (default) = as 1 with the
 * diagonal block factored in panels of 16 columns (one warp, shuffles; bit-identical to 1); 1 = register-tiled diagonal
 * block, right-looking solve, chol(S) on a side stream under the Gram; 0 = the first versions.  Same results to rounding
 * (scripts/check_dense_v2.py; all are run by tests/test_gpu_c5.py). */
int mcm_dense_use_v2(int on);
/* transform_moments(X, m, S; preserve_latin) (sigmapoints.jl:63-72) entirely on device:
 * outs[j] = m[j] + sum_k A[j,k] (xs[k] - mean(xs[k])),  A = cholesky(S).L / xl,
 * xl = cholesky(cov(X)).L  or  Diagonal(std(X)) when preserve_latin != 0.  S_host: row-major M x M. */
int mcm_transform_moments(int dtype, const void* const* xs, int M, size_t n, const double* m_host,
                          const double* S_host, int preserve_latin, void* const* outs);

/* ======================= sampling / resampling ============================= */
/* quantile(Normal(0,1), p) elementwise on device */
int mcm_norminvcdf(int dtype, const void* p_in, void* out, size_t n);
/* systematic_sample(N, Normal(mu,sigma); permute) (sampling.jl:6-13):
 * out[j] = mu + sigma*Phi^-1((2*i+1)/(2*n_total)) for global index i = i0 + j (or
 * i = perm[j] when perm != NULL; perm is a 0-based device int64 array of length n). */
int mcm_systematic_normal(int dtype, void* out, size_t n, double mu, double sigma,
                          const int64_t* perm_dev, uint64_t i0, uint64_t n_total);
/* a..b  Uniform systematic sample: out = a + (b-a)*(2i+1)/(2 n_total) */
int mcm_systematic_uniform(int dtype, void* out, size_t n, double a, double b,
                           const int64_t* perm_dev, uint64_t i0, uint64_t n_total);
/* ---- any continuous univariate Distribution (sampling.jl:10 `quantile.((d,), y)`; types.jl:67-90; `a ⊠ d`, `a ⊞ d`
 * particles.jl:42,50).  Families and parameters (a, b, c) follow Distributions.jl:
 *   NORMAL(mu, sigma) UNIFORM(a, b) EXPONENTIAL(theta) LOGNORMAL(mu, sigma) CAUCHY(mu, sigma) LAPLACE(mu, theta)
 *   LOGISTIC(mu, theta) RAYLEIGH(sigma) WEIBULL(alpha, theta) PARETO(alpha, theta) GUMBEL(mu, theta)
 *   TRIANGULAR(a, b, c) FRECHET(alpha, theta) KUMARASWAMY(a, b) GAMMA(alpha, theta) CHISQ(nu) INVERSEGAMMA(alpha, theta)
 *   BETA(alpha, beta) TDIST(nu) FDIST(nu1, nu2)
 * Closed forms where they exist; Halley iterations on the regularised incomplete gamma / beta functions otherwise
 * (csrc/mcm_dist.cuh), solved on the smaller of p and 1 - p: <= 1e-12 relative in both tails (tolerance tier).
 * Anything else goes through a host-supplied quantile table (host mirror: GenericDistribution). */
enum mcm_dist_kind {
  MCM_D_NORMAL = 0, MCM_D_UNIFORM, MCM_D_EXPONENTIAL, MCM_D_LOGNORMAL, MCM_D_CAUCHY, MCM_D_LAPLACE, MCM_D_LOGISTIC,
  MCM_D_RAYLEIGH, MCM_D_WEIBULL, MCM_D_PARETO, MCM_D_GUMBEL, MCM_D_TRIANGULAR, MCM_D_FRECHET, MCM_D_KUMARASWAMY,
  MCM_D_GAMMA, MCM_D_CHISQ, MCM_D_INVERSEGAMMA, MCM_D_BETA, MCM_D_TDIST, MCM_D_FDIST, MCM_D_COUNT
};
/* systematic_sample(N, d; permute): out[j] = quantile(d, (2i+1)/(2 n_total)), i = i0 + j or perm[j] */
int mcm_systematic_dist(int dtype, int kind, double a, double b, double c, void* out, size_t n,
                        const int64_t* perm_dev, uint64_t i0, uint64_t n_total);
/* iid draws by inversion: out[j] = quantile(d, u_j), u_j the Philox uniform of mcm_rand_real — the device stand-in
 * for rand(rng, d, N) (types.jl:71,79, systematic = false) */
int mcm_rand_dist(int dtype, int kind, double a, double b, double c, void* out, size_t n, uint64_t seed, uint64_t offset);
/* quantile.(d, p) for a device vector of probabilities */
int mcm_quantile_dist(int dtype, int kind, double a, double b, double c, const void* p_in, void* out, size_t n);
/* randperm(rng, N) stand-in: device permutation from a counter-based generator
 * (Philox-4x32-10 keys + stable radix sort).  Stream differs from Julia's RNG (H5). */
int mcm_randperm(int64_t* perm_out, size_t n, uint64_t seed);
/* Keyed pseudo-random permutation of [0, n_total) evaluated per index (4-round Feistel network +
 * cycle walking; no sort, no storage): out[j] = pi_{seed,column}(i0 + j).  The communication-free
 * replacement for randperm at sampling.jl:11 when the sample is large or sharded. */
int mcm_index_bijection(int64_t* out, size_t n, uint64_t seed, uint64_t column, uint64_t i0, uint64_t n_total);
/* M systematic columns in ONE launch (the multivariate systematic construction, docs index.md:141 +
 * types.jl:67-90 per column): outs[v][j] = quantile(d_v, grid[pi_v(i0 + j)]), d_v = Normal(a[v], b[v])
 * (kind 0) or Uniform(a[v], b[v]) (kind 1); a_host/b_host NULL = standard parameters; permute = 0
 * keeps the grid order, 1 applies the keyed bijection of column v. */
int mcm_systematic_batch(int dtype, int kind, void* const* outs, int M, size_t n, const double* a_host,
                         const double* b_host, int permute, uint64_t seed, uint64_t i0, uint64_t n_total);
/* A/B switch: 1 (default) = whole permuted columns of <= 32 MB are written in scatter form (grid order in,
 * pi^-1 out: branch-coherent Phi^-1), 0 = always the gather form.  Identical results. */
int mcm_systematic_use_scatter(int on);
/* iid draws: out[j] = a + b*Phi^-1(u_j) (kind 0, Normal(a,b)) or a + (b-a)*u_j (kind 1, Uniform(a,b)) with
 * u_j = (top 53 bits of Philox(seed, offset+j) + 1/2) 2^-53 — the device stand-in for rand(rng, d, N)
 * (types.jl:71,79,93,97; optimize.jl:5).  Julia's stream is not reproducible (H5). */
int mcm_rand_real(int dtype, int kind, void* out, size_t n, double a, double b, uint64_t seed, uint64_t offset);
/* StatsBase.autocor(p.particles, 1:maxlag) (demean = true): out_host[k-1] = lag-k autocorrelation, maxlag <= 256
 * — the reduction behind ess (sampling.jl:26-38). */
int mcm_autocor(int dtype, const void* in, size_t n, int maxlag, double* out_host);
/* ---- sharded bootstrap over peer memory (resampling.jl:76-84 on particle-sharded variables) -----
 * A window is peer-visible device memory (plain cudaMalloc + CUDA IPC, handles exchanged through the library's
 * communicator).  mcm_window_create / mcm_window_free are COLLECTIVE (every rank, same order, same size).
 * mcm_comm_fence is a stream-ordered fence across ranks: work enqueued after it on any rank runs after the work
 * enqueued before it on every rank (no-op without a communicator). */
int mcm_window_create(size_t bytes, int* id_out, void** local_out);
int mcm_window_free(int id);
int mcm_comm_fence(void);
/* Window layout: variables interleaved in groups of four, so the four values one remote read needs share a 32-byte
 * sector (NVLink moves sectors).  mcm_window_bytes = size to create for M variables of `per` elements per rank;
 * mcm_window_fill interleaves this rank's slices vars[v][0..n_local) into its window. */
size_t mcm_window_bytes(int M, uint64_t per, int dtype);
int mcm_window_fill(int dtype, int window, const void* const* vars, int M, size_t n_local, uint64_t per);
/* outs[v][j] = variable v of GLOBAL particle idx[j] (0-based, < n_global), loaded straight from the HBM of its owner
 * idx[j] / per over NVLink.  Bracket with mcm_comm_fence (after the fills, and before the next fill). */
int mcm_gather_peer(int dtype, int window, int M, uint64_t per, uint64_t n_global, const int64_t* idx, size_t n_out,
                    void* const* outs);
/* Systematic resampling of a particle-sharded weight vector: this rank holds w[global_lo ..
 * global_lo + n_local) of n_global weights.  Same rule and the same fixed-point scale (from the GLOBAL maximum and
 * n_global) as mcm_resample_systematic, so the indices equal the single-process ones.  2 doubles + 1 integer per rank
 * are all-gathered; then every source writes its run of output positions (GLOBAL source indices, int64) into the
 * index window of the rank that owns them: output k lives at slot k % per_out of rank k / per_out's window.
 * Collective.  Follow with mcm_comm_fence() before reading the window. */
int mcm_resample_systematic_sharded(const double* w_local, size_t n_local, uint64_t n_global, uint64_t global_lo,
                                    uint64_t n_out, double u, int window, uint64_t per_out);
/* ---- decision layer leftovers ----------------------------------------------------------
 * argmin(abs.(p.particles .- val)) (nominal.jl:21): 0-based FIRST index of the minimum; a NaN distance beats every
 * number, as in Julia's findmin. */
int mcm_argmin_absdiff(int dtype, const void* in, size_t n, double val, uint64_t* idx_host);
/* argmin_i sum_j (x_ij - val_j)^2 over M variables, row sums accumulated left to right (the joint nearest
 * particle; nominal.jl:26 sums over dims=1 instead — see with_nominal in the host mirror). */
int mcm_argmin_sqdist(int dtype, const void* const* vars, int M, size_t n, const double* val_host, uint64_t* idx_host);
/* shuffle_and_insert (nominal.jl:1-6): out = copy(in); out[ind] = in[0]; out[0] = val.  in may equal out. */
int mcm_shuffle_insert(int dtype, const void* in, size_t n, uint64_t ind, double val, void* out);
/* filter(x -> lo <= x <= hi, p.particles) (particles.jl:375-380, Base.intersect), order preserved: survivors
 * are written to out[out_offset ...]; *count_host = how many.  out needs room for out_offset + n elements. */
int mcm_filter_range(int dtype, const void* in, size_t n, double lo, double hi, void* out, size_t out_offset,
                     size_t* count_host);
/* ---- per-particle small-matrix operations ---------------------------------------------------
 * Device replacement for the host loops of R^n2R^n_function / R^n2C^n_function (particles.jl:241-282), eigvals
 * (:395-406), exp / exp! / lyap (:603-607) and `\` on particle matrices: entry (r,c) of the uncertain nmat x nmat
 * matrix is the length-n device vector A[r*nmat + c]; particle i of every entry forms matrix i.  One thread per
 * particle.  nmat <= 16 (lyap: <= 8); up to 8 the matrix is held in registers, above in per-thread local memory.  Results agree with LAPACK to rounding (tolerance tier), not bit for bit.
 *   SOLVE        out[r*m + c] = (A \ B)[r,c], B = nmat x m table (LU, partial pivoting)
 *   INV, EXP     out[r*nmat + c]; EXP is Higham's scaling and squaring as LinearAlgebra.exp!, without balancing
 *   DET          out[0]
 *   EIGVALS      out[k] + i out_im[k], k < nmat, ordered by (real, imaginary) — LAPACK's order is not reproducible
 *   EIGVALS_SYM  out[k] ascending (the upper triangle is read)
 *   LYAP         out = X with A X + X A' + C = 0, C = B (nmat x nmat table)
 *   SVDVALS      out[k] descending
 * A singular system or a non-converged iteration raises MCM_FLAG_DOMAIN (LAPACKException in the reference). */
enum mcm_la_op {
  MCM_LA_SOLVE = 0, MCM_LA_INV = 1, MCM_LA_DET = 2, MCM_LA_EXP = 3, MCM_LA_EIGVALS = 4, MCM_LA_EIGVALS_SYM = 5,
  MCM_LA_LYAP = 6, MCM_LA_SVDVALS = 7
};
int mcm_batched_linalg(int op, int dtype, int nmat, int m, const void* const* A, const void* const* B,
                       void* const* out, void* const* out_im, size_t n);
/* The same for matrices with Complex{Particles} entries (complex.jl:104-181: exp, exp!, log through ℂⁿ2ℂⁿ_function; det
 * through ℂⁿ2ℂ_function; eigvals; svdvals) and for log of a REAL uncertain matrix, whose result is complex
 * (particles.jl:603, Base.log(::Matrix{<:AbstractParticles}) = ℝⁿ2ℂⁿ_function(log, p)): A_im == NULL.  Tables are
 * row-major nmat x nmat (inputs; EXP / LOG outputs), one entry (DET) or nmat entries (EIGVALS ordered by (re, im);
 * SVDVALS descending, real: out_im unused).  Algorithms: csrc/mcm_small_linalg_c.cuh (LU, Higham expm, complex Schur +
 * inverse scaling and squaring log, one-sided Jacobi).  A singular matrix under LOG or a QR iteration that does not
 * converge stores NaN and raises MCM_FLAG_DOMAIN. */
enum mcm_lac_opcode { MCM_LAC_EXP = 0, MCM_LAC_LOG, MCM_LAC_DET, MCM_LAC_EIGVALS, MCM_LAC_SVDVALS, MCM_LAC_COUNT };
int mcm_batched_linalg_complex(int op, int dtype, int nmat, const void* const* A_re, const void* const* A_im,
                               void* const* out_re, void* const* out_im, size_t n);
/* ---- Complex{Particles} (src/complex.jl:1-93) ----------------------------------------------------------------
 * A complex uncertain number is two device vectors (re, im).  One streaming kernel per operation replaces the host
 * loop of C2C_function (complex.jl:30-67).  Operand components: a null pointer selects the constant beside it, so
 * real particles (im = NULL, 0.0), real and complex scalars are all expressible.  + - * follow Base bit for bit
 * ((ac - bd) + (ad + bc)i, no contraction); / ^ and the functions are tolerance tier (1e-13 relative). */
enum mcm_complex_opcode {
  MCM_C_ADD = 0, MCM_C_SUB = 1, MCM_C_MUL = 2, MCM_C_DIV = 3, MCM_C_POW = 4,
  MCM_C_EXP = 16, MCM_C_LOG = 17, MCM_C_SQRT = 18, MCM_C_SIN = 19, MCM_C_COS = 20, MCM_C_TAN = 21,
  MCM_C_EXP10 = 22, MCM_C_LOG10 = 23
};
int mcm_complex_op(int op, int dtype, const void* a_re, const void* a_im, double sa_re, double sa_im,
                   const void* b_re, const void* b_im, double sb_re, double sb_im, void* out_re, void* out_im, size_t n);
/* [rand(rng,1:N) for _ in 1:n] stand-in (resampling.jl:77,81): 0-based indices in [0,N) */
int mcm_rand_indices(int64_t* idx_out, size_t n_draw, uint64_t N, uint64_t seed, uint64_t offset);
/* p.particles[inds] (resampling.jl:77,82): out[j] = src[idx[j]] */
int mcm_gather(int dtype, const void* src, const int64_t* idx, void* out, size_t n_draw, size_t n_src);
/* same index vector applied to M variables (resampling.jl:80-84) */
int mcm_gather_multi(int dtype, const void* const* srcs, void* const* outs, int M,
                     const int64_t* idx, size_t n_draw, size_t n_src);
/* systematic resampling by prefix scan (restates the dead reference code at
 * resampling.jl:45-67): weights w (f64, >=0), offset u in [0,1):
 * idx_out[i] = smallest b with cumsum(w)[b] > (i+u)*sum(w)/n.  Exact fixed-point scan. */
int mcm_resample_systematic(const double* w, size_t n, double u, int64_t* idx_out);
/* A/B switch: 0 (default) = expansion (each source writes its run of outputs; one exact 128-bit division per
 * source), 1 = one binary search per output over the scan (first version).  Identical indices. */
int mcm_resample_use_search(int on);

/* ======================= multi-GPU (one process per GPU) ==================== */
/* NCCL is dlopen'ed; the 128-byte unique id is created on rank 0 and distributed
 * by the host program (torch.distributed / MPI / sockets). */
int mcm_comm_unique_id(void* id128);
int mcm_comm_init(const void* id128, int nranks, int rank);
int mcm_comm_destroy(void);
int mcm_comm_info(int* nranks, int* rank);
/* in-place sum all-reduce of `count` doubles on the library stream */
int mcm_comm_allreduce_sum_f64(double* dptr, size_t count);
/* sharded statistics: every rank passes its local shard; all ranks get the global
 * {n, mean, M2, min, max, sum} (all-gather of partials + fixed-order Chan merge) */
int mcm_moments_sharded(int dtype, const void* in, size_t n_local, double* out_host);
/* merge host-resident per-rank moments (e.g. mcm_chain's fused statistics) across ranks, in place */
int mcm_moments_allmerge(double* inout_host6);
/* sharded pquantile (particles.jl:316-323 over a particle index spread across GPUs): every rank passes its
 * shard (balanced to 1 %, >= 8192 particles, <= 64 ranks); all ranks return the quantiles of the
 * GLOBAL particle set, out_host[m*nq + i].  Sampled pivots (all-gather of 8192 sample keys per variable),
 * local streaming count/collect, radix select on all-reduced histograms; only the <= 2048 finalists per rank
 * group ever leave a GPU.  With a single rank this is mcm_quantile_batch. */
int mcm_quantile_sharded(int dtype, const void* const* vars, int M, size_t n_local, const double* q_host, int nq,
                         double* out_host);
/* sharded pcov: global means by all-reduce, local centred Gram, all-reduce of the
 * M x M block, scaled by 1/(N_global - corrected) */
int mcm_gram_sharded(int dtype, const void* const* vars, int M, size_t n_local, int corrected,
                     double* cov_out_dev);

#ifdef __cplusplus
}
#endif
#endif /* MCM_B200_H */
