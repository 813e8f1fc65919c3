// ref_cpu.cpp — "reference-shaped" CPU baseline (tier B1/B2).
//
// TEST / BENCH INFRASTRUCTURE ONLY (see oracle/mcm_oracle.c header).  The real
// reference (Julia + SLEEFPirates) cannot run here, so this stand-in mirrors how
// it executes the hot path: ONE SIMD pass per registered primitive, every
// temporary materialised in memory (src/register_primitive.jl:44-58,94-97), SLEEF
// kernels for sin/exp (src/sleefpirates.jl:7-16,26-35; SLEEF 3.8 from libtorch_cpu:
// sin u35 == the "_fast" class, exp u10), Base-class pow (SLEEF pow u10), then
// mean (one pass) and std (two passes) (src/particles.jl:309-322).
// `threads` = 1 reproduces the reference's single-threaded execution; more
// threads give the "all host cores" line.
#include <immintrin.h>
#include <math.h>
#include <thread>
#include <vector>
#include <functional>
#include <stddef.h>
#include <sleef.h>

#if defined(__AVX512F__)
typedef __m512d vd;
#define VL 8
#define vload _mm512_loadu_pd
#define vstore _mm512_storeu_pd
#define vsin Sleef_sind8_u35
#define vexp Sleef_expd8_u10
#define vpow Sleef_powd8_u10
#define vmul _mm512_mul_pd
#define vadd _mm512_add_pd
#define vdiv _mm512_div_pd
#define vsqrt _mm512_sqrt_pd
#define vset1 _mm512_set1_pd
static inline vd vneg(vd x) { return _mm512_sub_pd(_mm512_setzero_pd(), x); }
static inline vd vabs(vd x) { return _mm512_abs_pd(x); }
#else
typedef __m256d vd;
#define VL 4
#define vload _mm256_loadu_pd
#define vstore _mm256_storeu_pd
#define vsin Sleef_sind4_u35
#define vexp Sleef_expd4_u10
#define vpow Sleef_powd4_u10
#define vmul _mm256_mul_pd
#define vadd _mm256_add_pd
#define vdiv _mm256_div_pd
#define vsqrt _mm256_sqrt_pd
#define vset1 _mm256_set1_pd
static inline vd vneg(vd x) { return _mm256_sub_pd(_mm256_setzero_pd(), x); }
static inline vd vabs(vd x) { return _mm256_andnot_pd(_mm256_set1_pd(-0.0), x); }
#endif

// Static partition of [0, n) over a PERSISTENT pool of worker threads (libgomp's spec file is not in this image).
// The first version spawned and joined `threads` std::threads per pass — 13 passes per chain10 evaluation; the pool
// keeps the workers parked on a condition variable between passes, so a pass costs two notifications.
#include <condition_variable>
#include <mutex>
class Pool {
 public:
  explicit Pool(int n) : n_(n) {
    for (int t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
  }
  ~Pool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return n_; }
  void run(const std::function<void(int)>& job) {      // job(t) for t in [0, n); the caller is worker 0
    { std::lock_guard<std::mutex> lk(mu_); job_ = &job; pending_ = n_ - 1; ++gen_; }
    cv_.notify_all();
    job(0);
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }
 private:
  void loop(int t) {
    unsigned long seen = 0;
    for (;;) {
      const std::function<void(int)>* job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        job = job_;
      }
      (*job)(t);
      { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) done_.notify_one(); }
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  int pending_ = 0;
  unsigned long gen_ = 0;
  bool stop_ = false;
};
static Pool* g_pool = nullptr;
static Pool& pool(int threads) {
  if (!g_pool || g_pool->size() != threads) { delete g_pool; g_pool = new Pool(threads); }
  return *g_pool;
}

template <class F>
static void parallel_for(long n, int threads, F&& body) {
  if (threads <= 1) { body(0, n, 0); return; }
  const long chunk = (n + threads - 1) / threads;
  const std::function<void(int)> job = [&](int t) {
    const long b = t * chunk, e = b + chunk < n ? b + chunk : n;
    if (b < e) body(b, e, t);
  };
  pool(threads).run(job);
}

#define PASS1(dst, src, expr)                                      \
  parallel_for(nv, threads, [&](long b_, long e_, int) {           \
    for (long i = b_; i < e_; ++i) {                               \
      vd x = vload(src + i * VL);                                  \
      vstore(dst + i * VL, expr);                                  \
    }                                                              \
  });
#define PASS2(dst, s1, s2, expr)                                   \
  parallel_for(nv, threads, [&](long b_, long e_, int) {           \
    for (long i = b_; i < e_; ++i) {                               \
      vd x = vload(s1 + i * VL);                                   \
      vd y = vload(s2 + i * VL);                                   \
      vstore(dst + i * VL, expr);                                  \
    }                                                              \
  });

extern "C" {

int ref_vector_width(void) { return VL; }
int ref_max_threads(void) { unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }

// chain10 of SURVEY.md §8(d): t1=sin(a); t2=a*a; t3=t1*t2; t4=-a; t5=exp(t4); t6=t3+t5;
// t7=abs(t6); t8=sqrt(t7); t9=a^1.5; y=t8/t9; then pmean(y), pstd(y).
// work: 4 scratch vectors of n doubles (t1..t4 reused as the reference's GC would).
// n must be a multiple of the vector width (callers pad).
int ref_chain10(const double* a, double* yout, double* w1, double* w2, double* w3, long n, int threads,
                double* mean_out, double* std_out) {
  const long nv = n / VL;
  const vd c15 = vset1(1.5);
  PASS1(w1, a, vsin(x))            // t1 = sin.(a)
  PASS1(w2, a, vmul(x, x))         // t2 = a.*a       (literal_pow 2)
  PASS2(w3, w1, w2, vmul(x, y))    // t3 = t1.*t2
  PASS1(w1, a, vneg(x))            // t4 = -a
  PASS1(w2, w1, vexp(x))           // t5 = exp.(t4)
  PASS2(w1, w3, w2, vadd(x, y))    // t6 = t3 + t5
  PASS1(w2, w1, vabs(x))           // t7 = abs.(t6)
  PASS1(w3, w2, vsqrt(x))          // t8 = sqrt.(t7)
  PASS1(w1, a, vpow(x, c15))       // t9 = a.^1.5
  PASS2(yout, w3, w1, vdiv(x, y))     // y  = t8./t9
  // pmean: one pass; pstd: mean pass + squared-deviation pass (Statistics.std)
  std::vector<double> part(threads > 0 ? threads : 1, 0.0);
  auto psum = [&](auto f) {
    for (auto& p : part) p = 0.0;
    parallel_for(n, threads, [&](long b_, long e_, int t) {
      double acc = 0.0;
      for (long i = b_; i < e_; ++i) acc += f(yout[i]);
      part[t] = acc;
    });
    double tot = 0.0;
    for (double p : part) tot += p;
    return tot;
  };
  const double mean = psum([](double v) { return v; }) / (double)n;                 // pmean
  const double mean2 = psum([](double v) { return v; }) / (double)n;                // std recomputes the mean
  const double s2 = psum([mean2](double v) { const double d = v - mean2; return d * d; });
  *mean_out = mean;
  *std_out = sqrt(s2 / (double)(n - 1));
  return 0;
}

}  // extern "C"
