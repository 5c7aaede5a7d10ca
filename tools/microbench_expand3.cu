// Round-2 study: a FLAT block expansion.  pk_expand_cols walks block columns (one thread = one
// (interval, column), `rows` scalar 8-byte stores 8*n bytes apart).  Here a thread owns a 32-byte-
// aligned group of 4 consecutive output slots of a list and issues ONE 256-bit store
// (st.global.v4.f64 -> STG.E.ENL2.256 on sm_100a); slot -> (interval, row, column) is index
// arithmetic, the list value / unit entry / multiplier / width come through L1 (or shared memory).
// Sector alignment then no longer depends on the block shape (LGL n = 10: 90 slots per block).
// Same synthetic problem as microbench_expand.cu, plus the LGL shape; outputs are compared with the
// column walk bit for bit.
//   nvcc -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o tools/_bin/microbench_expand3 tools/microbench_expand3.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Geo {
  int n, rows, step, lists;
  unsigned nK;        // intervals
  long long Lm;       // nodes per W row
  long long list_stride;  // slots between list bases (>= nK * n * rows), lets bases be misaligned
  int mis;            // list l starts (1 + (l & 3)) * mis doubles later
};

__device__ __forceinline__ long long list_base(const Geo& g, int l) { return (long long)l * g.list_stride + (g.mis ? 1 + (l & 3) : 0); }

// ---- baseline: the parameter-driven column walk (pk_expand_cols)
template <int ST>
__device__ __forceinline__ void store(double* p, double v) {
  if (ST == 1) __stcs(p, v);        // streaming (evict-first) store
  else if (ST == 2) __stwt(p, v);   // write-through
  else if (ST == 3) __stcg(p, v);   // cache at L2 only
  else *p = v;
}

// variants of the column walk: block size, unroll depth, store cache operator
template <bool LAMF, int THREADS, int UNROLL, int ST>
__global__ void __launch_bounds__(THREADS) cols_v(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                                  const double* __restrict__ unit, const double* __restrict__ width, Geo g) {
  extern __shared__ double sm[];
  const int n = g.n, rows = g.rows, bn = n * rows;
  double* u_s = sm;
  double* lam_s = sm + bn;
  const unsigned pairs = g.nK * n;
  const unsigned t0 = blockIdx.x * THREADS;
  if (t0 >= pairs) return;
  const unsigned t = t0 + threadIdx.x;
  const bool live = t < pairs;
  const unsigned tt = live ? t : pairs - 1;
  const unsigned K = tt / n, cc = tt - K * n, K0 = t0 / n;
  const int l = blockIdx.y;
  const double sv = W[(size_t)l * g.Lm + 1 + (size_t)K * g.step + cc];
  const double w = width[K];
  for (int q = threadIdx.x; q < bn; q += THREADS) u_s[q] = -1.0 * unit[q];
  if (LAMF) {
    unsigned tl = t0 + THREADS - 1;
    if (tl >= pairs) tl = pairs - 1;
    const int n_lam = (int)(tl / n - K0 + 1) * rows;
    for (int q = threadIdx.x; q < n_lam; q += THREADS) lam_s[q] = LAM[(size_t)K0 * rows + q];
  }
  __syncthreads();
  if (!live) return;
  double* __restrict__ out = out_all + list_base(g, l) + (size_t)K * bn + cc;
  const double* u = u_s + cc;
  const double* lm = lam_s + (K - K0) * rows;
#pragma unroll UNROLL
  for (int r = 0; r < rows; ++r) {
    double v = (u[r * n] * w) / 2.0;
    if (LAMF) v = v * lm[r];
    store<ST>(out + r * n, v * sv);
  }
}

// column walk with the rows of a block split over RS blocks (blockIdx.z): shorter blocks -> finer tail,
// MINB = minimum resident blocks per SM asked of the compiler (register cap)
template <bool LAMF, int THREADS, int RS, int MINB, int ST>
__global__ void __launch_bounds__(THREADS, MINB) cols_split(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                                            const double* __restrict__ unit, const double* __restrict__ width, Geo g) {
  extern __shared__ double sm[];
  const int n = g.n, rows = g.rows, bn = n * rows;
  const int r0 = (rows * (int)blockIdx.z) / RS, r1 = (rows * ((int)blockIdx.z + 1)) / RS;
  const int nr = r1 - r0;
  double* u_s = sm;                 // [nr * n] rows r0..r1 of the sign-folded unit block
  double* lam_s = sm + nr * n;      // [(THREADS / n + 2) * nr]
  const unsigned pairs = g.nK * n;
  const unsigned t0 = blockIdx.x * THREADS;
  if (t0 >= pairs) return;
  const unsigned t = t0 + threadIdx.x;
  const bool live = t < pairs;
  const unsigned tt = live ? t : pairs - 1;
  const unsigned K = tt / n, cc = tt - K * n, K0 = t0 / n;
  const int l = blockIdx.y;
  const double sv = W[(size_t)l * g.Lm + 1 + (size_t)K * g.step + cc];
  const double w = width[K];
  for (int q = threadIdx.x; q < nr * n; q += THREADS) u_s[q] = -1.0 * unit[r0 * n + q];
  if (LAMF) {
    unsigned tl = t0 + THREADS - 1;
    if (tl >= pairs) tl = pairs - 1;
    const int nKb = (int)(tl / n - K0 + 1);
    for (int q = threadIdx.x; q < nKb * nr; q += THREADS) lam_s[q] = LAM[(size_t)(K0 + q / nr) * rows + r0 + q % nr];
  }
  __syncthreads();
  if (!live) return;
  double* __restrict__ out = out_all + list_base(g, l) + (size_t)K * bn + (size_t)r0 * n + cc;
  const double* u = u_s + cc;
  const double* lm = lam_s + (K - K0) * nr;
#pragma unroll 4
  for (int r = 0; r < nr; ++r) {
    double v = (u[r * n] * w) / 2.0;
    if (LAMF) v = v * lm[r];
    store<ST>(out + r * n, v * sv);
  }
}

template <bool LAMF>
__global__ void __launch_bounds__(128) cols(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                            const double* __restrict__ unit, const double* __restrict__ width, Geo g) {
  extern __shared__ double sm[];
  const int n = g.n, rows = g.rows, bn = n * rows;
  double* u_s = sm;
  double* lam_s = sm + bn;
  const unsigned pairs = g.nK * n;
  const unsigned t0 = blockIdx.x * 128;
  if (t0 >= pairs) return;
  const unsigned t = t0 + threadIdx.x;
  const bool live = t < pairs;
  const unsigned tt = live ? t : pairs - 1;
  const unsigned K = tt / n, cc = tt - K * n, K0 = t0 / n;
  const int l = blockIdx.y;
  const double sv = W[(size_t)l * g.Lm + 1 + (size_t)K * g.step + cc];
  const double w = width[K];
  for (int q = threadIdx.x; q < bn; q += 128) u_s[q] = -1.0 * unit[q];
  if (LAMF) {
    unsigned tl = t0 + 127;
    if (tl >= pairs) tl = pairs - 1;
    const int n_lam = (int)(tl / n - K0 + 1) * rows;
    for (int q = threadIdx.x; q < n_lam; q += 128) lam_s[q] = LAM[(size_t)K0 * rows + q];
  }
  __syncthreads();
  if (!live) return;
  double* __restrict__ out = out_all + list_base(g, l) + (size_t)K * bn + cc;
  const double* u = u_s + cc;
  const double* lm = lam_s + (K - K0) * rows;
#pragma unroll 4
  for (int r = 0; r < rows; ++r) {
    double v = (u[r * n] * w) / 2.0;
    if (LAMF) v = v * lm[r];
    out[r * n] = v * sv;
  }
}

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ---- flat: one aligned 4-slot group per thread, operands through L1 (read-only path)
// SM = 1: unit block (sign folded) in shared memory instead
template <bool LAMF, int SM, int GPT>  // GPT groups per thread (strided by the block size)
__global__ void __launch_bounds__(256) flat(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                            const double* __restrict__ unit, const double* __restrict__ width, Geo g) {
  extern __shared__ double u_s[];
  const int n = g.n, rows = g.rows, bn = n * rows;
  if (SM) {
    for (int q = threadIdx.x; q < bn; q += 256) u_s[q] = -1.0 * unit[q];
    __syncthreads();
  }
  const int l = blockIdx.y;
  double* __restrict__ out0 = out_all + list_base(g, l);
  const long long total = (long long)g.nK * bn;
  const int pad = (int)(((unsigned long long)(size_t)out0 >> 3) & 3ull);  // slots past a 32-byte boundary
  const double* __restrict__ Wl = W + (size_t)l * g.Lm + 1;
#pragma unroll
  for (int it = 0; it < GPT; ++it) {
    const long long grp = ((long long)blockIdx.x * GPT + it) * 256 + threadIdx.x;
    const long long e0 = grp * 4 - pad;
    if (e0 >= total) return;
    long long ef = e0 < 0 ? 0 : e0;
    unsigned K = (unsigned)(ef / bn);
    unsigned q = (unsigned)(ef - (long long)K * bn);
    unsigned r = q / (unsigned)n, cc = q - r * (unsigned)n;
    double v[4];
    const bool full = e0 >= 0 && e0 + 3 < total;
    if (full && cc + 3 < (unsigned)n) {  // the four slots share interval and row
      const double w = __ldg(width + K);
      const double* up = SM ? u_s + r * n + cc : unit + r * n + cc;
      const double* sp = Wl + (size_t)K * g.step + cc;
      double f = 1.0;
      if (LAMF) f = __ldg(LAM + (size_t)K * rows + r);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double uu = SM ? up[k] : -1.0 * __ldg(up + k);
        double x = (uu * w) / 2.0;
        if (LAMF) x = x * f;
        v[k] = x * __ldg(sp + k);
      }
      st256(out0 + e0, v[0], v[1], v[2], v[3]);
      continue;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long e = e0 + k;
      if (e >= ef && e < total) {
        const double uu = SM ? u_s[r * n + cc] : -1.0 * __ldg(unit + r * n + cc);
        double x = (uu * __ldg(width + K)) / 2.0;
        if (LAMF) x = x * __ldg(LAM + (size_t)K * rows + r);
        v[k] = x * __ldg(Wl + (size_t)K * g.step + cc);
        if (++cc == (unsigned)n) {
          cc = 0;
          if (++r == (unsigned)rows) { r = 0; ++K; }
        }
      }
    }
    if (full) {
      st256(out0 + e0, v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e0 + k >= 0 && e0 + k < total) out0[e0 + k] = v[k];
    }
  }
}

// ---- flat with everything a block needs staged in shared memory: a block owns 256*4*GPT consecutive
// slots = a run of whole/partial intervals; list values, multipliers and widths of those intervals
// are staged with coalesced loads, the store loop reads shared memory only
template <bool LAMF, int GPT>
__global__ void __launch_bounds__(256) flat_staged(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                                   const double* __restrict__ unit, const double* __restrict__ width, Geo g, int maxK) {
  extern __shared__ double sm[];
  const int n = g.n, rows = g.rows, bn = n * rows;
  double* u_s = sm;                    // [bn]
  double* sv_s = u_s + bn;             // [maxK * n]
  double* lam_s = sv_s + maxK * n;     // [maxK * rows]
  double* w_s = lam_s + maxK * rows;   // [maxK]
  const int l = blockIdx.y;
  double* __restrict__ out0 = out_all + list_base(g, l);
  const long long total = (long long)g.nK * bn;
  const int pad = (int)(((unsigned long long)(size_t)out0 >> 3) & 3ull);
  const long long span = 256LL * 4 * GPT;
  const long long b0 = (long long)blockIdx.x * span - pad;  // first slot of the block (may be < 0)
  if (b0 >= total) return;
  const long long bf = b0 < 0 ? 0 : b0;
  long long bl = b0 + span - 1;
  if (bl >= total) bl = total - 1;
  const unsigned K0 = (unsigned)(bf / bn), K1 = (unsigned)(bl / bn);
  const int nKb = (int)(K1 - K0 + 1);
  const double* __restrict__ Wl = W + (size_t)l * g.Lm + 1;
  for (int q = threadIdx.x; q < bn; q += 256) u_s[q] = -1.0 * unit[q];
  if (g.step == n) {
    for (int q = threadIdx.x; q < nKb * n; q += 256) sv_s[q] = Wl[(size_t)K0 * n + q];
  } else {
    for (int q = threadIdx.x; q < nKb * n; q += 256) sv_s[q] = Wl[(size_t)(K0 + q / n) * g.step + q % n];
  }
  if (LAMF)
    for (int q = threadIdx.x; q < nKb * rows; q += 256) lam_s[q] = LAM[(size_t)K0 * rows + q];
  for (int q = threadIdx.x; q < nKb; q += 256) w_s[q] = width[K0 + q];
  __syncthreads();
#pragma unroll
  for (int it = 0; it < GPT; ++it) {
    const long long e0 = b0 + ((long long)it * 256 + threadIdx.x) * 4;
    if (e0 >= total) return;
    const long long ef = e0 < 0 ? 0 : e0;
    unsigned K = (unsigned)(ef / bn);
    unsigned q = (unsigned)(ef - (long long)K * bn);
    unsigned r = q / (unsigned)n, cc = q - r * (unsigned)n;
    K -= K0;
    double v[4];
    const bool full = e0 >= 0 && e0 + 3 < total;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long e = e0 + k;
      if (e >= ef && e < total) {
        double x = (u_s[r * n + cc] * w_s[K]) / 2.0;
        if (LAMF) x = x * lam_s[K * rows + r];
        v[k] = x * sv_s[K * n + cc];
        if (++cc == (unsigned)n) {
          cc = 0;
          if (++r == (unsigned)rows) { r = 0; ++K; }
        }
      }
    }
    if (full) {
      st256(out0 + e0, v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (e0 + k >= 0 && e0 + k < total) out0[e0 + k] = v[k];
    }
  }
}

static double* dalloc(size_t n, bool rnd, double lo = 0.5, double hi = 1.5) {
  double* d;
  cudaMalloc(&d, 8 * n);
  std::vector<double> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = rnd ? lo + (hi - lo) * (rand() / (double)RAND_MAX) : 0.0;
  cudaMemcpy(d, h.data(), 8 * n, cudaMemcpyHostToDevice);
  return d;
}

template <typename F>
static float timeit(F launch, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f, tot;
  for (int rep = 0; rep < 3; ++rep) {
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int it = 0; it < iters; ++it) launch(it & 1);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&tot, a, b);
    if (tot < best) best = tot;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return 1000.f * best / iters;
}

static void study(const char* name, int n, int rows, int step, unsigned nK, int lists, long long Lm) {
  for (int mis = 0; mis < 2; ++mis) {
    Geo g{n, rows, step, lists, nK, Lm, (long long)nK * n * rows + 8, mis};
    const size_t slots = (size_t)lists * g.list_stride + 16;
    double* buf[2];
    cudaMalloc(&buf[0], 8 * slots);
    cudaMalloc(&buf[1], 8 * slots);
    double* W = dalloc((size_t)lists * Lm + 8, true);
    double* LAM = dalloc((size_t)nK * rows + 8, true, -1.0, 1.0);
    double* unit = dalloc((size_t)n * rows, true);
    double* width = dalloc(nK, true, 1e-3, 2e-3);
    const int bn = n * rows;
    const unsigned pairs = nK * n;
    const long long total = (long long)nK * bn;
    const double bytes = 8.0 * lists * total;
    std::vector<double> ref(slots), got(slots);
    const int iters = 30;
    for (int lamf = 0; lamf < 2; ++lamf) {
      auto report = [&](const char* kernel, float us, bool check) {
        bool same = true;
        if (check) {
          cudaMemcpy(got.data(), buf[0], 8 * slots, cudaMemcpyDeviceToHost);
          for (int l = 0; l < lists && same; ++l) {
            const size_t b0 = (size_t)l * g.list_stride + (mis ? 1 + (l & 3) : 0);
            same = memcmp(&ref[b0], &got[b0], 8 * (size_t)total) == 0;
          }
        }
        printf("{\"shape\": \"%s\", \"misaligned\": %d, \"lam\": %d, \"kernel\": \"%s\", \"us\": %.2f, \"GBps\": %.0f, \"bit_identical\": %s}\n", name, mis,
               lamf, kernel, us, bytes / us / 1e3, check ? (same ? "true" : "false") : "null");
        fflush(stdout);
      };
      const size_t csm = 8 * (size_t)(bn + (128 / n + 2) * rows);
      auto l_cols = [&](int w) {
        dim3 grid((pairs + 127) / 128, lists);
        if (lamf) cols<true><<<grid, 128, csm>>>(buf[w], W, LAM, unit, width, g);
        else cols<false><<<grid, 128, csm>>>(buf[w], W, LAM, unit, width, g);
      };
      cudaMemset(buf[0], 0, 8 * slots);
      float us = timeit(l_cols, iters);
      cudaDeviceSynchronize();
      cudaMemcpy(ref.data(), buf[0], 8 * slots, cudaMemcpyDeviceToHost);
      report("cols (pk_expand_cols)", us, false);
#define COLSV(T, U, ST, label)                                                                 \
  {                                                                                            \
    const size_t vsm = 8 * (size_t)(bn + (T / n + 2) * rows);                                  \
    auto lf = [&](int w) {                                                                     \
      dim3 grid((pairs + T - 1) / T, lists);                                                   \
      if (lamf) cols_v<true, T, U, ST><<<grid, T, vsm>>>(buf[w], W, LAM, unit, width, g);      \
      else cols_v<false, T, U, ST><<<grid, T, vsm>>>(buf[w], W, LAM, unit, width, g);          \
    };                                                                                         \
    cudaMemset(buf[0], 0, 8 * slots);                                                          \
    us = timeit(lf, iters);                                                                    \
    report(label, us, true);                                                                   \
  }
      COLSV(64, 4, 0, "cols 64 threads");
      COLSV(256, 4, 0, "cols 256 threads");
      COLSV(128, 2, 0, "cols unroll 2");
      COLSV(128, 10, 0, "cols unroll 10");
      COLSV(128, 20, 0, "cols unroll 20");
      COLSV(128, 4, 1, "cols st.cs");
      COLSV(128, 4, 2, "cols st.wt");
      COLSV(128, 4, 3, "cols st.cg");
      COLSV(64, 10, 1, "cols 64 threads unroll 10 st.cs");
#define COLSS(T, RS, MINB, ST, label)                                                          \
  {                                                                                            \
    const int nrmax = (rows + RS - 1) / RS + 1;                                                \
    const size_t vsm = 8 * (size_t)(nrmax * n + (T / n + 2) * nrmax);                          \
    auto lf = [&](int w) {                                                                     \
      dim3 grid((pairs + T - 1) / T, lists, RS);                                               \
      if (lamf) cols_split<true, T, RS, MINB, ST><<<grid, T, vsm>>>(buf[w], W, LAM, unit, width, g); \
      else cols_split<false, T, RS, MINB, ST><<<grid, T, vsm>>>(buf[w], W, LAM, unit, width, g);     \
    };                                                                                         \
    cudaMemset(buf[0], 0, 8 * slots);                                                          \
    us = timeit(lf, iters);                                                                    \
    report(label, us, true);                                                                   \
  }
      COLSS(128, 1, 16, 1, "split 1 (regs<=32) st.cs");
      COLSS(128, 2, 16, 1, "split 2 st.cs");
      COLSS(128, 4, 16, 1, "split 4 st.cs");
      COLSS(128, 5, 16, 1, "split 5 st.cs");
      COLSS(64, 2, 32, 1, "split 2, 64 threads st.cs");
      COLSS(256, 4, 8, 1, "split 4, 256 threads st.cs");
      COLSS(128, 2, 16, 0, "split 2 plain stores");
      if (getenv("MB_COLS_ONLY")) continue;
#define FLAT(SMV, GPT, label)                                                                  \
  {                                                                                            \
    auto lf = [&](int w) {                                                                     \
      dim3 grid((unsigned)((total / 4 + 2 + 256LL * GPT - 1) / (256LL * GPT)), lists);         \
      if (lamf) flat<true, SMV, GPT><<<grid, 256, SMV ? 8 * bn : 0>>>(buf[w], W, LAM, unit, width, g); \
      else flat<false, SMV, GPT><<<grid, 256, SMV ? 8 * bn : 0>>>(buf[w], W, LAM, unit, width, g);     \
    };                                                                                         \
    cudaMemset(buf[0], 0, 8 * slots);                                                          \
    us = timeit(lf, iters);                                                                    \
    report(label, us, true);                                                                   \
  }
      FLAT(0, 1, "flat L1 1 group/thread");
      FLAT(0, 2, "flat L1 2 groups/thread");
      FLAT(0, 4, "flat L1 4 groups/thread");
      FLAT(1, 1, "flat smem-unit 1 group/thread");
      FLAT(1, 4, "flat smem-unit 4 groups/thread");
#define STAGED(GPT, label)                                                                     \
  {                                                                                            \
    const int maxK = (int)((256LL * 4 * GPT + bn - 1) / bn + 2);                               \
    const size_t ssm = 8 * (size_t)(bn + maxK * (n + rows + 1));                               \
    auto lf = [&](int w) {                                                                     \
      dim3 grid((unsigned)((total + 3 + 256LL * 4 * GPT - 1) / (256LL * 4 * GPT)), lists);     \
      if (lamf) flat_staged<true, GPT><<<grid, 256, ssm>>>(buf[w], W, LAM, unit, width, g, maxK); \
      else flat_staged<false, GPT><<<grid, 256, ssm>>>(buf[w], W, LAM, unit, width, g, maxK);  \
    };                                                                                         \
    cudaMemset(buf[0], 0, 8 * slots);                                                          \
    us = timeit(lf, iters);                                                                    \
    report(label, us, true);                                                                   \
  }
      STAGED(1, "flat staged 1 group/thread (8 KB/block)");
      STAGED(2, "flat staged 2 groups/thread (16 KB/block)");
      STAGED(4, "flat staged 4 groups/thread (32 KB/block)");
    }
    cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(W); cudaFree(LAM); cudaFree(unit); cudaFree(width);
  }
}

int main() {
  srand(1);
  study("robot_arm LGR 20x20, 1999 intervals, 16 lists", 20, 20, 20, 1999, 16, 40000);
  study("humanoid LGL 9x10, 11110 intervals, 16 lists", 10, 9, 9, 11110, 16, 100009);
  study("hp LGR 7x7, 20000 intervals, 12 lists", 7, 7, 7, 20000, 12, 140008);
  return 0;
}
