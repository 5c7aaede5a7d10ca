// What separates the block-expansion kernel from a plain column-pattern fill?  Same write pattern
// (20x20 blocks, 16 lists x 2000 intervals = 12.8 M slots = 102 MB, two buffers alternated so the
// writes are DRAM-bound), with the ingredients switched on one at a time:
//   bit 0  list value + interval width loaded from global memory before the stores
//   bit 1  unit block staged in shared memory (+ __syncthreads), read per store
//   bit 2  multipliers staged in shared memory, read per store
//   bit 3  the four dependent FP64 multiplies per store
//   nvcc -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o tools/_bin/microbench_expand tools/microbench_expand.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int F, int THREADS>
__global__ void __launch_bounds__(THREADS) expand(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                                 const double* __restrict__ unit, const double* __restrict__ width, unsigned pairs, int mis) {
  __shared__ double u_s[400];
  __shared__ double lam_s[(THREADS / 20 + 2) * 20];
  const int n = 20, rows = 20;
  const unsigned t0 = blockIdx.x * THREADS;
  const unsigned t = t0 + threadIdx.x;
  const bool live = t < pairs;
  const unsigned tt = live ? t : pairs - 1;
  const unsigned K = tt / n, cc = tt - K * n, K0 = t0 / n;
  const unsigned l = blockIdx.y;
  double sv = 1.0, w = 1.0;
  if (F & 1) {
    sv = W[(size_t)l * 40000 + 1 + K * 20 + cc];
    w = width[K];
  }
  if (F & 2) {
    for (int q = threadIdx.x; q < 400; q += THREADS) u_s[q] = unit[q];
  }
  if (F & 4) {
    unsigned tl = t0 + THREADS - 1;
    if (tl >= pairs) tl = pairs - 1;
    const int n_lam = (int)(tl / n - K0 + 1) * rows;
    for (int q = threadIdx.x; q < n_lam; q += THREADS) lam_s[q] = LAM[(size_t)K0 * rows + q];
  }
  if (F & 6) __syncthreads();
  if (!live) return;
  double* __restrict__ out = out_all + (size_t)l * pairs * rows + (size_t)K * 400 + cc + (mis ? 1 + (l & 3) : 0);
  const double* u = u_s + cc;
  const double* lm = lam_s + (K - K0) * rows;
#pragma unroll 4
  for (int r = 0; r < rows; ++r) {
    double v = (F & 2) ? u[r * n] : 1.0 + r;
    if (F & 8) v = (v * w) / 2.0;
    if (F & 4) v = (F & 8) ? v * lm[r] : lm[r];
    if (F & 8) v = v * sv; else v = v + sv;
    out[r * n] = v;
  }
}

template <int F, int THREADS>
float run(double* const* buf, const double* W, const double* LAM, const double* unit, const double* width, int iters, int mis = 0) {
  const unsigned pairs = 1999 * 20;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f, tot;
  for (int rep = 0; rep < 3; ++rep) {
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int it = 0; it < iters; ++it)
      expand<F, THREADS><<<dim3((pairs + THREADS - 1) / THREADS, 16), THREADS>>>(buf[it & 1], W, LAM, unit, width, pairs, mis);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&tot, a, b);
    if (tot < best) best = tot;
  }
  return 1000.f * best / iters;
}

int main() {
  const size_t n = 16ull * 1999 * 400;
  double* buf[2];
  cudaMalloc(&buf[0], 8 * n + 64);
  cudaMalloc(&buf[1], 8 * n + 64);
  double *W, *LAM, *unit, *width;
  cudaMalloc(&W, 8 * 16 * 40000);
  cudaMalloc(&LAM, 8 * 240000);
  cudaMalloc(&unit, 8 * 400);
  cudaMalloc(&width, 8 * 2000);
  cudaMemset(W, 0, 8 * 16 * 40000);
  cudaMemset(LAM, 0, 8 * 240000);
  cudaMemset(unit, 0, 8 * 400);
  cudaMemset(width, 0, 8 * 2000);
  const int iters = 40;
#define RUN(F, T) printf("{\"flags\": %d, \"threads\": %d, \"us\": %.2f, \"GBps\": %.0f}\n", F, T, run<F, T>(buf, W, LAM, unit, width, iters), 8.0 * n / run<F, T>(buf, W, LAM, unit, width, iters) / 1e3)
  RUN(0, 128); RUN(1, 128); RUN(2, 128); RUN(4, 128); RUN(8, 128); RUN(3, 128); RUN(7, 128); RUN(11, 128); RUN(15, 128);
#define RUNM(F, T) printf("{\"flags\": %d, \"threads\": %d, \"misaligned\": 1, \"us\": %.2f}\n", F, T, run<F, T>(buf, W, LAM, unit, width, iters, 1))
  RUNM(0, 128); RUNM(1, 128); RUNM(8, 128); RUNM(9, 128); RUNM(6, 128); RUNM(15, 128); RUNM(15, 256);
  return 0;
}
