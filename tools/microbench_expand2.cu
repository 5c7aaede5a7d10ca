// Candidate restructuring of the block expansion ("P in registers"): a thread owns one column
// (interval K, column c) of the block for LC lists of a state: it computes
// P[r] = ((u[r][c]*w_K)/2 [*lam_r]) once (registers), then for each list l: out_l[r][c] = P[r]*sv_l,
// i.e. one FP64 multiply and one store per slot, no shared-memory traffic in the store loop.
// Same synthetic problem as microbench_expand.cu (16 lists x 1999 intervals x 20x20, misaligned list bases).
//   nvcc -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o tools/_bin/microbench_expand2 tools/microbench_expand2.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int LC, int THREADS, bool LAMF>
__global__ void __launch_bounds__(THREADS) expand_p(double* __restrict__ out_all, const double* __restrict__ W, const double* __restrict__ LAM,
                                                   const double* __restrict__ unit, const double* __restrict__ width, unsigned pairs, int mis) {
  constexpr int N = 20, ROWS = 20;
  __shared__ double u_s[N * ROWS];
  __shared__ double lam_s[(THREADS / N + 2) * ROWS];
  const unsigned t0 = blockIdx.x * THREADS;
  const unsigned t = t0 + threadIdx.x;
  const bool live = t < pairs;
  const unsigned tt = live ? t : pairs - 1;
  const unsigned K = tt / N, cc = tt - K * N, K0 = t0 / N;
  const unsigned l0 = blockIdx.y * LC;
  double sv[LC];
#pragma unroll
  for (int i = 0; i < LC; ++i) sv[i] = W[(size_t)(l0 + i) * 40000 + 1 + K * N + cc];
  const double w = width[K];
  for (int q = threadIdx.x; q < N * ROWS; q += THREADS) u_s[q] = unit[q];
  if (LAMF) {
    unsigned tl = t0 + THREADS - 1;
    if (tl >= pairs) tl = pairs - 1;
    const int n_lam = (int)(tl / N - K0 + 1) * ROWS;
    for (int q = threadIdx.x; q < n_lam; q += THREADS) lam_s[q] = LAM[(size_t)K0 * ROWS + q];
  }
  __syncthreads();
  if (!live) return;
  double P[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    double v = (u_s[r * N + cc] * w) / 2.0;
    if (LAMF) v = v * lam_s[(K - K0) * ROWS + r];
    P[r] = v;
  }
#pragma unroll
  for (int i = 0; i < LC; ++i) {
    double* __restrict__ out = out_all + (size_t)(l0 + i) * pairs * ROWS + (size_t)K * (N * ROWS) + cc + (mis ? 1 + ((l0 + i) & 3) : 0);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) out[r * N] = P[r] * sv[i];
  }
}

template <int LC, int THREADS, bool LAMF>
float run(double* const* buf, const double* W, const double* LAM, const double* unit, const double* width, int iters, int mis) {
  const unsigned pairs = 1999 * 20;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f, tot;
  for (int rep = 0; rep < 3; ++rep) {
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int it = 0; it < iters; ++it)
      expand_p<LC, THREADS, LAMF><<<dim3((pairs + THREADS - 1) / THREADS, 16 / LC), THREADS>>>(buf[it & 1], W, LAM, unit, width, pairs, mis);
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
#define RUN(LC, T, L) printf("{\"lists_per_thread\": %d, \"threads\": %d, \"lam\": %d, \"us_aligned\": %.2f, \"us_misaligned\": %.2f}\n", LC, T, (int)L, \
                             run<LC, T, L>(buf, W, LAM, unit, width, iters, 0), run<LC, T, L>(buf, W, LAM, unit, width, iters, 1))
  RUN(1, 128, true); RUN(2, 128, true); RUN(4, 128, true); RUN(8, 128, true); RUN(16, 128, true);
  RUN(4, 64, true); RUN(4, 256, true); RUN(4, 128, false); RUN(2, 64, true); RUN(8, 64, true);
  return 0;
}
