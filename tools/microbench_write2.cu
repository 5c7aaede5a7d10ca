// Second pass of the write-ceiling study: which launch SHAPES reach the ~6 TB/s a plain fill gets
// at 102 MB (the size of one Jacobian / Hessian value array of robot_arm LGR 2000x20)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/microbench_write2 tools/microbench_write2.cu
#include <cuda_runtime.h>
#include <cstdio>

// grid-stride fill, `per` = elements per thread implied by the grid size
__global__ void fill_gs(double* p, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
// every block owns one contiguous chunk of `per` * 256 elements
__global__ void fill_chunk(double* p, long long n, int per, double v) {
  long long i = (long long)blockIdx.x * per * 256 + threadIdx.x;
  for (int k = 0; k < per; ++k, i += 256)
    if (i < n) p[i] = v;
}
// persistent, but every block walks its own contiguous range
__global__ void fill_range(double* p, long long n, double v) {
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = lo + per < n ? lo + per : n;
  for (long long i = lo + threadIdx.x; i < hi; i += 256) p[i] = v;
}
// column walk of 20x20 blocks: one unit (column) per thread, grid-stride
__global__ void cols(double* p, long long n_units, double v) {
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n_units; u += (long long)gridDim.x * blockDim.x) {
    const long long K = u / 20, c = u - K * 20;
    double* o = p + K * 400 + c;
#pragma unroll 4
    for (int r = 0; r < 20; ++r) o[r * 20] = v * (double)r;
  }
}
// column walk, rows split over `split` threads (more, shorter walks)
__global__ void cols_split(double* p, long long n_units, int split, double v) {
  const int rows = 20 / split;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < n_units * split; w += (long long)gridDim.x * blockDim.x) {
    const long long u = w % n_units, part = w / n_units;
    const long long K = u / 20, c = u - K * 20;
    double* o = p + K * 400 + c + part * rows * 20;
    for (int r = 0; r < rows; ++r) o[r * 20] = v * (double)r;
  }
}

int main() {
  const long long n = 12800000LL;
  double* buf[2];
  cudaMalloc(&buf[0], sizeof(double) * n);
  cudaMalloc(&buf[1], sizeof(double) * n);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 40;
  struct V { const char* name; int kind; int arg; };
  const V vs[] = {
      {"fill grid-stride 1 blk/SM", 0, 1}, {"fill grid-stride 2 blk/SM", 0, 2}, {"fill grid-stride 4 blk/SM", 0, 4},
      {"fill grid-stride 8 blk/SM", 0, 8}, {"fill grid-stride 16 el/thread", 1, 16}, {"fill grid-stride 8 el/thread", 1, 8},
      {"fill grid-stride 4 el/thread", 1, 4}, {"fill grid-stride 2 el/thread", 1, 2}, {"fill grid-stride 1 el/thread", 1, 1},
      {"fill chunk 4 el/thread", 2, 4}, {"fill chunk 16 el/thread", 2, 16}, {"fill chunk 64 el/thread", 2, 64},
      {"fill range 8 blk/SM", 3, 8}, {"cols 6 blk/SM", 4, 6}, {"cols 8 blk/SM", 4, 8}, {"cols 1 unit/thread", 5, 1},
      {"cols 2 units/thread", 5, 2}, {"cols split2 1/thread", 6, 2}, {"cols split4 1/thread", 6, 4}, {"cols split4 8blk/SM", 7, 4},
  };
  for (const V& v : vs) {
    float best = 1e9f, tot = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaDeviceSynchronize();
      cudaEventRecord(a);
      for (int it = 0; it < iters; ++it) {
        double* p = buf[it & 1];
        const long long units = n / 20;
        switch (v.kind) {
          case 0: fill_gs<<<sms * v.arg, 256>>>(p, n, 1.0); break;
          case 1: fill_gs<<<(unsigned)((n + 256LL * v.arg - 1) / (256LL * v.arg)), 256>>>(p, n, 1.0); break;
          case 2: fill_chunk<<<(unsigned)((n + 256LL * v.arg - 1) / (256LL * v.arg)), 256>>>(p, n, v.arg, 1.0); break;
          case 3: fill_range<<<sms * v.arg, 256>>>(p, n, 1.0); break;
          case 4: cols<<<sms * v.arg, 256>>>(p, units, 1.0); break;
          case 5: cols<<<(unsigned)((units + 256LL * v.arg - 1) / (256LL * v.arg)), 256>>>(p, units, 1.0); break;
          case 6: cols_split<<<(unsigned)((units * v.arg + 255) / 256), 256>>>(p, units, v.arg, 1.0); break;
          case 7: cols_split<<<sms * 8, 256>>>(p, units, v.arg, 1.0); break;
        }
      }
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&tot, a, b);
      if (tot < best) best = tot;
    }
    const double us = 1000.0 * best / iters;
    printf("{\"bytes_MB\": %.1f, \"variant\": \"%s\", \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", 8.0 * n / 1e6, v.name, us, 8.0 * n / us / 1e3);
  }
  return 0;
}
