// Write-bandwidth ceiling at the sizes the expansion kernels work at (robot_arm LGR 2000x20:
// 100 MB Jacobian + 102 MB Hessian values per evaluation set).  Plain fills with 8- and 16-byte
// stores, one resident wave (persistent, grid-stride) vs one block per chunk, alternating between
// two output buffers like the engine does, timed with CUDA events over back-to-back launches.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/microbench_write tools/microbench_write.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void fill8(double* p, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill16(double2* p, long long n2, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) p[i] = make_double2(v, v);
}
__global__ void fill8_cs(double* p, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) __stcs(p + i, v);
}
// strided like the column-walking expansion: thread owns column c of a 20x20 block, writes 20 rows
__global__ void fill_cols(double* p, long long n_units, double v) {
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < n_units; u += (long long)gridDim.x * blockDim.x) {
    const long long K = u / 20, c = u - K * 20;
    double* o = p + K * 400 + c;
#pragma unroll 4
    for (int r = 0; r < 20; ++r) o[r * 20] = v * (double)r;
  }
}

int main() {
  const long long sizes[] = {12800000LL, 25600000LL, 33554432LL};  // doubles: 102 MB, 205 MB, 268 MB
  double* buf[2];
  cudaMalloc(&buf[0], sizeof(double) * sizes[2]);
  cudaMalloc(&buf[1], sizeof(double) * sizes[2]);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 40;
  for (long long n : sizes) {
    for (int variant = 0; variant < 6; ++variant) {
      const char* name[] = {"fill8 wave(8/SM)", "fill16 wave(8/SM)", "fill8 blocks", "fill16 blocks", "fill8.cs wave", "cols20x20 wave(6/SM)"};
      float best = 1e9f, tot = 0.f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        for (int it = 0; it < iters; ++it) {
          double* p = buf[it & 1];
          switch (variant) {
            case 0: fill8<<<sms * 8, 256>>>(p, n, 1.0); break;
            case 1: fill16<<<sms * 8, 256>>>((double2*)p, n / 2, 1.0); break;
            case 2: fill8<<<(unsigned)((n + 1023) / 1024), 256>>>(p, n, 1.0); break;
            case 3: fill16<<<(unsigned)((n / 2 + 1023) / 1024), 256>>>((double2*)p, n / 2, 1.0); break;
            case 4: fill8_cs<<<sms * 8, 256>>>(p, n, 1.0); break;
            case 5: fill_cols<<<sms * 6, 256>>>(p, n / 20, 1.0); break;
          }
        }
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&tot, a, b);
        if (tot < best) best = tot;
      }
      const double us = 1000.0 * best / iters;
      printf("{\"bytes_MB\": %.1f, \"variant\": \"%s\", \"us_per_launch\": %.2f, \"GBps\": %.0f}\n", 8.0 * n / 1e6, name[variant], us,
             8.0 * n / us / 1e3);
    }
  }
  return 0;
}
