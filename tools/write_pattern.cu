// tools/write_pattern.cu — how fast can 148 CTAs write a (P x T) fp32 block in the access patterns of the
// dense log-likelihood kernels?  Pure stores, no compute: the ceiling the output stream puts on K1.
//   pattern 0  pdf-major out[p][t], frame-stationary order (CTA c: frame tiles c, c+148, ...; all P rows each)
//   pattern 1  pdf-major out[p][t], Gaussian-stationary order (CTA c: its ~P/148 rows, frame tiles 0, 1, 2, ... in lock step)
//   pattern 2  tile-major out[t/128][p][128], frame-stationary order (a CTA writes one contiguous P x 512 B region per tile)
//   pattern 3  tile-major, Gaussian-stationary order (CTA c: its rows' contiguous ~14 KB of every frame tile)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/write_pattern.cu -o tools/write_pattern.bin
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(768) wp(float *out, int P, long long ld, int n_ft, int pattern, int pace) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, eg = warp >> 2, quad = warp & 3;
  const int C = gridDim.x, c = blockIdx.x;
  const int row = quad * 32 + lane;
  if (pattern == 0 || pattern == 2) {
    for (int f = c; f < n_ft; f += C)
      for (int p = eg; p < P; p += 6) {
        float *dst = pattern == 0 ? out + (long long)p * ld + (long long)f * 128 + row : out + ((long long)f * P + p) * 128 + row;
        *dst = (float)p;
        if (pace) __nanosleep(pace);
      }
  } else {
    const int p0 = (int)((long long)P * c / C), p1 = (int)((long long)P * (c + 1) / C);
    for (int f = 0; f < n_ft; ++f)
      for (int p = p0 + eg; p < p1; p += 6) {
        float *dst = pattern == 1 ? out + (long long)p * ld + (long long)f * 128 + row : out + ((long long)f * P + p) * 128 + row;
        *dst = (float)p;
        if (pace) __nanosleep(pace);
      }
  }
}

int main(int argc, char **argv) {
  const int P = argc > 1 ? atoi(argv[1]) : 4200;
  const int n_ft = argc > 2 ? atoi(argv[2]) : 148 * 16;
  const long long ld = (long long)n_ft * 128;
  float *out;
  cudaMalloc(&out, sizeof(float) * (size_t)P * ld);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int pattern = 0; pattern < 4; ++pattern) {
    for (int rep = 0; rep < 2; ++rep) wp<<<148, 768>>>(out, P, ld, n_ft, pattern, 0);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int rep = 0; rep < reps; ++rep) wp<<<148, 768>>>(out, P, ld, n_ft, pattern, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double gb = (double)P * ld * 4 / 1e9;
    printf("pattern %d: %.3f ms per %.2f GB block = %.0f GB/s  (%s)\n", pattern, ms, gb, gb / (ms * 1e-3),
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
