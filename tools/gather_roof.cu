// tools/gather_roof.cu — the ceiling the GATHER puts on the statistics kernels (K3 / K3t): how fast can the chip read
// T rows of D floats in bucketed (sorted-by-pdf) order, nothing else?  One thread per row (as K3t reads them: 256-bit or
// 128-bit loads, 32 different rows per warp instruction) or LPR lanes per row (coalesced within the row), rows summed and
// one float written per row so that nothing is optimised away.
//   order 0  sequential (the streaming roof, for comparison)
//   order 1  stable sort by a random pdf id in [0, P): what K2 hands to K3t
//   order 2  a random permutation
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gather_roof.cu -o tools/gather_roof.bin
// run:   tools/gather_roof.bin [T=8000000] [D=40] [P=4200]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>
#include <cuda_runtime.h>

template <int LPR>  // lanes per row: 1 = a thread reads its whole row, 5 = 32 bytes per lane (D = 40), 10 = 16 bytes per lane
__global__ void __launch_bounds__(256) gather(const float *__restrict__ feats, const int *__restrict__ order, int T, int D, float *__restrict__ out,
                                              int prefetch_ahead) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  if (LPR == 1) {
    for (int i = tid; i < T; i += nthr) {
      if (prefetch_ahead > 0 && i + prefetch_ahead * nthr < T) {
        const char *r = reinterpret_cast<const char *>(feats + (size_t)__ldg(order + i + prefetch_ahead * nthr) * D);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 4 * D - 4));
      }
      const float *src = feats + (size_t)__ldg(order + i) * D;
      float s = 0.f;
      for (int q = 0; q < D / 8; ++q) {
        float v[8];
        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(src + 8 * q));
        for (int j = 0; j < 8; ++j) s += v[j];
      }
      out[i] = s;
    }
  } else {
    const int lane = tid % LPR, rows_per_pass = nthr / LPR;
    constexpr int W = 40 / LPR;  // floats per lane (D = 40)
    if (tid / LPR >= rows_per_pass) return;
    for (int i = tid / LPR; i < T; i += rows_per_pass) {
      const float *src = feats + (size_t)__ldg(order + i) * D + lane * W;
      float s = 0.f;
      if (W == 8) {
        float v[8];
        asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(src));
        for (int j = 0; j < 8; ++j) s += v[j];
      } else {
        const float4 v = __ldg(reinterpret_cast<const float4 *>(src));
        s = v.x + v.y + v.z + v.w;
      }
      if (s == 12345.678f) out[i] = s;  // (never: keeps the loads alive without a write stream per lane)
      if (lane == 0) out[i] = s;
    }
  }
}

int main(int argc, char **argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 8000000, D = argc > 2 ? atoi(argv[2]) : 40, P = argc > 3 ? atoi(argv[3]) : 4200;
  if (D != 40) { printf("D = 40 only\n"); return 1; }
  float *feats, *out;
  int *order;
  cudaMalloc(&feats, sizeof(float) * (size_t)T * D);
  cudaMalloc(&out, sizeof(float) * (size_t)T);
  cudaMalloc(&order, sizeof(int) * (size_t)T);
  cudaMemset(feats, 0, sizeof(float) * (size_t)T * D);
  std::mt19937 rng(7);
  std::vector<int> ord(T), key(T);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const char *names[3] = {"sequential", "sorted by pdf (K2's order)", "random permutation"};
  for (int o = 0; o < 3; ++o) {
    std::iota(ord.begin(), ord.end(), 0);
    if (o == 1) {
      for (int i = 0; i < T; ++i) key[i] = (int)(rng() % (unsigned)P);
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return key[a] < key[b]; });
    } else if (o == 2) {
      std::shuffle(ord.begin(), ord.end(), rng);
    }
    cudaMemcpy(order, ord.data(), sizeof(int) * (size_t)T, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 5; ++variant) {
      // variants: thread per row at 8 / 4 CTAs of 256 per SM, the same with an L2 prefetch 8 passes ahead, 5 and 10 lanes per row
      const int ctas = (variant == 1 ? 4 : 8) * sms;
      auto launch = [&]() {
        if (variant == 0 || variant == 1) gather<1><<<ctas, 256>>>(feats, order, T, D, out, 0);
        else if (variant == 2) gather<1><<<ctas, 256>>>(feats, order, T, D, out, 8);
        else if (variant == 3) gather<5><<<ctas, 256>>>(feats, order, T, D, out, 0);
        else gather<10><<<ctas, 256>>>(feats, order, T, D, out, 0);
      };
      launch();
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      const int reps = 5;
      for (int r = 0; r < reps; ++r) launch();
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= reps;
      const char *vn[5] = {"thread per row, 64 warps/SM", "thread per row, 32 warps/SM", "thread per row + L2 prefetch", "5 lanes per row (32 B each)",
                           "10 lanes per row (16 B each)"};
      printf("%-28s %-30s %7.3f ms  %6.2f G rows/s  %7.1f GB/s (164 B/row)\n", names[o], vn[variant], ms, T / ms / 1e6, T * 164.0 / ms / 1e6);
    }
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
