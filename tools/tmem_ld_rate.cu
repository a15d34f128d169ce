// tools/tmem_ld_rate.cu — microbenchmark: tcgen05.ld throughput (TMEM -> registers) per SM.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_ld_rate.bin tools/tmem_ld_rate.cu
// W warps per CTA (warp w reads lane quadrant w % 4), each issues `reps` loads of 32 lanes x X
// columns (32x32b.xX), waiting after every `depth` loads.  Prints bytes/clk/SM.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<8>(uint32_t taddr) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
  return r[0] ^ r[7];
}
template <>
__device__ __forceinline__ uint32_t ld<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
  return r[0] ^ r[15];
}
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                 "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
  return r[0] ^ r[31];
}

template <int X>
__global__ void __launch_bounds__(1024, 1) tmem_ld_kernel(int reps, int depth, unsigned *out, unsigned *sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; r += depth) {
    for (int k = 0; k < depth; ++k) acc ^= ld<X>(trow + (uint32_t)(((r + k) * X + (warp >> 2) * 64) & 255));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = (unsigned)(t1 - t0);
  if (acc == 0x12345) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int X>
static void run(int warps, int depth, int grid) {
  unsigned *d_out;
  cudaMalloc(&d_out, sizeof(unsigned) * (grid + 1));
  const int reps = 4096;
  for (int it = 0; it < 2; ++it) tmem_ld_kernel<X><<<grid, 32 * warps>>>(reps, depth, d_out, d_out + grid);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  unsigned h[256];
  cudaMemcpy(h, d_out, sizeof(unsigned) * grid, cudaMemcpyDeviceToHost);
  double cyc = 0;
  for (int i = 0; i < grid; ++i) cyc += h[i];
  cyc /= grid;
  const double bytes = (double)warps * reps * 32 * X * 4;
  printf("x%-2d warps=%2d depth=%d grid=%3d: %8.0f cycles  %6.1f B/clk/SM  (%5.1f cyc per warp-load)\n", X, warps, depth, grid, cyc,
         bytes / cyc, cyc / reps);
  cudaFree(d_out);
}

int main() {
  for (int grid : {1, 148})
    for (int warps : {4, 8, 16, 32})
      for (int depth : {1, 2, 4}) {
        run<8>(warps, depth, grid);
        run<16>(warps, depth, grid);
        run<32>(warps, depth, grid);
      }
  return 0;
}
