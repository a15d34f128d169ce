// tools/mma_rate.cu — microbenchmark: cycles per tcgen05.mma for the operand modes K1 can use.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mma_rate.bin tools/mma_rate.cu
// Operands are whatever is in shared memory / TMEM (values do not matter for the rate).
// Variants: SS (A and B from smem), TS (A from TMEM), cta_group::1 (M=128) and ::2 (M=256 over a CTA pair).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}

template <int CG, bool TS, bool F16>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_lo_or_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
  if (TS) {
    if (CG == 1) {
      if (F16)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}"
                   ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    }
  } else {
    if (CG == 1) {
      if (F16)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
                     ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
                   ::"r"(d), "r"(a_lo_or_tmem), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
    }
  }
}

// One CTA (or CTA pair) per SM.  Thread 0 of the (leader) CTA issues `reps` groups of `per_group`
// MMAs that walk over `kslices` K slices of the smem operands (32 B apart), alternating between two
// accumulators, then commits and waits.  out[cta] = cycles per MMA * 100.
template <int CG, bool TS, bool F16, int KS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int reps, unsigned *out) {
  constexpr int per_group = 16, kslices = KS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *bp = smem_raw + (base - raw);
  // layout: A 4 chunks x 16 KB (128 rows x 128 B), B 4 chunks x 32 KB (256 rows x 128 B), barrier
  const uint32_t sA = base, sB = base + 4 * 16384, sBar = sB + 4 * 32768, sSlot = sBar + 8;
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(bp)[i] = 0x3c003c00u;
  uint32_t cta_rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sBar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sSlot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sSlot), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(bp + (sSlot - base));
  const int M = CG == 2 ? 256 : 128;
  const uint32_t fmt = F16 ? 0u : 2u;
  const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  // the whole warp runs the issue loop (warp-uniform control flow keeps descriptors in uniform
  // registers); one elected lane issues
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_u == 0 && cta_rank == 0) {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t tmem = tmem_u;
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t a0 = TS ? (tmem + 2 * 256 - 64) : ((sA >> 4) & 0x3FFF), b0 = (sB >> 4) & 0x3FFF;
    long long t0 = clock64();
    int n = 0;
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tmem + (r & 1) * (TS ? 224 : 256);
#pragma unroll
      for (int k = 0; k < per_group; ++k) {
        const int ks = k % kslices;  // K slice: chunk = ks / 4, 32-byte step within the 128-byte atom = ks % 4
        const uint32_t aoff = TS ? (uint32_t)(ks * 8) : (uint32_t)((ks >> 2) * (16384 >> 4) + (ks & 3) * 2);
        const uint32_t boff = (uint32_t)((ks >> 2) * (32768 >> 4) + (ks & 3) * 2);
        if (leader) mma<CG, TS, F16>(d, a0 + aoff, b0 + boff, idesc, k ? 1u : 0u);
        ++n;
      }
    }
    if (leader) {
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(sBar), "h"((uint16_t)1) : "memory");
    }
    mbar_wait(sBar, 0);
    long long t1 = clock64();
    if (leader) out[blockIdx.x] = (unsigned)((t1 - t0) * 100 / n);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG, bool TS, bool F16, int KS>
static void run(const char *name, int N, int grid) {
  const int per_group = 16, kslices = KS;
  unsigned *d_out;
  cudaMalloc(&d_out, sizeof(unsigned) * grid);
  cudaMemset(d_out, 0, sizeof(unsigned) * grid);
  const size_t smem = 4 * 16384 + 4 * 32768 + 64 + 1024;
  cudaFuncSetAttribute(mma_rate_kernel<CG, TS, F16, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int reps = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, mma_rate_kernel<CG, TS, F16, KS>, N, reps, d_out);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (err != cudaSuccess || e2 != cudaSuccess) {
      printf("%-34s N=%3d grid=%3d: ERROR %s / %s\n", name, N, grid, cudaGetErrorString(err), cudaGetErrorString(e2));
      exit(1);
    }
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned h[512];
  cudaMemcpy(h, d_out, sizeof(unsigned) * grid, cudaMemcpyDeviceToHost);
  double sum = 0;
  int cnt = 0;
  for (int i = 0; i < grid; ++i)
    if (h[i]) { sum += h[i]; ++cnt; }
  const double cyc = sum / cnt / 100.0;
  const int M = CG == 2 ? 256 : 128, K = F16 ? 16 : 8;
  const double floor_cyc = (double)128 * N / 256.0 * (F16 ? 1 : 1);  // per SM: 4096 f16 MAC/clk (tf32: half, K is half)
  const double macs = (double)M * N * K * reps * per_group * cnt;
  printf("%-34s N=%3d grid=%3d k-slices=%2d: %7.1f cyc/MMA (floor %5.1f, %4.0f %%)  %7.1f TFLOP/s (%.3f ms)\n", name, N, grid,
         kslices, cyc, floor_cyc, 100.0 * floor_cyc / cyc, 2.0 * macs / (ms * 1e-3) / 1e12, ms);
  cudaFree(d_out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int grid : {1, sms}) {
    for (int N : {240, 256, 128}) {
      run<1, false, true, 16>("SS f16 cta_group::1 M=128", N, grid);
      run<1, false, true, 1>("SS f16 cta_group::1 M=128 same-k", N, grid);
      run<1, true, true, 8>("TS f16 cta_group::1 M=128", N == 256 ? 224 : N, grid);
      run<1, false, false, 16>("SS tf32 cta_group::1 M=128", N, grid);
      run<1, true, false, 8>("TS tf32 cta_group::1 M=128", N == 256 ? 224 : N, grid);
    }
    const int g2 = grid == 1 ? 2 : (sms / 2) * 2;
    for (int N : {240, 256, 128}) {
      run<2, false, true, 16>("SS f16 cta_group::2 M=256", N, g2);
      run<2, true, true, 8>("TS f16 cta_group::2 M=256", N == 256 ? 224 : N, g2);
    }
  }
  return 0;
}
