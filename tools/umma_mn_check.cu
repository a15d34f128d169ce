// tools/umma_mn_check.cu — checks the operand trick the tensor-core statistics kernel (K3t) rests on, and
// measures the cost of small-N tcgen05.mma instructions.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/umma_mn_check.bin tools/umma_mn_check.cu
//
// One shared-memory tile X[t][k] (128 frames x 128 columns of fp16, K-major 128-byte swizzle: chunk c holds
// columns 64c..64c+63, row t at byte t*128, 16-byte units XOR-swizzled with t & 7) is used TWICE:
//   phase A  L[t][g] = sum_k X[t][k] * B[g][k]     A = X as a K-major operand      (M = t, K = k)
//   phase B  S[k][g] = sum_t X[t][k] * P[g][t]     A = X as an MN-major operand    (M = k, K = t)
// The MN-major canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units with Swizzle<3,4,3> is the
// same byte image: 8 frames x 128 bytes per atom, SBO = 1024 B to the next 8 frames, LBO = 16384 B to the
// next 64 columns.  Both results are compared with the CPU.  Then: cycles per MMA for N = 16, 32 (f16).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mma_f16(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
               ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi) : "memory");
}
__device__ __forceinline__ void mma_f16_desc(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
// byte offset of element (row, col) in a K-major 128B-swizzle operand of `rows` rows (chunk = 64 fp16 columns)
__host__ __device__ inline uint32_t sw_off(int rows, int row, int col) {
  return (uint32_t)((col >> 6) * rows * 128 + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1));
}

// X: 128 x 128 fp16 (row-major, plain), B: N x 128, P: N x 128 (row g, column t).  out_a[t][g], out_b[k][g].
__global__ void __launch_bounds__(128, 1) check_kernel(const __half *X, const __half *B, const __half *P, int N, float *out_a, float *out_b, float *out_c,
                                                       int reps, unsigned *cyc) {
  extern __shared__ uint8_t raw_[];
  const uint32_t raw = smem_u32(raw_);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *bp = raw_ + (base - raw);
  // X: 2 chunks x 16 KB | B: 2 chunks x N*128 | P: 2 chunks x N*128 | barrier, slot
  const uint32_t oX = 0, oB = 32768, oP = oB + 2 * 32 * 128, oP2 = oP + 2 * 32 * 128, oBar = oP2 + 2 * 4096, oSlot = oBar + 8;
  for (int i = threadIdx.x; i < 128 * 128; i += 128) {
    const int r = i >> 7, c = i & 127;
    *reinterpret_cast<__half *>(bp + oX + sw_off(128, r, c)) = X[i];
  }
  for (int i = threadIdx.x; i < N * 128; i += 128) {
    const int r = i >> 7, c = i & 127;
    *reinterpret_cast<__half *>(bp + oB + sw_off(N, r, c)) = B[i];
    *reinterpret_cast<__half *>(bp + oP + sw_off(N, r, c)) = P[i];
    // the same posterior matrix, frames as rows: P2[t = c][g = r], MN-major (g contiguous), 32-byte swizzle
    *reinterpret_cast<__half *>(bp + oP2 + (r >> 4) * 4096 + c * 32 + (((((r & 15) >> 3) ^ ((c >> 2) & 1))) << 4) + ((r & 7) << 1)) = P[i];
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(base + oBar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + oSlot), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(bp + oSlot);
  const uint32_t idesc_k = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // A, B K-major
  const uint32_t idesc_mn = idesc_k | (1u << 15);                                                 // A MN-major
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  uint32_t phase = 0;
  if (warp_u == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t xa = ((base + oX) >> 4) & 0x3FFF, ba = ((base + oB) >> 4) & 0x3FFF, pa = ((base + oP) >> 4) & 0x3FFF;
    // phase A: D0[t][g], K = 128 columns = 8 steps of 16: chunk = s / 4, 32-byte step (s % 4) inside the atom
    for (int s = 0; s < 8; ++s) {
      const uint32_t ao = (uint32_t)((s >> 2) * (16384 >> 4) + (s & 3) * 2), bo = (uint32_t)((s >> 2) * ((N * 128) >> 4) + (s & 3) * 2);
      if (leader) mma_f16(tmem, xa + ao, ba + bo, idesc_k, s ? 1u : 0u);
    }
    // phase B: D1[k][g], K = 128 frames = 8 steps of 16 frames = 2048 bytes of X each; the A descriptor carries
    // LBO = 16384 B (next 64 columns) in bits 16..29; the P operand is K-major over frames
    for (int s = 0; s < 8; ++s) {
      const uint32_t ao = (uint32_t)(s * (2048 >> 4)) | ((uint32_t)(16384 >> 4) << 16);
      const uint32_t bo = (uint32_t)((s >> 2) * ((N * 128) >> 4) + (s & 3) * 2);
      if (leader) mma_f16(tmem + 32, xa + ao, pa + bo, idesc_mn, s ? 1u : 0u);
    }
    // phase B': the same product with P as an MN-major operand (frames x Gaussians, 16 Gaussians = 32 bytes per row,
    // 32-byte swizzle: 8 frames x 32 B atoms, SBO = 256 B to the next 8 frames, LBO = 4096 B to the next 16 Gaussians)
    {
      const uint64_t hiA = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t hiB = ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)6 << 61);
      const uint32_t idesc_mn2 = idesc_mn | (1u << 16);
      for (int s = 0; s < 8; ++s) {
        const uint64_t da = hiA | (uint64_t)(((base + oX + s * 2048) >> 4) & 0x3FFF) | ((uint64_t)(16384 >> 4) << 16);
        const uint64_t db = hiB | (uint64_t)(((base + oP2 + s * 512) >> 4) & 0x3FFF) | ((uint64_t)(4096 >> 4) << 16);
        if (leader) mma_f16_desc(tmem + 64, da, db, idesc_mn2, s ? 1u : 0u);
      }
    }
    if (leader) commit(base + oBar);
  }
  mbar_wait(base + oBar, phase);
  phase ^= 1;
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int w = threadIdx.x >> 5, row = threadIdx.x;
    for (int c0 = 0; c0 < N; c0 += 16) {
      uint32_t r[16];
      ld16(tmem + ((uint32_t)(w * 32) << 16) + c0, r);
      for (int i = 0; i < 16; ++i) out_a[row * N + c0 + i] = __uint_as_float(r[i]);
      ld16(tmem + ((uint32_t)(w * 32) << 16) + 32 + c0, r);
      for (int i = 0; i < 16; ++i) out_b[row * N + c0 + i] = __uint_as_float(r[i]);
      ld16(tmem + ((uint32_t)(w * 32) << 16) + 64 + c0, r);
      for (int i = 0; i < 16; ++i) out_c[row * N + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- rate: `reps` groups of 16 phase-A-like MMAs (K-major) and 24 phase-B-like MMAs (MN-major A)
  if (warp_u == 0) {
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const uint32_t xa = ((base + oX) >> 4) & 0x3FFF, ba = ((base + oB) >> 4) & 0x3FFF, pa = ((base + oP) >> 4) & 0x3FFF;
    for (int mode = 0; mode < 2; ++mode) {
      long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          if (mode == 0) {
            const uint32_t ao = (uint32_t)((s >> 2) * (16384 >> 4) + (s & 3) * 2), bo = (uint32_t)((s >> 2) * ((N * 128) >> 4) + (s & 3) * 2);
            if (leader) mma_f16(tmem, xa + ao, ba + bo, idesc_k, s ? 1u : 0u);
          } else {
            const uint32_t ao = (uint32_t)(s * (2048 >> 4)) | ((uint32_t)(16384 >> 4) << 16);
            const uint32_t bo = (uint32_t)((s >> 2) * ((N * 128) >> 4) + (s & 3) * 2);
            if (leader) mma_f16(tmem + 32, xa + ao, pa + bo, idesc_mn, s ? 1u : 0u);
          }
        }
      }
      if (leader) commit(base + oBar);
      mbar_wait(base + oBar, phase);
      phase ^= 1;
      long long t1 = clock64();
      if (leader) cyc[mode] = (unsigned)((t1 - t0) * 100 / (reps * 8));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
  for (int N : {16, 32}) {
    std::vector<__half> X(128 * 128), B(N * 128), P(N * 128);
    srand(7 + N);
    auto rnd = [] { return (float)(rand() % 2001 - 1000) / 1000.0f; };
    for (auto &v : X) v = __float2half(rnd());
    for (auto &v : B) v = __float2half(rnd());
    for (auto &v : P) v = __float2half(fabsf(rnd()));
    __half *dX, *dB, *dP;
    float *da, *db, *dcc;
    unsigned *dc;
    cudaMalloc(&dX, X.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dP, P.size() * 2);
    cudaMalloc(&da, 128 * N * 4); cudaMalloc(&db, 128 * N * 4); cudaMalloc(&dcc, 128 * N * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dP, P.data(), P.size() * 2, cudaMemcpyHostToDevice);
    const size_t smem = 32768 + 4 * 32 * 128 + 2 * 4096 + 64 + 1024;
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    check_kernel<<<1, 128, smem>>>(dX, dB, dP, N, da, db, dcc, 2000, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: ERROR %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<float> ha(128 * N), hb(128 * N), hcc(128 * N);
    unsigned hc[2];
    cudaMemcpy(ha.data(), da, ha.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hb.data(), db, hb.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hcc.data(), dcc, hcc.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, dc, 8, cudaMemcpyDeviceToHost);
    double ea = 0, eb = 0, ec = 0;
    for (int t = 0; t < 128; ++t)
      for (int g = 0; g < N; ++g) {
        double ra = 0, rb = 0;
        for (int k = 0; k < 128; ++k) ra += (double)__half2float(X[t * 128 + k]) * __half2float(B[g * 128 + k]);
        for (int u = 0; u < 128; ++u) rb += (double)__half2float(X[u * 128 + t]) * __half2float(P[g * 128 + u]);  // row index t plays k
        ea = fmax(ea, fabs(ra - ha[t * N + g]));
        eb = fmax(eb, fabs(rb - hb[t * N + g]));
        ec = fmax(ec, fabs(rb - hcc[t * N + g]));
      }
    printf("N=%d: phase A (K-major X) max abs err %.3g; phase B (same tile as MN-major A) max abs err %.3g  -> %s\n", N, ea, eb,
           (ea < 1e-3 && eb < 1e-3) ? "OK" : "MISMATCH");
    printf("N=%d: phase B with P as an MN-major 32B-swizzle operand: max abs err %.3g -> %s\n", N, ec, ec < 1e-3 ? "OK" : "MISMATCH");
    printf("N=%d: %.1f cycles per K-major MMA, %.1f cycles per MN-major-A MMA (M=128, K=16, one CTA)\n", N, hc[0] / 100.0, hc[1] / 100.0);
  }
  return 0;
}
