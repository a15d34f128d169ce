// kaldi-hmm-gmm_b200/csrc/khg_stats_tc.cu — K3t: posteriors + sufficient statistics of frames bucketed by pdf on the
// 5th-generation tensor cores (tcgen05 / TMEM).  Same contract as stats_kernel (khg_kernels.cuh), i.e. per frame
// AccumAmDiagGmm::AccumulateForGmm (csrc/mle-am-diag-gmm.cc:41-52) -> AccumDiagGmm::AccumulateFromDiag
// (csrc/mle-diag-gmm.cc:145-158) -> DiagGmm::ComponentPosteriors (csrc/diag-gmm.cc:368-392) ->
// AccumulateFromPosteriors (csrc/mle-diag-gmm.cc:123-143), for a work item = up to 128 frames of ONE pdf.
//
// Both halves of the work are small GEMMs over the SAME shared-memory tile X[t][k] = [x | x^2 | 1 1] of the
// item's gathered frames (fp16 hi / lo split, per-dimension power-of-two scaling as in the dense fp16 kernel):
//   phase A  L[t][g] = sum_k X[t][k] * B_p[g][k]       log-likes of the pdf's Gaussians   (M = frames, K = columns)
//   phase B  S[k][g] = sum_t X[t][k] * post[t][g]      occ / mean / var statistics         (M = columns, K = frames)
// Phase A reads X as a K-major UMMA operand, phase B reads the same bytes as an MN-major operand (the canonical
// 128-byte-swizzle layouts of the two coincide: 8 frames x 128 bytes per atom; tools/umma_mn_check.cu checks this
// on the device), so the tile is written once.  Three-term split products as in K1: hi.hi + lo.hi + hi.lo, the
// two products that share the left operand ride in one instruction of doubled N ([P_hi ; P_lo]),
// because a tcgen05.mma with M = 128 costs ~51 (K-major A) / ~66 (MN-major A) cycles whether N is 16 or 64.
// S accumulates in TMEM (fp32) over consecutive items of a pdf and leaves as ONE fp64 atomic per statistic.
//
// A CTA is 4 warps (thread = frame row = TMEM lane); warp 0 also issues the MMAs.  Items whose features or
// weights do not fit fp16 after scaling (|x 2^-k| > 128, |w| > 8, non-finite) are not touched: their indices
// go to a device list that the fp32 kernel (stats_list_kernel) processes afterwards — no host round trip.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "khg_tc_common.cuh"

namespace khg {

// X tile, three 16 KB chunks of 64 fp16 columns: "virtual" columns 0..95 = hi parts of [x | x^2 | 1 1 | 0..],
// 96..191 = lo' parts (residual x 2^12) of [x | x^2 | 0..].  Phase A (screening) reads the hi columns only.
constexpr int kStkXBytes = 3 * kAChunkBytes;
constexpr int kStkLoCol0 = 96;                   // first virtual column of the lo' parts
constexpr int kStkMaxDP = 40;                    // 2 DP + 2 <= 96
constexpr int kStkMaxTilesPerFlush = 16;         // fp32 accumulation in TMEM (truncating adds) spans at most 2048 frames
constexpr float kStkPostScale = 4096.f;          // posteriors are stored x 2^12: keeps small ones off fp16's subnormals
constexpr float kStkLoScale = 4096.f;            // so are the residuals of x and x^2 (unscaled when S is flushed)
constexpr float kStkWeightLimit = 8.f;           // |w| * 4096 must stay inside fp16
#ifndef KHG_STK_REG_PREFETCH
#define KHG_STK_REG_PREFETCH 1
#endif
#ifndef KHG_STK_L2PF
#define KHG_STK_L2PF 1  // rows about four tiles ahead go to L2: 0 = no (-14 % at C4), 1 = prefetch.global.L2 per line, 2 = one bulk prefetch per row (-1 % at C4, -8 % at C5: profiles/r3q)
#endif
#ifndef KHG_STK_CTAS
#define KHG_STK_CTAS 3
#endif
#ifdef KHG_STK_TIMING  // per-phase cycle counters of thread 0, printed by one CTA (tools/ab_variant.sh x -DKHG_STK_TIMING)
#define STK_T(i) do { const long long _t = clock64(); if (tid == 0) tacc[i] += _t - tlast; tlast = _t; } while (0)
#else
#define STK_T(i) do { } while (0)
#endif
constexpr bool kStkRegPrefetch = KHG_STK_REG_PREFETCH != 0;  // 1: the next item's rows wait in registers (40 more of them)
constexpr float kStkScreen = 16.f;               // Gaussians within this of the frame's best (screened) log-like are re-evaluated in fp32

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// tcgen05.mma kind::f16 with the high words of the two shared-memory descriptors given separately (the posterior
// operand is MN-major with a 32-byte swizzle: other stride / layout fields than the X tile's)
__device__ __forceinline__ void stk_mma(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kDescHi), "r"(b_hi)
      : "memory");
}
constexpr uint32_t kStkPtDescHi = (uint32_t)(256 >> 4) | (1u << 14) | (6u << 29);  // SBO = 256 B (8 frames), version 1, SWIZZLE_32B
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2 *>(&u)); }

// hi unit and lo' unit (8 columns = 16 bytes each) of 8 values: hi = fp16(v), lo' = fp16((v - hi) * 2^12)
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack_h2(v[2 * j], v[2 * j + 1]);
    const float2 f = unpack_h2(h[j]);
    l[j] = pack_h2((v[2 * j] - f.x) * kStkLoScale, (v[2 * j + 1] - f.y) * kStkLoScale);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of the 16-byte unit holding virtual columns 8u..8u+7 of row t
__device__ __forceinline__ uint32_t x_unit_off(int u, int t) { return (uint32_t)((u >> 3) * kAChunkBytes + t * 128 + (((u & 7) ^ (t & 7)) << 4)); }

// Row pitch (floats) of the fp32 copy of the current pdf's parameters: a multiple of 4 whose 16-byte stride is odd,
// so that threads reading DIFFERENT Gaussians with LDS.128 spread over the banks
__host__ __device__ inline int stk_pitch(int dp) {
  int p = (dp + 3) & ~3;
  if (((p >> 2) & 1) == 0) p += 4;
  return p;
}

// The reference's fp32 arithmetic for ONE Gaussian (csrc/diag-gmm.cc:167-175: gconst + means_invvars . x
// - 0.5 inv_vars . x^2, sequential fused multiply-adds over the dimensions — the same sequence as stats_kernel).
// x and the parameters carry the per-dimension power-of-two scaling of the tile (x 2^-k, means_invvars 2^k,
// inv_vars 2^2k): every product and every rounding is the unscaled one.
template <int DP>
__device__ __forceinline__ float stk_exact_ll(const float *__restrict__ mrow, const float *__restrict__ vrow, float gc, const float (&x)[DP]) {
  float aa = 0.f, bb = 0.f;
#pragma unroll
  for (int q = 0; q < DP / 4; ++q) {
    const float4 m4 = *reinterpret_cast<const float4 *>(mrow + 4 * q);
    const float4 v4 = *reinterpret_cast<const float4 *>(vrow + 4 * q);
    aa = fmaf(m4.x, x[4 * q], aa);     bb = fmaf(v4.x, x[4 * q] * x[4 * q], bb);
    aa = fmaf(m4.y, x[4 * q + 1], aa); bb = fmaf(v4.y, x[4 * q + 1] * x[4 * q + 1], bb);
    aa = fmaf(m4.z, x[4 * q + 2], aa); bb = fmaf(v4.z, x[4 * q + 2] * x[4 * q + 2], bb);
    aa = fmaf(m4.w, x[4 * q + 3], aa); bb = fmaf(v4.w, x[4 * q + 3] * x[4 * q + 3], bb);
  }
  return (gc + aa) - 0.5f * bb;
}

// Softmax of the frame's NP (16 or 32) log-likes (csrc/eigen.cc:20-32), post *= w (csrc/mle-diag-gmm.cc:153), and
// the [P_hi ; P_lo] column of the frame in the posterior tile.  The tensor cores' log-likes (one-term fp16
// products, truncating fp32 accumulation: ~0.1 absolute at |terms| ~ 100) only SCREEN: every Gaussian within
// kStkScreen of the frame's best is re-evaluated in the reference's fp32 arithmetic; the others (posterior
// below 2e-7) keep the screened value.
template <int NP, int DP>
__device__ __forceinline__ float stk_softmax_store(uint32_t trow, int ng, float w, bool live, uint8_t *pt, int t, const float (&x)[DP],
                                                   const float *__restrict__ mf, const float *__restrict__ vf, const float *__restrict__ gcf, int pitch) {
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  float ll[NP];
#pragma unroll
  for (int c = 0; c < NP / 16; ++c) {
    TReg16 ra;
    tc_ld16_issue(trow + 16 * c, ra);
    tc_ld16_wait(ra);
#pragma unroll
    for (int i = 0; i < 16; ++i) ll[16 * c + i] = __uint_as_float(ra.r[i]);
  }
  float mx = -CUDART_INF_F;
#pragma unroll
  for (int g = 0; g < NP; ++g) mx = fmaxf(mx, ll[g]);  // (rows beyond the pdf's Gaussians are -inf: see stk_pack_kernel)
  if (live) {
    uint32_t todo = 0;
#pragma unroll
    for (int g = 0; g < NP; ++g)
      if (ll[g] >= mx - kStkScreen) todo |= 1u << g;
    while (todo) {
      const int g = __ffs(todo) - 1;
      todo &= todo - 1;
      const float e = stk_exact_ll<DP>(mf + g * pitch, vf + g * pitch, gcf[g], x);
#pragma unroll
      for (int j = 0; j < NP; ++j)
        if (j == g) ll[j] = e;
    }
    mx = ll[0];
#pragma unroll
    for (int g = 1; g < NP; ++g) mx = fmaxf(mx, ll[g]);
  }
  const float ml = mx * kLog2e;
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < NP; ++g) {
    ll[g] = fast_exp2(fmaf(ll[g], kLog2e, -ml));
    s += ll[g];
  }
  const float lse = fmaf(fast_log2(s), kLn2, mx);
  const float f = live ? (w * kStkPostScale) * __frcp_rn(s) : 0.f;
  // The posterior tile is an MN-major operand of the statistics GEMM (frames = K rows, 16 Gaussians = 32 contiguous
  // bytes per row, 32-byte swizzle: the canonical layout ((2,n),(8,k)):((1,LBO),(2,SBO)) of 16-byte units under
  // Swizzle<1,4,3>; tools/umma_mn_check.cu): atom r (Gaussians 16r..16r+15) at r * 4096 bytes, frame t at t * 32, the
  // two 16-byte units of a row swapped when bit 2 of t is set.  Atoms [0, NP/16) = P_hi, [NP/16, NP/8) = P_lo: the
  // frame's column is 2 (4) 128-bit stores per part instead of 16 (32) 16-bit ones.
  uint8_t *row = pt + t * 32;
  const uint32_t sw = (uint32_t)((t >> 2) & 1) << 4;
#pragma unroll
  for (int r = 0; r < NP / 16; ++r) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = ll[16 * r + 2 * j] * f, p1 = ll[16 * r + 2 * j + 1] * f;  // (dead rows: f = 0 and the exponentials are <= 1)
      h[j] = pack_h2(p0, p1);
      const float2 hf = unpack_h2(h[j]);
      l[j] = pack_h2(p0 - hf.x, p1 - hf.y);
    }
    *reinterpret_cast<uint4 *>(row + r * 4096 + (0u ^ sw)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(row + r * 4096 + (16u ^ sw)) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4 *>(row + (NP / 16 + r) * 4096 + (0u ^ sw)) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4 *>(row + (NP / 16 + r) * 4096 + (16u ^ sw)) = make_uint4(l[4], l[5], l[6], l[7]);
  }
  return lse;
}

// Per-pdf model image (one bulk copy): fp16 operand rows (2 chunks x NP rows x 128 B, swizzled), then the same
// parameters in fp32 with the tile's scaling (NP x pitch means_invvars 2^k, NP x pitch inv_vars 2^2k, 32 gconsts).
__host__ __device__ inline int stk_f16_bytes(int np) { return 2 * np * 128; }
__host__ __device__ inline int stk_f32_bytes(int np, int dp) { return (2 * np * stk_pitch(dp) + 32) * 4; }
__host__ __device__ inline int stk_img_bytes(int np, int dp) { return stk_f16_bytes(np) + stk_f32_bytes(np, dp); }
__host__ __device__ inline int stk_img_units(int np, int dp) { return (stk_img_bytes(np, dp) + 1023) / 1024; }  // 1024-byte units in HBM
__host__ __device__ inline int stk_pt_bytes(int np) { return 2 * 2 * np * 128; }
constexpr int kStkDescStage = 128;  // item descriptors staged in shared memory at a time (+ 2 of look-ahead)
// shared-memory bytes: X tile | posterior tile | model image | tables, staged descriptors, barriers
__host__ __device__ inline int stk_smem_bytes(int np_max, int dp) {
  return kStkXBytes + 1024 + stk_pt_bytes(np_max) + stk_img_bytes(np_max, dp) + 128 * 4 + (kStkDescStage + 8) * 16 + 64 + 64;
}

// NU = DP / 8, DP = D rounded up to 8: columns [0, DP) = x, [DP, 2 DP) = x^2, 2 DP and 2 DP + 1 = 1;
// NPM = 16 or 32: operand rows of the model's largest pdf (models of small pdfs get the leaner instantiation)
template <int NU, int NPM>
__global__ void __launch_bounds__(128, NPM == 16 ? KHG_STK_CTAS : 2) stats_tc_kernel(StatsTcArgs a) {
  constexpr int DP = 8 * NU;
  constexpr int kHiSteps = (2 * DP + 2 + 15) / 16;  // K steps (16 columns) of phase A
  static_assert(DP <= kStkMaxDP, "hi columns must fit 96");
  extern __shared__ uint8_t stk_smem_raw[];
  const uint32_t raw = smem_u32(stk_smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *bp = stk_smem_raw + (base - raw);
  const int pitch = stk_pitch(DP);
  // (1 KB of slack after X: the second statistics MMA reads one chunk past the lo' columns; those rows are never used)
  uint8_t *X = bp, *Pt = X + kStkXBytes + 1024, *Bm = Pt + stk_pt_bytes(a.np_max);
  float *uns = reinterpret_cast<float *>(Bm + stk_img_bytes(a.np_max, DP));  // 128 floats
  int4 *s_desc = reinterpret_cast<int4 *>(uns + 128);           // kStkDescStage + 8 item descriptors
  double *s_red = reinterpret_cast<double *>(s_desc + kStkDescStage + 8);  // 8 doubles
  const uint32_t sX = base, sPt = sX + kStkXBytes + 1024, sBm = sPt + stk_pt_bytes(a.np_max);
  const uint32_t sBarA = smem_u32(s_red + 8), sBarB = sBarA + 8, sBarM = sBarB + 8, sSlot = sBarM + 8;
  volatile int *s_flag = reinterpret_cast<volatile int *>(s_red + 8) + 7;  // set by a thread whose row does not fit fp16
  const int tid = threadIdx.x, D = a.D;
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);

  // ---- one-off: zero X and what the MMAs may read past it, the two 1-columns, tables
  for (int i = tid; i < (kStkXBytes + 1024 + stk_pt_bytes(a.np_max) + stk_img_bytes(a.np_max, DP)) / 16; i += 128)
    reinterpret_cast<uint4 *>(X)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  {
    const int c0 = 2 * DP, c1 = 2 * DP + 1;
    *reinterpret_cast<__half *>(X + x_unit_off(c0 >> 3, tid) + ((c0 & 7) << 1)) = __float2half(1.f);
    *reinterpret_cast<__half *>(X + x_unit_off(c1 >> 3, tid) + ((c1 & 7) << 1)) = __float2half(1.f);
  }
  uns[tid] = a.unscale[tid];
  if (tid == 0) {
    *s_flag = 0;
    mbar_init(sBarA, 1);
    mbar_init(sBarB, 1);
    mbar_init(sBarM, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp_u == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sSlot), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t *>(reinterpret_cast<uint8_t *>(s_red + 8) + 24);
  // TMEM columns: [0, 32) screened log-likes L, [32, 96) S_a = (hi | lo'[0:32])^T [P_hi ; P_lo], [96, 128) S_c = lo'[32:96]^T P_hi
  const uint32_t trow = tmem + ((uint32_t)(warp_u * 32) << 16);  // this warp's lane quadrant

  // which statistic virtual column `tid` is: x columns -> mean, x^2 columns -> var, the first 1-column -> occ
  double *row_dst = nullptr;
  int row_stride = 0;
  if (tid < D) { row_dst = a.mean ? a.mean + tid : nullptr; row_stride = D; }
  else if (tid >= DP && tid < DP + D) { row_dst = a.var ? a.var + (tid - DP) : nullptr; row_stride = D; }
  else if (tid == 2 * DP) { row_dst = a.occ; row_stride = 1; }
  const float row_unscale = uns[tid];

  const int n_items = a.item_start[a.P];
  const int i0 = (int)((int64_t)blockIdx.x * n_items / gridDim.x), i1 = (int)((int64_t)(blockIdx.x + 1) * n_items / gridDim.x);
  const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.feats) & 15) == 0);
  const bool vec256 = (D & 7) == 0 && ((reinterpret_cast<uintptr_t>(a.feats) & 31) == 0);
  const bool chunked = !vec && ((reinterpret_cast<uintptr_t>(a.feats) & 15) == 0);  // 4-byte aligned rows, 16-byte aligned buffer

  // Gather pipeline.  While item `it` is computed: the ROWS of item it+1 are in flight into xr[] (their index was
  // loaded one item earlier: the address depends on a loaded value and a thread cannot issue past a dependent
  // instruction), the INDEX of item it+2 is in flight into idx2, and the rows about four tiles ahead in the sorted
  // order are being pulled into L2 (prefetch: position-based, item boundaries do not matter for it), so that the
  // register loads are L2 hits.
  float xr[DP + 4];  // (rows that are only 4-byte aligned arrive as aligned 16-byte chunks: up to 3 leading words)
  int xoff1 = 0;     // word offset of the row inside its first chunk (0 on the aligned paths)
  float w1 = 0.f;
  int idx1 = -1, idx2 = -1;
  // item descriptors of [stage0, stage0 + kStkDescStage + 2) wait in shared memory: a descriptor read never
  // stalls a thread on a global load
  int stage0 = i0;
  auto stage_descs = [&](int first) {
    stage0 = first;
    for (int k = tid; k < kStkDescStage + 2; k += 128)
      if (first + k < i1) s_desc[k] = __ldg(a.item_desc + first + k);
    __syncthreads();
  };
  auto load_index = [&](int it) -> int {
    if (it >= i1) return -1;
    const int4 d = s_desc[it - stage0];
    return tid < d.z ? __ldg(a.order + d.y + tid) : -1;
  };
  auto load_rows = [&](int idx) {
    w1 = 0.f;
    if (idx < 0) return;
    w1 = a.weights ? __ldg(a.weights + idx) : 1.0f;
    const float *src = a.feats + (size_t)idx * D;
    if (vec256) {
      // 256-bit loads: the 32 lanes of a warp read 32 different rows, so every load instruction costs 32 tag lookups in
      // L1TEX whatever its width — half as many instructions, half the lookups (1280 -> 640 per 128-frame item)
#pragma unroll
      for (int q = 0; q < DP / 8; ++q) {
        if (8 * q < D) {
          asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=f"(xr[8 * q]), "=f"(xr[8 * q + 1]), "=f"(xr[8 * q + 2]), "=f"(xr[8 * q + 3]), "=f"(xr[8 * q + 4]),
                         "=f"(xr[8 * q + 5]), "=f"(xr[8 * q + 6]), "=f"(xr[8 * q + 7])
                       : "l"(src + 8 * q));
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) xr[8 * q + j] = 0.f;
        }
      }
    } else if (vec) {
#pragma unroll
      for (int q = 0; q < DP / 4; ++q) {
        if (4 * q < D) {
          const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + q);
          xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
        } else {
          xr[4 * q] = xr[4 * q + 1] = xr[4 * q + 2] = xr[4 * q + 3] = 0.f;
        }
      }
    } else if (chunked && idx != a.n_frames - 1) {
      // rows of dim 39 (13 x 3) and the like start on any 4-byte boundary: 39 scalar loads per row cost 39 x 32 tag
      // lookups per warp.  The ALIGNED 16-byte chunks that cover the row are loaded instead (11 at dim 39), the row's
      // word offset inside the first chunk is kept, and the words are picked by two levels of selects when the row is
      // consumed.  (The buffer's last row keeps the scalar loads: its last chunk would end past the buffer.)
      const int o = (int)(((size_t)idx * D) & 3);
      xoff1 = o;
      const float4 *c4 = reinterpret_cast<const float4 *>(src - o);
#pragma unroll
      for (int q = 0; q < DP / 4 + 1; ++q) {
        if (4 * q < o + D) {
          const float4 v = __ldg(c4 + q);
          xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w;
        } else {
          xr[4 * q] = xr[4 * q + 1] = xr[4 * q + 2] = xr[4 * q + 3] = 0.f;
        }
      }
    } else {
      xoff1 = 0;
#pragma unroll
      for (int d2 = 0; d2 < DP; ++d2) xr[d2] = d2 < D ? __ldg(src + d2) : 0.f;
    }
  };
  int pf_idx = -1;  // index whose row goes to L2 next
  auto l2_pipeline = [&](int pos) {
    if (KHG_STK_L2PF != 0 && pf_idx >= 0) {
      const char *r = reinterpret_cast<const char *>(a.feats + (size_t)pf_idx * D);
      if (KHG_STK_L2PF == 2 && vec) {
        // one bulk prefetch of the whole row (the TMA unit's path: no tag lookups in L1TEX, where the 32 rows of a warp's
        // prefetch instruction cost 32 each)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(r), "r"(4 * D) : "memory");
      } else {
        prefetch_l2(r);
        prefetch_l2(r + 4 * D - 4);
      }
      if (a.weights) prefetch_l2(a.weights + pf_idx);
    }
    pf_idx = pos < a.n_frames ? __ldg(a.order + pos) : -1;
  };
#pragma unroll
  for (int d2 = 0; d2 < DP + 4; ++d2) xr[d2] = 0.f;

  const float *mf = nullptr, *vf = nullptr, *gcf = nullptr;  // fp32 parameters of the current pdf inside its image
  int cur_p = -1, g0 = 0, ng = 0, NP = 16, acc_tiles = 0;  // (NP stays 16 when NPM == 16: the compiler folds it)
  uint32_t ph_a = 0, ph_b = 0, ph_m = 0;
  bool b_pending = false, m_pending = false, big_pdf = false;
  double my_like = 0.0, my_w = 0.0;

  auto wait_b = [&]() {
    if (b_pending) {
      mbar_wait(sBarB, ph_b);
      ph_b ^= 1;
      b_pending = false;
      tc_fence_after();
    }
  };
  // S -> the fp64 accumulators.  Lane r of S_a holds virtual column r (hi parts for r < 96, lo'[r - 96] above), lane
  // r < 64 of S_c holds lo'[32 + r]: the lo' rows go through shared memory (the posterior tile is free) to the lanes
  // of their hi parts; every statistic leaves as ONE atomic.
  auto flush = [&]() {
    wait_b();
    float *scr = reinterpret_cast<float *>(Pt);  // [96][NP]
    float tot[NPM];
#pragma unroll
    for (int c = 0; c < NPM; c += 16) {
      if (c >= NP) break;  // (warp-uniform: the loads are .sync.aligned)
      TReg16 ra, rb, rc;
      tc_ld16_issue(trow + 32 + c, ra);
      tc_ld16_issue(trow + 32 + NP + c, rb);
      tc_ld16_wait2(ra, rb);
      tc_ld16_issue(trow + 96 + c, rc);
      tc_ld16_wait(rc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float v = __uint_as_float(ra.r[i]) + __uint_as_float(rb.r[i]);
        tot[c + i] = v;
        if (tid >= kStkLoCol0) scr[(tid - kStkLoCol0) * NP + c + i] = v;
        if (tid < 64) scr[(32 + tid) * NP + c + i] = __uint_as_float(rc.r[i]);
      }
    }
    tc_fence_before();
    __syncthreads();
    if (row_dst != nullptr) {
#pragma unroll
      for (int i = 0; i < NPM; ++i)
        if (i < ng) {
          const float v = fmaf(scr[tid * NP + i], 1.0f / kStkLoScale, tot[i]);
          atomicAdd(row_dst + (size_t)(g0 + i) * row_stride, (double)(v * row_unscale));
        }
    }
    acc_tiles = 0;
  };

  stage_descs(i0);
  idx1 = load_index(i0);
  if (kStkRegPrefetch) load_rows(idx1);
  idx2 = load_index(i0 + 1);
  int pf_pos = i0 < i1 ? s_desc[0].y + 3 * 128 + tid : 0;
#ifdef KHG_STK_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
  for (int it = i0; it < i1; ++it) {
    STK_T(7);
    if (it - stage0 == kStkDescStage) {
      __syncthreads();  // (every thread has read its descriptors of the previous stage)
      stage_descs(it);
    }
    const int4 desc = s_desc[it - stage0];
    const int p = desc.x, n = desc.z;
    if (p != cur_p) {
      if (acc_tiles > 0) flush();
      if (m_pending) {  // (a skipped item may have left its pdf's image in flight)
        mbar_wait(sBarM, ph_m);
        ph_m ^= 1;
        m_pending = false;
      }
      cur_p = p;
      g0 = a.offsets[p];
      ng = a.offsets[p + 1] - g0;
      NP = (NPM == 16 || ng <= 16) ? 16 : 32;
      __syncthreads();  // every thread is past the previous pdf's phase A, softmax and flush: the image and Pt are free
      const int ioff = __ldg(a.img_off + p);
      big_pdf = ioff < 0;  // more than 32 Gaussians: no image; the pdf's items are declined (fp32 kernel), like out-of-range rows
      if (!big_pdf) {
        if (tid == 0) {
          const uint32_t bytes = (uint32_t)stk_img_bytes(NP, DP);
          mbar_expect_tx(sBarM, bytes);
          bulk_load(sBm, a.img + (size_t)ioff * 1024, bytes, sBarM);
        }
        m_pending = true;
      }
      mf = reinterpret_cast<const float *>(Bm + stk_f16_bytes(NP));  // (the fp32 part follows the operand rows)
      vf = mf + NP * pitch;
      gcf = vf + NP * pitch;
    }
    STK_T(0);  // pdf change: flush, image load
    // ---- this item's rows (prefetched), scaled; range check (NaN and Inf fail it too)
    const bool live = tid < n;
    if (!kStkRegPrefetch) load_rows(idx1);  // (an L2 hit: the row was prefetched about four tiles ago)
    const float w = w1;
    const int idx = idx1;
    float xs[DP];
    bool bad = false;
    if (chunked) {  // word d of the row is word d + offset of the loaded chunks: two levels of selects (offset bit 0, bit 1)
      const bool o1 = (xoff1 & 1) != 0, o2 = (xoff1 & 2) != 0;
      float u1[DP + 2];
#pragma unroll
      for (int d2 = 0; d2 < DP + 2; ++d2) u1[d2] = o1 ? xr[d2 + 1] : xr[d2];
#pragma unroll
      for (int d2 = 0; d2 < DP; ++d2) {
        const float v = o2 ? u1[d2 + 2] : u1[d2];
        xs[d2] = (d2 < D ? v : 0.f) * a.asc_c[d2];
      }
    } else {
#pragma unroll
      for (int d2 = 0; d2 < DP; ++d2) xs[d2] = xr[d2] * a.asc_c[d2];  // (a kernel parameter: a constant-bank operand of the FMUL, no load)
    }
#pragma unroll
    for (int d2 = 0; d2 < DP; ++d2) bad |= !(fabsf(xs[d2]) <= kF16FeatLimit);
    bad = live && (bad || big_pdf || !(fabsf(w) <= kStkWeightLimit));
    // next item's rows (index loaded one item ago), the index after that, L2 prefetch further ahead
    idx1 = idx2;
    if (kStkRegPrefetch) load_rows(idx1);
    idx2 = load_index(it + 2);
    l2_pipeline(pf_pos);
    pf_pos += n;
    STK_T(1);  // scale, check, next loads issued
    wait_b();  // phase B of the previous item has read X and Pt
    STK_T(2);  // wait for phase B of the previous item
    if (bad) *s_flag = 1;  // (the item is declined as a whole, below; an out-of-range row is never written to the tile)
    if (live && !bad) {
      uint8_t *xrow = X + tid * 128;
      const uint32_t t7s = (uint32_t)(tid & 7) << 4;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        float v[8], q[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = xs[8 * u + j];
          q[j] = v[j] * v[j];  // data.array().square(), csrc/diag-gmm.cc:175 (the scaling is a power of two: exact)
        }
        // unit index -> chunk (u >> 3) and swizzled 16-byte slot ((u & 7) << 4) ^ t7s of the thread's row
        constexpr int ul = kStkLoCol0 / 8;
        uint4 hi, lo;
        split8(v, hi, lo);
        *reinterpret_cast<uint4 *>(xrow + (u >> 3) * kAChunkBytes + ((((u) & 7) << 4) ^ t7s)) = hi;                  // x: virtual columns 8u..
        *reinterpret_cast<uint4 *>(xrow + ((ul + u) >> 3) * kAChunkBytes + ((((ul + u) & 7) << 4) ^ t7s)) = lo;
        split8(q, hi, lo);
        *reinterpret_cast<uint4 *>(xrow + ((NU + u) >> 3) * kAChunkBytes + ((((NU + u) & 7) << 4) ^ t7s)) = hi;      // x^2: virtual columns DP + 8u..
        *reinterpret_cast<uint4 *>(xrow + ((ul + NU + u) >> 3) * kAChunkBytes + ((((ul + NU + u) & 7) << 4) ^ t7s)) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (m_pending) {
      mbar_wait(sBarM, ph_m);
      ph_m ^= 1;
      m_pending = false;
    }
    tc_fence_before();
    __syncthreads();
    STK_T(3);  // build + barrier
    if (*s_flag != 0) {  // some row of the item is outside fp16's range: the whole item goes to the fp32 kernel
      __syncthreads();   // (every thread has read the flag)
      if (tid == 0) {
        *s_flag = 0;
        a.fb_items[atomicAdd(a.fb_count, 1)] = it;
      }
      __syncthreads();
      continue;
    }
    // ---- phase A (screening): L = X_hi . B_hi^T, N = NP
    if (warp_u == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xa = umma_desc_lo(sX), bm = umma_desc_lo(sBm);
        const uint32_t bchunk = (uint32_t)(NP * 128) >> 4;
        const uint32_t id1 = (1u << 4) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
        for (int s = 0; s < kHiSteps; ++s)
          tc_mma<true>(tmem, xa + (s >> 2) * (kAChunkBytes >> 4) + (s & 3) * 2, bm + (s >> 2) * bchunk + (s & 3) * 2, id1, s ? 1u : 0u);
        tc_commit(sBarA);
      }
      __syncwarp();
    }
    mbar_wait(sBarA, ph_a);
    ph_a ^= 1;
    tc_fence_after();
    STK_T(4);  // phase A: issue + wait
    float lse;
    if (NPM == 16 || NP == 16) lse = stk_softmax_store<16, DP>(trow, ng, w, live, Pt, tid, xs, mf, vf, gcf, pitch);
    else lse = stk_softmax_store<32, DP>(trow, ng, w, live, Pt, tid, xs, mf, vf, gcf, pitch);
    if (live) {
      if (!(fabsf(lse) <= 3.402823466e38f)) atomicOr(a.err, ERR_NONFINITE);
      if (a.per_frame) a.per_frame[idx] = lse;
      my_like += (double)(lse * w);   // csrc/mle-am-diag-gmm.cc:49-50
      my_w += (double)w;
    }
    tc_fence_before();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    STK_T(5);  // softmax, exact re-evaluation, posterior tile + barrier
    // ---- phase B: S_a (+)= X[chunks 0, 1]^T . [P_hi | P_lo] (N = 2 NP), S_c (+)= X[chunk 2, ...]^T . P_hi (N = NP); both
    // operands MN-major: X 16 frames = 2048 bytes per K step, LBO = one chunk (the next 64 virtual columns)
    if (warp_u == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t lbo = (uint32_t)(kAChunkBytes >> 4) << 16;
        const uint32_t xa = umma_desc_lo(sX) | lbo, xc = umma_desc_lo(sX + 2 * kAChunkBytes) | lbo;
        const uint32_t pt = umma_desc_lo(sPt) | ((uint32_t)(4096 >> 4) << 16);  // LBO = 4096 B: the next 16 Gaussians
        const uint32_t id2 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(2 * NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t id1 = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t acc0 = acc_tiles > 0 ? 1u : 0u;
#pragma unroll
        for (int s = 0; s < 8; ++s)  // 16 frames per K step: 2048 bytes of X, 512 bytes of the posterior tile
          stk_mma(tmem + 32, xa + s * (2048 >> 4), pt + s * (512 >> 4), kStkPtDescHi, id2, s ? 1u : acc0);
#pragma unroll
        for (int s = 0; s < 8; ++s)
          stk_mma(tmem + 96, xc + s * (2048 >> 4), pt + s * (512 >> 4), kStkPtDescHi, id1, s ? 1u : acc0);
        tc_commit(sBarB);
      }
      __syncwarp();
    }
    STK_T(6);  // phase B issue (the issuing thread stalls behind the tensor pipe's queue: 16 MMAs x ~66 cycles)
    b_pending = true;
    if (++acc_tiles == kStkMaxTilesPerFlush) {
      flush();
      __syncthreads();  // (the flush scratch is the posterior tile)
    }
  }
  if (acc_tiles > 0) flush();
  wait_b();
  if (m_pending) mbar_wait(sBarM, ph_m);
#ifdef KHG_STK_TIMING
  if (tid == 0 && blockIdx.x == 7 && i1 > i0)
    printf("K3t CTA 7: %d items; cycles per item: pdf-change %lld | scale+issue %lld | wait B %lld | build+bar %lld | phase A %lld | softmax+bar %lld | issue B %lld | loop %lld\n",
           i1 - i0, tacc[0] / (i1 - i0), tacc[1] / (i1 - i0), tacc[2] / (i1 - i0), tacc[3] / (i1 - i0), tacc[4] / (i1 - i0), tacc[5] / (i1 - i0),
           tacc[6] / (i1 - i0), tacc[7] / (i1 - i0));
#endif

  for (int off = 16; off > 0; off >>= 1) {
    my_like += __shfl_xor_sync(0xffffffffu, my_like, off);
    my_w += __shfl_xor_sync(0xffffffffu, my_w, off);
  }
  if ((tid & 31) == 0) {
    s_red[tid >> 5] = my_like;
    s_red[4 + (tid >> 5)] = my_w;
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    const double L = s_red[0] + s_red[1] + s_red[2] + s_red[3], W = s_red[4] + s_red[5] + s_red[6] + s_red[7];
    if (W != 0.0 || L != 0.0) {
      atomicAdd(&a.totals[0], L);
      atomicAdd(&a.totals[1], W);
      if (a.call_like) atomicAdd(a.call_like, L);
    }
  }
  if (warp_u == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

// ------------------------------------------------------------------ model images --
// k_d = round(log2(rms of feature d implied by the model)), as the dense fp16 pack (khg_loglikes_tc.cu)
__global__ void __launch_bounds__(256) stk_scale_kernel(int G, int D, int DP, const float *__restrict__ miv, const float *__restrict__ iv,
                                                        float *__restrict__ ascale, float *__restrict__ unscale) {
  const int d = blockIdx.x;
  __shared__ double red[256];
  double s = 0.0;
  for (int g = threadIdx.x; g < G; g += 256) {
    const double var = 1.0 / (double)iv[(size_t)g * D + d], mean = (double)miv[(size_t)g * D + d] * var;
    s += mean * mean + var;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double rms = sqrt(red[0] / G);
    int k = (rms > 0.0 && isfinite(rms)) ? (int)lround(log2(rms)) : 0;
    k = max(-20, min(20, k));
    ascale[d] = ldexpf(1.0f, -k);
    unscale[d] = ldexpf(1.0f, k) / kStkPostScale;
    unscale[DP + d] = ldexpf(1.0f, 2 * k) / kStkPostScale;
    if (d == 0) {
      unscale[2 * DP] = 1.0f / kStkPostScale;
      unscale[2 * DP + 1] = 0.f;
    }
  }
}

// image of pdf p: chunk c (columns 64c..64c+63) at c * NP * 128 bytes, row g = fp16 of the Gaussian's operand row
// [miv 2^k | -iv/2 2^2k | gconst, its fp16 residual | 0], 128-byte swizzle
__global__ void __launch_bounds__(128) stk_pack_kernel(int P, int D, int DP, const int32_t *__restrict__ offsets, const int32_t *__restrict__ img_off,
                                                       const float *__restrict__ miv, const float *__restrict__ iv, const float *__restrict__ gconsts,
                                                       const float *__restrict__ ascale, uint8_t *__restrict__ img, int *__restrict__ flag) {
  const int p = blockIdx.x;
  const int g0 = offsets[p], ng = offsets[p + 1] - g0, NP = ng <= 16 ? 16 : 32;
  if (img_off[p] < 0) return;  // (more than 32 Gaussians: the fp32 kernel's pdf)
  uint8_t *o = img + (size_t)img_off[p] * 1024;
  bool bad = false;
  for (int e = threadIdx.x; e < NP * 128; e += 128) {
    const int g = e >> 7, k = e & 127;
    float v = 0.f;
    bool gc_lo = false;
    if (g < ng) {
      const size_t r = (size_t)(g0 + g) * D;
      if (k < D) v = miv[r + k] / ascale[k];
      else if (k >= DP && k < DP + D) v = (-0.5f * iv[r + k - DP]) / (ascale[k - DP] * ascale[k - DP]);
      else if (k == 2 * DP) v = gconsts[g0 + g];
      else if (k == 2 * DP + 1) { v = gconsts[g0 + g]; gc_lo = true; }
    }
    if (!(fabsf(v) <= 3.0e4f)) bad = true;  // also -inf gconsts and NaN: such models stay on the fp32 kernel
    __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    // rows beyond the pdf's Gaussians: gconst = -inf, so that their screened log-like is -inf (x 1 in the tile's
    // 1-column; every other entry of the row is 0) and the softmax needs no index mask
    if (g >= ng && k == 2 * DP) hi = __float2half_rn(-CUDART_INF_F);
    const uint32_t off = (uint32_t)((k >> 6) * (NP * 128) + g * 128 + ((((k & 63) >> 3) ^ (g & 7)) << 4) + ((k & 7) << 1));
    *reinterpret_cast<__half *>(o + off) = gc_lo ? lo : hi;
  }
  // fp32 part (exact power-of-two scaling): means_invvars 2^k, inv_vars 2^2k, gconsts; pad entries zero
  const int pitch = stk_pitch(DP);
  float *f = reinterpret_cast<float *>(o + stk_f16_bytes(NP));
  for (int e = threadIdx.x; e < 2 * NP * pitch + 32; e += 128) {
    float v = 0.f;
    if (e < 2 * NP * pitch) {
      const int which = e / (NP * pitch), r = e - which * NP * pitch, g = r / pitch, d = r - g * pitch;
      if (g < ng && d < D) v = which == 0 ? miv[(size_t)(g0 + g) * D + d] / ascale[d] : iv[(size_t)(g0 + g) * D + d] / (ascale[d] * ascale[d]);
    } else if (e - 2 * NP * pitch < ng) {
      v = gconsts[g0 + e - 2 * NP * pitch];
    }
    f[e] = v;
  }
  if (bad) atomicOr(flag, 1);
}

void stats_tc_free(khg_model *m) {
  StatsTcPack &t = m->stk;
  cudaFree(t.img); cudaFree(t.img_off); cudaFree(t.ascale); cudaFree(t.unscale); cudaFree(t.fb_count);
  t = StatsTcPack();
}

// dim <= 40 and at least half of the Gaussians in pdfs of at most 32 (larger pdfs keep the fp32 kernel, item by item)
bool stats_tc_shape_ok(const khg_model *m) {
  if (m->dim < 1 || m->dim > kStkMaxDP) return false;
  if (m->max_gp <= 32) return true;
  int64_t small = 0;
  for (int p = 0; p < m->P; ++p) {
    const int ng = m->h_offsets[p + 1] - m->h_offsets[p];
    if (ng <= 32) small += ng;
  }
  return 2 * small >= (int64_t)m->G;
}

khg_status stats_tc_build(khg_model *m) {
  StatsTcPack &t = m->stk;
  if (t.tried) return KHG_OK;
  t.tried = true;
  if (!stats_tc_shape_ok(m)) return KHG_OK;
  const int D = m->dim, P = m->P;
  t.DP = (D + 7) / 8 * 8;
  std::vector<int32_t> off(P + 1);
  int64_t run = 0;
  t.np_max = 16;
  t.partial = false;
  for (int p = 0; p < P; ++p) {
    const int ng = m->h_offsets[p + 1] - m->h_offsets[p];
    if (ng > 32) {
      off[p] = -1;
      t.partial = true;
      continue;
    }
    off[p] = (int32_t)run;
    if (ng > 16) t.np_max = 32;
    run += stk_img_units(ng <= 16 ? 16 : 32, t.DP);
  }
  off[P] = (int32_t)run;
  if (run == 0) return KHG_OK;
  if (run > (int64_t)1 << 30) return KHG_OK;
  int *d_flag = nullptr;
  cudaStream_t st = m->stream;
  KHG_CUDA_TRY(cudaMalloc(&t.img, (size_t)run * 1024));
  KHG_CUDA_TRY(cudaMalloc(&t.img_off, sizeof(int32_t) * (P + 1)));
  KHG_CUDA_TRY(cudaMalloc(&t.ascale, sizeof(float) * 64));
  KHG_CUDA_TRY(cudaMalloc(&t.unscale, sizeof(float) * 128));
  KHG_CUDA_TRY(cudaMalloc(&t.fb_count, sizeof(int) * 2));
  KHG_CUDA_TRY(cudaMemcpyAsync(t.img_off, off.data(), sizeof(int32_t) * (P + 1), cudaMemcpyHostToDevice, st));
  KHG_CUDA_TRY(cudaMemsetAsync(t.unscale, 0, sizeof(float) * 128, st));
  KHG_CUDA_TRY(cudaMemsetAsync(t.ascale, 0, sizeof(float) * 64, st));
  KHG_CUDA_TRY(cudaMemsetAsync(t.fb_count, 0, sizeof(int) * 2, st));
  d_flag = t.fb_count + 1;
  stk_scale_kernel<<<D, 256, 0, st>>>(m->G, D, t.DP, m->d_miv, m->d_iv, t.ascale, t.unscale);
  stk_pack_kernel<<<P, 128, 0, st>>>(P, D, t.DP, m->d_offsets, t.img_off, m->d_miv, m->d_iv, m->d_gconsts, t.ascale, t.img, d_flag);
  g_launch_count += 2;
  KHG_CUDA_TRY(cudaGetLastError());
  int flag = 0;
  KHG_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(t.h_ascale, t.ascale, sizeof(float) * kStkMaxDP, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaStreamSynchronize(st));  // (pageable `off` and `flag` are done with)
  t.ready = flag == 0;
  return KHG_OK;
}

template <int NU, int NPM>
static khg_status stk_launch_nu(khg_model *m, const StatsTcArgs &a, cudaStream_t st) {
  const size_t smem = (size_t)stk_smem_bytes(a.np_max, 8 * NU) + 1024;  // (+ alignment slack)
  KHG_CUDA_TRY(cudaFuncSetAttribute(stats_tc_kernel<NU, NPM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  KHG_CUDA_TRY(cudaFuncSetAttribute(stats_tc_kernel<NU, NPM>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // persistent CTAs over contiguous item ranges; two per SM so that one's gather / softmax runs under the other's MMAs
  int per_sm = a.np_max <= 16 ? KHG_STK_CTAS : 2;
  if (const char *e = getenv("KHG_STATS_TC_CTAS_PER_SM")) per_sm = std::max(1, std::min(4, atoi(e)));  // experiments
  if (getenv("KHG_STATS_TC_DEBUG")) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stats_tc_kernel<NU, NPM>, 128, smem);
    fprintf(stderr, "stats_tc_kernel<%d,%d>: smem %zu B, occupancy %d CTAs/SM, grid %d\n", NU, NPM, smem, occ, m->sm_count * per_sm);
  }
  stats_tc_kernel<NU, NPM><<<m->sm_count * per_sm, 128, smem, st>>>(a);
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  return KHG_OK;
}

// fills a.img / img_off / ascale / unscale / fb_count from the model's pack and launches; the caller provides
// fb_items (capacity = number of items) and runs the fp32 list kernel on {fb_items, fb_count} afterwards
khg_status stats_tc_launch(khg_model *m, StatsTcArgs a, cudaStream_t st) {
  const StatsTcPack &t = m->stk;
  a.img = t.img;
  a.img_off = t.img_off;
  a.ascale = t.ascale;
  a.unscale = t.unscale;
  a.fb_count = t.fb_count;
  for (int d = 0; d < kStkMaxDP; ++d) a.asc_c[d] = t.h_ascale[d];
  a.miv = m->d_miv;
  a.iv = m->d_iv;
  a.gconsts = m->d_gconsts;
  a.np_max = t.np_max;
  KHG_CUDA_TRY(cudaMemsetAsync(t.fb_count, 0, sizeof(int), st));
  switch (t.DP / 8) {
    case 1: return a.np_max == 16 ? stk_launch_nu<1, 16>(m, a, st) : stk_launch_nu<1, 32>(m, a, st);
    case 2: return a.np_max == 16 ? stk_launch_nu<2, 16>(m, a, st) : stk_launch_nu<2, 32>(m, a, st);
    case 3: return a.np_max == 16 ? stk_launch_nu<3, 16>(m, a, st) : stk_launch_nu<3, 32>(m, a, st);
    case 4: return a.np_max == 16 ? stk_launch_nu<4, 16>(m, a, st) : stk_launch_nu<4, 32>(m, a, st);
    case 5: return a.np_max == 16 ? stk_launch_nu<5, 16>(m, a, st) : stk_launch_nu<5, 32>(m, a, st);
  }
  set_error("stats_tc_launch: unsupported dimension");
  return KHG_ERR_UNSUPPORTED;
}

}  // namespace khg
