// kaldi-hmm-gmm_b200/csrc/khg_tc_common.cuh — device-side building blocks shared by the two tcgen05 dense
// log-likelihood kernels (khg_loglikes_tc.cu: frame-stationary, streams the model operand;
// khg_loglikes_gs.cu: Gaussian-stationary, streams the pre-split feature operand): PTX wrappers
// (mbarrier, TMA, tcgen05.mma / ld / commit), the per-pdf log-sum-exp epilogue driven by host-built
// run tables, UMMA descriptors and the stage table of the packed [hi | lo] operand rows.
#ifndef KHG_TC_COMMON_CUH_
#define KHG_TC_COMMON_CUH_
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstring>
#include <utility>

#include "khg_internal.h"

namespace khg {

#ifndef KHG_FEAT_LOAD
#define KHG_FEAT_LOAD __ldcg
#endif

constexpr int kTileM = 128;         // frames per CTA tile (UMMA M)
constexpr int kTileN = 240;         // Gaussians per accumulator tile (UMMA N)
// Operand element: tf32 (4 B, 32 per 128-byte swizzle atom, UMMA_K = 8) or fp16 (2 B, 64
// per atom, UMMA_K = 16).  Either way one UMMA K-step is 32 bytes of each operand row.
template <bool F16> struct Elem {
  static constexpr int kBytes = F16 ? 2 : 4;
  static constexpr int kChunkK = 128 / kBytes;  // elements per 128-byte swizzle atom
  static constexpr int kUmmaK = 32 / kBytes;
};
constexpr int kAChunkBytes = kTileM * 128;   // 16384
constexpr int kBStageBytes = kTileN * 128;   // 30720 (multiple of 1024)
constexpr int kMaxChunks = 5;       // 128-byte K chunks per operand row held in smem
#ifndef KHG_EPI_GROUPS
#define KHG_EPI_GROUPS 6
#endif
constexpr int kEpiGroups = KHG_EPI_GROUPS;       // epilogue warpgroups (4 warps each, one per TMEM lane quadrant); <= 8
constexpr int kBuilderThreads = 64;
constexpr int kEpiWarps = 4 * kEpiGroups;      // warps 0..4*kEpiGroups-1: epilogue (TMEM lane quadrant = warp % 4)
constexpr int kBuilderWarp0 = kEpiWarps;        // warps 16-17: A builders
constexpr int kProducerWarp = kEpiWarps + 2;    // warp 18: TMA producer
constexpr int kMmaWarp = kEpiWarps + 3;         // warp 19: MMA issuer (+ TMEM alloc); the warp scheduler
                                                // favours high warp ids, and this warp must never starve
constexpr int kTcThreads = 32 * (kEpiWarps + 4);
// fp16 path: |x * 2^-k| beyond this keeps x^2 (and x) from fitting fp16 with margin
constexpr float kF16FeatLimit = 128.0f;
constexpr float kNegSentinel = -1.0e30f;  // stands in for gconst = -inf (0 * inf = NaN in the split)

// ---------------------------------------------------------------- PTX wrappers --
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait with a suspend-time hint: the warp is parked by the hardware until the phase
  // completes (or the hint expires) instead of polling, so waiting warps do not take issue
  // slots (or power) from the warps that have work.
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
// Cluster (CTA pair) forms: the pair streams ONE copy of the operand B' out of L2 — each CTA loads half
// of every stage and multicasts it into both CTAs' shared memory (same offsets, each CTA's own "full"
// barrier gets the bytes), and a stage is free again once BOTH CTAs' MMAs have read it.
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// One lane of a converged warp.  The MMA-issuing and TMA-issuing warps run their loops with all 32
// lanes (warp-uniform control flow: counters and descriptors then live in uniform registers and
// UTCHMMA / UTMALDG take them directly); a loop run by `if (lane == 0)` makes the compiler wrap
// every such instruction in a register-to-uniform "waterfall" loop (~165 cycles per MMA measured
// with tools/mma_rate.cu, against the 120-cycle floor of a 128x240x16 MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t p;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
  return p != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// The smem descriptors differ only in their low word (start address); the high word
// (stride, version, swizzle) is a constant, so the issuing thread does 32-bit adds only.
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
template <bool F16>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t idesc, uint32_t accum) {
  if (F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(accum), "r"(kDescHi)
        : "memory");
  }
}
// TMEM -> registers: 32 lanes x 16 consecutive fp32 columns; thread i of the warp gets
// lane (warp%4)*32+i.  Issue and wait are separate so that the next segment's load is in
// flight while the current one is reduced.  The wait names the destination registers as
// in/out operands: consumers then depend on the wait, not just on the (asynchronous) load.
struct TReg16 { uint32_t r[16]; };
__device__ __forceinline__ void tc_ld16_issue(uint32_t taddr, TReg16 &t) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(t.r[0]), "=r"(t.r[1]), "=r"(t.r[2]), "=r"(t.r[3]), "=r"(t.r[4]), "=r"(t.r[5]), "=r"(t.r[6]), "=r"(t.r[7]),
        "=r"(t.r[8]), "=r"(t.r[9]), "=r"(t.r[10]), "=r"(t.r[11]), "=r"(t.r[12]), "=r"(t.r[13]), "=r"(t.r[14]), "=r"(t.r[15])
      : "r"(taddr)
      : "memory");
}
// Same load under a warp-uniform predicate (no branch: the instruction keeps its place in the
// schedule between the two halves of the segment arithmetic).
__device__ __forceinline__ void tc_ld16_issue_if(uint32_t taddr, TReg16 &t, bool pred) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %17, 0;\n\t"
      "@p tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t}"
      : "=r"(t.r[0]), "=r"(t.r[1]), "=r"(t.r[2]), "=r"(t.r[3]), "=r"(t.r[4]), "=r"(t.r[5]), "=r"(t.r[6]), "=r"(t.r[7]),
        "=r"(t.r[8]), "=r"(t.r[9]), "=r"(t.r[10]), "=r"(t.r[11]), "=r"(t.r[12]), "=r"(t.r[13]), "=r"(t.r[14]), "=r"(t.r[15])
      : "r"(taddr), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ void tc_ld16_wait(TReg16 &t) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(t.r[0]), "+r"(t.r[1]), "+r"(t.r[2]), "+r"(t.r[3]), "+r"(t.r[4]), "+r"(t.r[5]), "+r"(t.r[6]), "+r"(t.r[7]),
                 "+r"(t.r[8]), "+r"(t.r[9]), "+r"(t.r[10]), "+r"(t.r[11]), "+r"(t.r[12]), "+r"(t.r[13]), "+r"(t.r[14]), "+r"(t.r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tc_ld16_wait2(TReg16 &t, TReg16 &u) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(t.r[0]), "+r"(t.r[1]), "+r"(t.r[2]), "+r"(t.r[3]), "+r"(t.r[4]), "+r"(t.r[5]), "+r"(t.r[6]), "+r"(t.r[7]),
                 "+r"(t.r[8]), "+r"(t.r[9]), "+r"(t.r[10]), "+r"(t.r[11]), "+r"(t.r[12]), "+r"(t.r[13]), "+r"(t.r[14]), "+r"(t.r[15]),
                 "+r"(u.r[0]), "+r"(u.r[1]), "+r"(u.r[2]), "+r"(u.r[3]), "+r"(u.r[4]), "+r"(u.r[5]), "+r"(u.r[6]), "+r"(u.r[7]),
                 "+r"(u.r[8]), "+r"(u.r[9]), "+r"(u.r[10]), "+r"(u.r[11]), "+r"(u.r[12]), "+r"(u.r[13]), "+r"(u.r[14]), "+r"(u.r[15])
               :
               : "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Max-subtracted log-sum-exp (csrc/eigen.cc:14-18) of the first L of 16 accumulator
// columns held in registers, in two parts so that the TMEM load of the NEXT segment can be
// issued into the same registers between them.  L is a compile-time constant so that no
// issue slot is spent on masked-off columns.
//   part 1: M = max, e[i] = x[i]*log2(e) - M*log2(e)  (after it the loaded registers are dead)
//   part 2: M + ln(sum 2^e[i])
template <int L>
struct SegLse {
  static constexpr int kPairs = L / 2;
  float2 e2[kPairs > 0 ? kPairs : 1];
  float e1;
  float M;
  // OFF: first of the L registers (several short segments can share one 16-column load)
  template <int OFF = 0>
  __device__ __forceinline__ void part1(const TReg16 &t) {
    constexpr float kLog2e = 1.4426950408889634f;
    float m = __uint_as_float(t.r[OFF]);
#pragma unroll
    for (int i = 1; i < L; ++i) m = fmaxf(m, __uint_as_float(t.r[OFF + i]));
    M = m;
    if (L == 1) return;
    const float ml = m * kLog2e;
    // packed fp32x2 math (FFMA2 / FADD2 on sm_100): one issue slot per two columns
    const float2 k2 = make_float2(kLog2e, kLog2e), nm2 = make_float2(-ml, -ml);
#pragma unroll
    for (int i = 0; i < kPairs; ++i)
      e2[i] = __ffma2_rn(make_float2(__uint_as_float(t.r[OFF + 2 * i]), __uint_as_float(t.r[OFF + 2 * i + 1])), k2, nm2);
    if (L & 1) e1 = fmaf(__uint_as_float(t.r[OFF + L - 1]), kLog2e, -ml);
  }
  // sum of 2^e[i] (the caller combines two chunks of a long segment)
  __device__ __forceinline__ float part2_sum() {
    if (L == 1) return 1.f;
    float2 s2;
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const float2 v = make_float2(fast_exp2(e2[i].x), fast_exp2(e2[i].y));
      s2 = i == 0 ? v : __fadd2_rn(s2, v);
    }
    float s = s2.x + s2.y;
    if (L & 1) s += fast_exp2(e1);
    return s;
  }
  __device__ __forceinline__ float part2() {
    constexpr float kLn2 = 0.6931471805599453f;
    if (L == 1) return M;
    float2 s2;
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const float2 v = make_float2(fast_exp2(e2[i].x), fast_exp2(e2[i].y));
      s2 = i == 0 ? v : __fadd2_rn(s2, v);
    }
    float s = s2.x + s2.y;
    if (L & 1) s += fast_exp2(e1);
    return fmaf(fast_log2(s), kLn2, M);
  }
};

// Generic LSE of a segment of `len` (> 16) columns at TMEM address taddr: two passes.
__device__ __forceinline__ float seg_lse_long(uint32_t taddr, int len) {
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  TReg16 w;
  float m = -CUDART_INF_F;
  for (int w0 = 0; w0 < len; w0 += 16) {
    tc_ld16_issue(taddr + w0, w);
    tc_ld16_wait(w);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (w0 + i < len) m = fmaxf(m, __uint_as_float(w.r[i]));
  }
  const float ml = m * kLog2e;
  float s = 0.f;
  for (int w0 = 0; w0 < len; w0 += 16) {
    tc_ld16_issue(taddr + w0, w);
    tc_ld16_wait(w);
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (w0 + i < len) s += fast_exp2(fmaf(__uint_as_float(w.r[i]), kLog2e, -ml));
  }
  return fmaf(fast_log2(s), kLn2, m);
}

// Epilogue state of one warp inside one accumulator tile.  The warp walks ITS list of
// segments (pdfs) of the tile, seg[k...], ordered in runs of equal length.  Invariant between
// segments: the TMEM load of segment k (descriptor d) has been issued into t.  A tcgen05.ld ->
// wait::ld round trip is ~170 cycles (tools/tmem_ld_rate.cu), so the load of segment k+1 is
// issued as soon as part 1 of segment k has consumed the registers, across run boundaries, and
// completes under part 2 (exponentials, log, store).  Every list ends with two sentinel
// descriptors (column 0), so "the next" and "the one after" always exist: no predicates.
//   trow  TMEM address of this warp's lane quadrant in the current accumulator buffer
//   sp    &seg[k]; a descriptor is  column | pdf << 8
//   out_t byte address of out[0][t] — or of a scratch word with ld_bytes = 0 for rows beyond
//         T, so that the store needs no predicate
//   nan_acc collects r*0 (NaN for a non-finite r, the reference's "Invalid answer"): one FFMA
//         instead of a compare/select/or per segment
// A descriptor is  column | two_chunks << 8 | pdf << 9: segments of 17..32 columns are read as two
// 16-column loads (t and t2); the flag tells whoever issues the NEXT segment's loads — possibly the
// last segment of a run of another length — that t2 is wanted too.
constexpr uint32_t kSegTwoChunks = 0x100u;
constexpr int kSegPdfShift = 9;
struct EpiState {
  TReg16 t, t2;
  uint32_t d, dn;   // descriptors of segment k (load in flight) and k+1
  const uint32_t *sp;
  uint32_t trow;
  char *out_t;
  uint32_t ld_bytes;
  float scale, nan_acc;
};

// TWO = the model has segments of 17..32 columns somewhere, i.e. the next segment may want its
// second load issued too; models without such pdfs run the variant without that instruction.
template <int L, bool TWO>
__device__ __forceinline__ void epi_run(EpiState &e, int cnt) {
#pragma unroll 1
  for (; cnt > 0; --cnt) {
    const uint32_t dcur = e.d, dnext = e.dn;
    e.dn = __ldg(e.sp + 2);  // two ahead: its latency never gates the TMEM load
    ++e.sp;
    tc_ld16_wait(e.t);
    SegLse<L> lse;
    lse.part1(e.t);
    tc_ld16_issue(e.trow + (dnext & 0xffu), e.t);  // after the last segment: the sentinel (unused)
    if (TWO) tc_ld16_issue_if(e.trow + (dnext & 0xffu) + 16, e.t2, (dnext & kSegTwoChunks) != 0);
    const float r = lse.part2();
    e.nan_acc = fmaf(r, 0.f, e.nan_acc);
    *reinterpret_cast<float *>(e.out_t + (uint64_t)(dcur >> kSegPdfShift) * e.ld_bytes) = e.scale * r;
    e.d = dnext;
  }
}

// Groups of NS = 16 / L column-adjacent segments (consecutive pdfs of L <= 8 Gaussians each) under
// ONE 16-column load: one descriptor, one wait and one load per group instead of per segment — what
// models with few Gaussians per pdf (a freshly initialised monophone system has one) are made of.
// The descriptor names the group's first column and first pdf.
template <int L, int... Ks>
__device__ __forceinline__ void multi_part1(SegLse<L> (&s)[sizeof...(Ks)], const TReg16 &t, std::integer_sequence<int, Ks...>) {
  (s[Ks].template part1<Ks * L>(t), ...);
}
template <int L, bool TWO>
__device__ __forceinline__ void epi_run_multi(EpiState &e, int cnt) {
  constexpr int NS = 16 / L;
#pragma unroll 1
  for (; cnt > 0; --cnt) {
    const uint32_t dcur = e.d, dnext = e.dn;
    e.dn = __ldg(e.sp + 2);
    ++e.sp;
    tc_ld16_wait(e.t);
    SegLse<L> lse[NS];
    multi_part1<L>(lse, e.t, std::make_integer_sequence<int, NS>());
    tc_ld16_issue(e.trow + (dnext & 0xffu), e.t);
    if (TWO) tc_ld16_issue_if(e.trow + (dnext & 0xffu) + 16, e.t2, (dnext & kSegTwoChunks) != 0);
    char *o = e.out_t + (uint64_t)(dcur >> kSegPdfShift) * e.ld_bytes;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
      const float r = lse[k].part2();
      e.nan_acc = fmaf(r, 0.f, e.nan_acc);
      *reinterpret_cast<float *>(o + (uint64_t)k * e.ld_bytes) = e.scale * r;
    }
    e.d = dnext;
  }
}

// Segments of 16 + LB columns (17..32 Gaussians): both 16-column loads were issued by the previous
// segment and complete under ONE wait; chunk a is reduced, the next segment's first load goes out,
// chunk b is reduced, the next segment's second load goes out, and the two (max, sum) pairs are
// merged: M = max(Ma, Mb), s = sa * 2^((Ma - M) log2 e) + sb * 2^((Mb - M) log2 e).
template <int LB>
__device__ __forceinline__ void epi_run2(EpiState &e, int cnt) {
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
#pragma unroll 1
  for (; cnt > 0; --cnt) {
    const uint32_t dcur = e.d, dnext = e.dn;
    e.dn = __ldg(e.sp + 2);
    ++e.sp;
    tc_ld16_wait2(e.t, e.t2);
    SegLse<16> la;
    la.part1(e.t);
    tc_ld16_issue(e.trow + (dnext & 0xffu), e.t);
    const float sa = la.part2_sum();
    SegLse<LB> lb;
    lb.part1(e.t2);
    tc_ld16_issue_if(e.trow + (dnext & 0xffu) + 16, e.t2, (dnext & kSegTwoChunks) != 0);
    const float sb = lb.part2_sum();
    const float M = fmaxf(la.M, lb.M);
    const float s = fmaf(sa, fast_exp2((la.M - M) * kLog2e), sb * fast_exp2((lb.M - M) * kLog2e));
    const float r = fmaf(fast_log2(s), kLn2, M);
    e.nan_acc = fmaf(r, 0.f, e.nan_acc);
    *reinterpret_cast<float *>(e.out_t + (uint64_t)(dcur >> kSegPdfShift) * e.ld_bytes) = e.scale * r;
    e.d = dnext;
  }
}

__device__ __forceinline__ void epi_run_long(EpiState &e, int cnt, int len) {  // (only models with two-chunk segments set the flag)
  for (; cnt > 0; --cnt) {
    const uint32_t dcur = e.d, dnext = e.dn;
    e.dn = __ldg(e.sp + 2);
    ++e.sp;
    tc_ld16_wait(e.t);  // the pending 16-column load is not used by the two-pass form
    const float r = seg_lse_long(e.trow + (dcur & 0xffu), len);
    tc_ld16_issue(e.trow + (dnext & 0xffu), e.t);
    tc_ld16_issue_if(e.trow + (dnext & 0xffu) + 16, e.t2, (dnext & kSegTwoChunks) != 0);
    e.nan_acc = fmaf(r, 0.f, e.nan_acc);
    *reinterpret_cast<float *>(e.out_t + (uint64_t)(dcur >> kSegPdfShift) * e.ld_bytes) = e.scale * r;
    e.d = dnext;
  }
}

// UMMA shared-memory descriptor: K-major, 128B swizzle, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return (saddr >> 4) & 0x3FFF; }
static_assert((((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61)) >> 32 == kDescHi, "descriptor high word");
// Instruction descriptor: D=f32, A=B=tf32 (format 2) or f16 (format 0), both K-major, M=128, N=240.
template <bool F16, int M = kTileM>
__host__ __device__ constexpr uint32_t make_idesc() {
  return (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((uint32_t)(kTileN >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// Writes the hi/lo split of v at (row, col) of the A operand (UMMA K-major, 128B swizzle):
// hi = round-to-nearest in the operand format, lo = v - hi (exact in fp32), also stored in
// the operand format.
template <bool F16>
__device__ __forceinline__ void a_store_split(uint8_t *a_hi, uint32_t lo_offset, int row, int col, float v);
template <>
__device__ __forceinline__ void a_store_split<false>(uint8_t *a_hi, uint32_t lo_offset, int row, int col, float v) {
  const float hi = tf32_rna(v), lo = v - hi;
  const uint32_t off = (col >> 5) * kAChunkBytes + row * 128 + ((((col & 31) >> 2) ^ (row & 7)) << 4) + ((col & 3) << 2);
  *reinterpret_cast<float *>(a_hi + off) = hi;
  *reinterpret_cast<float *>(a_hi + lo_offset + off) = lo;
}
template <>
__device__ __forceinline__ void a_store_split<true>(uint8_t *a_hi, uint32_t lo_offset, int row, int col, float v) {
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const uint32_t off = (col >> 6) * kAChunkBytes + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1);
  *reinterpret_cast<__half *>(a_hi + off) = hi;
  *reinterpret_cast<__half *>(a_hi + lo_offset + off) = lo;
}


// ------------------------------------------------------------------ B pack (K4) --
// The streamed operand B' is a sequence of 128-byte chunks (= TMA stages).  A chunk holds nh K steps
// of the hi part starting at hi step h0, then nl K steps of the lo part starting at lo step l0
// (a K step = 32 bytes = 16 fp16 / 8 tf32 columns).  The same table drives the pack kernels and
// the MMA issuer.  Two layouts are built on the host:
//   packed       [hi, Khh columns | lo, Kc columns]: fewest bytes (fp16, D=40: 3 chunks, not 4)
//   interleaved  hi chunk 0, lo chunk 0, hi chunk 1, ...: even MMA work per stage (tf32, whose
//                stage ring is only 4 deep and whose chunk count is the same either way)
// logical operand column of B' column kk (-1 = padding); lo_part says which half
template <bool F16>
__host__ __device__ inline int stage_logical_column(const StageTab &tab, int kk, bool &lo_part) {
  constexpr int ck = Elem<F16>::kChunkK, uk = Elem<F16>::kUmmaK;
  const uint32_t v = tab.e[kk / ck];
  const int h0 = v & 0xff, nh = (v >> 8) & 0xf, l0 = (v >> 12) & 0xff, nl = (v >> 20) & 0xf;
  const int s = (kk % ck) / uk, r = kk % uk;
  lo_part = s >= nh;
  if (s < nh) return (h0 + s) * uk + r;
  if (s - nh < nl) return (l0 + s - nh) * uk + r;
  return -1;
}
inline StageTab make_stage_tab(int Khh, int Kc, int ck, int uk, bool packed) {
  StageTab t;
  memset(&t, 0, sizeof(t));
  const int hs = Khh / uk, ls = Kc / uk, spc = ck / uk;  // K steps of each part, steps per chunk
  if (packed) {
    for (int q = 0; q < hs + ls; q += spc, ++t.n) {
      const int nh = std::max(0, std::min(spc, hs - q)), l0 = std::max(0, q - hs), nl = std::max(0, std::min(spc - nh, ls - l0));
      t.e[t.n] = (uint32_t)(nh ? q : 0) | (uint32_t)nh << 8 | (uint32_t)l0 << 12 | (uint32_t)nl << 20;
    }
  } else {
    for (int q = 0; q < hs; q += spc) {
      t.e[t.n++] = (uint32_t)q | (uint32_t)std::min(spc, hs - q) << 8;
      if (q < ls) t.e[t.n++] = (uint32_t)q << 12 | (uint32_t)std::min(spc, ls - q) << 20;
    }
  }
  return t;
}
}  // namespace khg

#endif  // KHG_TC_COMMON_CUH_
