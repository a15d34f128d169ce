// kaldi-hmm-gmm_b200/csrc/khg_loglikes_gs.cu
//
// K1g: dense all-pdf log-likelihoods, Gaussian-stationary form of the tcgen05 kernel (sm_100a).
//
// Same contract and arithmetic as khg_loglikes_tc.cu (DecodableAmDiagGmmUnmapped::
// LogLikelihoodZeroBased for a whole block of frames and all pdfs, reference
// kaldi-hmm-gmm/csrc/decodable-am-diag-gmm.cc:29-71; LogSumExp csrc/eigen.cc:14-18): the
// contraction  A[t,:] . B[g,:]  with  A = [x, x^2, 1, 1],  B = [means_invvars, -0.5*inv_vars,
// gconst, gconst residual],  3-term fp16 hi/lo split  A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  with
// fp32 accumulation in TMEM, fused per-pdf log-sum-exp epilogue, out[p][t] pdf-major.
//
// What is different is which operand stays put.  khg_loglikes_tc.cu keeps 128 frames in shared
// memory and streams the whole model past them: 3 x 30 KB per 128 x 240 accumulator tile, with the
// hi/lo split of the frames built inside the kernel.  Here
//   * K0 (gs_build_a_kernel) writes the split feature operand A' = [A_hi | A_lo] (fp16, the same
//     packed K-step order as the model operand B') for a block of frames ONCE, 384 bytes per
//     frame at D = 40, L2-resident for the launch that follows;
//   * every CTA keeps ONE 240-Gaussian tile of B' (3 x 30 KB) in shared memory for a long run of
//     frame tiles and streams A' (3 x 16 KB per 128 x 240 tile: about half the bytes per MMA) through
//     an 8-deep TMA ring.  The SMs walk the frames in lock step (tile j of the model on CTA j), so
//     an A' tile is fetched from HBM once and read from L2 by everybody else; the model tiles left
//     over after whole rounds (n_tiles mod n_CTAs) are spread over the CTAs by frame range.
//   * a CTA writes the rows of ITS pdfs for consecutive frame tiles: per pdf a sequential stream.
// Warp roles (28 warps): 0-23 epilogue (6 groups x 4 TMEM lane quadrants, thread = frame row; the
// same run-table-driven LSE as the frame-stationary kernel), 24 A' TMA producer, 25 B' TMA producer
// (one reload per segment), 27 MMA issuer + TMEM owner (2 x 256 columns, double buffered).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "khg_internal.h"
#include "khg_tc_common.cuh"

namespace khg {

constexpr int kGsProdAWarp = kEpiWarps;       // streams A' tiles
constexpr int kGsProdBWarp = kEpiWarps + 1;   // (re)loads the stationary B' tile of a segment
constexpr int kGsMaxStages = 8;               // A' ring depth (16 KB stages)
constexpr int kGsMaxBChunks = 4;              // B' tile = up to 4 x 30 KB

// ------------------------------------------------------------------ K0: A' --
// Row t of A' = the packed K steps [hi steps | lo steps] of [x*s, x^2*s^2, 1, 1] (hi = fp16 round-to-
// nearest, lo = fp16 of the exact fp32 residual; s = the per-dimension power-of-two scale of the fp16
// model pack), laid out by the same stage table as B'.  One thread = 8 consecutive fp16 (16 bytes).
// Rows [T, rows_padded) are zero.
__global__ void gs_build_a_kernel(const float *__restrict__ feats, int64_t T, int64_t rows_padded, int D, StageTab tab, int KPA,
                                  const float *__restrict__ ascale, __half *__restrict__ ap) {
  const int gpr = KPA >> 3;  // 16-byte groups per row
  const int64_t total = rows_padded * gpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / gpr;
    const int grp = (int)(i - row * gpr);
    __align__(16) __half o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      bool lo_part;
      const int k = stage_logical_column<true>(tab, grp * 8 + e, lo_part);
      float v = 0.f;
      if (row < T && k >= 0) {
        if (k < 2 * D) {
          const float x = __ldg(feats + row * D + (k < D ? k : k - D));
          v = (k < D ? x : x * x) * __ldg(ascale + k);  // data.array().square(), csrc/decodable-am-diag-gmm.cc:57
        } else if (k <= 2 * D + 1 && !lo_part) {
          v = 1.f;  // the two columns the gconst and its residual ride on
        }
      }
      const __half hi = __float2half_rn(v);
      o[e] = lo_part ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    *reinterpret_cast<uint4 *>(ap + i * 8) = *reinterpret_cast<const uint4 *>(o);
  }
}

// ------------------------------------------------------------------ K1g --
struct GsArgs {
  int64_t T;               // frames of this launch (rows of A' beyond are zero)
  int n_ft;                // frame tiles
  int n_tiles;             // model tiles (240 Gaussians, pdf-aligned)
  StageTab tab;            // chunks of a packed operand row (same for A' and B') and the K steps each holds
  int hs, ls;              // K steps of the hi part (2D+2 columns) and of the lo part (2D columns)
  int stages;              // A' ring depth
  const int32_t *tile_g0;  // first Gaussian (operand row) of every model tile
  const int2 *epi_hdr;     // epilogue tables (khg_loglikes_tc.cu)
  const uint32_t *runs;
  const uint32_t *seg;
  float scale;
  float *out;              // pdf-major, out[p*ld + t]
  int64_t ld;
  float *scratch;          // one device word: store target of rows beyond T
  int *err;
  int debug_mode;          // experiments (KHG_EXPERIMENTS builds only)
};

// The CTA's work: a list of segments (model tile j, frame tiles [f0, f1)).  Whole rounds first — tile
// r * C + c for ALL frame tiles, every CTA walking the frames in lock step — then the n_tiles mod C left-
// over tiles, whose (tile, frame tile) units are dealt out in contiguous ranges.  Every warp role
// enumerates the same list.
struct GsSegIter {
  int C, c, n_ft, rounds, r;
  int64_t u, u1;
  __device__ GsSegIter(int n_tiles, int n_ft_) : C((int)gridDim.x), c((int)blockIdx.x), n_ft(n_ft_), rounds(n_tiles / (int)gridDim.x), r(0) {
    const int64_t U2 = (int64_t)(n_tiles - rounds * C) * n_ft;
    u = U2 * c / C;
    u1 = U2 * (c + 1) / C;
  }
  __device__ bool next(int &j, int &f0, int &f1) {
    if (r < rounds) {
      j = r * C + c;
      f0 = 0;
      f1 = n_ft;
      ++r;
      return true;
    }
    if (u >= u1) return false;
    const int l = (int)(u / n_ft);
    f0 = (int)(u - (int64_t)l * n_ft);
    f1 = (int)min((int64_t)n_ft, f0 + (u1 - u));
    j = rounds * C + l;
    u += f1 - f0;
    return true;
  }
};

// TWO / GRP: as in khg_loglikes_tc.cu (pdfs of 17..32 Gaussians; groups of short pdfs under one load).
template <bool TWO, bool GRP>
__global__ void __launch_bounds__(kTcThreads, 1)
loglikes_gs_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GsArgs a) {
  constexpr uint32_t kIdesc = make_idesc<true>();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *base_ptr = smem_raw + (base - raw);
  const int NB = a.tab.n, S = a.stages;
  const uint32_t sB = base;                          // the stationary model tile: NB chunks of 240 rows x 128 B
  const uint32_t sA = base + NB * kBStageBytes;      // A' ring: S stages of 128 rows x 128 B
  const uint32_t sBar = sA + S * kAChunkBytes;
  auto a_full = [&](uint32_t s) { return sBar + 8u * s; };
  auto a_empty = [&](uint32_t s) { return sBar + 8u * (kGsMaxStages + s); };
  const uint32_t b_full = sBar + 8u * (2 * kGsMaxStages), b_free = b_full + 8u;
  auto acc_full = [&](uint32_t b) { return sBar + 8u * (2 * kGsMaxStages + 2 + b); };
  auto acc_empty = [&](uint32_t b) { return sBar + 8u * (2 * kGsMaxStages + 4 + b); };
  const uint32_t tmem_slot = sBar + 8u * (2 * kGsMaxStages + 6);
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + (tmem_slot - base));
  // per segment: the kEpiGroups pdf lists of the model tile {seg start, first run, runs, first two descriptors,
  // first run word}, so that a frame tile starts from one shared-memory read instead of three dependent global loads
  struct EpiRole { uint32_t seg0, run0, nr, d0, d1, runw, pad0, pad1; };
  EpiRole *roles = reinterpret_cast<EpiRole *>(base_ptr + (sBar + 8u * (2 * kGsMaxStages + 8) - base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(a_full(s), 1);
        mbar_init(a_empty(s), 1);
      }
      mbar_init(b_full, 1);
      mbar_init(b_free, 1);
      for (int b = 0; b < 2; ++b) {
        mbar_init(acc_full(b), 1);
        mbar_init(acc_empty(b), 4 * kEpiGroups);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp == kGsProdAWarp) {
    // ===================== A' producer: one 16 KB chunk of a frame tile per stage =====================
    const bool leader = elect_one();
    uint32_t st = 0, ph = 1;  // waits on "empty" with the inverted phase
    GsSegIter it(a.n_tiles, a.n_ft);
    int j, f0, f1;
    while (it.next(j, f0, f1)) {
      for (int f = f0; f < f1; ++f) {
        for (int c = 0; c < NB; ++c) {
          mbar_wait(a_empty(st), ph);
          if (leader) {
            mbar_expect_tx(a_full(st), kAChunkBytes);
#ifdef KHG_EXPERIMENTS
            // 2: every frame tile re-reads the launch's first A' tile (full TMA / shared-memory volume, hot L2 lines)
            tma_load_2d(sA + st * kAChunkBytes, &map_a, a_full(st), c * 64, (a.debug_mode == 2 ? 0 : f) * kTileM);
#else
            tma_load_2d(sA + st * kAChunkBytes, &map_a, a_full(st), c * 64, f * kTileM);
#endif
          }
          __syncwarp();
          if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == kGsProdBWarp) {
    // ===================== B' producer: the segment's model tile, once =====================
    const bool leader = elect_one();
    uint32_t seg_it = 0;
    GsSegIter it(a.n_tiles, a.n_ft);
    int j, f0, f1;
    while (it.next(j, f0, f1)) {
      mbar_wait(b_free, (seg_it & 1) ^ 1);  // all MMAs of the previous segment have read the old tile
      const int g0 = __shfl_sync(0xffffffffu, __ldg(a.tile_g0 + j), 0);
      if (leader) {
        mbar_expect_tx(b_full, (uint32_t)NB * kBStageBytes);
        for (int c = 0; c < NB; ++c) tma_load_2d(sB + c * kBStageBytes, &map_b, b_full, c * 64, g0);
      }
      __syncwarp();
      ++seg_it;
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the loop (uniform control flow, see elect_one); one elected lane issues.
    const bool leader = elect_one();
    uint32_t st = 0, ph = 0, acc_it = 0, seg_it = 0;
    const uint32_t a0 = umma_desc_lo(sA), b0 = umma_desc_lo(sB);
    constexpr uint32_t kAStageDesc = kAChunkBytes >> 4, kBChunkDesc = kBStageBytes >> 4;
    constexpr int kSpc = 4;  // K steps (32 bytes of every operand row) per 128-byte chunk
    // descriptor offset of packed K step i (hi step q -> i = q, lo step q -> i = hs + q) of the stationary tile
    auto b_step = [&](int i) { return b0 + (uint32_t)(i / kSpc) * kBChunkDesc + (uint32_t)(i % kSpc) * 2; };
    GsSegIter it(a.n_tiles, a.n_ft);
    int j, f0, f1;
    while (it.next(j, f0, f1)) {
      mbar_wait(b_full, seg_it & 1);
      tc_fence_after();
      for (int f = f0; f < f1; ++f, ++acc_it) {
        const uint32_t buf = acc_it & 1;
        mbar_wait(acc_empty(buf), ((acc_it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * 256;
        uint32_t accum = 0;
        for (int c = 0; c < NB; ++c) {
          const uint32_t e = a.tab.e[c];
          const int h0 = e & 0xff, nh = (e >> 8) & 0xf, l0 = (e >> 12) & 0xff, nl = (e >> 20) & 0xf;
          mbar_wait(a_full(st), ph);
          tc_fence_after();
          const uint32_t da = a0 + st * kAStageDesc;
          // hi steps of A' meet the hi step of B' (hi.hi) and, inside the feature columns, its lo step (hi.lo)
#pragma unroll 4
          for (int s = 0; s < nh; ++s) {
            const int q = h0 + s;
            if (leader) tc_mma<true>(tmem_d, da + 2 * s, b_step(q), kIdesc, accum);
            accum = 1;
            if (q < a.ls && leader) tc_mma<true>(tmem_d, da + 2 * s, b_step(a.hs + q), kIdesc, 1);
          }
          // lo steps of A' meet the hi step of B' (lo.hi)
#pragma unroll 4
          for (int s = 0; s < nl; ++s) {
            if (leader) tc_mma<true>(tmem_d, da + 2 * (nh + s), b_step(l0 + s), kIdesc, 1);
          }
          if (leader) tc_commit(a_empty(st));
          if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
        }
        if (leader) tc_commit(acc_full(buf));
      }
      if (leader) tc_commit(b_free);
      ++seg_it;
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue =====================
    const int eg = warp >> 2;    // epilogue group: its share of the tile's pdfs
    const int quad = warp & 3;   // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;
    uint32_t acc_it = 0;
    bool bad = false;
    EpiState e;
    e.scale = a.scale;
    GsSegIter it(a.n_tiles, a.n_ft);
    int j, f0, f1;
    while (it.next(j, f0, f1)) {
      // (re)build the role table of this model tile: all epilogue warps have left the previous segment
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      if (threadIdx.x < kEpiGroups) {
        const int2 h = __ldg(a.epi_hdr + (size_t)kEpiGroups * j + threadIdx.x);
        EpiRole r;
        r.seg0 = (uint32_t)h.x;
        r.run0 = (uint32_t)h.y & 0xffffffu;
        r.nr = (uint32_t)h.y >> 24;
        r.d0 = __ldg(a.seg + r.seg0);  // (a sentinel when the list is empty)
        r.d1 = __ldg(a.seg + r.seg0 + 1);
        r.runw = __ldg(a.runs + r.run0);
        r.pad0 = r.pad1 = 0;
        roles[threadIdx.x] = r;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      for (int f = f0; f < f1; ++f, ++acc_it) {
        const int buf = acc_it & 1;
        const int64_t t = (int64_t)f * kTileM + row;
        const bool valid = t < a.T;
#ifdef KHG_EXPERIMENTS
        // 10: every tile stores into the block's first 128 frames (L2-resident window); 11: no store traffic
        const bool to_scratch = !valid || a.debug_mode == 11;
        e.out_t = to_scratch ? reinterpret_cast<char *>(a.scratch) : reinterpret_cast<char *>(a.out + (a.debug_mode == 10 ? (int64_t)row : t));
        e.ld_bytes = to_scratch ? 0u : (uint32_t)(a.ld * 4);
#else
        // rows beyond T store into a scratch word (ld_bytes = 0): no predicate in the hot loop
        e.out_t = valid ? reinterpret_cast<char *>(a.out + t) : reinterpret_cast<char *>(a.scratch);
        e.ld_bytes = valid ? (uint32_t)(a.ld * 4) : 0u;
#endif
        e.nan_acc = 0.f;
        const EpiRole r = roles[eg];
        const uint32_t *rp = a.runs + r.run0;
        int nr = (int)r.nr;
        e.sp = a.seg + r.seg0;
        e.d = r.d0;
        e.dn = r.d1;
        uint32_t run = r.runw;
        mbar_wait(acc_full(buf), (acc_it >> 1) & 1);
        tc_fence_after();
        e.trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
#ifdef KHG_EXPERIMENTS
        if (a.debug_mode == 1) nr = 0;  // 1: the epilogue skips the LSE (pipeline ceiling)
#endif
        if (nr > 0) {
          tc_ld16_issue(e.trow + (e.d & 0xffu), e.t);
#define KHG_CASE(L) case L: epi_run<L, TWO>(e, cnt); break;
#define KHG_CASE2(L) case 16 + L: epi_run2<L>(e, cnt); break;
#define KHG_CASEM(L) case L: epi_run_multi<L, TWO>(e, cnt); break;
#define KHG_CASESM KHG_CASEM(1) KHG_CASEM(2) KHG_CASEM(3) KHG_CASEM(4) KHG_CASEM(5) KHG_CASEM(6) KHG_CASEM(7) KHG_CASEM(8)
#define KHG_CASES1 KHG_CASE(1) KHG_CASE(2) KHG_CASE(3) KHG_CASE(4) KHG_CASE(5) KHG_CASE(6) KHG_CASE(7) KHG_CASE(8) \
                   KHG_CASE(9) KHG_CASE(10) KHG_CASE(11) KHG_CASE(12) KHG_CASE(13) KHG_CASE(14) KHG_CASE(15) KHG_CASE(16)
#define KHG_CASES2 KHG_CASE2(1) KHG_CASE2(2) KHG_CASE2(3) KHG_CASE2(4) KHG_CASE2(5) KHG_CASE2(6) KHG_CASE2(7) KHG_CASE2(8) \
                   KHG_CASE2(9) KHG_CASE2(10) KHG_CASE2(11) KHG_CASE2(12) KHG_CASE2(13) KHG_CASE2(14) KHG_CASE2(15) KHG_CASE2(16)
          if constexpr (TWO) tc_ld16_issue_if(e.trow + (e.d & 0xffu) + 16, e.t2, (e.d & kSegTwoChunks) != 0);
          for (; nr > 0; --nr) {
            const int len = run & 0xff, cnt = (run >> 8) & 0x7fffff;
            const bool grouped = GRP && (run >> 31) != 0;
            if (nr > 1) run = __ldg(++rp);
            if (GRP && grouped) {
              switch (len) { KHG_CASESM }
            } else if constexpr (TWO) {
              switch (len) {
                KHG_CASES1 KHG_CASES2
                default: epi_run_long(e, cnt, len); break;
              }
            } else {
              switch (len) {
                KHG_CASES1
                default: epi_run_long(e, cnt, len); break;
              }
            }
          }
          // the (unused) sentinel load issued after the last segment
          if constexpr (TWO) tc_ld16_wait2(e.t, e.t2); else tc_ld16_wait(e.t);
#undef KHG_CASE
#undef KHG_CASE2
#undef KHG_CASES1
#undef KHG_CASES2
#undef KHG_CASEM
#undef KHG_CASESM
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
        const float nan_acc = e.nan_acc;
        if (valid && nan_acc != nan_acc) bad = true;  // r*0 is NaN exactly for a NaN/Inf result
      }
    }
    if (bad) atomicOr(a.err, ERR_NONFINITE);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host side --
// Frames per K0 + K1g launch pair: bounds the A' scratch (384 B per frame at D = 40) and keeps it
// L2-resident next to the output stream.  KHG_GS_CHUNK overrides (experiments).
static int64_t gs_chunk_frames(const khg_model *m) {
  int64_t c = 128LL * m->sm_count * 8;
  if (const char *e = getenv("KHG_GS_CHUNK")) c = std::max<int64_t>(128, atoll(e));
  return (c + 127) / 128 * 128;
}

bool gs_supported(const khg_model *m) {
  const TcPack &t = m->tc;
  return t.f16_ready && t.tab16.n <= kGsMaxBChunks;
}

void gs_free(khg_model *m) {
  m->tc.a_scr.release();
  m->tc.a_rows = 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static khg_status gs_reserve_a(khg_model *m, int64_t rows) {
  TcPack &t = m->tc;
  if (rows <= t.a_rows) return KHG_OK;
  if (t.a_rows > 0) KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));  // a launch may still read the old buffer
  KHG_TRY(t.a_scr.reserve(sizeof(__half) * (size_t)rows * t.KPB16));
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return KHG_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)t.KPB16, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)t.KPB16 * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&t.amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, t.a_scr.p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (A') failed with CUresult " + std::to_string((int)r));
    return KHG_ERR_CUDA;
  }
  t.a_rows = rows;
  return KHG_OK;
}

// One K0 + K1g pair over frames [0, T) (T <= the A' capacity).
static khg_status gs_launch(khg_model *m, const float *d_feats, int64_t T, float scale, float *d_out, int64_t ld_out) {
  TcPack &t = m->tc;
  const int D = m->dim;
  const int64_t rows = (T + kTileM - 1) / kTileM * kTileM;
  KHG_TRY(gs_reserve_a(m, rows));
  {
    const int64_t total = rows * (t.KPB16 / 8);
    const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, 16LL * m->sm_count);
    gs_build_a_kernel<<<grid, 256, 0, m->stream>>>(d_feats, T, rows, D, t.tab16, t.KPB16, t.ascale, t.a_scr.as<__half>());
    ++g_launch_count;
    KHG_CUDA_TRY(cudaGetLastError());
  }
  GsArgs a;
  a.T = T;
  a.n_ft = (int)(rows / kTileM);
  a.n_tiles = t.n_tiles;
  a.tab = t.tab16;
  a.hs = t.K16 / 16;
  a.ls = (2 * D + 15) / 16;
  const int b_bytes = a.tab.n * kBStageBytes;
  a.stages = std::min(kGsMaxStages, (int)((226 * 1024 - 512 - b_bytes) / kAChunkBytes));
  if (const char *e = getenv("KHG_GS_STAGES")) a.stages = std::max(2, std::min(a.stages, atoi(e)));  // experiments
  if (a.stages < 2) {
    set_error("feature dimension too large for the Gaussian-stationary tcgen05 kernel");
    return KHG_ERR_UNSUPPORTED;
  }
  a.tile_g0 = t.tile_g0;
  a.epi_hdr = static_cast<const int2 *>(t.epi_hdr);
  a.runs = t.runs;
  a.seg = t.seg;
  a.scale = scale;
  a.out = d_out;
  a.ld = ld_out;
  a.scratch = reinterpret_cast<float *>(m->d_scratch_int + 3);
  a.err = m->d_err;
  a.debug_mode = 0;
#ifdef KHG_EXPERIMENTS
  if (const char *dbg = getenv("KHG_TC_DEBUG_MODE")) a.debug_mode = atoi(dbg);
#endif
  if (ld_out * 4 >= ((int64_t)1 << 32)) {
    set_error("ld_out too large for the tensor-core kernel (needs ld_out < 2^30)");
    return KHG_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)b_bytes + (size_t)a.stages * kAChunkBytes + 512 + 1024;  // (512: barriers + the role table)
  auto kern = t.two_chunk_segs ? (t.grouped_segs ? loglikes_gs_kernel<true, true> : loglikes_gs_kernel<true, false>)
                               : (t.grouped_segs ? loglikes_gs_kernel<false, true> : loglikes_gs_kernel<false, false>);
  KHG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));  // (per device; cheap)
  unsigned grid = (unsigned)std::min<int64_t>((int64_t)a.n_tiles * a.n_ft, m->sm_count);
  if (const char *mc = getenv("KHG_TC_MAX_CTAS")) grid = std::min<unsigned>(grid, (unsigned)std::max(1, atoi(mc)));  // experiments
  kern<<<grid, kTcThreads, smem, m->stream>>>(t.amap, t.hmap_hi, a);
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  return KHG_OK;
}

// The fp16-split dense block of frames [0, T) in sub-blocks of gs_chunk_frames() (forced precision: the
// caller chose this kernel; out-of-range features give non-finite results, which are reported).
khg_status gs_loglikes(khg_model *m, const float *d_feats, int64_t T, float scale, float *d_out, int64_t ld_out) {
  const int64_t chunk = gs_chunk_frames(m);
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t n = std::min(chunk, T - t0);
    KHG_TRY(gs_launch(m, d_feats + t0 * m->dim, n, scale, d_out + t0, ld_out));
  }
  return KHG_OK;
}

}  // namespace khg
