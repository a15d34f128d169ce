// kaldi-hmm-gmm_b200/csrc/khg_loglikes_tc.cu
//
// K1: dense all-pdf log-likelihoods on the 5th-generation tensor cores (sm_100a).
//
// Replaces, for a whole block of frames and ALL pdfs at once, the per-(frame, pdf)
// arithmetic of DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased
// (reference kaldi-hmm-gmm/csrc/decodable-am-diag-gmm.cc:29-71):
//     ll(t,g)  = gconst_g + means_invvars_g . x_t - 0.5 * inv_vars_g . x_t^2     (:55-57)
//     out(t,p) = LogSumExp_{g in pdf p} ll(t,g)                                   (csrc/eigen.cc:14-18)
// as ONE contraction  A[t,:] . B[g,:]  with  A = [x, x^2, 1]  (T x K, K = 2D+1) and
// B = [means_invvars, -0.5*inv_vars, gconst]  (G x K), followed by a fused per-pdf
// log-sum-exp epilogue.
//
// Precision: a 3-term split.  Both operands are split  v = hi + lo  with hi = the value rounded to
// the operand container — tf32 (kind::tf32) or fp16 with exact per-column power-of-two scaling
// (kind::f16, twice the tensor rate) — and the tensor core accumulates
// A_hi.B_hi + A_lo.B_hi + A_hi.B_lo  in fp32 (the dropped lo.lo term is ~2^-22 relative), which
// keeps |error| well inside BASELINE.json's 1e-3 abs / 1e-4 rel bound on log-likelihoods of
// magnitude ~1e2.
//
// Structure (one persistent CTA per SM, 28 warps, warp-specialised; CTAs run in pairs that share
// the streamed operand by TMA multicast when there are enough frame tiles):
//   warps 0-23  epilogue (6 groups x 4 TMEM lane quadrants): tcgen05.ld the accumulator rows
//               (thread = frame), per-pdf max-subtracted log-sum-exp over the pdf's contiguous
//               Gaussians driven by host-built run tables, coalesced store of out[p][t] (pdf-major)
//   warps 24-25 A builders: load a 128-frame feature tile (16 loads in flight per thread), form
//               [x, x^2, 1, 1], split hi/lo and write it in the UMMA K-major 128B-swizzled layout
//               into one of TWO A buffers, one item ahead of the MMAs; A is stationary for all N
//               tiles of a work item
//   warp 26     TMA producer: streams the operand B' (240 Gaussians x one 128-byte K chunk per
//               stage) through a ring of smem stages (cp.async.bulk.tensor, 128B swizzle)
//   warp 27     MMA issuer: one elected lane issues tcgen05.mma, M=128 (frames), N=240
//               (Gaussians); accumulators live in TMEM (2 x 256 columns, double buffered against
//               the epilogue); this warp also allocates / frees the TMEM columns
// N tiles are aligned to pdf boundaries (tile table built on the host), so a pdf's
// Gaussians never straddle two accumulator tiles.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "khg_internal.h"
#include "khg_tc_common.cuh"

namespace khg {

__global__ void latch_error_kernel(int *err, int bits) { atomicOr(err, bits); }

// One operand row per Gaussian, laid out by the stage table (KPB = 128-byte chunks * n
// columns).  Logical columns: [means_invvars (D) |
// -0.5*inv_vars (D) | gconst | gconst residual].  The hi part spans all 2D+2 of them rounded up
// to the MMA K (Khh), the lo part only the 2D feature columns (Kc): the gconst rides entirely in
// the hi.hi product (its container-rounded part and the residual, both against a constant 1 in
// A_hi).  Storing hi and lo side by side instead of in two padded matrices cuts the bytes every
// CTA streams per tile (192 instead of 256 fp16 columns at D = 40).
__global__ void tc_pack_kernel(int G, int D, StageTab tab, int KPB, int rows, const float *__restrict__ miv,
                               const float *__restrict__ iv, const float *__restrict__ gconsts,
                               float *__restrict__ bp) {
  size_t total = (size_t)rows * KPB;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % KPB);
    const int g = (int)(i / KPB);
    bool lo_part;
    const int k = stage_logical_column<false>(tab, kk, lo_part);
    float out = 0.f;
    if (g < G && k >= 0) {
      if (k < 2 * D) {
        const float v = k < D ? miv[(size_t)g * D + k] : -0.5f * iv[(size_t)g * D + (k - D)];
        float hi = tf32_rna(v), lo = v - hi;
        if (!(fabsf(v) <= 3.0e38f)) { hi = v; lo = 0.f; }
        out = lo_part ? lo : hi;
      } else if (k <= 2 * D + 1 && !lo_part) {
        float v = gconsts[g];
        if (v == -CUDART_INF_F) v = kNegSentinel;  // zero-weight Gaussian (csrc/diag-gmm.cc:136-141)
        const float vh = tf32_rna(v);
        out = k == 2 * D ? vh : tf32_rna(v - vh);
        if (!(fabsf(v) <= 3.0e38f)) out = k == 2 * D ? v : 0.f;
      }
    }
    bp[i] = out;
  }
}

// fp16 variant: B scaled per dimension by powers of two (bscale[k], exact), split into fp16
// hi and lo.  flags[0] is set when a value does not fit fp16 comfortably or a gconst is
// -inf: the model then stays on the tf32 path.
__global__ void tc_pack_f16_kernel(int G, int D, StageTab tab, int KPB, int rows, const float *__restrict__ miv,
                                   const float *__restrict__ iv, const float *__restrict__ gconsts,
                                   const float *__restrict__ bscale, __half *__restrict__ bp, int *__restrict__ flags) {
  size_t total = (size_t)rows * KPB;
  bool bad = false;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % KPB);
    const int g = (int)(i / KPB);
    bool lo_part;
    const int k = stage_logical_column<true>(tab, kk, lo_part);
    float v = 0.f;
    bool is_gc = false;
    if (g < G && k >= 0) {
      if (k < D) v = miv[(size_t)g * D + k] * bscale[k];
      else if (k < 2 * D) v = -0.5f * iv[(size_t)g * D + (k - D)] * bscale[k];
      else if (k <= 2 * D + 1 && !lo_part) { v = gconsts[g]; is_gc = true; }
    }
    if (!(fabsf(v) <= 3.0e4f)) bad = true;  // also catches -inf gconsts and NaN
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    if (is_gc) bp[i] = k == 2 * D ? hi : lo;  // gconst: fp16 part in column 2D, residual in column 2D+1
    else bp[i] = lo_part ? lo : hi;
  }
  if (bad) atomicOr(flags, 1);
}

// max over the batch of |x_d| * ascale[d] (NaN counts as huge), as float bits.
__global__ void feat_absmax_kernel(const float *__restrict__ feats, int64_t n, int D,
                                   const float *__restrict__ ascale, unsigned *__restrict__ out) {
  unsigned m = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fabsf(feats[i]) * ascale[(int)(i % D)];
    m = max(m, __float_as_uint(v));  // non-negative floats and NaNs order like their bit patterns
  }
  for (int off = 16; off > 0; off >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// ------------------------------------------------------------------ the kernel --
struct TcArgs {
  const float *feats;      // T x D
  int64_t T;
  int D, K8, Kc, n_chunks; // K8 = 2D+2 rounded up to UMMA_K (hi.hi product), Kc = 2D rounded up (cross
                           // products); n_chunks = 128-byte chunks per A_hi / A_lo row held in smem
  StageTab tab;            // chunks (= TMA stages) of the streamed operand B' and what each holds
  const float *ascale;     // fp16 path: per-column power-of-two scale of [x | x^2] (2D floats); NULL = none
  const unsigned *gate;    // NULL, or device word with max |x*ascale| bits: see gate_limit
  float gate_limit;        // fp16 kernel runs iff *gate <= limit, tf32 kernel iff *gate > limit
  int gate_run_if_above;
  int stages;              // B ring depth
  int a_slots;             // 1 or 2 A buffers (2: the next item's A is built under the current item's MMAs)
  const int32_t *offsets;  // P+1
  const int32_t *tile_g0;  // n_tiles
  const int32_t *tile_p0;  // n_tiles+1
  int two_chunk_segs;        // some pdf has 17..32 Gaussians (descriptor flag kSegTwoChunks in use)
  const int2 *epi_hdr;       // per (tile, epilogue group): {start in seg[], first run | number of runs << 24}
  const uint32_t *runs;      // per run of equal-length segments: length | count << 8
  const uint32_t *seg;       // per segment: column | pdf << 8; per (tile, group) in run order, + 2 sentinels
  int n_tiles, n_splits, tiles_per_split;
  // optional tile subset (the batched aligner: only the model tiles that hold a pdf of the frames' utterances): frame
  // tiles 2q and 2q + 1 run the tiles sub_tiles[sub_off[q] .. sub_off[q + 1]) instead of all of them (n_splits == 1)
  const int32_t *sub_off;
  const int32_t *sub_tiles;
  int sub_shift;  // 1: one list per pair of frame tiles, 0: one per tile (no CTA pairs)
  int64_t n_items;
  float scale;
  float *out;              // pdf-major, out[p*ld + t]
  int64_t ld;
  float *scratch;          // one device word: store target of rows beyond T
  int *err;
  int debug_mode;          // 0 = normal; others: timing experiments, compiled in with -DKHG_EXPERIMENTS only (KHG_TC_DEBUG_MODE)
  int cluster;             // 1 = plain launch; 2 = CTA pairs sharing the streamed operand by TMA multicast
                           // (n_splits == 1: CTA rank r of pair k works on frame tile 2k + r)
};

// TWO: the model has pdfs of 17..32 Gaussians (two-load segments); GRP: it has groups of short pdfs
// read by one load (epi_run_multi).  Separate instantiations, so that models without them run
// exactly the plain single-load epilogue (its code layout is worth 2-4 % on the C4 shape).
template <bool F16, bool TWO, bool GRP>
__global__ void __launch_bounds__(kTcThreads, 1)
loglikes_tc_kernel(const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_b_half, TcArgs a) {
  constexpr int kChunkK = Elem<F16>::kChunkK, kUmmaK = Elem<F16>::kUmmaK;
  constexpr uint32_t kIdesc = make_idesc<F16>();
  if (a.gate != nullptr) {  // precision-path gate decided on the device (no host round trip)
    const bool above = !(__uint_as_float(*a.gate) <= a.gate_limit);
    if (above != (a.gate_run_if_above != 0)) return;
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *base_ptr = smem_raw + (base - raw);
  const int NCH = a.n_chunks, S = a.stages;
  const uint32_t kASlotBytes = 2u * NCH * kAChunkBytes;  // one A buffer: hi chunks, then lo chunks
  const uint32_t sA_hi = base;
  const uint32_t sA_lo = base + NCH * kAChunkBytes;
  const uint32_t sB = base + (uint32_t)a.a_slots * kASlotBytes;
  const uint32_t sBar = sB + S * kBStageBytes;
  // barrier slots (8 B each)
  auto b_full = [&](int s) { return sBar + 8u * s; };
  auto b_empty = [&](int s) { return sBar + 8u * (8 + s); };
  // A slot s (item it uses slot it % a_slots, phase (it / a_slots) & 1)
  auto a_full = [&](uint32_t sl) { return sBar + 8u * (sl == 0 ? 16 : 23); };
  auto a_free = [&](uint32_t sl) { return sBar + 8u * (sl == 0 ? 17 : 24); };
  const uint32_t a_two = a.a_slots == 2 ? 1u : 0u;
  auto acc_full = [&](int b) { return sBar + 8u * (18 + b); };
  auto acc_empty = [&](int b) { return sBar + 8u * (20 + b); };
  const uint32_t tmem_slot = sBar + 8u * 22;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(base_ptr + (tmem_slot - base));

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  // work items of this CTA: item(k) for k = k_first, k_first + k_stride, ... < k_end
  const bool clu = a.cluster >= 2;        // launched as CTA pairs
  const bool mcast = a.cluster == 2;      // each CTA runs its own MMAs; the operand stages are multicast
  const uint32_t crank = clu ? cluster_ctarank() : 0u;
  const int64_t k_first = clu ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t k_stride = clu ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x;
  const int64_t k_end = clu ? (a.n_items + 1) / 2 : a.n_items;  // (a phantom tile beyond T pads an odd count)
  auto item_of = [&](int64_t k) -> int64_t { return clu ? 2 * k + crank : k; };

  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(b_full(s), 1);
        mbar_init(b_empty(s), mcast ? 2 : 1);  // multicast pair: the MMA warps of both CTAs release a stage
      }
      for (uint32_t sl = 0; sl < 2; ++sl) {
        mbar_init(a_full(sl), 1);
        mbar_init(a_free(sl), 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(acc_full(b), 1);
        mbar_init(acc_empty(b), 4 * kEpiGroups);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else if (warp == kBuilderWarp0 || warp == kBuilderWarp0 + 1) {
    // one-time A init: zero everything, then the constant-1 column (k = 2D) of A_hi
    const int b = threadIdx.x - 32 * kBuilderWarp0;
    float4 *z = reinterpret_cast<float4 *>(base_ptr);
    for (int i = b; i < a.a_slots * 2 * NCH * kAChunkBytes / 16; i += kBuilderThreads) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("bar.sync 1, 64;" ::: "memory");
    for (int sl = 0; sl < a.a_slots; ++sl)
      for (int row = b; row < kTileM; row += kBuilderThreads) {
        a_store_split<F16>(base_ptr + sl * kASlotBytes, NCH * kAChunkBytes, row, 2 * a.D, 1.0f);
        a_store_split<F16>(base_ptr + sl * kASlotBytes, NCH * kAChunkBytes, row, 2 * a.D + 1, 1.0f);
      }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (clu) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_ptr, 0);

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    {
      const bool leader = elect_one();
      uint32_t st = 0, ph = 1;  // producer waits on "empty" with the inverted phase
      for (int64_t k = k_first; k < k_end; k += k_stride) {
        const int64_t item = item_of(k);
        const int split = (int)(item % a.n_splits);
        int j0 = split * a.tiles_per_split, j1 = min(a.n_tiles, j0 + a.tiles_per_split);
        if (a.sub_off) {
          j0 = __shfl_sync(0xffffffffu, __ldg(a.sub_off + (item >> a.sub_shift)), 0);
          j1 = __shfl_sync(0xffffffffu, __ldg(a.sub_off + (item >> a.sub_shift) + 1), 0);
        }
        for (int jj = j0; jj < j1; ++jj) {
          const int j = a.sub_off ? __shfl_sync(0xffffffffu, __ldg(a.sub_tiles + jj), 0) : jj;
          const int g0 = __shfl_sync(0xffffffffu, __ldg(a.tile_g0 + j), 0);
          for (int c = 0; c < a.tab.n; ++c) {
            mbar_wait(b_empty(st), ph);
            if (leader) {
#ifdef KHG_EXPERIMENTS
              // (timing only, results are garbage) 3 = MMA rate without operand traffic, 4 / 5 = only the first /
              // all but the second chunk of every tile is loaded, 6 / 7 = the 2nd / every chunk is always the SAME
              // box of the operand (full shared-memory write volume, no L2 footprint); plain launches only
              if (a.debug_mode == 3 || (a.debug_mode == 4 && c != 0) || (a.debug_mode == 5 && c == 1)) {
                mbar_arrive(b_full(st));
              } else if (!mcast && (a.debug_mode == 6 || a.debug_mode == 7)) {
                const bool same = a.debug_mode == 7 || c == 1;
                mbar_expect_tx(b_full(st), kBStageBytes);
                tma_load_2d(sB + st * kBStageBytes, &map_b, b_full(st), same ? 0 : c * kChunkK, same ? 0 : g0);
              } else
#endif
              if (mcast) {
                // this CTA's half of the stage's rows, into both CTAs; the other half arrives from the peer
                mbar_expect_tx(b_full(st), kBStageBytes);
                tma_load_2d_mc(sB + st * kBStageBytes + crank * (kBStageBytes / 2), &map_b_half, b_full(st), c * kChunkK,
                               g0 + (int)crank * (kTileN / 2), (uint16_t)3);
              } else {
                mbar_expect_tx(b_full(st), kBStageBytes);
                tma_load_2d(sB + st * kBStageBytes, &map_b, b_full(st), c * kChunkK, g0);
              }
            }
            __syncwarp();
            if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // All 32 lanes run the loop (uniform control flow, see elect_one); one elected lane issues.
    // The instruction stream is kept as short as possible (no divisions, 32-bit descriptor
    // arithmetic): this warp shares an SM sub-partition with four busy epilogue warps.
    {
      const bool leader = elect_one();
      uint32_t st = 0, ph = 0, acc_it = 0, a_it = 0;
      const uint32_t a_hi_base = umma_desc_lo(sA_hi), a_lo_base = umma_desc_lo(sA_lo), b0 = umma_desc_lo(sB);
      const uint32_t a_slot_desc = kASlotBytes >> 4;
      constexpr uint32_t kAChunkDesc = kAChunkBytes >> 4, kBStageDesc = kBStageBytes >> 4;
      for (int64_t k = k_first; k < k_end; k += k_stride, ++a_it) {
        const int64_t item = item_of(k);
        const int split = (int)(item % a.n_splits);
        int j0 = split * a.tiles_per_split, j1 = min(a.n_tiles, j0 + a.tiles_per_split);
        if (a.sub_off) {  // (only the number of tiles matters here)
          j0 = __shfl_sync(0xffffffffu, __ldg(a.sub_off + (item >> a.sub_shift)), 0);
          j1 = __shfl_sync(0xffffffffu, __ldg(a.sub_off + (item >> a.sub_shift) + 1), 0);
        }
        const uint32_t asl = a_it & a_two, aph = (a_two ? a_it >> 1 : a_it) & 1;
        const uint32_t a_hi0 = a_hi_base + asl * a_slot_desc, a_lo0 = a_lo_base + asl * a_slot_desc;
        mbar_wait(a_full(asl), aph);
        tc_fence_after();
        for (int j = j0; j < j1; ++j, ++acc_it) {
          const uint32_t buf = acc_it & 1;
          mbar_wait(acc_empty(buf), ((acc_it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * 256;
          uint32_t accum = 0;
          for (int c = 0; c < a.tab.n; ++c) {
            const uint32_t e = a.tab.e[c];
            const int h0 = e & 0xff, nh = (e >> 8) & 0xf, l0 = (e >> 12) & 0xff, nl = (e >> 20) & 0xf;
            mbar_wait(b_full(st), ph);
            tc_fence_after();
            const uint32_t db = b0 + st * kBStageDesc;
            constexpr int kSpc = kChunkK / kUmmaK;  // K steps per chunk (4)
#ifdef KHG_EXPERIMENTS
            if (a.debug_mode != 2)  // (2 = TMA rate without MMAs)
#endif
            {
              // hi steps of B feed A_hi (hi.hi) and, inside the feature columns, A_lo (lo.hi)
#pragma unroll 4
              for (int s = 0; s < nh; ++s) {
                const int q = h0 + s;
                const uint32_t ao = (uint32_t)(q / kSpc) * kAChunkDesc + (uint32_t)(q % kSpc) * 2;
                if (leader) tc_mma<F16>(tmem_d, a_hi0 + ao, db + 2 * s, kIdesc, accum);
                accum = 1;
                if (q * kUmmaK < a.Kc && leader) tc_mma<F16>(tmem_d, a_lo0 + ao, db + 2 * s, kIdesc, 1);
              }
              // lo steps of B feed A_hi (hi.lo)
#pragma unroll 4
              for (int s = 0; s < nl; ++s) {
                const int q = l0 + s;
                const uint32_t ao = (uint32_t)(q / kSpc) * kAChunkDesc + (uint32_t)(q % kSpc) * 2;
                if (leader) tc_mma<F16>(tmem_d, a_hi0 + ao, db + 2 * (nh + s), kIdesc, 1);
              }
            }
            if (leader) {
              if (mcast) tc_commit_mc(b_empty(st), (uint16_t)3);
              else tc_commit(b_empty(st));
            }
            if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
          }
          if (leader) tc_commit(acc_full(buf));
        }
        if (leader) tc_commit(a_free(asl));
      }
    }
  } else if (warp == kBuilderWarp0 || warp == kBuilderWarp0 + 1) {
    // ===================== A builders =====================
    const int b = threadIdx.x - 32 * kBuilderWarp0;
    const int D = a.D;
    uint32_t a_it = 0;
    for (int64_t k = k_first; k < k_end; k += k_stride, ++a_it) {
      const int64_t item = item_of(k);
      const int64_t t0 = (item / a.n_splits) * kTileM;
      const int64_t valid = max((int64_t)0, min((int64_t)kTileM, a.T - t0)) * D;  // (0 for the pair's phantom tile)
#ifdef KHG_EXPERIMENTS
      // (9, timing only: every item re-reads the CTA's first tile — the full A build on repeating data)
      const float *src = a.feats + (a.debug_mode == 9 ? item_of(k_first) / a.n_splits * kTileM : t0) * D;
#else
      const float *src = a.feats + t0 * D;
#endif
      // The tile's features come straight from HBM (every frame is read once per E-step), so the loads
      // are issued kABatch at a time per thread — a one-load-per-iteration loop pays the full DRAM
      // latency ~80 times in a row, during which the tensor pipe waits for A — and the first batch is
      // already in registers when the MMAs of the previous item release the A buffer.
      constexpr int kABatch = 16;
#ifdef KHG_EXPERIMENTS
      const int n_el = (a.debug_mode == 8 && a_it > 0) ? 0 : kTileM * D;  // (8, timing only: A built once)
#else
      const int n_el = kTileM * D;
#endif
      const uint32_t asl = a_it & a_two, aph = (a_two ? a_it >> 1 : a_it) & 1;
      uint8_t *a_ptr = base_ptr + asl * kASlotBytes;
      float xs[kABatch], ys[kABatch];
      auto load_batch = [&](float (&v)[kABatch], int e0) {
#pragma unroll
        for (int j = 0; j < kABatch; ++j) {
          const int e = e0 + j * kBuilderThreads;
          // L2-only load (ld.global.cg): every frame is read once, nothing to keep in L1 (measured: no
          // difference to ld.global.nc or .cs)
          v[j] = e < valid ? KHG_FEAT_LOAD(src + e) : 0.f;
        }
      };
      auto store_batch = [&](const float (&v)[kABatch], int e0) {
#pragma unroll
        for (int j = 0; j < kABatch; ++j) {
          const int e = e0 + j * kBuilderThreads;
          if (e < n_el) {
            const float x = v[j];
            const int row = e / D, d = e - row * D;
            const float q = x * x;  // data.array().square(), csrc/decodable-am-diag-gmm.cc:57
            if (F16) {  // exact power-of-two column scaling keeps both operands inside fp16's range
              a_store_split<F16>(a_ptr, NCH * kAChunkBytes, row, d, x * __ldg(a.ascale + d));
              a_store_split<F16>(a_ptr, NCH * kAChunkBytes, row, D + d, q * __ldg(a.ascale + D + d));
            } else {
              a_store_split<F16>(a_ptr, NCH * kAChunkBytes, row, d, x);
              a_store_split<F16>(a_ptr, NCH * kAChunkBytes, row, D + d, q);
            }
          }
        }
      };
      constexpr int kStep = kBuilderThreads * kABatch;
      load_batch(xs, b);
      if (b == 0) mbar_wait(a_free(asl), aph ^ 1);
      asm volatile("bar.sync 1, 64;" ::: "memory");
      for (int e0 = b; e0 < n_el; e0 += 2 * kStep) {  // the next batch's loads are in flight under the stores
        const int e1 = e0 + kStep, e2 = e1 + kStep;
        if (e1 < n_el) load_batch(ys, e1);
        store_batch(xs, e0);
        if (e2 < n_el) load_batch(xs, e2);
        if (e1 < n_el) store_batch(ys, e1);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 64;" ::: "memory");
      if (b == 0) mbar_arrive(a_full(asl));
    }
  } else if (warp < kEpiWarps) {
    // ===================== epilogue =====================
    const int eg = warp >> 2;           // epilogue group: handles segments eg, eg+4, ...
    const int quad = warp & 3;                // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;
    uint32_t acc_it = 0;
    bool bad = false;
    for (int64_t k = k_first; k < k_end; k += k_stride) {
      const int64_t item = item_of(k);
      const int split = (int)(item % a.n_splits);
      int j0 = split * a.tiles_per_split, j1 = min(a.n_tiles, j0 + a.tiles_per_split);
      if (a.sub_off) {
        j0 = __ldg(a.sub_off + (item >> a.sub_shift));
        j1 = __ldg(a.sub_off + (item >> a.sub_shift) + 1);
      }
      const int64_t t = (item / a.n_splits) * kTileM + row;
      const bool valid = t < a.T;
      // rows beyond T store into a scratch word (ld_bytes = 0): no predicate in the hot loop
#ifdef KHG_EXPERIMENTS
      // (timing only: 10 = every tile stores into the block's first 128 frames — an L2-resident window, no DRAM
      // write stream; 11 = every store goes to the scratch word)
      // 12 = tile-major addressing out[t / 128][p][128] inside the same block: every CTA's stores of an item fall into ONE
      // contiguous P x 512-byte region instead of P pieces 4 * ld bytes apart (address-translation locality of the store stream)
      const bool to_scratch = !valid || a.debug_mode == 11;
      char *out_t = to_scratch ? reinterpret_cast<char *>(a.scratch) : reinterpret_cast<char *>(a.out + (a.debug_mode == 10 ? (int64_t)row : t));
      uint32_t ld_bytes = to_scratch ? 0u : (uint32_t)(a.ld * 4);
      if (a.debug_mode == 12 && valid) {
        const int64_t P = __ldg(a.tile_p0 + a.n_tiles);
        out_t = reinterpret_cast<char *>(a.out + (item / a.n_splits) * P * kTileM + row);
        ld_bytes = (uint32_t)(kTileM * 4);
      }
#else
      // rows beyond T store into a scratch word (ld_bytes = 0): no predicate in the hot loop
      char *out_t = valid ? reinterpret_cast<char *>(a.out + t) : reinterpret_cast<char *>(a.scratch);
      const uint32_t ld_bytes = valid ? (uint32_t)(a.ld * 4) : 0u;
#endif
      EpiState e;
      e.scale = a.scale;
      e.nan_acc = 0.f;
      e.ld_bytes = ld_bytes;
      e.out_t = out_t;
      for (int jj = j0; jj < j1; ++jj, ++acc_it) {
        const int buf = acc_it & 1;
        const int j = a.sub_off ? __ldg(a.sub_tiles + jj) : jj;
        const int2 h = __ldg(a.epi_hdr + (size_t)kEpiGroups * j + eg);
        const uint32_t *rp = a.runs + (h.y & 0xffffff);
        int nr = (int)((uint32_t)h.y >> 24);
        e.sp = a.seg + h.x;
        e.d = __ldg(e.sp);  // (a sentinel when the list is empty)
        e.dn = __ldg(e.sp + 1);
        uint32_t run = __ldg(rp);
        mbar_wait(acc_full(buf), (acc_it >> 1) & 1);
        tc_fence_after();
        e.trow = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256;
#ifdef KHG_EXPERIMENTS
        if (a.debug_mode == 1) nr = 0;  // (1 = the epilogue skips the LSE: pipeline ceiling)
#endif
        if (nr > 0) {
          tc_ld16_issue(e.trow + (e.d & 0xffu), e.t);
#define KHG_CASE(L) case L: epi_run<L, KHG_TWO>(e, cnt); break;
#define KHG_CASE2(L) case 16 + L: epi_run2<L>(e, cnt); break;
#define KHG_CASEM(L) case L: epi_run_multi<L, KHG_TWO>(e, cnt); break;
#define KHG_CASESM KHG_CASEM(1) KHG_CASEM(2) KHG_CASEM(3) KHG_CASEM(4) KHG_CASEM(5) KHG_CASEM(6) KHG_CASEM(7) KHG_CASEM(8)
#define KHG_CASES1 KHG_CASE(1) KHG_CASE(2) KHG_CASE(3) KHG_CASE(4) KHG_CASE(5) KHG_CASE(6) KHG_CASE(7) KHG_CASE(8) \
                   KHG_CASE(9) KHG_CASE(10) KHG_CASE(11) KHG_CASE(12) KHG_CASE(13) KHG_CASE(14) KHG_CASE(15) KHG_CASE(16)
#define KHG_CASES2 KHG_CASE2(1) KHG_CASE2(2) KHG_CASE2(3) KHG_CASE2(4) KHG_CASE2(5) KHG_CASE2(6) KHG_CASE2(7) KHG_CASE2(8) \
                   KHG_CASE2(9) KHG_CASE2(10) KHG_CASE2(11) KHG_CASE2(12) KHG_CASE2(13) KHG_CASE2(14) KHG_CASE2(15) KHG_CASE2(16)
          if constexpr (TWO) {
            tc_ld16_issue_if(e.trow + (e.d & 0xffu) + 16, e.t2, (e.d & kSegTwoChunks) != 0);
            for (; nr > 0; --nr) {
              const int len = run & 0xff, cnt = (run >> 8) & 0x7fffff;
              const bool grouped = GRP && (run >> 31) != 0;
              if (nr > 1) run = __ldg(++rp);
#define KHG_TWO true
              if (GRP && grouped) {
                switch (len) { KHG_CASESM }
              } else {
                switch (len) {
                  KHG_CASES1 KHG_CASES2
                  default: epi_run_long(e, cnt, len); break;
                }
              }
#undef KHG_TWO
            }
            tc_ld16_wait2(e.t, e.t2);  // the (unused) sentinel load issued after the last segment
          } else {
            for (; nr > 0; --nr) {
              const int len = run & 0xff, cnt = (run >> 8) & 0x7fffff;
              const bool grouped = GRP && (run >> 31) != 0;
              if (nr > 1) run = __ldg(++rp);
#define KHG_TWO false
              if (GRP && grouped) {
                switch (len) { KHG_CASESM }
              } else {
                switch (len) {
                  KHG_CASES1
                  default: epi_run_long(e, cnt, len); break;
                }
              }
#undef KHG_TWO
            }
            tc_ld16_wait(e.t);  // the (unused) sentinel load issued after the last segment
          }
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
      }
      const float nan_acc = e.nan_acc;
      if (valid && nan_acc != nan_acc) bad = true;  // r*0 is NaN exactly for a NaN/Inf result
    }
    if (bad) atomicOr(a.err, ERR_NONFINITE);
  }

  tc_fence_before();
  __syncthreads();
  if (clu) cluster_sync_all();  // no CTA leaves while its peer can still multicast into it
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------ host side --
static bool tc_shape_ok(const khg_model *m, bool f16) {
  const int K = 2 * m->dim + 2;  // feature columns + gconst + its residual
  const int ck = f16 ? Elem<true>::kChunkK : Elem<false>::kChunkK;
  return (K + ck - 1) / ck <= kMaxChunks;
}
// pdfs of more than 240 Gaussians run as virtual sub-pdfs; that form has no gated fp32 fall-back, so it needs the
// tf32 split as the out-of-range path of the fp16 split
bool tc_supported(const khg_model *m) {
  return (tc_shape_ok(m, false) || tc_shape_ok(m, true)) && (m->max_gp <= kTileN || tc_shape_ok(m, false));
}

void tc_pack_free(khg_model *m) {
  TcPack &t = m->tc;
  gs_free(m);
  cudaFree(t.bhi); cudaFree(t.blo); cudaFree(t.tile_g0); cudaFree(t.tile_p0);
  cudaFree(t.hhi); cudaFree(t.hlo); cudaFree(t.ascale); cudaFree(t.gate);
  cudaFree(t.epi_hdr); cudaFree(t.runs); cudaFree(t.seg); cudaFree(t.d_vfirst);
  t.d_vfirst = nullptr;
  t.epi_hdr = nullptr;
  t.runs = nullptr;
  t.seg = nullptr;
  t.bhi = t.blo = nullptr;
  t.hhi = t.hlo = nullptr;
  t.ascale = nullptr;
  t.gate = nullptr;
  t.tile_g0 = t.tile_p0 = nullptr;
  t.ready = t.f16_ready = t.tf32_ready = false;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D tensor map over a rows x KP operand matrix: box = one 128-byte K chunk x 240 rows.
static khg_status make_map(CUtensorMap *map, void *ptr, int KP, int rows, bool f16, int box_rows = kTileN) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return KHG_ERR_CUDA;
  }
  const int eb = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)KP, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)KP * eb};
  cuuint32_t box[2] = {(cuuint32_t)(128 / eb), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return KHG_ERR_CUDA;
  }
  return KHG_OK;
}

// fp16-split operand pack.  Column d of [x | x^2] is scaled by 2^-k_d / 2^-2k_d and the
// matching model columns by the inverse (exact), with 2^k_d ~ the rms magnitude the model
// implies for feature d, so that both operands sit comfortably inside fp16's range.
static khg_status tc_pack_build_f16(khg_model *m) {
  TcPack &t = m->tc;
  const int D = m->dim, G = m->G;
  t.K16 = (t.K + 1 + 15) / 16 * 16;  // hi.hi product spans 2D+2 columns (gconst + its residual)
  t.KP16 = (t.K16 + 63) / 64 * 64;   // A_hi / A_lo row width in smem
  const int Kc16 = (2 * D + 15) / 16 * 16;
  t.tab16 = make_stage_tab(t.K16, Kc16, 64, 16, true);
  t.KPB16 = t.tab16.n * 64;
  std::vector<float> miv((size_t)G * D), iv((size_t)G * D);
  KHG_CUDA_TRY(cudaMemcpyAsync(miv.data(), m->d_miv, sizeof(float) * miv.size(), cudaMemcpyDeviceToHost, m->stream));
  KHG_CUDA_TRY(cudaMemcpyAsync(iv.data(), m->d_iv, sizeof(float) * iv.size(), cudaMemcpyDeviceToHost, m->stream));
  KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  std::vector<double> m2(D, 0.0);
  for (int g = 0; g < G; ++g)
    for (int d = 0; d < D; ++d) {
      const double var = 1.0 / (double)iv[(size_t)g * D + d], mean = (double)miv[(size_t)g * D + d] * var;
      m2[d] += mean * mean + var;
    }
  std::vector<float> ascale(2 * D), bscale(2 * D);
  for (int d = 0; d < D; ++d) {
    double rms = std::sqrt(m2[d] / G);
    int k = (rms > 0.0 && std::isfinite(rms)) ? (int)std::lround(std::log2(rms)) : 0;
    k = std::max(-20, std::min(20, k));
    ascale[d] = std::ldexp(1.0f, -k);
    ascale[D + d] = std::ldexp(1.0f, -2 * k);
    bscale[d] = std::ldexp(1.0f, k);
    bscale[D + d] = std::ldexp(1.0f, 2 * k);
  }
  float *d_bscale = nullptr;
  int *d_flag = nullptr;
  KHG_CUDA_TRY(cudaMalloc(&t.ascale, sizeof(float) * 2 * D));
  KHG_CUDA_TRY(cudaMalloc(&t.gate, sizeof(unsigned)));
  KHG_CUDA_TRY(cudaMalloc(&d_bscale, sizeof(float) * 2 * D));
  KHG_CUDA_TRY(cudaMalloc(&d_flag, sizeof(int)));
  KHG_CUDA_TRY(cudaMemcpyAsync(t.ascale, ascale.data(), sizeof(float) * 2 * D, cudaMemcpyHostToDevice, m->stream));
  KHG_CUDA_TRY(cudaMemcpyAsync(d_bscale, bscale.data(), sizeof(float) * 2 * D, cudaMemcpyHostToDevice, m->stream));
  KHG_CUDA_TRY(cudaMemsetAsync(d_flag, 0, sizeof(int), m->stream));
  KHG_CUDA_TRY(cudaMalloc(&t.hhi, sizeof(__half) * (size_t)t.rows * t.KPB16));
  size_t total = (size_t)t.rows * t.KPB16;
  tc_pack_f16_kernel<<<(unsigned)std::min<size_t>(2048, (total + 255) / 256), 256, 0, m->stream>>>(
      G, D, t.tab16, t.KPB16, t.rows, m->d_miv, m->d_iv, m->d_gconsts, d_bscale, static_cast<__half *>(t.hhi), d_flag);
  ++g_launch_count;
  int flag = 0;
  KHG_CUDA_TRY(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  cudaFree(d_bscale);
  cudaFree(d_flag);
  if (flag) return KHG_OK;  // model does not fit the fp16 split: stay on tf32 (f16_ready = false)
  KHG_TRY(make_map(&t.hmap_hi, t.hhi, t.KPB16, t.rows, true));
  KHG_TRY(make_map(&t.hmap_lo, t.hhi, t.KPB16, t.rows, true, kTileN / 2));  // half-stage box of the CTA-pair form
  t.f16_ready = true;
  return KHG_OK;
}

khg_status tc_pack_build(khg_model *m) {
  TcPack &t = m->tc;
  tc_pack_free(m);
  const int D = m->dim, G = m->G;
  t.K = 2 * D + 1;
  t.K8 = (t.K + 1 + 7) / 8 * 8;
  t.KP = (t.K8 + 31) / 32 * 32;      // A_hi / A_lo row width in smem
  const int Kc8 = (2 * D + 7) / 8 * 8;
  t.tab8 = make_stage_tab(t.K8, Kc8, 32, 8, false);
  t.KPB = t.tab8.n * 32;
  t.rows = (G + kTileN + 15) / 16 * 16;
  // virtual pdfs: a pdf of more than 240 Gaussians becomes ceil(len / 240) pieces of (almost) equal size
  std::vector<int32_t> vfirst(m->P + 1);
  t.v_off.assign(1, 0);
  for (int q = 0; q < m->P; ++q) {
    vfirst[q] = (int32_t)t.v_off.size() - 1;
    const int len = m->h_offsets[q + 1] - m->h_offsets[q], k = std::max(1, (len + kTileN - 1) / kTileN);
    for (int i = 1; i <= k; ++i) t.v_off.push_back(m->h_offsets[q] + (int32_t)((int64_t)len * i / k));
  }
  t.Pv = (int)t.v_off.size() - 1;
  vfirst[m->P] = t.Pv;
  t.h_vfirst = vfirst;
  const int P = t.Pv;                         // from here on "pdf" means virtual pdf
  const std::vector<int32_t> &off = t.v_off;
  if (t.Pv != m->P) {
    KHG_CUDA_TRY(cudaMalloc(&t.d_vfirst, sizeof(int32_t) * (m->P + 1)));
    KHG_CUDA_TRY(cudaMemcpy(t.d_vfirst, vfirst.data(), sizeof(int32_t) * (m->P + 1), cudaMemcpyHostToDevice));
  }
  // pdf-aligned N tiles (greedy)
  t.h_tile_g0.clear();
  t.h_tile_p0.clear();
  int p = 0;
  while (p < P) {
    int g0 = off[p];
    t.h_tile_g0.push_back(g0);
    t.h_tile_p0.push_back(p);
    int q = p;
    while (q < P && off[q + 1] - g0 <= kTileN) ++q;
    if (q == p) {
      set_error("internal: a virtual pdf has more than 240 Gaussians");
      return KHG_ERR_UNSUPPORTED;
    }
    p = q;
  }
  t.h_tile_p0.push_back(P);
  t.n_tiles = (int)t.h_tile_g0.size();
  // epilogue tables: per tile, its pdfs grouped by Gaussian count (one dispatch per class)
  {
    std::vector<int2> hdr((size_t)kEpiGroups * t.n_tiles);
    std::vector<uint32_t> runs, seg;
    seg.reserve(P + 2 * hdr.size());
    t.two_chunk_segs = false;
    t.grouped_segs = false;
    for (int j = 0; j < t.n_tiles; ++j) {
      const int pa = t.h_tile_p0[j], pb = t.h_tile_p0[j + 1], g0 = t.h_tile_g0[j];
      // work items of the tile: groups of 16 / len column-adjacent pdfs of the same length <= 8 (one
      // TMEM load per group), single pdfs otherwise; key = grouped << 8 | len
      std::vector<std::pair<int, int>> by_len;  // (key, first pdf)
      for (int q = pa; q < pb;) {
        const int len = off[q + 1] - off[q];
        const int ns = len <= 8 ? 16 / len : 1;
        bool grp = ns > 1 && q + ns <= pb;
        for (int k = 1; grp && k < ns; ++k) grp = off[q + k + 1] - off[q + k] == len;
        if (grp) t.grouped_segs = true;
        by_len.emplace_back((grp ? 256 : 0) | len, q);
        q += grp ? ns : 1;
      }
      std::stable_sort(by_len.begin(), by_len.end(), [](const std::pair<int, int> &x, const std::pair<int, int> &y) { return x.first < y.first; });
      // item i of the sorted order goes to epilogue group i % kEpiGroups (balanced)
      for (int eg = 0; eg < kEpiGroups; ++eg) {
        int2 h;
        h.x = (int)seg.size();
        const size_t r0 = runs.size();
        for (size_t i = eg; i < by_len.size(); i += kEpiGroups) {
          const int key = by_len[i].first, len = key & 255, q = by_len[i].second;
          const bool grp = key >= 256;
          // a run: equal keys (pdfs of > 254 Gaussians: one run each); length | count << 8 | grouped << 31
          const uint32_t tag = (uint32_t)std::min(len, 255) | (grp ? 1u << 31 : 0u);
          if (runs.size() > r0 && (runs.back() & 0x800000ffu) == tag && len < 255 && ((runs.back() >> 8) & 0x7fffffu) < 0x7fffffu)
            runs.back() += 1u << 8;
          else
            runs.push_back(tag | 1u << 8);
          if (len > 16 && len <= 32) t.two_chunk_segs = true;
          seg.push_back((uint32_t)(off[q] - g0) | (len > 16 && len <= 32 ? kSegTwoChunks : 0u) | (uint32_t)q << kSegPdfShift);
        }
        seg.push_back(0);  // two sentinels: the epilogue prefetches two descriptors ahead
        seg.push_back(0);
        h.y = (int)(r0 | (runs.size() - r0) << 24);
        hdr[(size_t)kEpiGroups * j + eg] = h;
      }
    }
    runs.push_back(0);  // read (unused) by groups without segments
    if (P >= (1 << (32 - kSegPdfShift)) || runs.size() >= (1u << 24)) {
      set_error("too many pdfs for the tensor-core epilogue tables");
      return KHG_ERR_UNSUPPORTED;
    }
    KHG_CUDA_TRY(cudaMalloc(&t.epi_hdr, sizeof(int2) * hdr.size()));
    KHG_CUDA_TRY(cudaMalloc(&t.runs, sizeof(uint32_t) * runs.size()));
    KHG_CUDA_TRY(cudaMalloc(&t.seg, sizeof(uint32_t) * seg.size()));
    KHG_CUDA_TRY(cudaMemcpy(t.epi_hdr, hdr.data(), sizeof(int2) * hdr.size(), cudaMemcpyHostToDevice));
    KHG_CUDA_TRY(cudaMemcpy(t.runs, runs.data(), sizeof(uint32_t) * runs.size(), cudaMemcpyHostToDevice));
    KHG_CUDA_TRY(cudaMemcpy(t.seg, seg.data(), sizeof(uint32_t) * seg.size(), cudaMemcpyHostToDevice));
    // a pdf whose Gaussians all have gconst = -inf makes LogSumExp NaN in the reference
    // (csrc/eigen.cc:14-18 -> throw at csrc/decodable-am-diag-gmm.cc:63-65) for every frame
    std::vector<float> gc(G);
    KHG_CUDA_TRY(cudaMemcpy(gc.data(), m->d_gconsts, sizeof(float) * G, cudaMemcpyDeviceToHost));
    t.dead_pdf = false;
    for (int q = 0; q < m->P && !t.dead_pdf; ++q) {
      bool all_dead = true;
      for (int g = m->h_offsets[q]; g < m->h_offsets[q + 1]; ++g) all_dead = all_dead && std::isinf(gc[g]) && gc[g] < 0;
      t.dead_pdf = all_dead;
    }
  }
  KHG_CUDA_TRY(cudaMalloc(&t.tile_g0, sizeof(int32_t) * t.n_tiles));
  KHG_CUDA_TRY(cudaMalloc(&t.tile_p0, sizeof(int32_t) * (t.n_tiles + 1)));
  KHG_CUDA_TRY(cudaMemcpyAsync(t.tile_g0, t.h_tile_g0.data(), sizeof(int32_t) * t.n_tiles, cudaMemcpyHostToDevice, m->stream));
  KHG_CUDA_TRY(cudaMemcpyAsync(t.tile_p0, t.h_tile_p0.data(), sizeof(int32_t) * (t.n_tiles + 1), cudaMemcpyHostToDevice, m->stream));
  if (tc_shape_ok(m, false)) {
    KHG_CUDA_TRY(cudaMalloc(&t.bhi, sizeof(float) * (size_t)t.rows * t.KPB));
    size_t total = (size_t)t.rows * t.KPB;
    tc_pack_kernel<<<(unsigned)std::min<size_t>(2048, (total + 255) / 256), 256, 0, m->stream>>>(G, D, t.tab8, t.KPB, t.rows, m->d_miv, m->d_iv, m->d_gconsts, t.bhi);
    ++g_launch_count;
    KHG_CUDA_TRY(cudaGetLastError());
    KHG_TRY(make_map(&t.map_hi, t.bhi, t.KPB, t.rows, false));
    KHG_TRY(make_map(&t.map_lo, t.bhi, t.KPB, t.rows, false, kTileN / 2));  // half-stage box of the CTA-pair form
    t.tf32_ready = true;
  }
  KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  if (tc_shape_ok(m, true)) KHG_TRY(tc_pack_build_f16(m));
  t.ready = t.tf32_ready || t.f16_ready;
  return KHG_OK;
}

void tc_pdf_tile_range(const khg_model *m, int p, int *ja, int *jb) {
  const TcPack &t = m->tc;
  const int v0 = t.h_vfirst[p], v1 = t.h_vfirst[p + 1] - 1;  // virtual pdfs of p; tile j owns virtual pdfs [h_tile_p0[j], h_tile_p0[j + 1])
  *ja = (int)(std::upper_bound(t.h_tile_p0.begin(), t.h_tile_p0.end(), v0) - t.h_tile_p0.begin()) - 1;
  *jb = (int)(std::upper_bound(t.h_tile_p0.begin(), t.h_tile_p0.end(), v1) - t.h_tile_p0.begin()) - 1;
}
int tc_num_tiles(const khg_model *m) { return m->tc.n_tiles; }

template <bool F16>
static khg_status tc_launch(khg_model *m, const float *d_feats, int64_t T, float scale, float *d_out, int64_t ld_out,
                            const unsigned *gate, int gate_run_if_above, const TileSubset *subset = nullptr) {
  TcPack &t = m->tc;
  TcArgs a;
  a.feats = d_feats;
  a.T = T;
  a.D = m->dim;
  a.K8 = F16 ? t.K16 : t.K8;
  a.Kc = (2 * m->dim + Elem<F16>::kUmmaK - 1) / Elem<F16>::kUmmaK * Elem<F16>::kUmmaK;
  a.n_chunks = (F16 ? t.KP16 : t.KP) / Elem<F16>::kChunkK;
  a.tab = F16 ? t.tab16 : t.tab8;
  a.ascale = F16 ? t.ascale : nullptr;
  a.gate = gate;
  a.gate_limit = kF16FeatLimit;
  a.gate_run_if_above = gate_run_if_above;
  const int a_bytes = 2 * a.n_chunks * kAChunkBytes;
  // two A buffers when they leave room for >= 3 operand stages (3, 4 and 5 stages run at the same
  // rate, profiles/r1u_*): the A build of the next item then overlaps the MMAs of the current one
  a.a_slots = (225 * 1024 - 2 * a_bytes) / kBStageBytes >= 3 ? 2 : 1;
  if (const char *e = getenv("KHG_TC_A_SLOTS")) a.a_slots = std::max(1, std::min(a.a_slots, atoi(e)));  // experiments
  a.stages = std::min(F16 ? 6 : 4, (int)((225 * 1024 - a.a_slots * a_bytes) / kBStageBytes));
  if (const char *e = getenv("KHG_TC_STAGES")) a.stages = std::max(2, std::min(a.stages, atoi(e)));  // experiments: ring depth
  if (a.stages < 2) {
    set_error("feature dimension too large for the tcgen05 kernel");
    return KHG_ERR_UNSUPPORTED;
  }
  a.offsets = m->d_offsets;
  a.tile_g0 = t.tile_g0;
  a.tile_p0 = t.tile_p0;
  a.epi_hdr = static_cast<const int2 *>(t.epi_hdr);
  a.two_chunk_segs = t.two_chunk_segs ? 1 : 0;
  a.runs = t.runs;
  a.seg = t.seg;
  a.n_tiles = t.n_tiles;
  const int64_t n_m = (T + kTileM - 1) / kTileM;
  // Split the N range when there are too few frame tiles to fill the SMs; keep >= 8
  // tiles per item so that building A (once per item) stays amortised.
  int64_t splits = std::max<int64_t>(1, (2LL * m->sm_count + n_m - 1) / n_m);
  splits = std::min<int64_t>(splits, std::max(1, t.n_tiles / 8));
  a.tiles_per_split = (int)((t.n_tiles + splits - 1) / splits);
  a.n_splits = (t.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  a.n_items = n_m * a.n_splits;
  const bool sub = subset && subset->off && a.n_splits == 1;
  a.sub_off = sub ? subset->off : nullptr;
  a.sub_tiles = sub ? subset->tiles : nullptr;
  a.sub_shift = sub ? subset->shift : 1;
  a.scale = scale;
  a.out = d_out;
  a.ld = ld_out;
  a.err = m->d_err;
  a.scratch = reinterpret_cast<float *>(m->d_scratch_int + 3);
  if (ld_out * 4 >= ((int64_t)1 << 32)) {
    set_error("ld_out too large for the tensor-core kernel (needs ld_out < 2^30)");
    return KHG_ERR_UNSUPPORTED;
  }
  a.debug_mode = 0;
#ifdef KHG_EXPERIMENTS
  if (const char *dbg = getenv("KHG_TC_DEBUG_MODE")) a.debug_mode = atoi(dbg);
#endif
  const size_t smem = (size_t)a.a_slots * a_bytes + (size_t)a.stages * kBStageBytes + 256 + 1024;
  auto kern = t.two_chunk_segs ? (t.grouped_segs ? loglikes_tc_kernel<F16, true, true> : loglikes_tc_kernel<F16, true, false>)
                               : (t.grouped_segs ? loglikes_tc_kernel<F16, false, true> : loglikes_tc_kernel<F16, false, false>);
  // (an attribute of the function on the CURRENT device: set on every launch, so that khg_set_device works)
  KHG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  unsigned grid = (unsigned)std::min<int64_t>(a.n_items, m->sm_count);
  if (const char *mc = getenv("KHG_TC_MAX_CTAS")) grid = std::min<unsigned>(grid, (unsigned)std::max(1, atoi(mc)));  // experiments
  // CTA pairs (clusters of 2) share the streamed operand through TMA multicast when every CTA has
  // whole frame tiles to itself and there are enough of them (KHG_TC_CLUSTER=0 forces the plain form)
  const char *cl = getenv("KHG_TC_CLUSTER");
  a.cluster = (a.n_splits == 1 && n_m >= 2LL * m->sm_count && grid >= 2 && !(cl && atoi(cl) == 0) && !(sub && a.sub_shift == 0)) ? 2 : 1;
#ifdef KHG_EXPERIMENTS
  if (a.debug_mode == 6 || a.debug_mode == 7) a.cluster = 1;
#endif
  if (a.cluster >= 2) {
    grid &= ~1u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, F16 ? t.hmap_hi : t.map_hi, F16 ? t.hmap_lo : t.map_lo, a) != cudaSuccess) {
      (void)cudaGetLastError();  // pairs cannot be scheduled here (e.g. a partitioned GPU): the plain form
      a.cluster = 1;
      kern<<<grid, kTcThreads, smem, m->stream>>>(F16 ? t.hmap_hi : t.map_hi, F16 ? t.hmap_lo : t.map_lo, a);
    }
  } else {
    kern<<<grid, kTcThreads, smem, m->stream>>>(F16 ? t.hmap_hi : t.map_hi, F16 ? t.hmap_lo : t.map_lo, a);
  }
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  return KHG_OK;
}

// precision: 0 = automatic (fp16 split when the model fits and, decided on the device per
// call, the features fit; tf32 split otherwise), 1 = force the tf32 split, 2 = force fp16,
// 3 = force fp16 in the Gaussian-stationary form.
khg_status tc_loglikes(khg_model *m, const float *d_feats, int64_t T, float scale, float *d_out, int64_t ld_out,
                       int precision, const unsigned **simt_gate, float *gate_limit, const TileSubset *subset) {
  TcPack &t = m->tc;
  *simt_gate = nullptr;
  *gate_limit = kF16FeatLimit;
  if (!t.ready) {
    set_error("tcgen05 model pack not built");
    return KHG_ERR_UNSUPPORTED;
  }
  if (t.dead_pdf) {
    latch_error_kernel<<<1, 1, 0, m->stream>>>(m->d_err, ERR_NONFINITE);
    ++g_launch_count;
  }
  if (precision >= 2 && !t.f16_ready) {
    set_error("the fp16-split tensor-core path does not fit this model (parameter range or -inf gconsts)");
    return KHG_ERR_UNSUPPORTED;
  }
  if (precision == 1 && !t.tf32_ready) {
    set_error("the tf32-split tensor-core path does not fit this model (2*dim+2 > 160)");
    return KHG_ERR_UNSUPPORTED;
  }
  if (precision == 1 || !t.f16_ready) return tc_launch<false>(m, d_feats, T, scale, d_out, ld_out, nullptr, 0, subset);
  if (precision == 3) {  // the Gaussian-stationary form of the fp16 split (khg_loglikes_gs.cu), forced
    if (!gs_supported(m)) {
      set_error("the Gaussian-stationary fp16 kernel does not fit this model (fp16 range, or 2*dim+2 > 128)");
      return KHG_ERR_UNSUPPORTED;
    }
    return gs_loglikes(m, d_feats, T, scale, d_out, ld_out);
  }
  if (precision == 2) return tc_launch<true>(m, d_feats, T, scale, d_out, ld_out, nullptr, 0, subset);
  // automatic: one pass over the features finds max |x * 2^-k|; both kernels are launched and
  // exactly one of them runs, chosen on the device (no host round trip, stays asynchronous)
  KHG_CUDA_TRY(cudaMemsetAsync(t.gate, 0, sizeof(unsigned), m->stream));
  const int64_t n = T * m->dim;
  feat_absmax_kernel<<<(unsigned)std::min<int64_t>(4 * m->sm_count, (n + 255) / 256), 256, 0, m->stream>>>(
      d_feats, n, m->dim, t.ascale, t.gate);
  ++g_launch_count;
  KHG_TRY(tc_launch<true>(m, d_feats, T, scale, d_out, ld_out, t.gate, 0, subset));
  if (t.tf32_ready) return tc_launch<false>(m, d_feats, T, scale, d_out, ld_out, t.gate, 1, subset);
  *simt_gate = t.gate;  // no tf32 operands for this shape: the caller launches the gated SIMT kernel
  return KHG_OK;
}

}  // namespace khg
