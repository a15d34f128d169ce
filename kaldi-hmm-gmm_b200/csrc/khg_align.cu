// kaldi-hmm-gmm_b200/csrc/khg_align.cu — khg_align_batch: gmm-align-compiled for a batch of
// utterances on the device (BASELINE config C5).
//
// Reference behaviour restated (paths relative to kaldi-hmm-gmm/csrc/ of the reference):
//   AlignUtteranceWrapper      decoder-wrappers.cc:16-108  (beam, retry beam, like)
//   FasterDecoder::Decode      faster-decoder.cc:41-152    (InitDecoding, ProcessNonemitting)
//   ::ProcessEmitting          faster-decoder.cc:154-228
//   ::GetCutoff                faster-decoder.cc:230-320   (beam / min_active, beam_delta)
//   ::ReachedFinal/GetBestPath faster-decoder.cc:346-425
//   DecodableAmDiagGmmScaled   decodable-am-diag-gmm.h:94-98 (scale * loglike(frame, tid2pdf[tid]))
//
// Shape of the computation: the dense kernel (K1) writes the scaled all-pdf block of a chunk of
// utterances; `viterbi_kernel` then runs ONE CTA per utterance.  The reference's hash of tokens
// becomes one cost per graph state in shared memory (double, like Token::cost_) and the token
// passing becomes a pull over each state's incoming arcs (the host transposes every graph
// once), so there are no atomics and ties are broken by the lowest arc id.  Back-pointers
// (one int32 arc id per (frame, state)) go to HBM, coalesced; thread 0 walks them back.
//
// Exactness.  The reference prunes new tokens against a RUNNING next_weight_cutoff while it walks its
// token list (faster-decoder.cc:196-216), so for one frame it can keep tokens above the frame's final
// cutoff — which ones depends on the list order (hash-list-inl.h).  The device search applies the final
// cutoff (order-independent) and, per frame, checks a certificate that those extra tokens cannot exist
// or cannot matter:
//   * a token can only be "extra" if its cost lies in [final cutoff, c0), c0 = the bound the reference
//     gets from expanding its best token first (:177-190); x_min = the cheapest such candidate;
//   * extras are never cheaper than any kept token, so with more than min_active kept tokens GetCutoff
//     (:230-320) returns the same cutoffs; they are expanded on the next frame only if x_min < that
//     frame's weight_cutoff; with <= min_active kept tokens they could change the token count or be
//     expanded (the cutoff is +inf then), and on the last frame they could be the best final token.
// An utterance for which any of these cannot be ruled out — or that has an exact cost tie between
// competing predecessors / final states (the reference then keeps whichever its list order met first), or
// a negative epsilon weight — is FLAGGED and re-aligned by the exact host restatement of the reference
// (khg_align_exact.cu) on the same likelihood block.  Everything else is provably the reference's result.
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include <math_constants.h>

#include "khg_host_pool.h"
#include "khg_internal.h"

namespace khg {

// one utterance of the batch, as the kernels see it
struct UttDesc {
  int64_t frame0;  // first row of feats / first entry of alignment
  int64_t bp0;     // first back-pointer of the utterance inside the chunk's buffer
  int32_t T, S, state0, start;
  int32_t ed0, n_ed;    // states with incoming epsilon arcs: ed_state[ed0 .. ed0+n_ed)
  int32_t pdf0, n_pdf;  // distinct pdfs on the graph: utt_pdfs[pdf0 .. pdf0+n_pdf)
  int32_t col0;         // first column of the utterance in the chunk's likelihood block
  int32_t pad[3];
};
static_assert(sizeof(UttDesc) == 64, "UttDesc layout");

struct AlignDev {
  const UttDesc *utts;
  const int32_t *order;     // chunk-local launch order (longest first)
  const int32_t *in_off;    // total_states+1, absolute into in_arcs
  const int4 *in_arcs;      // {src (local), local pdf, arc id (absolute), weight bits}
  const int32_t *ed_state;  // local state ids
  const int32_t *ed_off;    // n_ed_total+1, absolute into e_arcs
  const int4 *e_arcs;       // {src (local), arc id (absolute), weight bits, 0}
  const float *final_cost;  // total_states
  const int32_t *arc_src;   // per arc: local source state
  const int32_t *arc_il;    // per arc: ilabel
  const float *arc_w;       // per arc: weight
  const int32_t *utt_pdfs;
  const int32_t *tid2pdf;
};

struct UttOut {
  int32_t status, best_state, path_len, pad;
  float like;
  int32_t pad2[3];
};

int g_align_prep_hit = 0;                  // the last call reused the previous call's graph preparation
double g_align_tile_fraction = 1.0;        // (tile, frame tile) units the last call's dense kernel computed / all of them
int64_t g_align_exact_utts = 0;           // utterances of the last khg_align_batch call that took the exact host pass
constexpr int kAlignMinActive = 20;       // faster-decoder.h:42
constexpr float kAlignBeamDelta = 0.5f;   // faster-decoder.h:43
constexpr int kMaxWarps = 16;

struct RedScratch {
  double v[2][kMaxWarps];
  double v2[2][kMaxWarps];
  int n[2][kMaxWarps];
  int s[2][kMaxWarps];
};

// Order-preserving map double -> uint64 (so that a min over costs is two 32-bit redux.sync).
__device__ __forceinline__ unsigned long long cost_key(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_cost(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
// (min, min, sum) over the CTA; one __syncthreads per call (two alternating slots).  Warp stage: the
// minimum of the 64-bit keys as redux.sync.min over the high words, then over the low words of the
// lanes that hold the minimal high word; the count as redux.sync.add.
__device__ __forceinline__ double warp_min_f64(double v) {
  const unsigned long long k = cost_key(v);
  const unsigned hi = (unsigned)(k >> 32), mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? (unsigned)k : 0xffffffffu);
  return key_cost(((unsigned long long)mhi << 32) | mlo);
}
__device__ __forceinline__ void block_min2_sum(double &v, double &v2, int &n, RedScratch *rs, int &slot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const double wv = warp_min_f64(v), wv2 = warp_min_f64(v2);
  n = __reduce_add_sync(0xffffffffu, n);
  if (lane == 0) { rs->v[slot][warp] = wv; rs->v2[slot][warp] = wv2; rs->n[slot][warp] = n; }
  __syncthreads();
  double bv = rs->v[slot][0], bv2 = rs->v2[slot][0];
  int bn = rs->n[slot][0];
  for (int w = 1; w < nw; ++w) { bv = fmin(bv, rs->v[slot][w]); bv2 = fmin(bv2, rs->v2[slot][w]); bn += rs->n[slot][w]; }
  v = bv;
  v2 = bv2;
  n = bn;
  slot ^= 1;
}
__device__ __forceinline__ void block_min_sum(double &v, int &n, RedScratch *rs, int &slot) {
  double dummy = CUDART_INF;
  block_min2_sum(v, dummy, n, rs, slot);
}

// lexicographic min over (v, s): lowest state among equal costs
__device__ __forceinline__ void block_argmin(double &v, int &s, RedScratch *rs, int &slot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int os = __shfl_xor_sync(0xffffffffu, s, o);
    if (ov < v || (ov == v && os < s)) { v = ov; s = os; }
  }
  if (lane == 0) { rs->v[slot][warp] = v; rs->s[slot][warp] = s; }
  __syncthreads();
  double bv = rs->v[slot][0];
  int bs = rs->s[slot][0];
  for (int w = 1; w < nw; ++w) {
    double ov = rs->v[slot][w];
    int os = rs->s[slot][w];
    if (ov < bv || (ov == bv && os < bs)) { bv = ov; bs = os; }
  }
  v = bv;
  s = bs;
  slot ^= 1;
}

// ProcessNonemitting (faster-decoder.cc:57-123) as a Jacobi relaxation over the states that have
// incoming epsilon arcs: X <- min(X, min over eps arcs (X[src] + w)), a source or a result
// above `cutoff` does not propagate.  Converges to the same least fixed point as the
// reference's work queue; on exact ties the earlier token (then the lowest arc id) stays.
__device__ __forceinline__ void eps_closure(double *X, double *Y, const AlignDev &g, const UttDesc &u, double cutoff,
                                            int32_t *bp_row, int *s_flag, bool &tie) {
  if (u.n_ed == 0) return;
  for (;;) {
    __syncthreads();  // X complete; previous round's flag consumed
    if (threadIdx.x == 0) *s_flag = 0;
    __syncthreads();
    bool changed = false;
    for (int i = threadIdx.x; i < u.n_ed; i += blockDim.x) {
      const int d = g.ed_state[u.ed0 + i];
      double bc = X[d];
      int ba = -1;
      const int k1 = g.ed_off[u.ed0 + i + 1];
      for (int k = g.ed_off[u.ed0 + i]; k < k1; ++k) {
        const int4 e = g.e_arcs[k];
        const double cs = X[e.x];
        if (!(cs > cutoff)) {
          const double nc = cs + (double)__int_as_float(e.z);
          if (!(nc > cutoff)) {
            if (nc < bc) { bc = nc; ba = e.y; }
            else if (nc == bc) {
              // two epsilon arrivals at the same cost: the reference's queue order decides which one stays
              // (an emitting token of the same cost always stays, in both: faster-decoder.cc:105-115)
              const int prev = ba >= 0 ? ba : bp_row[d];
              if (prev >= 0 && prev != e.y && g.arc_il[prev] == 0) tie = true;
            }
          }
        }
      }
      Y[d] = bc;
      if (ba >= 0) { bp_row[d] = ba; changed = true; }
    }
    if (changed) *s_flag = 1;
    __syncthreads();
    const bool any = *s_flag != 0;
    if (!any) break;
    for (int i = threadIdx.x; i < u.n_ed; i += blockDim.x) {
      const int d = g.ed_state[u.ed0 + i];
      X[d] = Y[d];
    }
  }
  __syncthreads();
}

// One CTA per utterance.  Dynamic shared memory: 3 x S_cap doubles (costs) + n_pdf_cap x FC
// floats (the utterance's slice of the likelihood block for FC frames); FC == 0 reads the
// block directly, use_gcost keeps the costs in global scratch (very large graphs).
__global__ void __launch_bounds__(512) viterbi_kernel(AlignDev g, int u_base, const float *__restrict__ ll, int64_t ld,
                                                      float beam, float retry_beam, int S_cap, int n_pdf_cap, int FC,
                                                      double *gcost, int32_t *__restrict__ bp, int32_t *__restrict__ alignment,
                                                      int32_t *__restrict__ pdf_ids, UttOut *__restrict__ outs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ RedScratch rs;
  __shared__ int s_flag, s_cnt;
  __shared__ float s_mac;

  const int ui = u_base + g.order[blockIdx.x];
  const UttDesc u = g.utts[ui];
  const int S = u.S, T = u.T, tid = threadIdx.x, NT = blockDim.x;
  double *cbase = gcost ? gcost + (size_t)blockIdx.x * 3 * S_cap : reinterpret_cast<double *>(smem_raw);
  double *cur = cbase, *nxt = cbase + S_cap, *alt = cbase + 2 * S_cap;
  float *tile = reinterpret_cast<float *>(smem_raw + (gcost ? 0 : sizeof(double) * 3 * (size_t)S_cap));
  int32_t *ubp = bp + u.bp0;
  const int32_t *in_off = g.in_off + u.state0;
  const float *fin = g.final_cost + u.state0;
  const int32_t *upd = g.utt_pdfs + u.pdf0;
  const double kInf = CUDART_INF;
  int slot = 0;
  UttOut res;
  res.status = KHG_ALIGN_FAILED;
  res.best_state = -1;
  res.path_len = 0;
  res.like = 0.f;
  bool flagged = false;  // block-uniform: the reference's order-dependent pruning could have mattered
  bool tie = false;      // per thread: an exact cost tie between competing predecessors

  if (u.start < 0 || S <= 0) {  // empty graph: decoder-wrappers.cc:36-42
    if (tid == 0) outs[ui] = res;
    return;
  }

  for (int attempt = 0; attempt < 2; ++attempt) {
    const float cfg_beam = attempt == 0 ? beam : retry_beam;
    if (attempt == 1 && retry_beam == 0.f) break;
    // ---- InitDecoding: start token, then the epsilon closure with cutoff FLT_MAX
    for (int s = tid; s < S; s += NT) {
      cur[s] = s == u.start ? 0.0 : kInf;
      ubp[s] = -1;
    }
    eps_closure(cur, alt, g, u, (double)FLT_MAX, ubp, &s_flag, tie);
    __syncthreads();
    bool dead = false;
    double x_local = kInf;  // cheapest candidate of the previous frame that the reference may have kept above the cutoff
    for (int t = 0; t < T; ++t) {
      const int tf = FC ? (t & (FC - 1)) : 0;  // FC is 32, 16, 8 or 0
      if (FC && tf == 0) {  // stage the next FC frames of this utterance's pdfs (ordered by the sync in the reduction)
        const int nf = min(FC, T - t);
        const int fsh = 31 - __clz(FC);
        // (eight independent loads in flight per thread: one load -> store per iteration paid the latency of the
        // pdf-id load and of the likelihood load n_pdf * FC / NT times in a row, 9 % of the kernel's stall samples)
        const int n_el = u.n_pdf * FC;
        for (int i0 = tid; i0 < n_el; i0 += 8 * NT) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * NT;
            const int j = i >> fsh, f = i & (FC - 1);
            v[q] = (i < n_el && f < nf) ? ll[(int64_t)__ldg(upd + j) * ld + u.col0 + t + f] : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int i = i0 + q * NT;
            if (i < n_el) tile[i] = v[q];
          }
        }
      }
      // ---- GetCutoff
      double best = kInf;
      int ntok = 0;
      for (int s = tid; s < S; s += NT) {
        const double c = cur[s];
        if (c < kInf) { best = fmin(best, c); ++ntok; }
      }
      double x_min = x_local;
      block_min2_sum(best, x_min, ntok, &rs, slot);
      x_local = kInf;
      // extras of the previous frame exist at best at cost x_min: with <= min_active kept tokens they can change
      // the token count GetCutoff sees, or be expanded under its +inf cutoff
      if (x_min < kInf && ntok <= kAlignMinActive) flagged = true;
      if (ntok == 0) { dead = true; break; }
      const double beam_cutoff = best + (double)cfg_beam;
      // with at most min_active tokens the reference's min_active_cutoff stays +inf, which is
      // "looser than the beam": nothing is pruned on this frame (faster-decoder.cc:283-310)
      double weight_cutoff = kInf;
      float adaptive_beam = CUDART_INF_F;
      if (ntok > kAlignMinActive) {
        weight_cutoff = beam_cutoff;
        adaptive_beam = cfg_beam;
        // min_active_cutoff = tmp_array_[min_active] (costs rounded to float) exceeds the beam
        // cutoff iff at most min_active tokens are at or below it
        int c_le = 0;
        double dummy = 0.0;
        for (int s = tid; s < S; s += NT) {
          const double c = cur[s];
          if (c < kInf && (double)(float)c <= beam_cutoff) ++c_le;
        }
        block_min_sum(dummy, c_le, &rs, slot);
        if (c_le <= kAlignMinActive) {
          float *list = reinterpret_cast<float *>(alt);
          if (tid == 0) s_cnt = 0;
          __syncthreads();
          for (int s = tid; s < S; s += NT) {
            const double c = cur[s];
            if (c < kInf) list[atomicAdd(&s_cnt, 1)] = (float)c;
          }
          __syncthreads();
          const int n = s_cnt;
          for (int i = tid; i < n; i += NT) {
            const float v = list[i];
            int r = 0;
            for (int j = 0; j < n; ++j) {
              const float w = list[j];
              r += (w < v || (w == v && j < i)) ? 1 : 0;
            }
            if (r == kAlignMinActive) s_mac = v;
          }
          __syncthreads();
          const double mac = (double)s_mac;
          if (mac > beam_cutoff) {
            weight_cutoff = mac;
            adaptive_beam = (float)(mac - best + (double)kAlignBeamDelta);
          }
        }
      }
      // kept extras would be expanded on this frame iff they are below its cutoff
      if (x_min < weight_cutoff) flagged = true;
      // ---- ProcessEmitting as a pull over incoming arcs
      int32_t *bp_row = ubp + (size_t)(t + 1) * S;
      double lmin = kInf, lmin0 = kInf;  // lowest new cost; lowest new cost out of the best token (the reference's first bound)
      int n_best = 0;
      for (int d = tid; d < S; d += NT) {
        double bc = kInf;
        int ba = -1, bsrc = -1;
        n_best += cur[d] == best ? 1 : 0;
        const int k1 = in_off[d + 1];
        for (int k = in_off[d]; k < k1; ++k) {
          const int4 e = g.in_arcs[k];
          const double cs = cur[e.x];
          if (cs < weight_cutoff) {
            const float lk = FC ? tile[e.y * FC + tf] : ll[(int64_t)upd[e.y] * ld + u.col0 + t];
            const float ac = -1.f * lk;
            const double nw = ((double)__int_as_float(e.w) + cs) + (double)ac;
            if (cs == best) lmin0 = fmin(lmin0, nw);
            if (nw < bc) { bc = nw; ba = e.z; bsrc = e.x; }
            else if (nw == bc && e.x != bsrc) tie = true;  // equal cost from two source tokens: list order decides in the reference
          }
        }
        nxt[d] = bc;
        bp_row[d] = ba;
        lmin = fmin(lmin, bc);
      }
      block_min2_sum(lmin, lmin0, n_best, &rs, slot);
      const double next_cutoff = lmin + (double)adaptive_beam;
      // the reference's bound after expanding its best token first (faster-decoder.cc:177-190); with several
      // equally good best tokens it is whichever comes first in its list: unknown here, so no bound
      const double c0 = n_best > 1 ? kInf : lmin0 + (double)adaptive_beam;
      for (int d = tid; d < S; d += NT) {
        const double c = nxt[d];
        if (!(c < next_cutoff)) {
          if (c < c0) x_local = fmin(x_local, c);  // the running cutoff may have let this one in
          nxt[d] = kInf;
          bp_row[d] = -1;
        }
      }
      eps_closure(nxt, alt, g, u, next_cutoff, bp_row, &s_flag, tie);
      double *tmp = cur; cur = nxt; nxt = tmp;
    }
    __syncthreads();
    // ---- ReachedFinal + the best final token (faster-decoder.cc:346-388)
    double bf = kInf;
    int bs = 0x7fffffff;
    if (!dead) {
      for (int s = tid; s < S; s += NT) {
        const double c = cur[s];
        const float f = fin[s];
        if (c != kInf && f != CUDART_INF_F) {
          const double tc = c + (double)f;
          if (tc != kInf && (tc < bf || (tc == bf && s < bs))) { bf = tc; bs = s; }
        }
      }
    }
    block_argmin(bf, bs, &rs, slot);
    {
      // extras of the LAST frame could be final tokens; two final states at exactly the best total cost are
      // resolved by list order in the reference (faster-decoder.cc:377-384)
      double xl = x_local, dummy = kInf;
      int n_tie = 0;
      if (!dead && bs != 0x7fffffff)
        for (int s = tid; s < S; s += NT)
          if (cur[s] != kInf && fin[s] != CUDART_INF_F && cur[s] + (double)fin[s] == bf) ++n_tie;
      block_min2_sum(xl, dummy, n_tie, &rs, slot);
      if (xl < kInf || n_tie > 1) flagged = true;
    }
    if (bs != 0x7fffffff) {
      res.status = attempt == 0 ? KHG_ALIGN_OK : KHG_ALIGN_RETRIED;
      res.best_state = bs;
      break;
    }
    __syncthreads();
  }

  res.pad = (flagged || __syncthreads_or(tie ? 1 : 0)) ? 1 : 0;
  if (tid == 0) {
    if (res.status != KHG_ALIGN_FAILED) {
      // traceback: GetBestPath + GetLinearSymbolSequence; like = -(graph + acoustic) / scale is
      // finished on the host (needs the scale)
      int t = T, s = res.best_state, n = 0;
      double graph = (double)fin[s], ac = 0.0;
      for (;;) {
        const int a = ubp[(size_t)t * S + s];
        if (a < 0) break;
        ++n;
        const int il = g.arc_il[a];
        graph += (double)g.arc_w[a];
        if (il != 0) {
          --t;
          alignment[u.frame0 + t] = il;
          const int pdf = g.tid2pdf[il];
          if (pdf_ids) pdf_ids[u.frame0 + t] = pdf;
          ac -= (double)ll[(int64_t)pdf * ld + u.col0 + t];
        }
        s = g.arc_src[a];
      }
      if (t != 0 || s != u.start) res.status = KHG_ALIGN_FAILED;  // cannot happen (KHG_ASSERT in the reference)
      res.path_len = n;
      res.like = (float)(graph + ac);
    }
    outs[ui] = res;
  }
}

// Second walk over the back-pointers: the arcs of the best path in forward order.
__global__ void path_kernel(AlignDev g, int u_base, int n, const int32_t *__restrict__ bp, const UttOut *__restrict__ outs,
                            const int64_t *__restrict__ path_off, int64_t off0, int32_t *__restrict__ path) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const UttDesc u = g.utts[u_base + i];
  const UttOut o = outs[u_base + i];
  if (o.status == KHG_ALIGN_FAILED || path_off[u_base + i] < 0) return;  // (< 0: re-aligned on the host)
  int32_t *dst = path + (path_off[u_base + i] - off0);
  const int32_t *ubp = bp + u.bp0;
  int t = u.T, s = o.best_state, k = o.path_len;
  while (k > 0) {
    const int a = ubp[(size_t)t * u.S + s];
    if (a < 0) break;
    dst[--k] = a;
    if (g.arc_il[a] != 0) --t;
    s = g.arc_src[a];
  }
}

// The likelihood rows of flagged utterances, compacted for the host: entry i = (utterance, first float of its
// n_pdf x T block in dst).
__global__ void gather_ll_kernel(AlignDev g, const int2 *__restrict__ list, const int64_t *__restrict__ dst_off,
                                 const float *__restrict__ ll, int64_t ld, float *__restrict__ dst) {
  const UttDesc u = g.utts[list[blockIdx.x].x];
  const int32_t *upd = g.utt_pdfs + u.pdf0;
  float *o = dst + dst_off[blockIdx.x];
  for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < (int64_t)u.n_pdf * u.T; i += (int64_t)gridDim.y * blockDim.x) {
    const int j = (int)(i / u.T), t = (int)(i - (int64_t)j * u.T);
    o[i] = ll[(int64_t)upd[j] * ld + u.col0 + t];
  }
}



// What the host preparation of khg_align_batch produces and later phases need, kept on the model between calls: the
// realignment passes of an EM recipe align the SAME graphs again and again (egs/yesno/train.py:165-206), and the
// preparation — transposing 2000 graphs — is the longest host phase of a call.  A call whose graphs, frame offsets,
// tid2pdf and chunking hash to the key of the previous call reuses the device copy of the graphs (w_al_graph) and
// these host tables.  KHG_ALIGN_PREP_CACHE=0 disables it.
struct AlignPrepCache {
  bool valid = false;
  uint64_t key[2] = {0, 0};
  std::vector<UttDesc> desc;
  std::vector<int32_t> neg_eps, arc_lp;
  std::vector<std::vector<int32_t>> updf;
  int S_max = 1, n_pdf_max = 1;
  size_t n_in = 0, n_eds = 0, n_edo = 0, eps_total = 0, n_updf = 0;
  // tile lists of the dense launches (they depend on the graphs' pdfs, the frame offsets and the pack only): the device
  // copies in w_al_tiles stay valid for an identical batch
  struct TileLists { int u0, u1, shift; size_t first_word, n_off, n_list; const void *base; };
  std::vector<TileLists> lists;
};
void align_cache_free(khg_model *m) {
  delete static_cast<AlignPrepCache *>(m->al_cache);
  m->al_cache = nullptr;
}
// 2 x 64-bit content hash of a byte range (multiply-xorshift over 8-byte words, blocks hashed by the pool's workers)
static void hash_bytes(const void *ptr, size_t bytes, uint64_t (&key)[2]) {
  if (!ptr || !bytes) return;
  const size_t kBlock = 1 << 16, nb = (bytes + kBlock - 1) / kBlock;
  std::vector<uint64_t> h0(nb), h1(nb);
  parallel_for((int)nb, [&](int b, int) {
    const unsigned char *q = static_cast<const unsigned char *>(ptr) + (size_t)b * kBlock;
    const size_t n = std::min(kBlock, bytes - (size_t)b * kBlock);
    uint64_t a = 0x9E3779B97F4A7C15ull ^ (uint64_t)b, c = 0xC2B2AE3D27D4EB4Full + (uint64_t)b;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
      uint64_t w;
      memcpy(&w, q + i, 8);
      a = (a ^ w) * 0xFF51AFD7ED558CCDull;
      a ^= a >> 32;
      c = (c + w) * 0x9FB21C651E98DF25ull;
      c ^= c >> 29;
    }
    uint64_t w = 0;
    if (i < n) memcpy(&w, q + i, n - i);
    a = (a ^ w ^ (uint64_t)n) * 0xFF51AFD7ED558CCDull;
    c = (c + w + (uint64_t)n) * 0x9FB21C651E98DF25ull;
    h0[b] = a ^ (a >> 31);
    h1[b] = c ^ (c >> 33);
  }, 8);
  for (size_t b = 0; b < nb; ++b) {
    key[0] = (key[0] ^ h0[b]) * 0xD6E8FEB86659FD93ull + 0x2545F4914F6CDD1Dull;
    key[1] = (key[1] + h1[b]) * 0xA0761D6478BD642Full ^ (key[1] >> 27);
  }
}

}  // namespace khg

using namespace khg;

extern "C" khg_status khg_align_batch(khg_model *m, const khg_graph_batch *gb, const float *feats, int32_t feats_loc,
                                      const int32_t *tid2pdf, int32_t n_tids, float acoustic_scale, float beam,
                                      float retry_beam, int32_t *alignment, int32_t *utt_status, float *utt_like,
                                      int32_t *path_arcs, int64_t *path_offsets, int64_t path_capacity,
                                      int32_t *pdf_ids_dev) {
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  KHG_REQUIRE(gb && gb->n_utts >= 0, "null graph batch");
  // decoder-wrappers.cc:29-33
  if ((retry_beam != 0 && retry_beam <= beam) || beam <= 0.0f) {
    set_error("Beams do not make sense: beam " + std::to_string(beam) + ", retry-beam " + std::to_string(retry_beam));
    return KHG_ERR_INVALID;
  }
  KHG_REQUIRE(acoustic_scale != 0.f, "acoustic_scale must not be 0");
  const int U = gb->n_utts;
  if (path_offsets) path_offsets[0] = 0;
  if (U == 0) return KHG_OK;
  KHG_REQUIRE(gb->frame_offsets && gb->state_offsets && gb->arc_offsets && gb->start_state && tid2pdf && n_tids > 0,
              "null graph array");
  KHG_REQUIRE(!path_arcs || path_offsets, "path_arcs needs path_offsets");
  const int64_t T_all = gb->frame_offsets[U] - gb->frame_offsets[0];
  const int32_t S_all = gb->state_offsets[U];
  KHG_REQUIRE(gb->state_offsets[0] == 0 && gb->arc_offsets[0] == 0, "offsets must start at 0");
  const int32_t A_all = gb->arc_offsets[S_all];
  KHG_REQUIRE(A_all == 0 || (gb->arc_ilabel && gb->arc_nextstate && gb->arc_weight), "null arc array");
  KHG_REQUIRE(S_all == 0 || gb->final_cost, "null final_cost");
  KHG_REQUIRE(T_all == 0 || feats, "null feats");
  for (int32_t i = 0; i < n_tids; ++i)
    KHG_REQUIRE(i == 0 || (tid2pdf[i] >= 0 && tid2pdf[i] < m->P), "tid2pdf entry out of range");
  const int D = m->dim, P = m->P;
  cudaStream_t st = m->stream;
  // KHG_ALIGN_TIMING=1: host-preparation / dense-kernel / search times of the call on stderr
  const bool timing = getenv("KHG_ALIGN_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms_dense = 0.f, ms_search = 0.f;
  if (timing)
    for (auto &e : ev) cudaEventCreate(&e);

  // ---------------- chunks of consecutive utterances: bounded likelihood block and back-pointers
  // (a quarter of the free HBM each, at most 24 GB: more utterances per launch = more CTAs per SM
  // for the latency-bound search).  Needs only the frame and state counts, so the dense kernel of
  // the first chunk is launched BEFORE the host transposes the graphs and runs under that work.
  std::vector<int64_t> bp0_v(U), col0_v(U);
  std::vector<int> chunk_start(1, 0);
  size_t bp_bytes_max = 0;
  int64_t chunk_frames_max = 0;
  int max_chunk_utts = 0;
  {
    size_t mem_free = 0, mem_total = 0;
    KHG_CUDA_TRY(cudaMemGetInfo(&mem_free, &mem_total));
    mem_free += m->w_al_block.cap + m->w_al_bp.cap;  // what an earlier call already holds is reusable
    const int64_t budget = std::max<int64_t>(256LL << 20, std::min<int64_t>(24LL << 30, (int64_t)(mem_free / 4)));
    int64_t max_chunk_frames = std::max<int64_t>(1024, budget / (4LL * P));
    if (const char *e = getenv("KHG_ALIGN_CHUNK_FRAMES")) max_chunk_frames = std::max<int64_t>(1, atoll(e));  // tests: force several chunks
    int64_t fr = 0, bpb = 0;
    for (int u = 0; u < U; ++u) {
      const int64_t Tu = gb->frame_offsets[u + 1] - gb->frame_offsets[u];
      const int64_t Su = gb->state_offsets[u + 1] - gb->state_offsets[u];
      KHG_REQUIRE(Tu >= 0 && Su >= 0 && Tu < (1LL << 31), "graph of utterance " + std::to_string(u) + ": frame / state offsets out of range");
      const int64_t ub = 4LL * (Tu + 1) * Su;
      if (u > chunk_start.back() && (fr + Tu > max_chunk_frames || bpb + ub > budget)) {
        chunk_start.push_back(u);
        fr = 0;
        bpb = 0;
      }
      bp0_v[u] = bpb / 4;
      col0_v[u] = fr;
      fr += Tu;
      bpb += ub;
      bp_bytes_max = std::max(bp_bytes_max, (size_t)bpb);
      chunk_frames_max = std::max(chunk_frames_max, fr);
      max_chunk_utts = std::max(max_chunk_utts, u - chunk_start.back() + 1);
    }
    chunk_start.push_back(U);
  }
  const int64_t ld = (chunk_frames_max + 3) & ~(int64_t)3;
  KHG_TRY(m->w_al_block.reserve(sizeof(float) * (size_t)P * std::max<int64_t>(ld, 4)));
  float *d_block = m->w_al_block.as<float>();
  // Dense kernel of a chunk of utterances.  Only the model tiles (240-Gaussian blocks of pdfs) that hold a pdf of some
  // graph of the frames' utterances are computed — the batched form of what the reference's decodable does lazily
  // (DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased, csrc/decodable-am-diag-gmm.cc:29-71, evaluates a pdf when
  // the decoder asks for it): per pair of 128-frame tiles the union of the tiles of the utterances it overlaps
  // (KHG_ALIGN_TILE_SUBSET=0: everything).  The search reads the rows of its utterance's pdfs only.
  const bool want_subset = !(getenv("KHG_ALIGN_TILE_SUBSET") && atoi(getenv("KHG_ALIGN_TILE_SUBSET")) == 0) && m->tc.ready &&
                           m->kernel != KHG_KERNEL_SIMT;
  // ---- preparation cache (see AlignPrepCache): key over everything the preparation reads
  if (!m->al_cache) m->al_cache = new AlignPrepCache();
  AlignPrepCache &pc = *static_cast<AlignPrepCache *>(m->al_cache);
  bool prep_hit = false;
  {
    uint64_t key[2] = {0x1234567ull + (uint64_t)U * 1315423911ull + (uint64_t)P, 0x89abcdefull + (uint64_t)n_tids + ((uint64_t)want_subset << 40)};
    hash_bytes(gb->frame_offsets, 8 * ((size_t)U + 1), key);
    hash_bytes(gb->state_offsets, 4 * ((size_t)U + 1), key);
    hash_bytes(gb->arc_offsets, 4 * ((size_t)S_all + 1), key);
    hash_bytes(gb->arc_ilabel, 4 * (size_t)A_all, key);
    hash_bytes(gb->arc_nextstate, 4 * (size_t)A_all, key);
    hash_bytes(gb->arc_weight, 4 * (size_t)A_all, key);
    hash_bytes(gb->start_state, 4 * (size_t)U, key);
    hash_bytes(gb->final_cost, 4 * (size_t)S_all, key);
    hash_bytes(tid2pdf, 4 * (size_t)n_tids, key);
    hash_bytes(chunk_start.data(), sizeof(int) * chunk_start.size(), key);
    const char *e = getenv("KHG_ALIGN_PREP_CACHE");
    prep_hit = !(e && atoi(e) == 0) && pc.valid && pc.key[0] == key[0] && pc.key[1] == key[1] && (int)pc.desc.size() == U;
    pc.valid = false;  // (set again once this call's tables and device copy are complete)
    pc.key[0] = key[0];
    pc.key[1] = key[1];
  }
  g_align_prep_hit = prep_hit ? 1 : 0;
  const bool lists_cacheable = chunk_start.size() == 2;  // (the lists of a later chunk overwrite those of the previous one)
  if (!prep_hit || !lists_cacheable) pc.lists.clear();
  std::vector<std::vector<int32_t>> &updf = pc.updf;  // distinct pdfs of every graph (filled by the first host pass)
  if (!prep_hit) updf.assign(U, std::vector<int32_t>());
  int64_t units_done = 0, units_all = 0;
  auto stage_feats = [&](int u0, int u1, const float **d_f) -> khg_status {
    const int64_t f0 = gb->frame_offsets[u0] - gb->frame_offsets[0], nfr = gb->frame_offsets[u1] - gb->frame_offsets[u0];
    *d_f = feats + (gb->frame_offsets[0] + f0) * D;
    if (nfr > 0 && feats_loc == KHG_HOST) {
      KHG_TRY(m->w_feats.reserve(sizeof(float) * (size_t)nfr * D));
      KHG_TRY(h2d_copy(m, m->w_feats.p, *d_f, sizeof(float) * (size_t)nfr * D));  // (pageable features: staged by several threads)
      *d_f = m->w_feats.as<float>();
    }
    return KHG_OK;
  };
  size_t list_used = 0;                          // int32 words of w_al_tiles handed out so far
  auto run_dense = [&](int u0, int u1, const float *d_f, bool subset, int64_t col = 0) -> khg_status {
    const int64_t nfr = gb->frame_offsets[u1] - gb->frame_offsets[u0];
    if (nfr <= 0) return KHG_OK;
    float *d_dst = d_block + col;
    TileSubset sub;
    const int n_tiles = m->tc.ready ? tc_num_tiles(m) : 0;
    // one tile list per frame tile (128 frames; the kernel then runs without CTA pairs: -1 % of operand sharing, but a tile
    // overlaps fewer utterances than a pair of tiles: 16.8 % instead of 18.6 % of the tile units at C5), or per pair of
    // tiles (KHG_ALIGN_SUBSET_SHIFT=1)
    int sub_shift = 0;
    if (const char *e = getenv("KHG_ALIGN_SUBSET_SHIFT")) sub_shift = atoi(e) != 0 ? 1 : 0;
    const int64_t unit_frames = 128LL << sub_shift;
    const int64_t n_pairs = (nfr + unit_frames - 1) / unit_frames;  // (lists: per pair of tiles or per tile)
    if (subset && n_tiles > 1) {
      for (const AlignPrepCache::TileLists &L : pc.lists)  // identical batch: the lists of this launch are still on the device
        if (prep_hit && lists_cacheable && L.u0 == u0 && L.u1 == u1 && L.shift == sub_shift && L.first_word == list_used && L.base == m->w_al_tiles.p) {
          sub.off = m->w_al_tiles.as<int32_t>() + L.first_word;
          sub.tiles = sub.off + L.n_off;
          sub.shift = sub_shift;
          list_used += L.n_off + L.n_list;
          bool used = false;
          KHG_TRY(dense_block(m, d_f, nfr, acoustic_scale, KHG_PDF_MAJOR, d_dst, ld, &sub, &used));
          units_all += n_pairs * n_tiles;
          units_done += used ? (int64_t)L.n_list : n_pairs * n_tiles;
          return KHG_OK;
        }
      // tiles of every utterance (bitmap), then per frame-tile pair the union over the utterances it overlaps
      const int words = (n_tiles + 63) / 64;
      std::vector<uint64_t> ubits((size_t)(u1 - u0) * words, 0);
      parallel_for(u1 - u0, [&](int i, int) {
        uint64_t *b = ubits.data() + (size_t)i * words;
        for (int32_t pdf : updf[u0 + i]) {
          int ja, jb;
          tc_pdf_tile_range(m, pdf, &ja, &jb);
          for (int j = ja; j <= jb; ++j) b[j >> 6] |= 1ull << (j & 63);
        }
      });
      std::vector<int32_t> off((size_t)n_pairs + 1, 0), tiles;
      tiles.reserve((size_t)n_pairs * 64);
      std::vector<uint64_t> acc(words);
      int u = u0;
      const int64_t base = gb->frame_offsets[u0];
      for (int64_t q = 0; q < n_pairs; ++q) {
        const int64_t fa = base + q * unit_frames, fb = std::min<int64_t>(fa + unit_frames, base + nfr);
        while (u + 1 < u1 && gb->frame_offsets[u + 1] <= fa) ++u;
        std::fill(acc.begin(), acc.end(), 0);
        for (int v = u; v < u1 && gb->frame_offsets[v] < fb; ++v)
          if (gb->frame_offsets[v + 1] > fa)
            for (int w2 = 0; w2 < words; ++w2) acc[w2] |= ubits[(size_t)(v - u0) * words + w2];
        for (int w2 = 0; w2 < words; ++w2)
          for (uint64_t x = acc[w2]; x; x &= x - 1) tiles.push_back(w2 * 64 + __builtin_ctzll(x));
        off[q + 1] = (int32_t)tiles.size();
      }
      // (w_al_tiles was sized for the whole call up front: launches of one call must not move it)
      if (4 * (list_used + off.size() + tiles.size() + 2) > m->w_al_tiles.cap) {
        units_all += n_pairs * n_tiles;
        units_done += n_pairs * n_tiles;
        return dense_block(m, d_f, nfr, acoustic_scale, KHG_PDF_MAJOR, d_dst, ld);
      }
      int32_t *d_off = m->w_al_tiles.as<int32_t>() + list_used, *d_tiles = d_off + off.size();
      {  // what these words held before is gone (lists of another form, or of a buffer that has moved)
        const size_t w0 = list_used, w1 = list_used + off.size() + tiles.size();
        pc.lists.erase(std::remove_if(pc.lists.begin(), pc.lists.end(), [&](const AlignPrepCache::TileLists &L) {
          return L.base != m->w_al_tiles.p || (L.first_word < w1 && w0 < L.first_word + L.n_off + L.n_list);
        }), pc.lists.end());
      }
      if (lists_cacheable) pc.lists.push_back({u0, u1, sub_shift, list_used, off.size(), tiles.size(), m->w_al_tiles.p});
      list_used += off.size() + tiles.size();
      {  // through the pinned image: a copy from pageable memory would block the host until the dense kernel of the
         // previous group, running on this stream, has finished — and with it the host pass that should run under it
        int32_t *h = m->pin_al_tiles.as<int32_t>() + (d_off - m->w_al_tiles.as<int32_t>());
        std::memcpy(h, off.data(), 4 * off.size());
        std::memcpy(h + off.size(), tiles.data(), 4 * tiles.size());
        KHG_CUDA_TRY(cudaMemcpyAsync(d_off, h, 4 * (off.size() + tiles.size()), cudaMemcpyHostToDevice, st));
      }
      const int64_t n_list = (int64_t)tiles.size();
      sub.off = d_off;
      sub.tiles = d_tiles;
      sub.shift = sub_shift;
      bool used = false;
      KHG_TRY(dense_block(m, d_f, nfr, acoustic_scale, KHG_PDF_MAJOR, d_dst, ld, &sub, &used));
      units_all += n_pairs * n_tiles;
      units_done += used ? n_list : n_pairs * n_tiles;
      return KHG_OK;
    }
    units_all += n_pairs * std::max(1, n_tiles);
    units_done += n_pairs * std::max(1, n_tiles);
    return dense_block(m, d_f, nfr, acoustic_scale, KHG_PDF_MAJOR, d_dst, ld);
  };
  auto launch_dense = [&](int u0, int u1, bool subset) -> khg_status {
    list_used = 0;  // (the previous chunk's kernel has finished: the host synchronised on its search)
    const float *d_f = nullptr;
    KHG_TRY(stage_feats(u0, u1, &d_f));
    return run_dense(u0, u1, d_f, subset);
  };
  if (want_subset && m->tc.ready)  // worst case: every pair of frame tiles lists every model tile (+ the offsets, per launch)
  {
    KHG_TRY(m->w_al_tiles.reserve(4 * ((size_t)((chunk_frames_max + 127) / 128 + 16) * (size_t)(tc_num_tiles(m) + 1) + 64)));
    m->pin_al_tiles.pinned = true;
    KHG_TRY(m->pin_al_tiles.reserve(m->w_al_tiles.cap));
  }
  // with the subset the kernel of the first chunk needs the graphs' pdf lists (first host pass below); without it, it is
  // launched right away and runs under the whole host preparation
  const float *d_f0 = nullptr;
  if (timing) cudaEventRecord(ev[0], st);
  // host features of the first chunk in subset mode go to the device group by group on the copy stream (below): the
  // copy of a group runs under the dense kernel of the previous one
  const bool grouped_copy = want_subset && feats_loc == KHG_HOST && T_all > 0;
  if (grouped_copy) {
    const int64_t nfr0 = gb->frame_offsets[chunk_start[1]] - gb->frame_offsets[chunk_start[0]];
    KHG_TRY(m->w_feats.reserve(sizeof(float) * (size_t)std::max<int64_t>(nfr0, 1) * D));
    KHG_TRY(ensure_copy_stream(m));
    d_f0 = m->w_feats.as<float>();
  } else {
    KHG_TRY(stage_feats(chunk_start[0], chunk_start[1], &d_f0));
  }
  if (!want_subset) {
    KHG_TRY(run_dense(chunk_start[0], chunk_start[1], d_f0, false));
    if (timing) cudaEventRecord(ev[1], st);
  }

  double t_mark[6] = {0, 0, 0, 0, 0, 0};
  t_mark[0] = now();
  // ---------------- host: transpose every graph (incoming emitting arcs per state, incoming
  // epsilon arcs per state), local pdf lists
  std::vector<UttDesc> &desc = pc.desc;
  std::vector<int32_t> &neg_eps = pc.neg_eps, &arc_lp = pc.arc_lp;
  int &S_max = pc.S_max, &n_pdf_max = pc.n_pdf_max;
  std::vector<int32_t> n_emit, n_eps, bad, in_off, arc_src, eps_deg;
  if (!prep_hit) {
    desc.assign(U, UttDesc());
    neg_eps.assign(U, 0);
    arc_lp.assign(A_all, 0);
    n_emit.assign(U, 0);
    n_eps.assign(U, 0);
    bad.assign(U, 0);
    in_off.assign((size_t)S_all + 1, 0);
    arc_src.assign(A_all, 0);
    eps_deg.assign((size_t)S_all, 0);
  }
  const int n_workers = 16;
  std::vector<std::vector<int32_t>> stamp, lidx;  // per worker: which utterance saw a pdf last, and its local index there
  if (!prep_hit) {
    stamp.assign(n_workers, std::vector<int32_t>(P, -1));
    lidx.assign(n_workers, std::vector<int32_t>(P, 0));
  }
  auto first_pass = [&](int u, int w) {
    const int32_t s0 = gb->state_offsets[u], s1 = gb->state_offsets[u + 1], S = s1 - s0;
    UttDesc &d = desc[u];
    memset(&d, 0, sizeof(d));
    d.frame0 = gb->frame_offsets[u] - gb->frame_offsets[0];
    d.T = (int32_t)(gb->frame_offsets[u + 1] - gb->frame_offsets[u]);
    d.S = S;
    d.state0 = s0;
    d.start = gb->start_state[u];
    d.bp0 = bp0_v[u];
    d.col0 = (int32_t)col0_v[u];
    if (d.T < 0 || S < 0 || d.start >= S) { bad[u] = 1; return; }
    auto &st_ = stamp[w];
    auto &li = lidx[w];
    // (per-utterance counters stay local until the end: neighbouring utterances are other threads' work, and a counter
    // bumped per arc in a shared array bounces its cache line between them)
    int32_t ne = 0, nep = 0, neg = 0;
    std::vector<int32_t> &up_ = updf[u];
    up_.clear();
    up_.reserve((size_t)S);
    for (int32_t s = s0; s < s1; ++s) {
      if (gb->arc_offsets[s + 1] < gb->arc_offsets[s]) { bad[u] = 1; return; }
      for (int32_t a = gb->arc_offsets[s]; a < gb->arc_offsets[s + 1]; ++a) {
        const int32_t ns = gb->arc_nextstate[a], il = gb->arc_ilabel[a];
        if (ns < 0 || ns >= S || il < 0 || il >= n_tids) { bad[u] = 1; return; }
        arc_src[a] = s - s0;
        if (il == 0) {
          if (gb->arc_weight[a] < 0.f) neg = 1;  // the exactness certificate assumes epsilon costs >= 0
          ++nep;
          ++eps_deg[s0 + ns];
        } else {
          ++ne;
          ++in_off[(size_t)s0 + ns + 1];
          const int32_t pdf = tid2pdf[il];
          if (st_[pdf] != u) { st_[pdf] = u; li[pdf] = (int32_t)up_.size(); up_.push_back(pdf); }
          arc_lp[a] = li[pdf];
        }
      }
    }
    n_emit[u] = ne;
    n_eps[u] = nep;
    neg_eps[u] = neg;
    d.n_pdf = (int32_t)up_.size();
  };
  if (want_subset) {
    // the first chunk in groups of utterances: the dense kernel of a group is launched as soon as the group's pdf
    // lists exist and runs under the host pass of the next groups (a group keeps >= 2 frame tiles per SM, what the
    // kernel's subset mode needs)
    const int c0 = chunk_start[0], c1 = chunk_start[1];
    const int64_t fr0 = gb->frame_offsets[c0], frc = gb->frame_offsets[c1] - fr0;
    const int n_groups = (int)std::max<int64_t>(1, std::min<int64_t>(4, frc / (4LL * m->sm_count * 128)));
    int ga = c0;
    for (int gidx = 0; gidx < n_groups; ++gidx) {
      int gb_ = c1;
      if (gidx + 1 < n_groups) {
        gb_ = ga;
        while (gb_ < c1 && gb->frame_offsets[gb_] - fr0 < frc * (gidx + 1) / n_groups) ++gb_;
      }
      if (grouped_copy && gb_ > ga) {
        const int64_t f_a = gb->frame_offsets[ga] - fr0, n_g = gb->frame_offsets[gb_] - gb->frame_offsets[ga];
        if (gidx == 0) {  // w_feats may still be read by an earlier call's kernels on the model stream
          KHG_CUDA_TRY(cudaEventRecord(m->ev_done[0], st));
          KHG_CUDA_TRY(cudaStreamWaitEvent(m->copy_stream, m->ev_done[0], 0));
        }
        KHG_TRY(h2d_copy(m, m->w_feats.as<float>() + f_a * D, feats + gb->frame_offsets[ga] * D, sizeof(float) * (size_t)n_g * D, m->copy_stream));
        KHG_CUDA_TRY(cudaEventRecord(m->ev_copy[gidx & 1], m->copy_stream));
        KHG_CUDA_TRY(cudaStreamWaitEvent(st, m->ev_copy[gidx & 1], 0));
      }
      const double tg0 = now();
      if (!prep_hit) parallel_for(gb_ - ga, [&](int i, int w) { first_pass(ga + i, w); });
      const double tg1 = now();
      KHG_TRY(run_dense(ga, gb_, d_f0 + (gb->frame_offsets[ga] - fr0) * D, true, gb->frame_offsets[ga] - fr0));
      if (timing) fprintf(stderr, "  group %d: first pass %.2f ms, lists + launch %.2f ms\n", gidx, tg1 - tg0, now() - tg1);
      ga = gb_;
    }
    if (timing) cudaEventRecord(ev[1], st);
    if (!prep_hit) parallel_for(U - c1, [&](int i, int w) { first_pass(c1 + i, w); });
  } else if (!prep_hit) {
    parallel_for(U, first_pass);
  }
  t_mark[1] = now();
  std::vector<int32_t> ed_state, ed_off, utt_pdfs;
  std::vector<int4> in_arcs, e_arcs;
  int64_t eps_total = 0, n_in = 0;
  if (!prep_hit) {
  for (int u = 0; u < U; ++u)
    if (bad[u]) {
      cudaStreamSynchronize(st);  // the first chunk's dense kernel is already running
      KHG_REQUIRE(false, "graph of utterance " + std::to_string(u) + ": state / label / offset out of range");
    }
  // prefix sums: in-arc CSR over all states; epsilon-destination lists
  for (size_t s = 0; s < (size_t)S_all; ++s) in_off[s + 1] += in_off[s];
  ed_off.assign(1, 0);
  S_max = 1;
  n_pdf_max = 1;
  for (int u = 0; u < U; ++u) {
    UttDesc &d = desc[u];
    d.ed0 = (int32_t)ed_state.size();
    for (int32_t s = 0; s < d.S; ++s)
      if (eps_deg[d.state0 + s]) {
        ed_state.push_back(s);
        eps_total += eps_deg[d.state0 + s];
        ed_off.push_back((int32_t)eps_total);
      }
    d.n_ed = (int32_t)ed_state.size() - d.ed0;
    d.pdf0 = (int32_t)utt_pdfs.size();
    utt_pdfs.insert(utt_pdfs.end(), updf[u].begin(), updf[u].end());
    S_max = std::max(S_max, d.S);
    n_pdf_max = std::max(n_pdf_max, d.n_pdf);
  }
  t_mark[2] = now();
  n_in = in_off[S_all];
  in_arcs.assign((size_t)n_in, make_int4(0, 0, 0, 0));
  e_arcs.assign((size_t)eps_total, make_int4(0, 0, 0, 0));
  parallel_for(U, [&](int u, int) {
    const UttDesc &d = desc[u];
    const int32_t s0 = d.state0;
    std::vector<int32_t> fill(d.S, 0), eslot(d.S, -1), efill(d.n_ed, 0);
    for (int i = 0; i < d.n_ed; ++i) eslot[ed_state[d.ed0 + i]] = i;
    for (int32_t s = s0; s < s0 + d.S; ++s)
      for (int32_t a = gb->arc_offsets[s]; a < gb->arc_offsets[s + 1]; ++a) {
        const int32_t ns = gb->arc_nextstate[a];
        int32_t wbits;
        memcpy(&wbits, &gb->arc_weight[a], 4);
        if (gb->arc_ilabel[a] != 0) {
          in_arcs[(size_t)in_off[s0 + ns] + fill[ns]++] = make_int4(s - s0, arc_lp[a], a, wbits);
        } else {
          const int i = eslot[ns];
          e_arcs[(size_t)ed_off[d.ed0 + i] + efill[i]++] = make_int4(s - s0, a, wbits, 0);
        }
      }
  });

  pc.n_in = (size_t)n_in;
  pc.n_eds = ed_state.size();
  pc.n_edo = ed_off.size();
  pc.eps_total = (size_t)eps_total;
  pc.n_updf = utt_pdfs.size();
  }  // !prep_hit
  const double t_prep = now();
  t_mark[3] = t_prep;
  // ---------------- device copy of the graphs
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t o_desc = 0, o_inoff = o_desc + al(sizeof(UttDesc) * U), o_in = o_inoff + al(4 * ((size_t)S_all + 1)),
         o_eds = o_in + al(16 * pc.n_in), o_edo = o_eds + al(4 * pc.n_eds),
         o_ea = o_edo + al(4 * pc.n_edo), o_fin = o_ea + al(16 * pc.eps_total),
         o_src = o_fin + al(4 * (size_t)S_all), o_il = o_src + al(4 * (size_t)A_all), o_w = o_il + al(4 * (size_t)A_all),
         o_pdf = o_w + al(4 * (size_t)A_all), o_t2p = o_pdf + al(4 * pc.n_updf),
         o_ord = o_t2p + al(4 * (size_t)n_tids), o_out = o_ord + al(4 * (size_t)U),
         o_poff = o_out + al(sizeof(UttOut) * U), o_end = o_poff + al(8 * ((size_t)U + 1));
  KHG_TRY(m->w_al_graph.reserve(o_end));
  unsigned char *gbase = m->w_al_graph.as<unsigned char>();
  // (on the copy stream: queued on the model's stream the copies — from pageable vectors — would wait behind the dense
  // kernel of the first chunk, and the host with them; the search waits for ev_copy[0] below)
  KHG_TRY(ensure_copy_stream(m));
  cudaStream_t up_st = m->copy_stream;
  auto up = [&](size_t off, const void *src, size_t bytes) -> khg_status {
    return h2d_copy(m, gbase + off, src, bytes, up_st);  // (the arc arrays are a few MB each: staged by the pool's threads)
  };

  if (!prep_hit) {
  // launch order inside each chunk: longest search first
  std::vector<int32_t> order(U);
  for (size_t c = 0; c + 1 < chunk_start.size(); ++c) {
    const int u0 = chunk_start[c], u1 = chunk_start[c + 1];
    for (int u = u0; u < u1; ++u) order[u] = u - u0;
    std::stable_sort(order.begin() + u0, order.begin() + u1, [&](int a, int b) {
      return (int64_t)desc[u0 + a].T * desc[u0 + a].S > (int64_t)desc[u0 + b].T * desc[u0 + b].S;
    });
  }
  KHG_TRY(up(o_desc, desc.data(), sizeof(UttDesc) * U));
  KHG_TRY(up(o_inoff, in_off.data(), 4 * in_off.size()));
  KHG_TRY(up(o_in, in_arcs.data(), 16 * in_arcs.size()));
  KHG_TRY(up(o_eds, ed_state.data(), 4 * ed_state.size()));
  KHG_TRY(up(o_edo, ed_off.data(), 4 * ed_off.size()));
  KHG_TRY(up(o_ea, e_arcs.data(), 16 * e_arcs.size()));
  KHG_TRY(up(o_fin, gb->final_cost, 4 * (size_t)S_all));
  KHG_TRY(up(o_src, arc_src.data(), 4 * (size_t)A_all));
  KHG_TRY(up(o_il, gb->arc_ilabel, 4 * (size_t)A_all));
  KHG_TRY(up(o_w, gb->arc_weight, 4 * (size_t)A_all));
  KHG_TRY(up(o_pdf, utt_pdfs.data(), 4 * utt_pdfs.size()));
  KHG_TRY(up(o_t2p, tid2pdf, 4 * (size_t)n_tids));
  KHG_TRY(up(o_ord, order.data(), 4 * (size_t)U));
  KHG_CUDA_TRY(cudaEventRecord(m->ev_copy[0], up_st));
  KHG_CUDA_TRY(cudaStreamWaitEvent(st, m->ev_copy[0], 0));
  }  // !prep_hit: the device copy of an identical batch is still in w_al_graph
  pc.valid = true;
  t_mark[4] = now();
  AlignDev g;
  g.utts = reinterpret_cast<const UttDesc *>(gbase + o_desc);
  g.in_off = reinterpret_cast<const int32_t *>(gbase + o_inoff);
  g.in_arcs = reinterpret_cast<const int4 *>(gbase + o_in);
  g.ed_state = reinterpret_cast<const int32_t *>(gbase + o_eds);
  g.ed_off = reinterpret_cast<const int32_t *>(gbase + o_edo);
  g.e_arcs = reinterpret_cast<const int4 *>(gbase + o_ea);
  g.final_cost = reinterpret_cast<const float *>(gbase + o_fin);
  g.arc_src = reinterpret_cast<const int32_t *>(gbase + o_src);
  g.arc_il = reinterpret_cast<const int32_t *>(gbase + o_il);
  g.arc_w = reinterpret_cast<const float *>(gbase + o_w);
  g.utt_pdfs = reinterpret_cast<const int32_t *>(gbase + o_pdf);
  g.tid2pdf = reinterpret_cast<const int32_t *>(gbase + o_t2p);
  UttOut *d_outs = reinterpret_cast<UttOut *>(gbase + o_out);
  int64_t *d_poff = reinterpret_cast<int64_t *>(gbase + o_poff);

  // ---------------- kernel shape
  const size_t cost_bytes = sizeof(double) * 3 * (size_t)S_max;
  const size_t smem_budget = 200 * 1024;
  // KHG_ALIGN_FORCE_GCOST / KHG_ALIGN_FORCE_FC: tests force the large-graph paths (costs in global
  // scratch; likelihood tile of 16 / 8 / 0 frames) on small inputs
  const bool use_gcost = cost_bytes > 160 * 1024 || getenv("KHG_ALIGN_FORCE_GCOST") != nullptr;
  // Launch shape.  The search is latency-bound per utterance (one CTA walks its frames in sequence), so its time is
  // (waves of CTAs) x (latency of one utterance): fewer threads per utterance and a shorter likelihood tile are slower
  // per utterance (measured at C5, 205 states: 128 / 64 / 32 threads = 1 : 1.65 : 2.2; 8 instead of 32 frames per tile
  // +5 %) but more utterances are resident per SM.  The shape with the lowest estimate wins (C5, 2000 utterances: 64
  // threads x 8 frames, one wave, 5.4 ms against 6.5 ms for two waves of 128 x 32: profiles/r4f_align_search.txt).
  // KHG_ALIGN_FORCE_FC caps the tile (tests: the large-graph paths), KHG_ALIGN_NT fixes the threads.
  const char *force_fc = getenv("KHG_ALIGN_FORCE_FC");
  const int nt_base = S_max <= 256 ? 128 : (S_max <= 2048 ? 256 : 512);
  KHG_CUDA_TRY(cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));  // (per device: set on every call)
  int FC = 0, NT = nt_base;
  size_t smem = (use_gcost ? 0 : cost_bytes) + 16;
  {
    double best_est = 0.0;
    bool have = false;
    int nt_forced = 0;
    if (const char *e = getenv("KHG_ALIGN_NT")) nt_forced = std::max(32, std::min(512, atoi(e) & ~31));
    for (int nt : {nt_base, nt_base / 2, nt_base / 4}) {
      if (nt_forced) nt = nt_forced;
      if (nt < 32) continue;
      for (int fc : {32, 16, 8}) {
        const size_t sm = (use_gcost ? 0 : cost_bytes) + 4 * (size_t)fc * n_pdf_max + 16;
        if (sm - 16 > smem_budget || (force_fc && fc > atoi(force_fc))) continue;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, viterbi_kernel, nt, sm) != cudaSuccess || per_sm < 1) {
          (void)cudaGetLastError();
          continue;
        }
        const int64_t slots = (int64_t)per_sm * m->sm_count;
        const double waves = (double)((max_chunk_utts + slots - 1) / slots);
        const double lat = (nt >= nt_base ? 1.0 : (nt >= nt_base / 2 ? 1.65 : 2.2)) * (fc == 32 ? 1.0 : (fc == 16 ? 1.03 : 1.05));
        const double est = waves * lat;
        if (!have || est < best_est - 1e-9) {
          have = true;
          best_est = est;
          NT = nt;
          FC = fc;
          smem = sm;
        }
      }
    }
    if (!have) {  // no tile fits: the likelihoods are read from the block directly
      FC = 0;
      NT = nt_base;
      smem = (use_gcost ? 0 : cost_bytes) + 16;
    }
  }
  double *d_gcost = nullptr;
  if (use_gcost) {
    KHG_TRY(m->w_al_cost.reserve(sizeof(double) * 3 * (size_t)S_max * max_chunk_utts));
    d_gcost = m->w_al_cost.as<double>();
  }
  KHG_TRY(m->w_al_bp.reserve(std::max<size_t>(bp_bytes_max, 256)));
  KHG_TRY(m->w_al_ali.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(T_all, 1)));
  int32_t *d_ali = m->w_al_ali.as<int32_t>();
  KHG_CUDA_TRY(cudaMemsetAsync(d_ali, 0, sizeof(int32_t) * (size_t)T_all, st));
  if (pdf_ids_dev) KHG_CUDA_TRY(cudaMemsetAsync(pdf_ids_dev, 0, sizeof(int32_t) * (size_t)T_all, st));
  int32_t *d_bp = m->w_al_bp.as<int32_t>();
  std::vector<UttOut> h_outs(U);
  std::vector<int64_t> h_poff((size_t)U + 1, 0);
  std::vector<std::vector<int32_t>> x_ali(U), x_path(U), x_pdf(U);  // results of the exact host pass
  std::vector<char> is_exact(U, 0);
  int exact_mode = 1, n_redo_total = 0;  // 1 = flagged utterances only
  if (const char *e = getenv("KHG_ALIGN_EXACT")) exact_mode = !strcmp(e, "none") ? 0 : (!strcmp(e, "all") ? 2 : 1);
  double ms_exact = 0.0;

  for (size_t c = 0; c + 1 < chunk_start.size(); ++c) {
    const int u0 = chunk_start[c], u1 = chunk_start[c + 1];
    if (c > 0) {
      if (timing) cudaEventRecord(ev[0], st);
      KHG_TRY(launch_dense(u0, u1, want_subset));
      if (timing) cudaEventRecord(ev[1], st);
    }
    if (timing) cudaEventRecord(ev[3], st);  // (chunk 0: ev[0] / ev[1] were recorded around the early launch)
    g.order = reinterpret_cast<const int32_t *>(gbase + o_ord) + u0;
    viterbi_kernel<<<u1 - u0, NT, smem, st>>>(g, u0, d_block, ld, beam, retry_beam, S_max, n_pdf_max, FC, d_gcost, d_bp,
                                              d_ali, pdf_ids_dev, d_outs);
    ++g_launch_count;
    KHG_CUDA_TRY(cudaGetLastError());
    if (timing) cudaEventRecord(ev[2], st);
    KHG_CUDA_TRY(cudaMemcpyAsync(h_outs.data() + u0, d_outs + u0, sizeof(UttOut) * (u1 - u0), cudaMemcpyDeviceToHost, st));
    KHG_TRY(sync_and_check(m));
    if (timing) {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, ev[0], ev[1]);
      cudaEventElapsedTime(&b, ev[3], ev[2]);
      ms_dense += a;
      ms_search += b;
    }
    // ---- utterances the device search could not certify: the reference's decoder on the host, on the same
    // likelihood block (KHG_ALIGN_EXACT=all / none forces every / no utterance through it: tests, timing)
    std::vector<int> redo;
    for (int u = u0; u < u1; ++u)
      if (exact_mode != 0 && desc[u].start >= 0 && desc[u].S > 0 && (exact_mode == 2 || h_outs[u].pad != 0 || neg_eps[u])) redo.push_back(u);
    n_redo_total += (int)redo.size();
    if (!redo.empty()) {
      const double t_x0 = now();
      std::vector<int2> list(redo.size());
      std::vector<int64_t> off(redo.size() + 1, 0);
      for (size_t i = 0; i < redo.size(); ++i) {
        list[i] = make_int2(redo[i], 0);
        off[i + 1] = off[i] + (int64_t)desc[redo[i]].n_pdf * desc[redo[i]].T;
      }
      KHG_TRY(m->w_al_xlist.reserve(sizeof(int2) * list.size() + 8 * off.size()));
      KHG_TRY(m->w_al_xll.reserve(sizeof(float) * (size_t)std::max<int64_t>(off.back(), 1)));
      int2 *d_list = m->w_al_xlist.as<int2>();
      int64_t *d_off = reinterpret_cast<int64_t *>(d_list + list.size());
      KHG_CUDA_TRY(cudaMemcpyAsync(d_list, list.data(), sizeof(int2) * list.size(), cudaMemcpyHostToDevice, st));
      KHG_CUDA_TRY(cudaMemcpyAsync(d_off, off.data(), 8 * off.size(), cudaMemcpyHostToDevice, st));
      gather_ll_kernel<<<dim3((unsigned)redo.size(), 8), 256, 0, st>>>(g, d_list, d_off, d_block, ld, m->w_al_xll.as<float>());
      ++g_launch_count;
      KHG_CUDA_TRY(cudaGetLastError());
      std::vector<float> xll((size_t)off.back());
      if (off.back() > 0)
        KHG_CUDA_TRY(cudaMemcpyAsync(xll.data(), m->w_al_xll.p, sizeof(float) * xll.size(), cudaMemcpyDeviceToHost, st));
      KHG_CUDA_TRY(cudaStreamSynchronize(st));
      std::vector<khg_status> rst(redo.size(), KHG_OK);
      parallel_for((int)redo.size(), [&](int i, int) {
        const int u = redo[i];
        x_ali[u].assign((size_t)desc[u].T, 0);
        int32_t status = KHG_ALIGN_FAILED;
        float cost = 0.f;
        rst[i] = align_exact_host(gb, u, xll.data() + off[i], desc[u].T, nullptr, arc_lp.data(), beam, retry_beam,
                                  x_ali[u].data(), &status, &cost, &x_path[u]);
        h_outs[u].status = status;
        h_outs[u].path_len = (int32_t)x_path[u].size();
        h_outs[u].like = cost;
        is_exact[u] = 1;
      }, 1);
      for (khg_status r : rst) KHG_TRY(r);
      // the per-frame pdf ids the statistics pass consumes, and the device copy of the alignment
      for (int u : redo) {
        if (desc[u].T == 0) continue;
        KHG_CUDA_TRY(cudaMemcpyAsync(d_ali + desc[u].frame0, x_ali[u].data(), sizeof(int32_t) * desc[u].T, cudaMemcpyHostToDevice, st));
        if (pdf_ids_dev) {
          x_pdf[u].resize(desc[u].T);
          for (int t = 0; t < desc[u].T; ++t) x_pdf[u][t] = x_ali[u][t] > 0 ? tid2pdf[x_ali[u][t]] : 0;
          KHG_CUDA_TRY(cudaMemcpyAsync(pdf_ids_dev + desc[u].frame0, x_pdf[u].data(), sizeof(int32_t) * desc[u].T, cudaMemcpyHostToDevice, st));
        }
      }
      KHG_CUDA_TRY(cudaStreamSynchronize(st));
      ms_exact += now() - t_x0;
    }
    if (path_arcs) {
      for (int u = u0; u < u1; ++u) h_poff[u + 1] = h_poff[u] + (h_outs[u].status == KHG_ALIGN_FAILED ? 0 : h_outs[u].path_len);
      const int64_t n_path = h_poff[u1] - h_poff[u0];
      KHG_REQUIRE(h_poff[u1] <= path_capacity, "path_capacity too small");
      if (n_path > 0) {
        KHG_TRY(m->w_al_path.reserve(sizeof(int32_t) * (size_t)n_path));
        std::vector<int64_t> poff_dev(h_poff.begin() + u0, h_poff.begin() + u1);
        for (int u = u0; u < u1; ++u)
          if (is_exact[u]) poff_dev[u - u0] = -1;  // path_kernel skips them
        KHG_CUDA_TRY(cudaMemcpyAsync(d_poff + u0, poff_dev.data(), 8 * (size_t)(u1 - u0), cudaMemcpyHostToDevice, st));
        path_kernel<<<(u1 - u0 + 63) / 64, 64, 0, st>>>(g, u0, u1 - u0, d_bp, d_outs, d_poff, h_poff[u0],
                                                        m->w_al_path.as<int32_t>());
        ++g_launch_count;
        KHG_CUDA_TRY(cudaGetLastError());
        KHG_CUDA_TRY(cudaMemcpyAsync(path_arcs + h_poff[u0], m->w_al_path.p, sizeof(int32_t) * (size_t)n_path,
                                     cudaMemcpyDeviceToHost, st));
        KHG_CUDA_TRY(cudaStreamSynchronize(st));
        for (int u = u0; u < u1; ++u)
          if (is_exact[u] && !x_path[u].empty()) std::copy(x_path[u].begin(), x_path[u].end(), path_arcs + h_poff[u]);
      }
    }
  }
  if (alignment && T_all > 0)
    KHG_CUDA_TRY(cudaMemcpyAsync(alignment, d_ali, sizeof(int32_t) * (size_t)T_all, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  if (timing) {
    fprintf(stderr, "khg_align_batch: utts %d frames %lld states %d arcs %d | S_max %d n_pdf_max %d FC %d NT %d smem %zu chunks %zu | "
            "host prep %.2f ms, dense %.2f ms (%.1f %% of the tile units), search %.2f ms, exact host pass %.2f ms (%d utterances), total %.2f ms\n", U,
            (long long)T_all, S_all, A_all, S_max, n_pdf_max, FC, NT, smem, chunk_start.size() - 1, t_prep - t_begin, ms_dense,
            units_all > 0 ? 100.0 * units_done / units_all : 100.0, ms_search, ms_exact, n_redo_total, now() - t_begin);
    fprintf(stderr, "khg_align_batch host: setup %.2f | pass 1 (+ dense launches) %.2f | serial %.2f | pass 2 %.2f | upload %.2f ms\n",
            t_mark[0] - t_begin, t_mark[1] - t_mark[0], t_mark[2] - t_mark[1], t_mark[3] - t_mark[2], t_mark[4] - t_mark[3]);
    for (auto &e : ev) cudaEventDestroy(e);
  }
  g_align_exact_utts = n_redo_total;
  g_align_tile_fraction = units_all > 0 ? (double)units_done / (double)units_all : 1.0;
  for (int u = 0; u < U; ++u) {
    if (utt_status) utt_status[u] = h_outs[u].status;
    // decoder-wrappers.cc:91: like = -(graph + acoustic) / acoustic_scale
    if (utt_like) utt_like[u] = h_outs[u].status == KHG_ALIGN_FAILED ? 0.f : -h_outs[u].like / acoustic_scale;
    if (path_offsets) path_offsets[u + 1] = h_poff[u + 1];
  }
  return KHG_OK;
}

extern "C" int64_t khg_align_last_exact_count(void) { return g_align_exact_utts; }
extern "C" double khg_align_last_tile_fraction(void) { return g_align_tile_fraction; }
extern "C" int32_t khg_align_last_prep_cached(void) { return g_align_prep_hit; }
