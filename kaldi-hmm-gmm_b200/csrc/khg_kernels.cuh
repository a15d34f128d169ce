// kaldi-hmm-gmm_b200/csrc/khg_kernels.cuh — SIMT (fp32 FMA) kernels of the diag-GMM
// E-step: model pack (K4), dense all-pdf log-likelihoods (K1-simt, the any-shape
// companion of the tcgen05 kernel in khg_loglikes_tc.cu), pdf bucketing helpers
// (K2) and posteriors + statistics (K3).  Reference citations are relative to
// kaldi-hmm-gmm/ in csukuangfj/kaldi-hmm-gmm v1.1.4.
#ifndef KHG_KERNELS_CUH_
#define KHG_KERNELS_CUH_

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "khg_internal.h"

namespace khg {

__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.402823466e38f; }

// ---------------------------------------------------------------------------
// K4a: DiagGmm::ComputeGconsts (csrc/diag-gmm.cc:103-147), one thread per
// Gaussian.  The right-hand side of `gc += ...` is evaluated in double (0.5 is
// a double literal) and gc is rounded back to float at every step, like the
// reference.  flags[0] += #inf gconsts ("num_bad"), flags[1] |= NaN seen.
// ---------------------------------------------------------------------------
__global__ void gconsts_kernel(int G, int D, const float *__restrict__ w,
                               const float *__restrict__ miv,
                               const float *__restrict__ iv, float *__restrict__ gc_out,
                               int *__restrict__ flags) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const double kLog2Pi = 1.8378770664093454835606594728112;  // csrc/kaldi-math.h:25
  float offset = (float)(-0.5 * kLog2Pi * D);
  float gc = logf(w[g]) + offset;
  const float *m = miv + (size_t)g * D, *v = iv + (size_t)g * D;
  for (int d = 0; d < D; ++d) {
    double rhs = 0.5 * (double)logf(v[d]) - 0.5 * (double)m[d] * (double)m[d] / (double)v[d];
    gc = (float)((double)gc + rhs);
  }
  if (isnan(gc)) atomicOr(&flags[1], 1);
  if (isinf(gc)) {
    atomicAdd(&flags[0], 1);
    if (gc > 0) gc = -gc;
  }
  gc_out[g] = gc;
}

// K4b: SIMT chunk pack, packT[chunk][which][d][g%32] (see khg_internal.h).
__global__ void pack_simt_kernel(int G, int D, int n_chunks, const float *__restrict__ miv,
                                 const float *__restrict__ iv, float *__restrict__ packT) {
  size_t total = (size_t)n_chunks * 2 * D * kSimtChunk;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int gl = i % kSimtChunk;
    size_t r = i / kSimtChunk;
    int d = r % D;
    r /= D;
    int which = r % 2;
    int chunk = r / 2;
    int g = chunk * kSimtChunk + gl;
    float v = 0.f;
    if (g < G) v = (which ? iv : miv)[(size_t)g * D + d];
    packT[i] = v;
  }
}

// K4c: per-pdf groups of 8 Gaussians for K3: pack8[grp][which][d][8] and
// gc8[grp][8] (-inf padded); grp_start[p] = first group of pdf p.
__global__ void pack8_kernel(int P, int D, const int32_t *__restrict__ offsets,
                             const int32_t *__restrict__ grp_start,
                             const float *__restrict__ miv, const float *__restrict__ iv,
                             const float *__restrict__ gconsts, float *__restrict__ pack8,
                             float *__restrict__ gc8) {
  int p = blockIdx.x;
  if (p >= P) return;
  int g0 = offsets[p], ng = offsets[p + 1] - g0;
  int grp0 = grp_start[p], ngrp = grp_start[p + 1] - grp0;
  int per_grp = 2 * D * 8;
  for (int i = threadIdx.x; i < ngrp * per_grp; i += blockDim.x) {
    int j = i % 8;
    int r = i / 8;
    int d = r % D;
    r /= D;
    int which = r % 2;
    int grp = r / 2;
    int gl = grp * 8 + j;
    float v = 0.f;
    if (gl < ng) v = (which ? iv : miv)[(size_t)(g0 + gl) * D + d];
    pack8[(size_t)grp0 * per_grp + i] = v;
  }
  for (int i = threadIdx.x; i < ngrp * 8; i += blockDim.x)
    gc8[(size_t)grp0 * 8 + i] = i < ng ? gconsts[g0 + i] : -CUDART_INF_F;
}

// ---------------------------------------------------------------------------
// Running log-sum-exp with the reference's max-subtracted form
// (csrc/eigen.cc:14-18) evaluated online: (M, s) with sum = s * exp(M).
// -inf terms (zero-weight Gaussians, csrc/diag-gmm.cc:132-141) contribute 0; a
// NaN term poisons s so the non-finite result is flagged like the reference's
// throw (csrc/diag-gmm.cc:160-162).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void lse_push(float &M, float &s, float v) {
  if (v > M) {
    s = s * __expf(M - v) + 1.0f;
    M = v;
  } else if (!(v == -CUDART_INF_F)) {
    s += __expf(v - M);
  }
}

// ---------------------------------------------------------------------------
// K1-simt: all-pdf log-likelihoods, fp32 FMA.
// Replaces DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased
// (csrc/decodable-am-diag-gmm.cc:29-71) for a whole block of frames and pdfs:
//   ll(t,g) = gconst_g + means_invvars_g . x_t - 0.5 * inv_vars_g . x_t^2
//   out(t,p) = scale * LogSumExp_{g in pdf p} ll(t,g)
// CTA = 128 threads x 2 frames = 256 frames, loops over the Gaussians of the pdf
// range [p0, p1) of blockIdx.y.  Features are staged once, transposed, in smem
// (xs[d][frame], pitch 257: conflict-free both ways); model chunks of 32
// Gaussians are staged with a contiguous copy and read with broadcast LDS.128.
// out[p*stride_p + t*stride_t].
// ---------------------------------------------------------------------------
constexpr int kDenseFrames = 256;
constexpr int kDenseXP = 257;

__global__ void __launch_bounds__(128)
loglikes_simt_kernel(const float *__restrict__ feats, int64_t T, int D,
                     const float *__restrict__ packT, const float *__restrict__ gconsts,
                     const int32_t *__restrict__ offsets, int P, int pdfs_per_group,
                     float scale, float *__restrict__ out, int64_t stride_p,
                     int64_t stride_t, int *__restrict__ err, const unsigned *__restrict__ gate,
                     float gate_limit) {
  // gate != NULL: this launch is the fall-back of the fp16-split tensor-core kernel and runs
  // only when the call's features are outside fp16's range (decided on the device)
  if (gate != nullptr && __uint_as_float(*gate) <= gate_limit) return;
  extern __shared__ float smem[];
  float *xs = smem;                       // D x 257
  float *ms = smem + (((size_t)D * kDenseXP + 3) & ~(size_t)3);  // 2 x D x 32, 16-byte aligned
  const int tid = threadIdx.x;
  const int64_t t0 = (int64_t)blockIdx.x * kDenseFrames;
  const int nfr = (int)min((int64_t)kDenseFrames, T - t0);

  // stage features: contiguous block of nfr*D floats
  {
    const float *src = feats + t0 * D;
    int total = kDenseFrames * D;
    int valid = nfr * D;
    for (int e = tid; e < total; e += 128) {
      int r = e / D, d = e - r * D;
      xs[d * kDenseXP + r] = e < valid ? src[e] : 0.f;
    }
  }
  const int p0 = blockIdx.y * pdfs_per_group;
  const int p1 = min(P, p0 + pdfs_per_group);
  if (p0 >= p1) return;
  const int g_begin = offsets[p0], g_end = offsets[p1];
  int p = p0;
  int pdf_end = offsets[p + 1];
  float M0 = -CUDART_INF_F, s0 = 0.f, M1 = -CUDART_INF_F, s1 = 0.f;
  const bool v0ok = tid < nfr, v1ok = tid + 128 < nfr;
  bool bad = false;

  for (int chunk = g_begin / kSimtChunk; chunk * kSimtChunk < g_end; ++chunk) {
    __syncthreads();
    {
      const float4 *src = reinterpret_cast<const float4 *>(packT + (size_t)chunk * 2 * D * kSimtChunk);
      float4 *dst = reinterpret_cast<float4 *>(ms);
      for (int e = tid; e < 2 * D * kSimtChunk / 4; e += 128) dst[e] = src[e];
    }
    __syncthreads();
#pragma unroll 1
    for (int gb = 0; gb < kSimtChunk / 8; ++gb) {
      const int gbase = chunk * kSimtChunk + gb * 8;
      if (gbase + 8 <= g_begin || gbase >= g_end) continue;
      float a0[8], b0[8], a1[8], b1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a0[j] = b0[j] = a1[j] = b1[j] = 0.f;
      const float *mm = ms + gb * 8;
      const float *vv = ms + D * kSimtChunk + gb * 8;
#pragma unroll 2
      for (int d = 0; d < D; ++d) {
        float x0 = xs[d * kDenseXP + tid], x1 = xs[d * kDenseXP + 128 + tid];
        float q0 = x0 * x0, q1 = x1 * x1;  // data.array().square(), csrc/diag-gmm.cc:175
        float4 ma = *reinterpret_cast<const float4 *>(mm + d * kSimtChunk);
        float4 mb = *reinterpret_cast<const float4 *>(mm + d * kSimtChunk + 4);
        float4 va = *reinterpret_cast<const float4 *>(vv + d * kSimtChunk);
        float4 vb = *reinterpret_cast<const float4 *>(vv + d * kSimtChunk + 4);
        float mj[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
        float vj[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a0[j] = fmaf(mj[j], x0, a0[j]);
          b0[j] = fmaf(vj[j], q0, b0[j]);
          a1[j] = fmaf(mj[j], x1, a1[j]);
          b1[j] = fmaf(vj[j], q1, b1[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = gbase + j;
        if (g < g_begin || g >= g_end) continue;
        const float gc = __ldg(gconsts + g);
        lse_push(M0, s0, (gc + a0[j]) - 0.5f * b0[j]);  // csrc/diag-gmm.cc:174-175
        lse_push(M1, s1, (gc + a1[j]) - 0.5f * b1[j]);
        if (g + 1 == pdf_end) {
          float r0 = M0 + logf(s0), r1 = M1 + logf(s1);  // csrc/eigen.cc:17
          if (v0ok) {
            if (!finite_f(r0)) bad = true;
            out[p * stride_p + (t0 + tid) * stride_t] = scale * r0;
          }
          if (v1ok) {
            if (!finite_f(r1)) bad = true;
            out[p * stride_p + (t0 + tid + 128) * stride_t] = scale * r1;
          }
          ++p;
          pdf_end = p < P ? offsets[p + 1] : 0x7fffffff;
          M0 = M1 = -CUDART_INF_F;
          s0 = s1 = 0.f;
        }
      }
    }
  }
  if (bad) atomicOr(err, ERR_NONFINITE);
}

// 32x32 smem transpose: src is rows x cols (ld_src), dst is cols x rows (ld_dst).
__global__ void transpose_kernel(const float *__restrict__ src, int64_t rows, int64_t cols,
                                 int64_t ld_src, float *__restrict__ dst, int64_t ld_dst) {
  __shared__ float tile[32][33];
  int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * ld_src + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * ld_dst + r] = tile[threadIdx.x][i];
  }
}

// pdfs of more than 240 Gaussians run on the tensor-core kernel as several VIRTUAL pdfs (<= 240 Gaussians each, see
// tc_pack_build); src holds the unscaled log-sum-exp of every virtual pdf, row v = src + v * ld_src (frames
// contiguous).  dst[p * sp + t * st] = scale * LogSumExp over the virtual rows [vfirst[p], vfirst[p + 1]) of pdf p
// (csrc/eigen.cc:14-18 applied to the partial sums: log sum exp is associative).  A 32 x 32 tile goes through shared
// memory so that both layouts of dst are written coalesced.
__global__ void merge_virtual_kernel(const float *__restrict__ src, int64_t ld_src, const int32_t *__restrict__ vfirst, int P, int64_t T,
                                     float scale, float *__restrict__ dst, int64_t sp, int64_t st, int *__restrict__ err) {
  __shared__ float tile[32][33];
  const int64_t t0 = (int64_t)blockIdx.x * 32;
  const int p0 = blockIdx.y * 32;
  const int64_t t = t0 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i;
    if (p < P && t < T) {
      const int v0 = vfirst[p], v1 = vfirst[p + 1];
      float r = src[(int64_t)v0 * ld_src + t];
      if (v1 - v0 > 1) {
        float mx = r;
        for (int v = v0 + 1; v < v1; ++v) mx = fmaxf(mx, src[(int64_t)v * ld_src + t]);
        float sum = 0.f;
        for (int v = v0; v < v1; ++v) sum += expf(src[(int64_t)v * ld_src + t] - mx);
        r = mx + logf(sum);
      }
      if (!finite_f(r)) atomicOr(err, ERR_NONFINITE);
      tile[i][threadIdx.x] = scale * r;
    }
  }
  __syncthreads();
  if (st == 1) {  // pdf-major: rows of dst are pdfs, frames contiguous
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
      if (p0 + i < P && t < T) dst[(int64_t)(p0 + i) * sp + t] = tile[i][threadIdx.x];
  } else {        // frame-major: pdfs contiguous
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
      if (t0 + i < T && p0 + (int)threadIdx.x < P) dst[(t0 + i) * st + (int64_t)(p0 + threadIdx.x) * sp] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------
// Per-pdf, per-Gaussian log-likelihoods: DiagGmm::LogLikelihoods /
// LogLikelihoodsMatrix (csrc/diag-gmm.cc:167-189).  One thread per (t, g);
// API-parity path for the single-frame class methods, not a throughput path.
// ---------------------------------------------------------------------------
__global__ void pdf_loglikes_kernel(const float *__restrict__ feats, int64_t T, int D,
                                    const float *__restrict__ miv, const float *__restrict__ iv,
                                    const float *__restrict__ gc, int ng, float *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * ng) return;
  int64_t t = i / ng;
  int g = (int)(i - t * ng);
  const float *x = feats + t * D, *m = miv + (size_t)g * D, *v = iv + (size_t)g * D;
  float a = 0.f, b = 0.f;
  for (int d = 0; d < D; ++d) {
    float xv = x[d];
    a = fmaf(m[d], xv, a);
    b = fmaf(v[d], xv * xv, b);
  }
  out[i] = (gc[g] + a) - 0.5f * b;
}

// Softmax (csrc/eigen.cc:20-32) over each row of ll (T x ng): post = exp(ll-max)/sum,
// loglike = log(sum)+max.  One thread per frame.  post may alias ll or be NULL.
__global__ void pdf_softmax_kernel(const float *__restrict__ ll, int64_t T, int ng,
                                   float *__restrict__ post, float *__restrict__ loglike,
                                   int *__restrict__ err) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float *r = ll + t * ng;
  float mx = r[0];
  for (int g = 1; g < ng; ++g) mx = fmaxf(mx, r[g]);
  float s = 0.f;
  for (int g = 0; g < ng; ++g) s += expf(r[g] - mx);
  float lse = logf(s) + mx;
  if (!finite_f(lse)) atomicOr(err, ERR_NONFINITE);
  if (loglike) loglike[t] = lse;
  if (post) {
    float *po = post + t * ng;
    for (int g = 0; g < ng; ++g) po[g] = expf(r[g] - mx) / s;
  }
}

// ---------------------------------------------------------------------------
// K2 helpers: exact integer bucketing of frames by pdf id.
// ---------------------------------------------------------------------------
__global__ void prep_keys_kernel(const int32_t *__restrict__ ids, int64_t n, int P,
                                 int32_t *__restrict__ keys, int32_t *__restrict__ vals,
                                 int *__restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t k = ids[i];
  if (k < 0 || k >= P) {  // AccumulateForGmm asserts the range (csrc/mle-am-diag-gmm.cc:44)
    atomicOr(err, ERR_BAD_INDEX);
    k = P;  // sentinel bucket that no work item covers: the frame contributes nothing, as in stats_direct_kernel
  }
  keys[i] = k;
  vals[i] = (int32_t)i;
}

// tid -> pdf (csrc/transition-information.h:71-73) + transition counts
// (csrc/transition-model.h:183-189 with prob 1).
__global__ void map_tids_kernel(const int32_t *__restrict__ tids, int64_t n,
                                const int32_t *__restrict__ tid2pdf, int num_tids,
                                int32_t *__restrict__ pdf_ids,
                                unsigned long long *__restrict__ counts, int *__restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t tid = tids[i];
  if (tid < 1 || tid > num_tids) {
    atomicOr(err, ERR_BAD_INDEX);
    pdf_ids[i] = -1;
    return;
  }
  pdf_ids[i] = tid2pdf[tid];
  if (counts) atomicAdd(&counts[tid], 1ULL);
}

// starts[p] = first position of key >= p in the sorted key array (starts[P] = n).
__global__ void bucket_starts_kernel(const int32_t *__restrict__ sorted_keys, int64_t n, int P,
                                     int32_t *__restrict__ starts) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int32_t kprev = i == 0 ? -1 : sorted_keys[i - 1];
  int32_t kcur = i == n ? P : sorted_keys[i];
  for (int32_t p = kprev + 1; p <= kcur; ++p) starts[p] = (int32_t)i;
}

#ifndef KHG_STATS_UNROLL_A
#define KHG_STATS_UNROLL_A 1
#endif
#ifndef KHG_STATS_UNROLL_B
#define KHG_STATS_UNROLL_B 2
#endif
constexpr int kStatsUnrollA = KHG_STATS_UNROLL_A;  // unroll factors of the two hot loops of stats_kernel
constexpr int kStatsUnrollB = KHG_STATS_UNROLL_B;
constexpr int kStatsFrames = 128;      // frames per work item (upper bound)
constexpr int kStatsPostCapMax = 8192; // max floats of smem for the posterior tile
constexpr int kStatsMaxGp = 1600;      // largest pdf the statistics kernel accepts

// Row pitch (floats) for a smem matrix whose rows are read with LDS.128 by consecutive
// threads: a multiple of 4 whose 16-byte stride is odd => conflict-free.
__host__ __device__ inline int stats_pitch(int n) {
  int p = (n + 3) & ~3;
  if (((p >> 2) & 1) == 0) p += 4;
  return p;
}
// Frames per work item of a pdf with ng Gaussians, given the posterior-tile capacity.
__host__ __device__ inline int stats_frames_for(int ng, int post_cap) {
  int f = post_cap / stats_pitch(ng);
  return f < 1 ? 1 : (f > kStatsFrames ? kStatsFrames : f);
}

// item_start[p] = exclusive scan of ceil(n_p / frames_for(g_p)); single block.
__global__ void __launch_bounds__(1024) item_scan_kernel(int P, const int32_t *__restrict__ offsets,
                                                         const int32_t *__restrict__ starts,
                                                         int32_t *__restrict__ item_start, int post_cap) {
  // every thread owns a contiguous slice of pdfs; slice sums are scanned with warp shuffles
  __shared__ int32_t warp_tot[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (P + 1023) / 1024;
  const int p0 = min(P, tid * per), p1 = min(P, p0 + per);
  int32_t mine = 0;
  for (int p = p0; p < p1; ++p) {
    const int f = stats_frames_for(offsets[p + 1] - offsets[p], post_cap);
    mine += (starts[p + 1] - starts[p] + f - 1) / f;
  }
  int32_t incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int32_t w = warp_tot[lane], wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    warp_tot[lane] = wi - w;  // exclusive
  }
  __syncthreads();
  int32_t run = warp_tot[warp] + incl - mine;
  for (int p = p0; p < p1; ++p) {
    item_start[p] = run;
    const int f = stats_frames_for(offsets[p + 1] - offsets[p], post_cap);
    run += (starts[p + 1] - starts[p] + f - 1) / f;
  }
  if (tid == 1023) item_start[P] = run;  // (threads past the last pdf carry the total)
}

// item_desc[item] = {pdf, first position in `order`, frames, 0}: one warp per pdf writes the
// descriptors of its work items, so that a statistics CTA starts from one 16-byte load instead of a
// binary search over item_start.
__global__ void item_table_kernel(int P, const int32_t *__restrict__ offsets, const int32_t *__restrict__ starts,
                                  const int32_t *__restrict__ item_start, int post_cap, int4 *__restrict__ desc) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const int f = stats_frames_for(offsets[p + 1] - offsets[p], post_cap);
  const int s0 = starts[p], s1 = starts[p + 1], i0 = item_start[p], ni = item_start[p + 1] - i0;
  for (int i = threadIdx.x & 31; i < ni; i += 32) desc[i0 + i] = make_int4(p, s0 + i * f, min(f, s1 - (s0 + i * f)), 0);
}

// ---------------------------------------------------------------------------
// K3: posteriors + sufficient statistics for frames bucketed by pdf.
// Replaces, per frame, AccumAmDiagGmm::AccumulateForGmm
// (csrc/mle-am-diag-gmm.cc:41-52) -> AccumDiagGmm::AccumulateFromDiag
// (csrc/mle-diag-gmm.cc:145-158) -> DiagGmm::ComponentPosteriors
// (csrc/diag-gmm.cc:368-392) -> AccumulateFromPosteriors
// (csrc/mle-diag-gmm.cc:123-143).
// One CTA (128 threads) per work item = up to 128 frames of one pdf.
//   stage   gathered feature rows -> smem X[t][.] (row pitch chosen so that one
//           LDS.128 per thread-row is conflict-free)
//   phase A thread = frame: log-likes of the pdf's Gaussians (groups of 8 or 4, model
//           staged in smem, broadcast LDS.128), max-subtracted softmax
//           (csrc/eigen.cc:20-32), post *= w, totals (csrc/mle-am-diag-gmm.cc:49-50)
//   phase B a (4 Gaussians x 4 dims) register tile per thread over a quarter of the
//           frames: occ += post, mean += post*x, var += post*x^2 in fp32 over <=128
//           frames, then ONE fp64 atomicAdd per statistic (the reference casts every
//           frame's fp32 product to double; the difference is ~1e-7 relative).
// ---------------------------------------------------------------------------
struct StatsArgs {
  const float *feats;       // T x D (original order)
  const int32_t *order;     // frame indices sorted by pdf
  const float *weights;     // T or NULL
  const int32_t *starts;    // P+1 positions in `order`
  const int32_t *item_start;  // P+1
  const int4 *item_desc;      // per work item: {pdf, first position, frames, 0}
  const int32_t *offsets;   // P+1
  const int32_t *grp_start;   // P+1
  const float *pack8;
  const float *gc8;
  double *occ, *mean, *var;  // packed stats (mean/var may be NULL)
  double *totals;            // [tot_like, tot_frames]
  double *call_like;         // this call's sum ll*w (may be NULL)
  float *per_frame;          // T or NULL (original order)
  int *err;
  int P, D, grp_batch, post_cap;
  // optional: process only the items item_list[0 .. *item_list_n) — what the tensor-core kernel (khg_stats_tc.cu)
  // left to this one — grid-strided, the count read on the device
  const int32_t *item_list;
  const int *item_list_n;
};

// log-likes of NG (8 or 4) Gaussians of one group for the frame row xr (D floats).  Packed fp32x2
// arithmetic (FFMA2 on sm_100: two IEEE fused multiply-adds per issue slot, bitwise the same
// results as scalar fmaf): one instruction per pair of Gaussians.
template <int NG>
__device__ __forceinline__ void stats_group_ll(const float *__restrict__ xr, const float *__restrict__ mm,
                                               const float *__restrict__ vv, int D, float (&aa)[8], float (&bb)[8]) {
  float2 a2[4], b2[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) a2[j] = b2[j] = make_float2(0.f, 0.f);
  const int D4 = D & ~3;
#pragma unroll(kStatsUnrollA)
  for (int d0 = 0; d0 < D4; d0 += 4) {
    const float4 xq = *reinterpret_cast<const float4 *>(xr + d0);
    const float xs4[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
    for (int dd = 0; dd < 4; ++dd) {
      const float x = xs4[dd], q = x * x;  // data.array().square(), csrc/diag-gmm.cc:175
      const float2 xx = make_float2(x, x), qq = make_float2(q, q);
      const float4 ma = *reinterpret_cast<const float4 *>(mm + (d0 + dd) * 8);
      const float4 va = *reinterpret_cast<const float4 *>(vv + (d0 + dd) * 8);
      a2[0] = __ffma2_rn(make_float2(ma.x, ma.y), xx, a2[0]);
      a2[1] = __ffma2_rn(make_float2(ma.z, ma.w), xx, a2[1]);
      b2[0] = __ffma2_rn(make_float2(va.x, va.y), qq, b2[0]);
      b2[1] = __ffma2_rn(make_float2(va.z, va.w), qq, b2[1]);
      if (NG == 8) {
        const float4 mb = *reinterpret_cast<const float4 *>(mm + (d0 + dd) * 8 + 4);
        const float4 vb = *reinterpret_cast<const float4 *>(vv + (d0 + dd) * 8 + 4);
        a2[2] = __ffma2_rn(make_float2(mb.x, mb.y), xx, a2[2]);
        a2[3] = __ffma2_rn(make_float2(mb.z, mb.w), xx, a2[3]);
        b2[2] = __ffma2_rn(make_float2(vb.x, vb.y), qq, b2[2]);
        b2[3] = __ffma2_rn(make_float2(vb.z, vb.w), qq, b2[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    aa[2 * j] = a2[j].x; aa[2 * j + 1] = a2[j].y;
    bb[2 * j] = b2[j].x; bb[2 * j + 1] = b2[j].y;
  }
  for (int d = D4; d < D; ++d) {
    const float x = xr[d], q = x * x;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      aa[j] = fmaf(mm[d * 8 + j], x, aa[j]);
      bb[j] = fmaf(vv[d * 8 + j], q, bb[j]);
    }
  }
}

// The same for TWO frames per thread: every broadcast load of the model feeds both, which halves the
// shared-memory wavefronts of phase A (the kernel's bound: 40 of them per frame before, ncu r2f).
template <int NG>
__device__ __forceinline__ void stats_group_ll2(const float *__restrict__ xr0, const float *__restrict__ xr1,
                                                const float *__restrict__ mm, const float *__restrict__ vv, int D,
                                                float (&aa0)[8], float (&bb0)[8], float (&aa1)[8], float (&bb1)[8]) {
  float2 a0[4], b0[4], a1[4], b1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) a0[j] = b0[j] = a1[j] = b1[j] = make_float2(0.f, 0.f);
  const int D4 = D & ~3;
#pragma unroll(kStatsUnrollA)
  for (int d0 = 0; d0 < D4; d0 += 4) {
    const float4 xq0 = *reinterpret_cast<const float4 *>(xr0 + d0);
    const float4 xq1 = *reinterpret_cast<const float4 *>(xr1 + d0);
    const float xs0[4] = {xq0.x, xq0.y, xq0.z, xq0.w}, xs1[4] = {xq1.x, xq1.y, xq1.z, xq1.w};
#pragma unroll
    for (int dd = 0; dd < 4; ++dd) {
      const float x0 = xs0[dd], q0 = x0 * x0, x1 = xs1[dd], q1 = x1 * x1;  // data.array().square(), csrc/diag-gmm.cc:175
      const float2 xx0 = make_float2(x0, x0), qq0 = make_float2(q0, q0), xx1 = make_float2(x1, x1), qq1 = make_float2(q1, q1);
      const float4 ma = *reinterpret_cast<const float4 *>(mm + (d0 + dd) * 8);
      const float4 va = *reinterpret_cast<const float4 *>(vv + (d0 + dd) * 8);
      const float2 m01 = make_float2(ma.x, ma.y), m23 = make_float2(ma.z, ma.w), v01 = make_float2(va.x, va.y), v23 = make_float2(va.z, va.w);
      a0[0] = __ffma2_rn(m01, xx0, a0[0]); a0[1] = __ffma2_rn(m23, xx0, a0[1]);
      b0[0] = __ffma2_rn(v01, qq0, b0[0]); b0[1] = __ffma2_rn(v23, qq0, b0[1]);
      a1[0] = __ffma2_rn(m01, xx1, a1[0]); a1[1] = __ffma2_rn(m23, xx1, a1[1]);
      b1[0] = __ffma2_rn(v01, qq1, b1[0]); b1[1] = __ffma2_rn(v23, qq1, b1[1]);
      if (NG == 8) {
        const float4 mb = *reinterpret_cast<const float4 *>(mm + (d0 + dd) * 8 + 4);
        const float4 vb = *reinterpret_cast<const float4 *>(vv + (d0 + dd) * 8 + 4);
        const float2 m45 = make_float2(mb.x, mb.y), m67 = make_float2(mb.z, mb.w), v45 = make_float2(vb.x, vb.y), v67 = make_float2(vb.z, vb.w);
        a0[2] = __ffma2_rn(m45, xx0, a0[2]); a0[3] = __ffma2_rn(m67, xx0, a0[3]);
        b0[2] = __ffma2_rn(v45, qq0, b0[2]); b0[3] = __ffma2_rn(v67, qq0, b0[3]);
        a1[2] = __ffma2_rn(m45, xx1, a1[2]); a1[3] = __ffma2_rn(m67, xx1, a1[3]);
        b1[2] = __ffma2_rn(v45, qq1, b1[2]); b1[3] = __ffma2_rn(v67, qq1, b1[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    aa0[2 * j] = a0[j].x; aa0[2 * j + 1] = a0[j].y; bb0[2 * j] = b0[j].x; bb0[2 * j + 1] = b0[j].y;
    aa1[2 * j] = a1[j].x; aa1[2 * j + 1] = a1[j].y; bb1[2 * j] = b1[j].x; bb1[2 * j + 1] = b1[j].y;
  }
  for (int d = D4; d < D; ++d) {
    const float x0 = xr0[d], q0 = x0 * x0, x1 = xr1[d], q1 = x1 * x1;
#pragma unroll
    for (int j = 0; j < NG; ++j) {
      const float mv = mm[d * 8 + j], vvv = vv[d * 8 + j];
      aa0[j] = fmaf(mv, x0, aa0[j]); bb0[j] = fmaf(vvv, q0, bb0[j]);
      aa1[j] = fmaf(mv, x1, aa1[j]); bb1[j] = fmaf(vvv, q1, bb1[j]);
    }
  }
}

__global__ void __launch_bounds__(128) stats_kernel(StatsArgs a) {
  extern __shared__ float smem[];
  const int D = a.D;
  const int XP = stats_pitch(D);
  float *X = smem;                                      // 128 x XP (tail of each row zero)
  float *post = X + kStatsFrames * XP;                  // post_cap floats: [t][PG]
  float *ms = post + a.post_cap;                        // grp_batch x 2 x D x 8
  float *gcs = ms + (size_t)a.grp_batch * 2 * D * 8;    // grp_batch x 8
  float *wsm = gcs + a.grp_batch * 8;                   // 128 weights
  __shared__ int s_idx[kStatsFrames];
  __shared__ double s_red[8];
  const int tid = threadIdx.x;
  const int n_work = a.item_list ? *a.item_list_n : a.item_start[a.P];
  for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
  if (work != (int)blockIdx.x) __syncthreads();  // (the previous item's shared memory is done with)
  const int item = a.item_list ? a.item_list[work] : work;
  const int4 desc = __ldg(a.item_desc + item);
  const int p = desc.x, pos0 = desc.y, n = desc.z;
  const int g0 = a.offsets[p], ng = a.offsets[p + 1] - g0;
  const int PG = stats_pitch(ng);
  const int grp0 = a.grp_start[p], ngrp = a.grp_start[p + 1] - grp0;

  // stage features (gathered rows): the item's row indices go to shared memory first; then 16
  // lanes per row issue the row's 16-byte chunks with cp.async (8 rows per pass, no per-row
  // warp-wide bookkeeping); the pad columns of every row are zeroed by the row's owner thread.
  {
    int idx = 0;
    if (tid < n) {
      idx = a.order[pos0 + tid];
      s_idx[tid] = idx;
      wsm[tid] = a.weights ? a.weights[idx] : 1.0f;
    }
    for (int d = D; d < XP; ++d) X[tid * XP + d] = 0.f;
    __syncthreads();
    const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.feats) & 15) == 0);
    if (vec) {
      // one thread per gathered row: its 16-byte chunks back to back (16 lanes per row cost 15 warp instructions per
      // frame in address arithmetic and loop control, ncu r2f; the rows are scattered, so nothing is lost in coalescing,
      // and consecutive rows land conflict-free in shared memory thanks to the row pitch)
      if (tid < n) {
        const float *src = a.feats + (size_t)idx * D;
        const uint32_t d32 = static_cast<uint32_t>(__cvta_generic_to_shared(X + tid * XP));
        const int chunks = D >> 2;
#pragma unroll 4
        for (int c = 0; c < chunks; ++c)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d32 + 16 * c), "l"(src + 4 * c) : "memory");
      }
    } else {
      const int warp = tid >> 5, lane = tid & 31;
      for (int r = warp; r < n; r += 4) {  // rows that are only 4-byte aligned (e.g. dim 39): 4-byte cp.async
        const float *src = a.feats + (size_t)s_idx[r] * D;
        for (int d = lane; d < D; d += 32) {
          const uint32_t d32 = static_cast<uint32_t>(__cvta_generic_to_shared(X + r * XP + d));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d32), "l"(src + d) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");

  // ---- phase A ----
  for (int gb0 = 0; gb0 < ngrp; gb0 += a.grp_batch) {
    const int nb = min(a.grp_batch, ngrp - gb0);
    __syncthreads();
    {
      const float4 *src = reinterpret_cast<const float4 *>(a.pack8 + (size_t)(grp0 + gb0) * 2 * D * 8);
      float4 *dst = reinterpret_cast<float4 *>(ms);
      for (int e = tid; e < nb * 2 * D * 8 / 4; e += 128) dst[e] = src[e];
      for (int e = tid; e < nb * 8; e += 128) gcs[e] = a.gc8[(size_t)(grp0 + gb0) * 8 + e];
    }
    __syncthreads();
    if (nb >= 2) {
      // warp pair k = warps (2k, 2k+1) takes groups k, k+2, ...; lane l of the pair's warp h handles frames
      // f = 32 h + l and f + 64
      const int pair = tid >> 6, f0 = tid & 63, f1 = f0 + 64;
      if (f0 < n) {
        const float *xr0 = X + f0 * XP, *xr1 = X + min(f1, n - 1) * XP;
        for (int gb = pair; gb < nb; gb += 2) {
          float aa0[8], bb0[8], aa1[8], bb1[8];
          const float *mm = ms + (size_t)gb * 2 * D * 8;
          const float *vv = mm + D * 8;
          const int gl0 = (gb0 + gb) * 8;
          const int cnt = min(8, ng - gl0);
          if (cnt > 4) stats_group_ll2<8>(xr0, xr1, mm, vv, D, aa0, bb0, aa1, bb1);
          else stats_group_ll2<4>(xr0, xr1, mm, vv, D, aa0, bb0, aa1, bb1);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < cnt) {
              const float gc = gcs[gb * 8 + j];
              post[f0 * PG + gl0 + j] = (gc + aa0[j]) - 0.5f * bb0[j];  // csrc/diag-gmm.cc:174-175
              if (f1 < n) post[f1 * PG + gl0 + j] = (gc + aa1[j]) - 0.5f * bb1[j];
            }
        }
      }
    } else if (tid < n) {
      const float *xr = X + tid * XP;
      for (int gb = 0; gb < nb; ++gb) {
        float aa[8], bb[8];
        const float *mm = ms + (size_t)gb * 2 * D * 8;
        const float *vv = mm + D * 8;
        const int gl0 = (gb0 + gb) * 8;
        const int cnt = min(8, ng - gl0);
        if (cnt > 4) stats_group_ll<8>(xr, mm, vv, D, aa, bb);
        else stats_group_ll<4>(xr, mm, vv, D, aa, bb);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j < cnt) post[tid * PG + gl0 + j] = (gcs[gb * 8 + j] + aa[j]) - 0.5f * bb[j];  // csrc/diag-gmm.cc:174-175
      }
    }
  }
  // softmax over the pdf's Gaussians (csrc/eigen.cc:20-32), then post *= weight
  // (csrc/mle-diag-gmm.cc:153); totals as csrc/mle-am-diag-gmm.cc:49-50.
  double my_like = 0.0, my_w = 0.0;
  __syncthreads();  // a frame's log-likes may have been written by several warps (one per Gaussian group)
  if (tid < n) {
    float *pr = post + tid * PG;
    const float w = wsm[tid];
    float lse;
    if (PG <= 20) {
      // up to 16 Gaussians (+ pad): the row goes through registers with 128-bit accesses — the scalar walk below
      // costs four conflicted passes over shared memory
      float v[20];
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (4 * q < PG) {
          const float4 t4 = *reinterpret_cast<const float4 *>(pr + 4 * q);
          v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
        }
      float mx = v[0];
#pragma unroll
      for (int g = 1; g < 20; ++g)
        if (g < ng) mx = fmaxf(mx, v[g]);
      float s = 0.f;
#pragma unroll
      for (int g = 0; g < 20; ++g) {
        if (g < ng) {
          v[g] = __expf(v[g] - mx);
          s += v[g];
        } else {
          v[g] = 0.f;
        }
      }
      lse = logf(s) + mx;
      const float rs = __frcp_rn(s);  // exp / sum within 1 ulp of the reference's division (csrc/eigen.cc:29-31)
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (4 * q < PG)
          *reinterpret_cast<float4 *>(pr + 4 * q) = make_float4((v[4 * q] * rs) * w, (v[4 * q + 1] * rs) * w, (v[4 * q + 2] * rs) * w,
                                                                (v[4 * q + 3] * rs) * w);
    } else {
      float mx = pr[0];
      for (int g = 1; g < ng; ++g) mx = fmaxf(mx, pr[g]);
      float s = 0.f;
      for (int g = 0; g < ng; ++g) {
        float e = __expf(pr[g] - mx);
        pr[g] = e;
        s += e;
      }
      lse = logf(s) + mx;
      const float rs = __frcp_rn(s);
      for (int g = 0; g < ng; ++g) pr[g] = (pr[g] * rs) * w;
      for (int g = ng; g < PG; ++g) pr[g] = 0.f;
    }
    if (!finite_f(lse)) atomicOr(a.err, ERR_NONFINITE);
    if (a.per_frame) a.per_frame[s_idx[tid]] = lse;
    my_like = (double)(lse * w);
    my_w = (double)w;
  }
  for (int off = 16; off > 0; off >>= 1) {
    my_like += __shfl_xor_sync(0xffffffffu, my_like, off);
    my_w += __shfl_xor_sync(0xffffffffu, my_w, off);
  }
  if ((tid & 31) == 0) {
    s_red[tid >> 5] = my_like;
    s_red[4 + (tid >> 5)] = my_w;
  }
  __syncthreads();
  if (tid == 0) {
    double L = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    double W = s_red[4] + s_red[5] + s_red[6] + s_red[7];
    atomicAdd(&a.totals[0], L);
    atomicAdd(&a.totals[1], W);
    if (a.call_like) atomicAdd(a.call_like, L);
  }
  // ---- phase B ----
  // (4 Gaussians x 4 dims) register tiles; when they fit the CTA, TQ = 128 / tiles frame slices per
  // tile, whose partial sums are added in shared memory so that every statistic leaves the CTA as
  // ONE fp64 atomic; pdfs with more than 128 tiles take the tiles in passes.
  const int n_gt = PG >> 2, n_dt = (D + 3) >> 2;
  const int tiles = n_gt * n_dt;
  const int TQ = tiles <= 128 ? 128 / tiles : 1;
  const int n_act = tiles * TQ;
  const bool reduce = TQ > 1 && 36 * n_act <= kStatsFrames * XP + a.post_cap;
  const int per = (n + TQ - 1) / TQ;
  for (int w0 = 0; w0 < n_act; w0 += 128) {  // (one pass when the tiles fit the CTA)
    const int w = w0 + tid;
    const bool act = w < n_act;
    const int tq = w / tiles, tile = w - tq * tiles;
    const int gt = tile / n_dt, dt = tile - gt * n_dt;
    // packed fp32x2 accumulators: (dims 0,1) and (dims 2,3) of the tile per Gaussian, occupancies in pairs
    float2 occ2[2], sm2[4][2], sv2[4][2];
    occ2[0] = occ2[1] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) sm2[i][0] = sm2[i][1] = sv2[i][0] = sv2[i][1] = make_float2(0.f, 0.f);
    if (act) {
      const int ta = tq * per, tb = min(n, ta + per);
      const float *pp = post + gt * 4, *xp = X + dt * 4;
#pragma unroll(kStatsUnrollB)
      for (int t = ta; t < tb; ++t) {
        const float4 pq = *reinterpret_cast<const float4 *>(pp + t * PG);
        const float4 xq = *reinterpret_cast<const float4 *>(xp + t * XP);
        const float2 x01 = make_float2(xq.x, xq.y), x23 = make_float2(xq.z, xq.w);
        const float2 q01 = __fmul2_rn(x01, x01), q23 = __fmul2_rn(x23, x23);
        occ2[0] = __fadd2_rn(occ2[0], make_float2(pq.x, pq.y));
        occ2[1] = __fadd2_rn(occ2[1], make_float2(pq.z, pq.w));
        const float pv[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 p2 = make_float2(pv[i], pv[i]);
          sm2[i][0] = __ffma2_rn(p2, x01, sm2[i][0]);
          sm2[i][1] = __ffma2_rn(p2, x23, sm2[i][1]);
          sv2[i][0] = __ffma2_rn(p2, q01, sv2[i][0]);
          sv2[i][1] = __ffma2_rn(p2, q23, sv2[i][1]);
        }
      }
    }
    const float occ[4] = {occ2[0].x, occ2[0].y, occ2[1].x, occ2[1].y};
    float sm[4][4], sv[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sm[i][0] = sm2[i][0].x; sm[i][1] = sm2[i][0].y; sm[i][2] = sm2[i][1].x; sm[i][3] = sm2[i][1].y;
      sv[i][0] = sv2[i][0].x; sv[i][1] = sv2[i][0].y; sv[i][2] = sv2[i][1].x; sv[i][3] = sv2[i][1].y;
    }
    if (reduce) {
      __syncthreads();  // every thread has finished reading X / post: they become the scratch
      float *scr = smem;  // [36][n_act]
      if (act) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          scr[i * n_act + w] = occ[i];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            scr[(4 + i * 4 + j) * n_act + w] = sm[i][j];
            scr[(20 + i * 4 + j) * n_act + w] = sv[i][j];
          }
        }
      }
      __syncthreads();
      for (int g = tid; g < ng; g += 128) {
        const int tl = (g >> 2) * n_dt, k = g & 3;
        float v = 0.f;
        for (int q = 0; q < TQ; ++q) v += scr[k * n_act + q * tiles + tl];
        atomicAdd(&a.occ[g0 + g], (double)v);
      }
      if (a.mean) {
        for (int o = tid; o < ng * D; o += 128) {
          const int g = o / D, d = o - g * D;
          const int tl = (g >> 2) * n_dt + (d >> 2), k = 4 + (g & 3) * 4 + (d & 3);
          float v = 0.f, u = 0.f;
          for (int q = 0; q < TQ; ++q) {
            v += scr[k * n_act + q * tiles + tl];
            u += scr[(k + 16) * n_act + q * tiles + tl];
          }
          atomicAdd(&a.mean[(size_t)(g0 + g) * D + d], (double)v);
          if (a.var) atomicAdd(&a.var[(size_t)(g0 + g) * D + d], (double)u);
        }
      }
    } else if (act) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int g = gt * 4 + i;
        if (g >= ng) continue;
        if (dt == 0) atomicAdd(&a.occ[g0 + g], (double)occ[i]);
        if (a.mean) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int d = dt * 4 + j;
            if (d >= D) continue;
            atomicAdd(&a.mean[(size_t)(g0 + g) * D + d], (double)sm[i][j]);
            if (a.var) atomicAdd(&a.var[(size_t)(g0 + g) * D + d], (double)sv[i][j]);
          }
        }
      }
    }
  }
  }  // work items of this CTA
}

// ---------------------------------------------------------------------------
// K3-direct: the same per-frame arithmetic for SMALL batches (one utterance): no bucketing,
// one warp per frame, lanes over feature dimensions.  It exists for launch latency — the
// reference's script calls gmm-acc-stats-ali once per utterance (scripts/gmm_acc_stats_ali.py)
// — where the sort + work-item pipeline of the batched path costs ~10 launches.  Every
// posterior-weighted product is rounded to fp32, cast to double and added with an fp64
// atomic: exactly csrc/mle-diag-gmm.cc:131-141.
// ---------------------------------------------------------------------------
constexpr int kDirectMaxGp = 256;   // Gaussians per pdf this path accepts
constexpr int kDirectWarps = 8;

struct DirectArgs {
  const float *feats;       // T x D
  const int32_t *ids;       // T pdf ids
  const float *weights;     // T or NULL
  const int32_t *offsets;   // P+1
  const float *miv, *iv, *gconsts;  // packed model, row-major
  double *occ, *mean, *var, *totals, *call_like;
  float *per_frame;
  int *err;
  int T, P, D;
};

__global__ void __launch_bounds__(32 * kDirectWarps) stats_direct_kernel(DirectArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * kDirectWarps + warp;
  if (t >= a.T) return;
  const int D = a.D;
  float *xs = smem + (size_t)warp * (D + kDirectMaxGp);  // this warp's frame
  float *ll = xs + D;                                      // and its log-likes / posteriors
  int p = a.ids[t];
  if (p < 0 || p >= a.P) {  // AccumulateForGmm asserts the range (csrc/mle-am-diag-gmm.cc:44)
    if (lane == 0) atomicOr(a.err, ERR_BAD_INDEX);
    return;
  }
  const int g0 = a.offsets[p], ng = a.offsets[p + 1] - g0;
  const float w = a.weights ? a.weights[t] : 1.0f;
  const float *x = a.feats + (size_t)t * D;
  for (int d = lane; d < D; d += 32) xs[d] = x[d];
  __syncwarp();
  for (int g = 0; g < ng; ++g) {
    const float *m = a.miv + (size_t)(g0 + g) * D, *v = a.iv + (size_t)(g0 + g) * D;
    float sa = 0.f, sb = 0.f;
    for (int d = lane; d < D; d += 32) {
      const float xv = xs[d];
      sa = fmaf(m[d], xv, sa);
      sb = fmaf(v[d], xv * xv, sb);
    }
    for (int off = 16; off > 0; off >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, off);
      sb += __shfl_xor_sync(0xffffffffu, sb, off);
    }
    if (lane == 0) ll[g] = (a.gconsts[g0 + g] + sa) - 0.5f * sb;  // csrc/diag-gmm.cc:174-175
  }
  __syncwarp();
  // Softmax (csrc/eigen.cc:20-32): lanes over Gaussians
  float mx = -CUDART_INF_F;
  for (int g = lane; g < ng; g += 32) mx = fmaxf(mx, ll[g]);
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float s = 0.f;
  for (int g = lane; g < ng; g += 32) {
    const float e = expf(ll[g] - mx);
    ll[g] = e;
    s += e;
  }
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float lse = logf(s) + mx;
  for (int g = lane; g < ng; g += 32) ll[g] = (ll[g] / s) * w;  // post *= weight (csrc/mle-diag-gmm.cc:153)
  __syncwarp();
  if (lane == 0) {
    if (!finite_f(lse)) atomicOr(a.err, ERR_NONFINITE);
    if (a.per_frame) a.per_frame[t] = lse;
    const double L = (double)(lse * w);  // csrc/mle-am-diag-gmm.cc:49-50
    atomicAdd(&a.totals[0], L);
    atomicAdd(&a.totals[1], (double)w);
    if (a.call_like) atomicAdd(a.call_like, L);
  }
  for (int g = lane; g < ng; g += 32) atomicAdd(&a.occ[g0 + g], (double)ll[g]);
  if (a.mean) {
    for (int g = 0; g < ng; ++g) {
      const float pg = ll[g];
      for (int d = lane; d < D; d += 32) {
        const float xv = xs[d];
        atomicAdd(&a.mean[(size_t)(g0 + g) * D + d], (double)__fmul_rn(pg, xv));
        if (a.var) atomicAdd(&a.var[(size_t)(g0 + g) * D + d], (double)__fmul_rn(pg, __fmul_rn(xv, xv)));
      }
    }
  }
}

// rows of a pdf-major block: dst[i][t] = src[subset[i]][t]
__global__ void gather_rows_kernel(const float *__restrict__ src, int64_t ld_src, const int32_t *__restrict__ subset,
                                   int n, int64_t T, float *__restrict__ dst, int64_t ld_dst, int P, int *err) {
  const int i = blockIdx.y;
  const int p = subset[i];
  if (p < 0 || p >= P) {
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicOr(err, ERR_BAD_INDEX);
    return;
  }
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x)
    dst[i * ld_dst + t] = src[p * ld_src + t];
}

// AccumDiagGmm::AccumulateFromPosteriors (csrc/mle-diag-gmm.cc:123-143) for T frames of
// one pdf with caller-supplied posteriors (T x ng).  One thread per (g, k) pair.
__global__ void acc_from_post_kernel(const float *__restrict__ feats, int64_t T, int D,
                                     const float *__restrict__ post, int ng, int g0,
                                     double *occ, double *mean, double *var, double *totals) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int KK = D + 1;
  if (e >= ng * KK) return;
  int g = e / KK, k = e - g * KK;
  if (k == D) {
    double o = 0.0;
    for (int64_t t = 0; t < T; ++t) o += (double)post[t * ng + g];
    occ[g0 + g] += o;
    atomicAdd(&totals[1], o);  // csrc/mle-am-diag-gmm.cc:85: total_frames_ += posteriors.sum()
  } else if (mean) {
    double sm = 0.0, sv = 0.0;
    for (int64_t t = 0; t < T; ++t) {
      float pv = post[t * ng + g], x = feats[t * D + k];
      sm += (double)(pv * x);          // fp32 product, then cast (csrc/mle-diag-gmm.cc:134)
      sv += (double)(pv * (x * x));
    }
    mean[(size_t)(g0 + g) * D + k] += sm;
    if (var) var[(size_t)(g0 + g) * D + k] += sv;
  }
}

// ---------------------------------------------------------------------------
// Device M-step: MleDiagGmmUpdate (csrc/mle-diag-gmm.cc:243-390) for every pdf.
// One CTA per pdf.  Pass 1 (mle_update_kernel): old objective, new parameters in the
// packed layout (same Gaussian slots), removal flags, counters.  gconsts of the updated
// (pre-removal) model come from gconsts_kernel; pass 2 (mle_objective_kernel) evaluates
// MlObjective (csrc/mle-diag-gmm.cc:479-499) per pdf; mle_compact_kernel gathers the
// kept Gaussians into the new model and renormalises weights like RemoveComponents
// (csrc/diag-gmm.cc:853-937: one removal at a time, fp32).
// ---------------------------------------------------------------------------
struct MleArgs {
  int P, D;
  uint16_t acc_flags, upd_flags;
  float min_w, min_occ;
  double min_var;
  int remove_low;
  const int32_t *offsets;
  const double *occ, *mean, *var;       // packed stats (mean/var may be NULL)
  const float *w_old, *miv_old, *iv_old;
  float *w_new, *miv_new, *iv_new;       // same slots as the old model
  int32_t *remove;                       // G flags
  int32_t *counters;                     // [0] floored elements, [1] floored gaussians, [2] removed
  double *pdf_occ;                       // P: sum of occupancies (count)
};

__device__ __forceinline__ double block_sum_128(double v, double *red) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(128) mle_update_kernel(MleArgs a) {
  __shared__ double red[4];
  __shared__ int s_cand;
  const int p = blockIdx.x, tid = threadIdx.x, D = a.D;
  const int g0 = a.offsets[p], ng = a.offsets[p + 1] - g0;
  double part = 0.0;
  for (int g = tid; g < ng; g += 128) part += a.occ[g0 + g];
  const double occ_sum = block_sum_128(part, red);
  if (tid == 0) {
    a.pdf_occ[p] = occ_sum;
    s_cand = 0;
  }
  __syncthreads();
  // candidates for removal are counted first: when every Gaussian of the pdf is a
  // candidate the last one is kept (to_remove.size() < num_gauss - 1, :338-339)
  int my_cand = 0;
  for (int g = tid; g < ng; g += 128) {
    const double occ = a.occ[g0 + g];
    const double prob = occ_sum > 0.0 ? occ / occ_sum : 1.0 / ng;
    if (!(occ > (double)a.min_occ && prob > (double)a.min_w)) ++my_cand;
  }
  if (my_cand) atomicAdd(&s_cand, my_cand);
  __syncthreads();
  const int n_cand = s_cand;
  int last_cand = -1;
  if (n_cand == ng) last_cand = ng - 1;  // all are candidates: the highest index survives
  for (int g = tid; g < ng; g += 128) {
    const size_t G = g0 + g;
    const double occ = a.occ[G];
    const double prob = occ_sum > 0.0 ? occ / occ_sum : 1.0 / ng;
    const float *miv_o = a.miv_old + G * D, *iv_o = a.iv_old + G * D;
    float *miv_n = a.miv_new + G * D, *iv_n = a.iv_new + G * D;
    float w = a.w_old[G];
    int rm = 0;
    if (occ > (double)a.min_occ && prob > (double)a.min_w) {
      if (a.upd_flags & KHG_GMM_WEIGHTS) w = (float)prob;
      int floored = 0;
      for (int d = 0; d < D; ++d) {
        // DiagGmmNormal (csrc/diag-gmm-normal.cc:14-20): double mean / variance form
        const double old_var = 1.0 / (double)iv_o[d];
        const double old_mean = (double)miv_o[d] * old_var;
        double mean = old_mean, var = old_var;
        if (a.acc_flags & (KHG_GMM_MEANS | KHG_GMM_VARIANCES)) mean = a.mean[G * D + d] / occ;
        if (a.acc_flags & KHG_GMM_VARIANCES) {
          var = a.var[G * D + d] / occ - mean * mean;
          if (!(a.upd_flags & KHG_GMM_MEANS)) {  // :300-304
            const double dm = old_mean - mean;
            var += dm * dm;
          }
          if (var < a.min_var) {
            var = a.min_var;
            ++floored;
          }
        }
        // CopyToDiagGmm (csrc/diag-gmm-normal.cc:22-48)
        float ivf = iv_o[d], mivf = miv_o[d];
        if (a.upd_flags & KHG_GMM_VARIANCES) {
          ivf = (float)(1.0 / var);
          if (!(a.upd_flags & KHG_GMM_MEANS)) mivf = (float)old_mean * ivf;
        }
        if (a.upd_flags & KHG_GMM_MEANS) mivf = (float)mean * ivf;
        iv_n[d] = ivf;
        miv_n[d] = mivf;
      }
      if (floored) {
        atomicAdd(&a.counters[0], floored);
        atomicAdd(&a.counters[1], 1);
      }
    } else {
      for (int d = 0; d < D; ++d) {
        iv_n[d] = iv_o[d];
        miv_n[d] = miv_o[d];
      }
      if (a.remove_low && g != last_cand && ng > 1) {
        rm = 1;
        atomicAdd(&a.counters[2], 1);
      } else if (a.upd_flags & KHG_GMM_WEIGHTS) {
        w = (float)fmax(prob, (double)a.min_w);  // :353-354
      }
    }
    a.w_new[G] = w;
    a.remove[G] = rm;
  }
}

// obj[p] = MlObjective of pdf p (csrc/mle-diag-gmm.cc:479-499), reference float roundings.
__global__ void __launch_bounds__(128)
mle_objective_kernel(int P, int D, uint16_t acc_flags, const int32_t *__restrict__ offsets,
                     const double *__restrict__ occ, const double *__restrict__ mean, const double *__restrict__ var,
                     const float *__restrict__ gc, const float *__restrict__ miv, const float *__restrict__ iv,
                     float *__restrict__ obj) {
  __shared__ double red[4];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int g0 = offsets[p], ng = offsets[p + 1] - g0;
  double o = 0.0, sm = 0.0, sv = 0.0;
  for (int g = tid; g < ng; g += 128) {
    o += occ[g0 + g] * (double)gc[g0 + g];
  }
  const size_t e0 = (size_t)g0 * D, ne = (size_t)ng * D;
  if (acc_flags & KHG_GMM_MEANS)
    for (size_t e = tid; e < ne; e += 128) sm += mean[e0 + e] * (double)miv[e0 + e];
  if (acc_flags & KHG_GMM_VARIANCES)
    for (size_t e = tid; e < ne; e += 128) sv += var[e0 + e] * (double)iv[e0 + e];
  o = block_sum_128(o, red);
  sm = block_sum_128(sm, red);
  sv = block_sum_128(sv, red);
  if (tid == 0) {
    float f = (float)o;
    if (acc_flags & KHG_GMM_MEANS) f = (float)((double)f + sm);
    if (acc_flags & KHG_GMM_VARIANCES) f = (float)((double)f - 0.5 * sv);
    obj[p] = f;
  }
}

// Gathers the kept Gaussians of pdf p into the new model; weights are renormalised the
// way RemoveComponents does it: after each single removal, in fp32, in index order.
__global__ void __launch_bounds__(128)
mle_compact_kernel(int P, int D, const int32_t *__restrict__ old_off, const int32_t *__restrict__ new_off,
                   const int32_t *__restrict__ remove, const float *__restrict__ w, const float *__restrict__ miv,
                   const float *__restrict__ iv, float *__restrict__ w_out, float *__restrict__ miv_out,
                   float *__restrict__ iv_out) {
  extern __shared__ float ws[];  // ng weights + ng flags (as float)
  const int p = blockIdx.x, tid = threadIdx.x;
  const int g0 = old_off[p], ng = old_off[p + 1] - g0, n0 = new_off[p];
  int *dst = reinterpret_cast<int *>(ws + ng);
  for (int g = tid; g < ng; g += 128) ws[g] = w[g0 + g];
  __syncthreads();
  if (tid == 0) {
    int k = 0, removed = 0;
    for (int g = 0; g < ng; ++g) dst[g] = remove[g0 + g] ? -1 : k++;
    // sequential renormalisation: one pass per removed component over the survivors
    // that are still present at that point (components removed later are still in)
    for (int g = 0; g < ng; ++g) {
      if (!remove[g0 + g]) continue;
      float s = 0.f;
      for (int h = 0; h < ng; ++h)
        if (h != g && !(remove[g0 + h] && h < g)) s += ws[h];
      for (int h = 0; h < ng; ++h)
        if (h != g && !(remove[g0 + h] && h < g)) ws[h] /= s;
      ++removed;
    }
  }
  __syncthreads();
  for (int g = tid; g < ng; g += 128)
    if (dst[g] >= 0) w_out[n0 + dst[g]] = ws[g];
  for (int e = tid; e < ng * D; e += 128) {
    const int g = e / D, d = e - g * D;
    if (dst[g] >= 0) {
      miv_out[(size_t)(n0 + dst[g]) * D + d] = miv[(size_t)(g0 + g) * D + d];
      iv_out[(size_t)(n0 + dst[g]) * D + d] = iv[(size_t)(g0 + g) * D + d];
    }
  }
}

__global__ void axpy_f64_kernel(double *dst, const double *src, double scale, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] += src[i] * scale;
}
__global__ void scale_f64_kernel(double *dst, double scale, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    dst[i] *= scale;
}

}  // namespace khg
#endif  // KHG_KERNELS_CUH_
