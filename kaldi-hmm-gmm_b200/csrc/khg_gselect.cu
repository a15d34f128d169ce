// kaldi-hmm-gmm_b200/csrc/khg_gselect.cu — khg_gaussian_selection: DiagGmm::GaussianSelection
// (reference kaldi-hmm-gmm/csrc/diag-gmm.cc:202-317) and GaussianSelectionPreselect (:319-366) for a
// batch of frames ("other users of the dense kernel").
//
// The per-Gaussian log-likelihoods are the dense kernel's accumulators before the per-pdf
// log-sum-exp, so the pdf's Gaussians are viewed as a model of one-Gaussian pdfs (same operand
// rows, gconsts carried over): K1 then writes the (Gaussians x frames) block, on the tensor cores
// when the shape allows, and `gselect_kernel` takes the top num_gselect per frame.
//
// Selection rule of the reference: threshold = the (n - k)-th order statistic, every component
// >= threshold, pairs (loglike, index) sorted with std::greater, first k kept, tot = LogAdd chain in
// that order (csrc/kaldi-math.h:60-78) = the k largest under the lexicographic order (loglike, index).
// For preselect lists with duplicates the list position is the last tie-break.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

#include <math_constants.h>

#include "khg_internal.h"

namespace khg {

constexpr int kGselWarps = 8;

struct GselKey {
  float v;
  int lab, pos;
};
__device__ __forceinline__ bool gsel_less(const GselKey &a, const GselKey &b) {  // a < b
  return a.v < b.v || (a.v == b.v && (a.lab < b.lab || (a.lab == b.lab && a.pos < b.pos)));
}

// block: per-Gaussian log-likes, Gaussian-major (block[g * ld + t]).  One CTA = FR consecutive
// frames; the tile is transposed into shared memory (row = frame, odd pitch) and each warp selects
// for one frame at a time: every lane keeps the best not-yet-selected item of its residue class,
// a warp arg-max picks the winner, only the winning lane rescans.
__global__ void __launch_bounds__(32 * kGselWarps) gselect_kernel(const float *__restrict__ block, int64_t ld, int64_t T, int n,
                                                                  const int32_t *__restrict__ presel, int kk, int FR,
                                                                  int32_t *__restrict__ out_idx, float *__restrict__ out_ll,
                                                                  float *__restrict__ frame_like, int *__restrict__ err) {
  extern __shared__ float tile[];
  const int ns = n | 1;
  const int64_t t0 = (int64_t)blockIdx.x * FR;
  const int nf = (int)min((int64_t)FR, T - t0);
  for (int e = threadIdx.x; e < n * FR; e += blockDim.x) {
    const int i = e / FR, f = e - i * FR;
    if (f < nf) tile[f * ns + i] = block[(int64_t)(presel ? presel[i] : i) * ld + t0 + f];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float kMinLogDiff = logf(FLT_EPSILON);  // kMinLogDiffFloat, csrc/kaldi-math.h:36
  for (int f = warp; f < nf; f += kGselWarps) {
    const float *row = tile + f * ns;
    GselKey bound = {CUDART_INF_F, 0x7fffffff, 0x7fffffff};  // everything is below it
    GselKey mine = {-CUDART_INF_F, -1, -1};
    bool have = false, bad = false;
    for (int i = lane; i < n; i += 32) {
      const GselKey c = {row[i], presel ? presel[i] : i, i};
      bad |= c.v != c.v;
      if (!have || gsel_less(mine, c)) { mine = c; have = true; }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(err, ERR_NONFINITE);
    float tot = -CUDART_INF_F;
    const int64_t o = (t0 + f) * kk;
    for (int j = 0; j < kk; ++j) {
      GselKey w = mine;
      bool wh = have;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        GselKey c;
        c.v = __shfl_xor_sync(0xffffffffu, w.v, s);
        c.lab = __shfl_xor_sync(0xffffffffu, w.lab, s);
        c.pos = __shfl_xor_sync(0xffffffffu, w.pos, s);
        const bool ch = __shfl_xor_sync(0xffffffffu, (int)wh, s) != 0;
        if (ch && (!wh || gsel_less(w, c))) { w = c; wh = true; }
      }
      if (lane == 0) {
        out_idx[o + j] = w.lab;
        if (out_ll) out_ll[o + j] = w.v;
        // LogAdd(tot, w.v), csrc/kaldi-math.h:60-78
        float x = tot, y = w.v, diff;
        if (x < y) { diff = x - y; x = y; } else { diff = y - x; }
        tot = diff >= kMinLogDiff ? x + log1pf(expf(diff)) : x;
      }
      if (have && mine.pos == w.pos) {  // the winner's lane: next best below the one just taken
        bound = w;
        have = false;
        for (int i = lane; i < n; i += 32) {
          const GselKey c = {row[i], presel ? presel[i] : i, i};
          if (gsel_less(c, bound) && (!have || gsel_less(mine, c))) { mine = c; have = true; }
        }
      }
      __syncwarp();
    }
    if (lane == 0 && frame_like) frame_like[t0 + f] = tot;
  }
}

khg_status gsel_shadow(khg_model *m, int32_t pdf, khg_model **out) {
  // the shadow shares the parent's scratch buffers, so it always works on the parent's CURRENT stream
  if (m->gsel_shadow && m->gsel_pdf == pdf) { m->gsel_shadow->stream = m->stream; *out = m->gsel_shadow; return KHG_OK; }
  if (m->gsel_shadow) { khg_model_destroy(m->gsel_shadow); m->gsel_shadow = nullptr; }
  const int g0 = m->h_offsets[pdf], ng = m->h_offsets[pdf + 1] - g0, D = m->dim;
  std::vector<float> miv((size_t)ng * D), iv((size_t)ng * D), gc(ng), w(ng, 1.0f);
  KHG_CUDA_TRY(cudaMemcpy(miv.data(), m->d_miv + (size_t)g0 * D, sizeof(float) * miv.size(), cudaMemcpyDeviceToHost));
  KHG_CUDA_TRY(cudaMemcpy(iv.data(), m->d_iv + (size_t)g0 * D, sizeof(float) * iv.size(), cudaMemcpyDeviceToHost));
  KHG_CUDA_TRY(cudaMemcpy(gc.data(), m->d_gconsts + g0, sizeof(float) * ng, cudaMemcpyDeviceToHost));
  std::vector<int32_t> offs(ng + 1);
  for (int i = 0; i <= ng; ++i) offs[i] = i;
  khg_model *s = nullptr;
  KHG_TRY(khg_model_create(D, ng, offs.data(), &s));
  // one Gaussian per pdf fits every dense kernel: automatic choice whatever the parent model uses
  khg_status st = khg_model_upload(s, w.data(), miv.data(), iv.data(), gc.data(), nullptr);
  if (st != KHG_OK) { khg_model_destroy(s); return st; }
  // a zero-weight Gaussian (gconst = -inf) is a legitimate, never-selected candidate here, not
  // the "all components dead" pdf that makes LogLikelihood throw
  s->tc.dead_pdf = false;
  s->stream = m->stream;
  m->gsel_shadow = s;
  m->gsel_pdf = pdf;
  *out = s;
  return KHG_OK;
}

}  // namespace khg

using namespace khg;

extern "C" khg_status khg_gaussian_selection(khg_model *m, int32_t pdf, const float *feats, int64_t T, int32_t feats_loc,
                                             const int32_t *preselect, int32_t n_preselect, int32_t num_gselect,
                                             int32_t *out_indices, float *out_loglikes, float *frame_loglike,
                                             double *tot_loglike) {
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  KHG_REQUIRE(pdf >= 0 && pdf < m->P, "pdf_index out of range");
  KHG_REQUIRE(T > 0 && feats && out_indices, "num_frames != 0");  // csrc/diag-gmm.cc:272
  KHG_REQUIRE(num_gselect > 0, "num_gselect > 0");
  const int ng = m->h_offsets[pdf + 1] - m->h_offsets[pdf], D = m->dim;
  KHG_REQUIRE(n_preselect >= 0 && (n_preselect == 0 || preselect), "null preselect");
  for (int i = 0; i < n_preselect; ++i) KHG_REQUIRE(preselect[i] >= 0 && preselect[i] < ng, "preselect index out of range");
  const int n = n_preselect > 0 ? n_preselect : ng;
  const int kk = std::min(num_gselect, n);
  khg_model *sh = nullptr;
  KHG_TRY(gsel_shadow(m, pdf, &sh));
  cudaStream_t st = sh->stream;
  int FR = 32;
  while (FR > 1 && sizeof(float) * (size_t)FR * (n | 1) > 96 * 1024) FR >>= 1;
  const size_t smem = sizeof(float) * (size_t)FR * (n | 1);
  if (smem > 200 * 1024) {
    set_error("too many candidates for the selection kernel");
    return KHG_ERR_UNSUPPORTED;
  }
  KHG_CUDA_TRY(cudaFuncSetAttribute(gselect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));  // (per device: set on every call)
  const int32_t *d_pre = nullptr;
  if (n_preselect > 0) {
    KHG_TRY(m->w_sub.reserve(sizeof(int32_t) * n_preselect));
    KHG_CUDA_TRY(cudaMemcpyAsync(m->w_sub.p, preselect, sizeof(int32_t) * n_preselect, cudaMemcpyHostToDevice, st));
    d_pre = m->w_sub.as<int32_t>();
  }
  int64_t chunk = std::max<int64_t>(256, std::min<int64_t>(T, (int64_t)(512e6 / (4.0 * ng))) & ~(int64_t)255);
  if (const char *e = getenv("KHG_GSEL_CHUNK_FRAMES")) chunk = std::max<int64_t>(1, atoll(e));  // tests: force several chunks
  double tot = 0.0;
  std::vector<float> h_like;
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t nT = std::min(chunk, T - t0);
    const float *d_f = feats + t0 * D;
    if (feats_loc == KHG_HOST) {
      KHG_TRY(m->w_feats.reserve(sizeof(float) * (size_t)nT * D));
      KHG_CUDA_TRY(cudaMemcpyAsync(m->w_feats.p, d_f, sizeof(float) * (size_t)nT * D, cudaMemcpyHostToDevice, st));
      d_f = m->w_feats.as<float>();
    }
    const int64_t ld = (nT + 3) & ~(int64_t)3;
    KHG_TRY(m->w_full.reserve(sizeof(float) * (size_t)ng * ld));
    KHG_TRY(dense_block(sh, d_f, nT, 1.0f, KHG_PDF_MAJOR, m->w_full.as<float>(), ld));
    KHG_TRY(m->w_ids.reserve(sizeof(int32_t) * (size_t)nT * kk));
    KHG_TRY(m->w_pf.reserve(sizeof(float) * (size_t)nT * (kk + 1)));
    int32_t *d_idx = m->w_ids.as<int32_t>();
    float *d_ll = m->w_pf.as<float>(), *d_fl = d_ll + (size_t)nT * kk;
    gselect_kernel<<<(unsigned)((nT + FR - 1) / FR), 32 * kGselWarps, smem, st>>>(m->w_full.as<float>(), ld, nT, n, d_pre, kk, FR,
                                                                                d_idx, out_loglikes ? d_ll : nullptr, d_fl, sh->d_err);
    ++g_launch_count;
    KHG_CUDA_TRY(cudaGetLastError());
    KHG_CUDA_TRY(cudaMemcpyAsync(out_indices + t0 * kk, d_idx, sizeof(int32_t) * (size_t)nT * kk, cudaMemcpyDeviceToHost, st));
    if (out_loglikes)
      KHG_CUDA_TRY(cudaMemcpyAsync(out_loglikes + t0 * kk, d_ll, sizeof(float) * (size_t)nT * kk, cudaMemcpyDeviceToHost, st));
    float *hl = frame_loglike ? frame_loglike + t0 : (h_like.resize(nT), h_like.data());
    KHG_CUDA_TRY(cudaMemcpyAsync(hl, d_fl, sizeof(float) * (size_t)nT, cudaMemcpyDeviceToHost, st));
    KHG_TRY(sync_and_check(sh));
    for (int64_t i = 0; i < nT; ++i) tot += (double)hl[i];  // "ans += tot_loglike", csrc/diag-gmm.cc:312
  }
  if (tot_loglike) *tot_loglike = tot;
  return KHG_OK;
}
