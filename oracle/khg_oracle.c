/* oracle/khg_oracle.c
 *
 * TEST INFRASTRUCTURE — see khg_oracle.h for the scope and pinning statement.
 * Plain C11, no dependencies.  Build: oracle/Makefile.
 *
 * Floating-point conventions follow the reference: the model is fp32, the
 * likelihood arithmetic is fp32 (Eigen FloatVector/FloatMatrix,
 * csrc/eigen.h:10-22), the statistics are fp64 and each posterior-weighted
 * product is rounded to fp32 BEFORE the cast to double
 * (csrc/mle-diag-gmm.cc:132-141).  Eigen's vectorised reduction order is
 * unspecified, so dot products use 8 interleaved partial sums (what an 8-lane
 * packet reduction does); bit-exactness with Eigen is neither defined nor
 * required — BASELINE.json's tolerances govern.
 */
#include "khg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KHG_M_LOG_2PI 1.8378770664093454835606594728112 /* csrc/kaldi-math.h:25 */

/* ---- csrc/eigen.cc:14-18 ---------------------------------------------- */
float khg_oracle_logsumexp(const float *v, int32_t n) {
  float max_v = v[0];
  for (int32_t i = 1; i < n; ++i)
    if (v[i] > max_v) max_v = v[i];
  float s = 0.0f;
  for (int32_t i = 0; i < n; ++i) s += expf(v[i] - max_v);
  return logf(s) + max_v;
}

/* ---- csrc/eigen.cc:20-32 ---------------------------------------------- */
void khg_oracle_softmax(const float *v, int32_t n, float *out,
                        float *log_sum_exp) {
  float max_v = v[0];
  for (int32_t i = 1; i < n; ++i)
    if (v[i] > max_v) max_v = v[i];
  float s = 0.0f;
  for (int32_t i = 0; i < n; ++i) {
    out[i] = expf(v[i] - max_v);
    s += out[i];
  }
  if (log_sum_exp) *log_sum_exp = logf(s) + max_v;
  for (int32_t i = 0; i < n; ++i) out[i] = out[i] / s;
}

/* ---- csrc/model-common.cc:72-84 --------------------------------------- */
uint16_t khg_oracle_augment_flags(uint16_t flags) {
  if (flags & KHG_ORACLE_GMM_VARIANCES) flags |= KHG_ORACLE_GMM_MEANS;
  if (flags & KHG_ORACLE_GMM_MEANS) flags |= KHG_ORACLE_GMM_WEIGHTS;
  if (!(flags & KHG_ORACLE_GMM_WEIGHTS)) flags |= KHG_ORACLE_GMM_WEIGHTS;
  return flags;
}

/* ---- csrc/diag-gmm.cc:103-147 ----------------------------------------- */
int32_t khg_oracle_compute_gconsts(int32_t nmix, int32_t dim,
                                   const float *weights,
                                   const float *means_invvars,
                                   const float *inv_vars, float *gconsts) {
  float offset = (float)(-0.5 * KHG_M_LOG_2PI * dim); /* :106 */
  int32_t num_bad = 0;
  for (int32_t mix = 0; mix < nmix; ++mix) {
    float gc = logf(weights[mix]) + offset; /* :119, may be -inf */
    for (int32_t d = 0; d < dim; ++d) {
      float iv = inv_vars[(size_t)mix * dim + d];
      float miv = means_invvars[(size_t)mix * dim + d];
      /* :122-125: 0.5 is a double literal, so the right-hand side is evaluated
       * in double and gc is rounded back to float at each step. */
      double rhs = 0.5 * (double)logf(iv) - 0.5 * (double)miv * (double)miv / (double)iv;
      gc = (float)((double)gc + rhs);
    }
    if (isnan(gc)) return -1; /* :132-135 KHG_ERR */
    if (isinf(gc)) {          /* :136-141 */
      num_bad++;
      if (gc > 0) gc = -gc;
    }
    gconsts[mix] = gc;
  }
  return num_bad;
}

/* fp32 dot with 8 interleaved partial sums (see header comment). */
static inline float dot8(const float *a, const float *b, int32_t n) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int32_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int32_t j = 0; j < 8; ++j) acc[j] += a[i + j] * b[i + j];
  float tail = 0.0f;
  for (; i < n; ++i) tail += a[i] * b[i];
  float s01 = acc[0] + acc[4], s23 = acc[1] + acc[5];
  float s45 = acc[2] + acc[6], s67 = acc[3] + acc[7];
  return ((s01 + s45) + (s23 + s67)) + tail;
}

#define KHG_MAX_STACK_DIM 512

/* ---- csrc/diag-gmm.cc:167-176 ----------------------------------------- */
static void loglikes_sq(int32_t nmix, int32_t dim, const float *gconsts,
                        const float *miv, const float *iv, const float *x,
                        const float *xsq, float *out) {
  for (int32_t g = 0; g < nmix; ++g) {
    float a = dot8(miv + (size_t)g * dim, x, dim);
    float b = dot8(iv + (size_t)g * dim, xsq, dim);
    out[g] = (gconsts[g] + a) - 0.5f * b; /* :174-175 */
  }
}

void khg_oracle_loglikes(int32_t nmix, int32_t dim, const float *gconsts,
                         const float *means_invvars, const float *inv_vars,
                         const float *x, float *loglikes) {
  float stack_sq[KHG_MAX_STACK_DIM] = {0};
  float *xsq = dim <= KHG_MAX_STACK_DIM ? stack_sq : (float *)malloc(sizeof(float) * dim);
  for (int32_t d = 0; d < dim; ++d) xsq[d] = x[d] * x[d];
  loglikes_sq(nmix, dim, gconsts, means_invvars, inv_vars, x, xsq, loglikes);
  if (xsq != stack_sq) free(xsq);
}

/* ---- csrc/diag-gmm.cc:177-189 ----------------------------------------- */
void khg_oracle_loglikes_matrix(int32_t nmix, int32_t dim, const float *gconsts,
                                const float *means_invvars,
                                const float *inv_vars, const float *feats,
                                int64_t T, float *out) {
  for (int64_t t = 0; t < T; ++t)
    khg_oracle_loglikes(nmix, dim, gconsts, means_invvars, inv_vars,
                        feats + t * dim, out + t * nmix);
}

/* ---- csrc/diag-gmm.cc:150-165 ----------------------------------------- */
int32_t khg_oracle_log_likelihood(int32_t nmix, int32_t dim,
                                  const float *gconsts,
                                  const float *means_invvars,
                                  const float *inv_vars, const float *x,
                                  float *log_like) {
  float *ll = (float *)malloc(sizeof(float) * nmix);
  khg_oracle_loglikes(nmix, dim, gconsts, means_invvars, inv_vars, x, ll);
  float s = khg_oracle_logsumexp(ll, nmix);
  free(ll);
  *log_like = s;
  return (isnan(s) || isinf(s)) ? 1 : 0; /* :160-162 */
}

/* ---- csrc/diag-gmm.cc:368-392 ----------------------------------------- */
int32_t khg_oracle_component_posteriors(int32_t nmix, int32_t dim,
                                        const float *gconsts,
                                        const float *means_invvars,
                                        const float *inv_vars, const float *x,
                                        float *posteriors, float *log_like) {
  float *ll = (float *)malloc(sizeof(float) * nmix);
  khg_oracle_loglikes(nmix, dim, gconsts, means_invvars, inv_vars, x, ll);
  float s;
  khg_oracle_softmax(ll, nmix, posteriors, &s);
  free(ll);
  *log_like = s;
  return (isnan(s) || isinf(s)) ? 1 : 0; /* :385-387 */
}

/* ---- csrc/mle-diag-gmm.cc:123-143 ------------------------------------- */
static inline void acc_from_post_sq(int32_t nmix, int32_t dim, uint16_t flags,
                                    const float *x, const float *xsq,
                                    const float *post, double *occ,
                                    double *mean_acc, double *var_acc) {
  for (int32_t g = 0; g < nmix; ++g) occ[g] += (double)post[g]; /* :131 */
  if (flags & KHG_ORACLE_GMM_MEANS) {
    for (int32_t g = 0; g < nmix; ++g) {
      float p = post[g];
      double *m = mean_acc + (size_t)g * dim;
      /* :134 — (posteriors * data^T) is an fp32 outer product, THEN cast */
      for (int32_t d = 0; d < dim; ++d) m[d] += (double)(p * x[d]);
    }
    if (flags & KHG_ORACLE_GMM_VARIANCES) { /* :136-141, inside the means branch */
      for (int32_t g = 0; g < nmix; ++g) {
        float p = post[g];
        double *v = var_acc + (size_t)g * dim;
        for (int32_t d = 0; d < dim; ++d) v[d] += (double)(p * xsq[d]);
      }
    }
  }
}

void khg_oracle_acc_from_posteriors(int32_t nmix, int32_t dim, uint16_t flags,
                                    const float *x, const float *posteriors,
                                    double *occ, double *mean_acc,
                                    double *var_acc) {
  flags = khg_oracle_augment_flags(flags);
  float *xsq = (float *)malloc(sizeof(float) * dim);
  for (int32_t d = 0; d < dim; ++d) xsq[d] = x[d] * x[d];
  acc_from_post_sq(nmix, dim, flags, x, xsq, posteriors, occ, mean_acc, var_acc);
  free(xsq);
}

/* ---- csrc/mle-diag-gmm.cc:100-121 ------------------------------------- */
void khg_oracle_acc_for_component(int32_t nmix, int32_t dim, uint16_t flags,
                                  const float *x, int32_t comp, float weight,
                                  double *occ, double *mean_acc,
                                  double *var_acc) {
  (void)nmix;
  flags = khg_oracle_augment_flags(flags);
  double wt = weight;
  occ[comp] += wt; /* :112 */
  if (flags & KHG_ORACLE_GMM_MEANS) {
    for (int32_t d = 0; d < dim; ++d) /* :115 data.cast<double>() * wt */
      mean_acc[(size_t)comp * dim + d] += (double)x[d] * wt;
    if (flags & KHG_ORACLE_GMM_VARIANCES) {
      for (int32_t d = 0; d < dim; ++d) {
        /* :117-119: (x.square() * wt) — float vector times double scalar;
         * Eigen promotes nothing here: the product expression has scalar type
         * double only through an explicit cast, which comes AFTER, so the
         * product is taken in fp32 with wt narrowed to float. */
        float sq = x[d] * x[d];
        var_acc[(size_t)comp * dim + d] += (double)(sq * (float)wt);
      }
    }
  }
}

/* ---- csrc/mle-diag-gmm.cc:145-158 ------------------------------------- */
int32_t khg_oracle_acc_from_diag(int32_t nmix, int32_t dim, uint16_t flags,
                                 const float *gconsts,
                                 const float *means_invvars,
                                 const float *inv_vars, const float *x,
                                 float weight, double *occ, double *mean_acc,
                                 double *var_acc, float *log_like) {
  flags = khg_oracle_augment_flags(flags);
  float *post = (float *)malloc(sizeof(float) * nmix);
  int32_t bad = khg_oracle_component_posteriors(nmix, dim, gconsts, means_invvars,
                                                inv_vars, x, post, log_like);
  for (int32_t g = 0; g < nmix; ++g) post[g] *= weight; /* :153 */
  khg_oracle_acc_from_posteriors(nmix, dim, flags, x, post, occ, mean_acc, var_acc);
  free(post);
  return bad;
}

/* ---- scripts/gmm_acc_stats_ali.py:46-56 -> csrc/mle-am-diag-gmm.cc:41-52 */
static int64_t acc_stats_range(int32_t dim, const int32_t *offsets,
                               const float *gconsts, const float *miv,
                               const float *iv, uint16_t flags,
                               const float *feats, int64_t t0, int64_t t1,
                               const int32_t *pdf_ids, const float *fw,
                               double *occ, double *mean_acc, double *var_acc,
                               double *totals, float *per_frame_ll,
                               int32_t max_g) {
  int64_t bad = 0;
  float *ll = (float *)malloc(sizeof(float) * (size_t)(max_g > 0 ? max_g : 1));
  float *post = (float *)malloc(sizeof(float) * (size_t)(max_g > 0 ? max_g : 1));
  float *xsq = (float *)malloc(sizeof(float) * dim);
  double tot_like = 0.0, tot_frames = 0.0;
  for (int64_t t = t0; t < t1; ++t) {
    int32_t p = pdf_ids[t];
    int32_t g0 = offsets[p], ng = offsets[p + 1] - g0;
    const float *x = feats + t * dim;
    float w = fw ? fw[t] : 1.0f;
    for (int32_t d = 0; d < dim; ++d) xsq[d] = x[d] * x[d];
    loglikes_sq(ng, dim, gconsts + g0, miv + (size_t)g0 * dim, iv + (size_t)g0 * dim, x, xsq, ll);
    float log_like;
    khg_oracle_softmax(ll, ng, post, &log_like);
    if (isnan(log_like) || isinf(log_like)) bad++;
    for (int32_t g = 0; g < ng; ++g) post[g] *= w; /* mle-diag-gmm.cc:153 */
    acc_from_post_sq(ng, dim, flags, x, xsq, post, occ + g0,
                     mean_acc ? mean_acc + (size_t)g0 * dim : NULL,
                     var_acc ? var_acc + (size_t)g0 * dim : NULL);
    tot_like += (double)(log_like * w); /* mle-am-diag-gmm.cc:49, fp32 product */
    tot_frames += (double)w;            /* :50 */
    if (per_frame_ll) per_frame_ll[t] = log_like;
  }
  totals[0] += tot_like;
  totals[1] += tot_frames;
  free(ll);
  free(post);
  free(xsq);
  return bad;
}

static int32_t max_gauss(int32_t num_pdfs, const int32_t *offsets) {
  int32_t m = 0;
  for (int32_t p = 0; p < num_pdfs; ++p)
    if (offsets[p + 1] - offsets[p] > m) m = offsets[p + 1] - offsets[p];
  return m;
}

int64_t khg_oracle_acc_stats_ali(int32_t dim, int32_t num_pdfs,
                                 const int32_t *offsets, const float *gconsts,
                                 const float *means_invvars,
                                 const float *inv_vars, uint16_t flags,
                                 const float *feats, int64_t T,
                                 const int32_t *pdf_ids,
                                 const float *frame_weights, double *occ,
                                 double *mean_acc, double *var_acc,
                                 double *totals, float *per_frame_ll) {
  flags = khg_oracle_augment_flags(flags);
  return acc_stats_range(dim, offsets, gconsts, means_invvars, inv_vars, flags,
                         feats, 0, T, pdf_ids, frame_weights, occ, mean_acc,
                         var_acc, totals, per_frame_ll,
                         max_gauss(num_pdfs, offsets));
}

int64_t khg_oracle_acc_stats_ali_mt(int32_t dim, int32_t num_pdfs,
                                    const int32_t *offsets,
                                    const float *gconsts,
                                    const float *means_invvars,
                                    const float *inv_vars, uint16_t flags,
                                    const float *feats, int64_t T,
                                    const int32_t *pdf_ids,
                                    const float *frame_weights, double *occ,
                                    double *mean_acc, double *var_acc,
                                    double *totals, int32_t threads) {
  flags = khg_oracle_augment_flags(flags);
  if (threads < 1) threads = 1;
  int32_t G = offsets[num_pdfs];
  int32_t mg = max_gauss(num_pdfs, offsets);
  int64_t bad = 0;
  size_t nocc = (size_t)G, nmat = (size_t)G * dim;
  double *po = (double *)calloc(nocc * threads, sizeof(double));
  double *pm = mean_acc ? (double *)calloc(nmat * threads, sizeof(double)) : NULL;
  double *pv = var_acc ? (double *)calloc(nmat * threads, sizeof(double)) : NULL;
  double *pt = (double *)calloc(2 * (size_t)threads, sizeof(double));
#pragma omp parallel for num_threads(threads) reduction(+ : bad) schedule(static, 1)
  for (int32_t j = 0; j < threads; ++j) {
    int64_t t0 = T * j / threads, t1 = T * (j + 1) / threads;
    bad += acc_stats_range(dim, offsets, gconsts, means_invvars, inv_vars, flags,
                           feats, t0, t1, pdf_ids, frame_weights, po + nocc * j,
                           pm ? pm + nmat * j : NULL, pv ? pv + nmat * j : NULL,
                           pt + 2 * j, NULL, mg);
  }
  /* AccumAmDiagGmm::Add(1.0, other), csrc/mle-am-diag-gmm.cc:119-128 */
  for (int32_t j = 0; j < threads; ++j) {
    for (size_t i = 0; i < nocc; ++i) occ[i] += po[nocc * j + i];
    if (pm) for (size_t i = 0; i < nmat; ++i) mean_acc[i] += pm[nmat * j + i];
    if (pv) for (size_t i = 0; i < nmat; ++i) var_acc[i] += pv[nmat * j + i];
    totals[0] += pt[2 * j];
    totals[1] += pt[2 * j + 1];
  }
  free(po);
  free(pm);
  free(pv);
  free(pt);
  return bad;
}

/* ---- csrc/decodable-am-diag-gmm.cc:29-71 for every (frame, pdf) -------- */
int64_t khg_oracle_loglikes_all_pdfs(int32_t dim, int32_t num_pdfs,
                                     const int32_t *offsets,
                                     const float *gconsts,
                                     const float *means_invvars,
                                     const float *inv_vars, const float *feats,
                                     int64_t T, float scale, int32_t pdf_major,
                                     float *out, int32_t threads) {
  int32_t mg = max_gauss(num_pdfs, offsets);
  int64_t bad = 0;
  if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(+ : bad)
  {
    float *ll = (float *)malloc(sizeof(float) * (size_t)(mg > 0 ? mg : 1));
    float *xsq = (float *)malloc(sizeof(float) * dim);
#pragma omp for schedule(static)
    for (int64_t t = 0; t < T; ++t) {
      const float *x = feats + t * dim;
      for (int32_t d = 0; d < dim; ++d) xsq[d] = x[d] * x[d];
      for (int32_t p = 0; p < num_pdfs; ++p) {
        int32_t g0 = offsets[p], ng = offsets[p + 1] - g0;
        loglikes_sq(ng, dim, gconsts + g0, means_invvars + (size_t)g0 * dim,
                    inv_vars + (size_t)g0 * dim, x, xsq, ll);
        float s = khg_oracle_logsumexp(ll, ng);
        if (isnan(s) || isinf(s)) bad++;
        s = scale * s; /* decodable-am-diag-gmm.h:95-98 */
        if (pdf_major)
          out[(size_t)p * T + t] = s;
        else
          out[(size_t)t * num_pdfs + p] = s;
      }
    }
    free(ll);
    free(xsq);
  }
  return bad;
}

/* ---- csrc/diag-gmm.cc:177-189 (LogLikelihoodsMatrix) for every pdf ------
 * The matrix form of the same arithmetic: a block of FB frames at a time,
 *   loglikes(FB x nmix) = data * means_invvars^T - 0.5 * data^2 * inv_vars^T + gconsts,
 * i.e. the frame-blocked GEMM the reference executes when it is handed a feature
 * matrix, so that the model streams through the cache once per FB frames instead
 * of once per frame.  This is the CPU *baseline* form (bench.py); the per-frame
 * function above stays the parity form.  The exponentials are a vectorisable
 * Cephes-style expf (what Eigen's packet exp is), the sums run over frames so the
 * compiler keeps them in SIMD registers.  */
#define KHG_FB 64
typedef float khg_v8 __attribute__((vector_size(32), aligned(32)));
typedef int32_t khg_i8 __attribute__((vector_size(32), aligned(32)));
#define KHG_V8(c) ((khg_v8){c, c, c, c, c, c, c, c})
static inline khg_v8 khg_max_v8(khg_v8 a, khg_v8 b) {
  khg_i8 m = a > b, ai, bi;
  memcpy(&ai, &a, sizeof(ai));
  memcpy(&bi, &b, sizeof(bi));
  khg_i8 r = (ai & m) | (bi & ~m);
  khg_v8 out;
  memcpy(&out, &r, sizeof(out));
  return out;
}
static inline khg_v8 khg_exp_v8(khg_v8 x) {
  /* exp(x) for x <= 0 (max-subtracted): x = n ln2 + r, degree-5 polynomial, 2^n by exponent bits */
  khg_v8 lo = KHG_V8(-87.0f);
  x = khg_max_v8(x, lo);
  khg_v8 t = x * KHG_V8(1.44269504088896341f) - KHG_V8(0.5f);
  khg_i8 n = __builtin_convertvector(t, khg_i8); /* truncation: round to nearest for non-positive arguments */
  khg_v8 fn = __builtin_convertvector(n, khg_v8);
  khg_v8 r = x - fn * KHG_V8(0.693359375f);
  r = r + fn * KHG_V8(2.12194440e-4f);
  khg_v8 p = KHG_V8(1.9875691500e-4f);
  p = p * r + KHG_V8(1.3981999507e-3f);
  p = p * r + KHG_V8(8.3334519073e-3f);
  p = p * r + KHG_V8(4.1665795894e-2f);
  p = p * r + KHG_V8(1.6666665459e-1f);
  p = p * r + KHG_V8(5.0000001201e-1f);
  p = p * r * r + r + KHG_V8(1.0f);
  khg_i8 e = (n + 127) << 23;
  khg_v8 scale;
  memcpy(&scale, &e, sizeof(scale));
  return p * scale;
}

int64_t khg_oracle_loglikes_all_pdfs_blocked(int32_t dim, int32_t num_pdfs,
                                             const int32_t *offsets,
                                             const float *gconsts,
                                             const float *means_invvars,
                                             const float *inv_vars,
                                             const float *feats, int64_t T,
                                             float scale, int32_t pdf_major,
                                             float *out, int32_t threads) {
  int32_t mg = max_gauss(num_pdfs, offsets);
  int64_t bad = 0;
  if (threads < 1) threads = 1;
  int64_t n_blocks = (T + KHG_FB - 1) / KHG_FB;
#pragma omp parallel num_threads(threads) reduction(+ : bad)
  {
    float *xt = (float *)aligned_alloc(64, sizeof(float) * (size_t)dim * KHG_FB);
    float *xs = (float *)aligned_alloc(64, sizeof(float) * (size_t)dim * KHG_FB);
    float *ll = (float *)aligned_alloc(64, sizeof(float) * (size_t)(mg > 0 ? mg : 1) * KHG_FB);
#pragma omp for schedule(dynamic, 1)
    for (int64_t b = 0; b < n_blocks; ++b) {
      int64_t t0 = b * KHG_FB;
      int32_t nf = (int32_t)(T - t0 < KHG_FB ? T - t0 : KHG_FB);
      for (int32_t d = 0; d < dim; ++d)
        for (int32_t f = 0; f < KHG_FB; ++f) {
          float x = f < nf ? feats[(size_t)(t0 + f) * dim + d] : 0.0f;
          xt[d * KHG_FB + f] = x;
          xs[d * KHG_FB + f] = x * x;
        }
      for (int32_t p = 0; p < num_pdfs; ++p) {
        int32_t g0 = offsets[p], ng = offsets[p + 1] - g0;
        for (int32_t g = 0; g < ng; ++g) {
          const float *mv = means_invvars + (size_t)(g0 + g) * dim;
          const float *iv = inv_vars + (size_t)(g0 + g) * dim;
          /* 8 packet accumulators (64 frames) stay in registers across the D loop */
          khg_v8 acc[KHG_FB / 8];
          float gc = gconsts[g0 + g];
          for (int32_t i = 0; i < KHG_FB / 8; ++i) acc[i] = (khg_v8){gc, gc, gc, gc, gc, gc, gc, gc};
          for (int32_t d = 0; d < dim; ++d) {
            float a1 = mv[d], h1 = -0.5f * iv[d];
            khg_v8 a = {a1, a1, a1, a1, a1, a1, a1, a1}, h = {h1, h1, h1, h1, h1, h1, h1, h1};
            const khg_v8 *xr = (const khg_v8 *)(xt + d * KHG_FB), *sr = (const khg_v8 *)(xs + d * KHG_FB);
            for (int32_t i = 0; i < KHG_FB / 8; ++i) {
              acc[i] += a * xr[i];
              acc[i] += h * sr[i];
            }
          }
          for (int32_t i = 0; i < KHG_FB / 8; ++i) ((khg_v8 *)ll)[g * (KHG_FB / 8) + i] = acc[i];
        }
        /* LogSumExp over the pdf's Gaussians, csrc/eigen.cc:14-18, for FB frames at once */
        khg_v8 mxv[KHG_FB / 8], smv[KHG_FB / 8];
        const khg_v8 *lv = (const khg_v8 *)ll;
        for (int32_t i = 0; i < KHG_FB / 8; ++i) mxv[i] = lv[i];
        for (int32_t g = 1; g < ng; ++g)
          for (int32_t i = 0; i < KHG_FB / 8; ++i) {
            khg_v8 l = lv[g * (KHG_FB / 8) + i];
            mxv[i] = khg_max_v8(l, mxv[i]);
          }
        for (int32_t i = 0; i < KHG_FB / 8; ++i) smv[i] = KHG_V8(0.0f);
        for (int32_t g = 0; g < ng; ++g)
          for (int32_t i = 0; i < KHG_FB / 8; ++i) smv[i] += khg_exp_v8(lv[g * (KHG_FB / 8) + i] - mxv[i]);
        const float *mx = (const float *)mxv, *sm = (const float *)smv;
        for (int32_t f = 0; f < nf; ++f) {
          float s = logf(sm[f]) + mx[f];
          if (isnan(s) || isinf(s)) bad++;
          s = scale * s;
          if (pdf_major)
            out[(size_t)p * T + t0 + f] = s;
          else
            out[(size_t)(t0 + f) * num_pdfs + p] = s;
        }
      }
    }
    free(xt);
    free(xs);
    free(ll);
  }
  return bad;
}

/* ---- csrc/mle-diag-gmm.cc:479-499 ------------------------------------- */
float khg_oracle_ml_objective(int32_t nmix, int32_t dim, uint16_t acc_flags,
                              const float *gconsts, const float *means_invvars,
                              const float *inv_vars, const double *occ,
                              const double *mean_acc, const double *var_acc) {
  double o = 0.0;
  for (int32_t g = 0; g < nmix; ++g) o += occ[g] * (double)gconsts[g];
  float obj = (float)o; /* :482 float obj = double dot */
  if (acc_flags & KHG_ORACLE_GMM_MEANS) {
    double s = 0.0;
    for (size_t i = 0; i < (size_t)nmix * dim; ++i) s += mean_acc[i] * (double)means_invvars[i];
    obj = (float)((double)obj + s); /* :486-488 obj += double */
  }
  if (acc_flags & KHG_ORACLE_GMM_VARIANCES) {
    double s = 0.0;
    for (size_t i = 0; i < (size_t)nmix * dim; ++i) s += var_acc[i] * (double)inv_vars[i];
    obj = (float)((double)obj - 0.5 * s); /* :493-495 */
  }
  return obj;
}

/* ---- csrc/mle-diag-gmm.cc:243-390 ------------------------------------- */
int32_t khg_oracle_mle_update(const khg_oracle_mle_opts *opts, int32_t nmix,
                              int32_t dim, uint16_t acc_flags,
                              uint16_t update_flags, const double *occ,
                              const double *mean_acc, const double *var_acc,
                              float *weights, float *means_invvars,
                              float *inv_vars, float *gconsts,
                              int32_t *nmix_out, float *obj_change_out,
                              float *count_out, int32_t *floored_elements_out,
                              int32_t *floored_gaussians_out,
                              int32_t *removed_gaussians_out) {
  if (update_flags & ~acc_flags) return -2; /* :253-255 KHG_ERR */
  size_t nd = (size_t)nmix * dim;
  double occ_sum = 0.0;
  for (int32_t g = 0; g < nmix; ++g) occ_sum += occ[g];
  int32_t elements_floored = 0, gauss_floored = 0;

  if (khg_oracle_compute_gconsts(nmix, dim, weights, means_invvars, inv_vars, gconsts) < 0)
    return -1;
  float obj_old = khg_oracle_ml_objective(nmix, dim, acc_flags, gconsts, means_invvars,
                                          inv_vars, occ, mean_acc, var_acc);

  /* DiagGmmNormal ngmm(*gmm): csrc/diag-gmm-normal.cc:14-20 */
  double *nw = (double *)malloc(sizeof(double) * nmix);
  double *nvars = (double *)malloc(sizeof(double) * nd);
  double *nmeans = (double *)malloc(sizeof(double) * nd);
  for (int32_t g = 0; g < nmix; ++g) nw[g] = (double)weights[g];
  for (size_t i = 0; i < nd; ++i) {
    nvars[i] = 1.0 / (double)inv_vars[i]; /* 1.0f / cast<double> -> double */
    nmeans[i] = (double)means_invvars[i] * nvars[i];
  }
  int32_t *to_remove = (int32_t *)malloc(sizeof(int32_t) * nmix);
  int32_t n_remove = 0;
  double *old_mean = (double *)malloc(sizeof(double) * dim);
  double *var = (double *)malloc(sizeof(double) * dim);

  for (int32_t i = 0; i < nmix; ++i) {
    double o = occ[i];
    double prob = occ_sum > 0.0 ? o / occ_sum : 1.0 / nmix;
    if (o > (double)opts->min_gaussian_occupancy && prob > (double)opts->min_gaussian_weight) {
      nw[i] = prob;
      for (int32_t d = 0; d < dim; ++d) old_mean[d] = nmeans[(size_t)i * dim + d];
      if (acc_flags & (KHG_ORACLE_GMM_MEANS | KHG_ORACLE_GMM_VARIANCES))
        for (int32_t d = 0; d < dim; ++d) nmeans[(size_t)i * dim + d] = mean_acc[(size_t)i * dim + d] / o;
      if (acc_flags & KHG_ORACLE_GMM_VARIANCES) {
        for (int32_t d = 0; d < dim; ++d) {
          double m = nmeans[(size_t)i * dim + d];
          var[d] = var_acc[(size_t)i * dim + d] / o - m * m;
        }
        if (!(update_flags & KHG_ORACLE_GMM_MEANS)) { /* :300-304 */
          for (int32_t d = 0; d < dim; ++d) {
            double dm = old_mean[d] - nmeans[(size_t)i * dim + d];
            var[d] += dm * dm;
          }
        }
        int32_t floored = 0;
        for (int32_t d = 0; d < dim; ++d) { /* :320-327 (no floor vector in the bindings) */
          if (var[d] < opts->min_variance) {
            var[d] = opts->min_variance;
            ++floored;
          }
        }
        if (floored) {
          elements_floored += floored;
          ++gauss_floored;
        }
        for (int32_t d = 0; d < dim; ++d) nvars[(size_t)i * dim + d] = var[d];
      }
    } else { /* :337-358 */
      if (opts->remove_low_count_gaussians && n_remove < nmix - 1) {
        to_remove[n_remove++] = i;
      } else {
        double mw = (double)opts->min_gaussian_weight;
        nw[i] = prob > mw ? prob : mw;
      }
    }
  }

  /* ngmm.CopyToDiagGmm(gmm, flags): csrc/diag-gmm-normal.cc:22-48 */
  if (update_flags & KHG_ORACLE_GMM_WEIGHTS)
    for (int32_t g = 0; g < nmix; ++g) weights[g] = (float)nw[g];
  if (update_flags & KHG_ORACLE_GMM_VARIANCES) {
    /* oldg means are needed only when means are not updated */
    if (!(update_flags & KHG_ORACLE_GMM_MEANS)) {
      for (size_t i = 0; i < nd; ++i) {
        double oldvar = 1.0 / (double)inv_vars[i];
        double oldmean = (double)means_invvars[i] * oldvar;
        float niv = (float)(1.0 / nvars[i]);
        means_invvars[i] = (float)oldmean * niv;
        inv_vars[i] = niv;
      }
    } else {
      for (size_t i = 0; i < nd; ++i) inv_vars[i] = (float)(1.0 / nvars[i]);
    }
  }
  if (update_flags & KHG_ORACLE_GMM_MEANS)
    for (size_t i = 0; i < nd; ++i) means_invvars[i] = (float)nmeans[i] * inv_vars[i];

  if (khg_oracle_compute_gconsts(nmix, dim, weights, means_invvars, inv_vars, gconsts) < 0) {
    free(nw); free(nvars); free(nmeans); free(to_remove); free(old_mean); free(var);
    return -1;
  }
  float obj_new = khg_oracle_ml_objective(nmix, dim, acc_flags, gconsts, means_invvars,
                                          inv_vars, occ, mean_acc, var_acc);
  if (obj_change_out) *obj_change_out = obj_new - obj_old;
  if (count_out) *count_out = (float)occ_sum;
  if (floored_elements_out) *floored_elements_out = elements_floored;
  if (floored_gaussians_out) *floored_gaussians_out = gauss_floored;

  /* RemoveComponents(to_remove, true): csrc/diag-gmm.cc:853-937 — one at a
   * time, renormalising the fp32 weights after each removal. */
  int32_t cur = nmix;
  for (int32_t r = 0; r < n_remove; ++r) {
    int32_t g = to_remove[r] - r;
    memmove(weights + g, weights + g + 1, sizeof(float) * (size_t)(cur - g - 1));
    memmove(means_invvars + (size_t)g * dim, means_invvars + (size_t)(g + 1) * dim,
            sizeof(float) * (size_t)(cur - g - 1) * dim);
    memmove(inv_vars + (size_t)g * dim, inv_vars + (size_t)(g + 1) * dim,
            sizeof(float) * (size_t)(cur - g - 1) * dim);
    cur--;
    float s = 0.0f;
    for (int32_t k = 0; k < cur; ++k) s += weights[k];
    for (int32_t k = 0; k < cur; ++k) weights[k] /= s;
  }
  if (n_remove > 0)
    if (khg_oracle_compute_gconsts(cur, dim, weights, means_invvars, inv_vars, gconsts) < 0) {
      free(nw); free(nvars); free(nmeans); free(to_remove); free(old_mean); free(var);
      return -1;
    }
  if (nmix_out) *nmix_out = cur;
  if (removed_gaussians_out) *removed_gaussians_out = n_remove;
  free(nw); free(nvars); free(nmeans); free(to_remove); free(old_mean); free(var);
  return 0;
}
