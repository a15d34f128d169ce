/* oracle/khg_oracle.h
 *
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement ("oracle") of the diagonal-GMM E-step hot path of
 * csukuangfj/kaldi-hmm-gmm v1.1.4.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (libkhg_b200.so) never links, loads or calls it.
 *
 * The reference's own implementation cannot be compiled in this image: its
 * arithmetic lives in Eigen 3.4.0, an un-vendored dependency fetched at
 * configure time (reference cmake/eigen.cmake:4-6), and Eigen is absent here.
 * So this file restates the reference's algorithm from its call sites; every
 * function cites the reference file:line it follows.  Pinning status:
 *   - LogSumExp / Softmax: pinned by the reference's known answers
 *     (kaldi-hmm-gmm/csrc/eigen-test.cc:460-474, 641-654).
 *   - gconsts, log-likelihoods, posteriors, AccumDiagGmm accumulate_*: pinned
 *     by the reference's closed-form unit tests re-expressed in
 *     tests/test_oracle.py (python/tests/test_diag_gmm.py:45-51,327-403,529-576;
 *     python/tests/test_mle_diag_gmm.py:48-285).
 *   - AccumAmDiagGmm::AccumulateForGmm totals, DecodableAmDiagGmm*,
 *     MleAmDiagGmmUpdate: "parity unpinned" — the reference has no asserting
 *     test for them; they are pinned only through this oracle.
 *
 * Model layout used by the packed ("am") functions: the pdfs of an AmDiagGmm
 * are concatenated; pdf p owns Gaussians [offsets[p], offsets[p+1]).
 */
#ifndef KHG_ORACLE_H_
#define KHG_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference csrc/model-common.h:18-26 */
enum {
  KHG_ORACLE_GMM_MEANS = 1,
  KHG_ORACLE_GMM_VARIANCES = 2,
  KHG_ORACLE_GMM_WEIGHTS = 4,
  KHG_ORACLE_GMM_TRANSITIONS = 8,
  KHG_ORACLE_GMM_ALL = 15
};

/* csrc/eigen.cc:14-18 */
float khg_oracle_logsumexp(const float *v, int32_t n);
/* csrc/eigen.cc:20-32; log_sum_exp may be NULL */
void khg_oracle_softmax(const float *v, int32_t n, float *out,
                        float *log_sum_exp);

/* csrc/model-common.cc:72-84 */
uint16_t khg_oracle_augment_flags(uint16_t flags);

/* csrc/diag-gmm.cc:103-147. Returns num_bad (>=0) or -1 if a NaN gconst is met
 * (the reference throws there). */
int32_t khg_oracle_compute_gconsts(int32_t nmix, int32_t dim,
                                   const float *weights,
                                   const float *means_invvars,
                                   const float *inv_vars, float *gconsts);

/* csrc/diag-gmm.cc:167-176 */
void khg_oracle_loglikes(int32_t nmix, int32_t dim, const float *gconsts,
                         const float *means_invvars, const float *inv_vars,
                         const float *x, float *loglikes);

/* csrc/diag-gmm.cc:177-189: out is (T, nmix) row-major */
void khg_oracle_loglikes_matrix(int32_t nmix, int32_t dim, const float *gconsts,
                                const float *means_invvars,
                                const float *inv_vars, const float *feats,
                                int64_t T, float *out);

/* csrc/diag-gmm.cc:150-165. Returns 0, or 1 if the result is NaN/Inf (the
 * reference throws there). */
int32_t khg_oracle_log_likelihood(int32_t nmix, int32_t dim,
                                  const float *gconsts,
                                  const float *means_invvars,
                                  const float *inv_vars, const float *x,
                                  float *log_like);

/* csrc/diag-gmm.cc:368-392. Same return convention. */
int32_t khg_oracle_component_posteriors(int32_t nmix, int32_t dim,
                                        const float *gconsts,
                                        const float *means_invvars,
                                        const float *inv_vars, const float *x,
                                        float *posteriors, float *log_like);

/* csrc/mle-diag-gmm.cc:123-143. mean_acc / var_acc may be NULL when the
 * (augmented) flags lack m / v. */
void khg_oracle_acc_from_posteriors(int32_t nmix, int32_t dim, uint16_t flags,
                                    const float *x, const float *posteriors,
                                    double *occ, double *mean_acc,
                                    double *var_acc);

/* csrc/mle-diag-gmm.cc:100-121 */
void khg_oracle_acc_for_component(int32_t nmix, int32_t dim, uint16_t flags,
                                  const float *x, int32_t comp, float weight,
                                  double *occ, double *mean_acc,
                                  double *var_acc);

/* csrc/mle-diag-gmm.cc:145-158 */
int32_t khg_oracle_acc_from_diag(int32_t nmix, int32_t dim, uint16_t flags,
                                 const float *gconsts,
                                 const float *means_invvars,
                                 const float *inv_vars, const float *x,
                                 float weight, double *occ, double *mean_acc,
                                 double *var_acc, float *log_like);

/* The gmm-acc-stats-ali inner loop, scripts/gmm_acc_stats_ali.py:46-56 ->
 * csrc/mle-am-diag-gmm.cc:41-52, one frame at a time, in frame order.
 *   offsets[P+1]; packed model arrays of G = offsets[P] Gaussians;
 *   pdf_ids[T] (already mapped tid->pdf, csrc/transition-information.h:71-73);
 *   frame_weights may be NULL (=1);
 *   occ[G], mean_acc[G*D], var_acc[G*D] are accumulated in place (may be NULL
 *   per flags); totals[0] += sum ll*w, totals[1] += sum w; per_frame_ll may be
 *   NULL.  Returns the number of frames whose log-like was NaN/Inf (the
 *   reference throws at the first one). */
int64_t khg_oracle_acc_stats_ali(int32_t dim, int32_t num_pdfs,
                                 const int32_t *offsets, const float *gconsts,
                                 const float *means_invvars,
                                 const float *inv_vars, uint16_t flags,
                                 const float *feats, int64_t T,
                                 const int32_t *pdf_ids,
                                 const float *frame_weights, double *occ,
                                 double *mean_acc, double *var_acc,
                                 double *totals, float *per_frame_ll);

/* Same computation sharded over `threads` OpenMP threads with private
 * accumulators merged by Add (csrc/mle-am-diag-gmm.cc:119-128, Kaldi's
 * multi-job + gmm-sum-accs).  CPU-baseline only. */
int64_t khg_oracle_acc_stats_ali_mt(int32_t dim, int32_t num_pdfs,
                                    const int32_t *offsets,
                                    const float *gconsts,
                                    const float *means_invvars,
                                    const float *inv_vars, uint16_t flags,
                                    const float *feats, int64_t T,
                                    const int32_t *pdf_ids,
                                    const float *frame_weights, double *occ,
                                    double *mean_acc, double *var_acc,
                                    double *totals, int32_t threads);

/* All-pdf log-likelihoods: what DecodableAmDiagGmmUnmapped::
 * LogLikelihoodZeroBased (csrc/decodable-am-diag-gmm.cc:29-71) returns for
 * every (frame, pdf).  out is (T, P) row-major when pdf_major==0, else (P, T).
 * Returns the count of NaN/Inf entries.  threads<=1 -> single thread. */
int64_t khg_oracle_loglikes_all_pdfs(int32_t dim, int32_t num_pdfs,
                                     const int32_t *offsets,
                                     const float *gconsts,
                                     const float *means_invvars,
                                     const float *inv_vars, const float *feats,
                                     int64_t T, float scale, int32_t pdf_major,
                                     float *out, int32_t threads);
/* The same block in the frame-blocked matrix form of csrc/diag-gmm.cc:177-189
 * (LogLikelihoodsMatrix): the CPU baseline bench.py times; agrees with the per-frame
 * form to fp32 rounding (tests/test_oracle.py). */
int64_t khg_oracle_loglikes_all_pdfs_blocked(int32_t dim, int32_t num_pdfs,
                                     const int32_t *offsets,
                                     const float *gconsts,
                                     const float *means_invvars,
                                     const float *inv_vars, const float *feats,
                                     int64_t T, float scale, int32_t pdf_major,
                                     float *out, int32_t threads);

/* csrc/mle-diag-gmm.cc:479-499 */
float khg_oracle_ml_objective(int32_t nmix, int32_t dim, uint16_t acc_flags,
                              const float *gconsts, const float *means_invvars,
                              const float *inv_vars, const double *occ,
                              const double *mean_acc, const double *var_acc);

/* csrc/mle-diag-gmm.cc:243-390 (MleDiagGmmUpdate) for ONE pdf, including
 * csrc/diag-gmm-normal.cc:14-48 and RemoveComponents with renormalisation.
 * In/out: weights[nmix], means_invvars/inv_vars[nmix*dim] are updated in place
 * and compacted to *nmix_out Gaussians; gconsts recomputed.
 * min_variance etc. follow MleDiagGmmOptions (csrc/mle-diag-gmm.h:23-45). */
typedef struct {
  float min_gaussian_weight;
  float min_gaussian_occupancy;
  double min_variance;
  int32_t remove_low_count_gaussians;
} khg_oracle_mle_opts;

int32_t khg_oracle_mle_update(const khg_oracle_mle_opts *opts, int32_t nmix,
                              int32_t dim, uint16_t acc_flags,
                              uint16_t update_flags, const double *occ,
                              const double *mean_acc, const double *var_acc,
                              float *weights, float *means_invvars,
                              float *inv_vars, float *gconsts,
                              int32_t *nmix_out, float *obj_change_out,
                              float *count_out, int32_t *floored_elements_out,
                              int32_t *floored_gaussians_out,
                              int32_t *removed_gaussians_out);

#ifdef __cplusplus
}
#endif
#endif /* KHG_ORACLE_H_ */
