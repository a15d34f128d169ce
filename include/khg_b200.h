/* include/khg_b200.h — C ABI of libkhg_b200.so
 *
 * B200-native (sm_100a) implementation of ONE hot path of
 * csukuangfj/kaldi-hmm-gmm v1.1.4: the E-step of diagonal-GMM acoustic-model
 * training (per-frame log-likelihoods + alignment-driven sufficient
 * statistics).  This header is the drop-in boundary: plain C types, caller-
 * owned buffers, opaque handles.  The reference has no FFI of its own (it is
 * C++ classes behind pybind11, python/csrc/kaldi-hmm-gmm.cc:35-68); each entry
 * point below names the reference C++ interface whose arithmetic it replaces
 * (paths relative to kaldi-hmm-gmm/ in the reference).  INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *  - Every function returns khg_status (0 = KHG_OK).  On failure
 *    khg_last_error() returns a thread-local message; the C++/pybind layer
 *    turns it into std::runtime_error, the reference's error convention
 *    (csrc/log.h:46-53).
 *  - There is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with KHG_ERR_CUDA.
 *  - `loc` arguments say where a caller buffer lives: KHG_HOST or KHG_DEVICE.
 *    Host buffers are copied inside the call; device buffers are used in place.
 *  - A call whose outputs are all device-resident is asynchronous on the
 *    handle's stream; data-dependent failures (a NaN/Inf log-likelihood, which
 *    makes the reference throw: csrc/diag-gmm.cc:160-162, 385-387,
 *    csrc/decodable-am-diag-gmm.cc:63-65) are latched in a device flag and
 *    reported by the next synchronising call or by khg_model_sync().
 *  - Handles are thread-compatible, not thread-safe, like the reference's
 *    classes.  One process drives one GPU (khg_set_device).
 *  - Matrices are row-major fp32 like the reference's FloatMatrix
 *    (csrc/eigen.h:10-22); statistics are fp64 (csrc/mle-diag-gmm.h:174-181).
 */
#ifndef KHG_B200_H_
#define KHG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t khg_status;
enum {
  KHG_OK = 0,
  KHG_ERR_INVALID = 1,    /* bad argument (the reference KHG_ASSERTs) */
  KHG_ERR_CUDA = 2,       /* CUDA runtime failure / no device */
  KHG_ERR_NONFINITE = 3,  /* NaN/Inf log-likelihood or NaN gconst */
  KHG_ERR_UNSUPPORTED = 4
};

enum { KHG_HOST = 0, KHG_DEVICE = 1 };

/* Layout of an all-pdf log-likelihood block. */
enum {
  KHG_FRAME_MAJOR = 0, /* out[t*ld + p]: what LogLikelihoodsMatrix-style callers expect */
  KHG_PDF_MAJOR = 1    /* out[p*ld + t]: native layout of the tensor-core kernel        */
};

/* GmmUpdateFlags, csrc/model-common.h:18-26 */
enum {
  KHG_GMM_MEANS = 0x001,
  KHG_GMM_VARIANCES = 0x002,
  KHG_GMM_WEIGHTS = 0x004,
  KHG_GMM_TRANSITIONS = 0x008,
  KHG_GMM_ALL = 0x00F
};

/* Which kernel computes the dense all-pdf log-likelihoods. */
enum {
  KHG_KERNEL_AUTO = 0,
  KHG_KERNEL_SIMT = 1,       /* fp32 FMA kernel (any dim, any pdf size) */
  KHG_KERNEL_TCGEN05 = 2,    /* tcgen05 kernel, 3xTF32 split (hi/lo tf32 operands) */
  KHG_KERNEL_TCGEN05_F16 = 3, /* tcgen05 kernel, 3xFP16 split: same 22-bit split precision at
                                twice the tensor rate; needs model and features in fp16 range
                                after a per-dimension power-of-two scaling.  AUTO uses it when
                                the model fits and, decided on the device for every call, the
                                features fit; otherwise the 3xTF32 split runs. */
  KHG_KERNEL_TCGEN05_F16_GS = 4 /* the same 3xFP16 arithmetic, Gaussian-stationary form: a 240-Gaussian model
                                tile stays in shared memory and the pre-split feature operand is streamed
                                (half the operand bytes per MMA; needs 2*dim+2 <= 128).  Same results; not
                                faster than 3 on B200, where the path is power-bound in the tensor pipe
                                 — kept selectable, never chosen by AUTO. */
};

typedef struct khg_model khg_model; /* device-resident packed AmDiagGmm      */
typedef struct khg_stats khg_stats; /* device-resident packed AccumAmDiagGmm */

const char *khg_last_error(void);
int32_t khg_abi_version(void); /* 4 (3 + khg_model_stats_kernel, khg_align_last_tile_fraction, khg_align_last_prep_cached) */

khg_status khg_device_count(int32_t *count);
khg_status khg_set_device(int32_t device);
/* AugmentGmmFlags, csrc/model-common.cc:72-84 */
uint16_t khg_augment_flags(uint16_t flags);

/* ---------------------------------------------------------------- model --
 * Replaces the storage of DiagGmm (csrc/diag-gmm.h:243-257) / AmDiagGmm
 * (csrc/am-diag-gmm.h:96): one contiguous pack of G = gauss_offsets[P]
 * Gaussians; pdf p owns [gauss_offsets[p], gauss_offsets[p+1]). */
khg_status khg_model_create(int32_t dim, int32_t num_pdfs,
                            const int32_t *gauss_offsets /* host, P+1 */,
                            khg_model **out);
/* Uploads parameters (host pointers).  gconsts==NULL -> computed on the device
 * from weights exactly as DiagGmm::ComputeGconsts (csrc/diag-gmm.cc:103-147):
 * *num_bad receives the count of +-inf gconsts; a NaN gconst fails with
 * KHG_ERR_NONFINITE.  weights may be NULL when gconsts is given. */
khg_status khg_model_upload(khg_model *m, const float *weights /* G */,
                            const float *means_invvars /* G x D */,
                            const float *inv_vars /* G x D */,
                            const float *gconsts /* G or NULL */,
                            int32_t *num_bad /* may be NULL */);
khg_status khg_model_get_gconsts(khg_model *m, float *gconsts /* host, G */);
khg_status khg_model_info(const khg_model *m, int32_t *dim, int32_t *num_pdfs,
                          int32_t *num_gauss);
/* Which dense kernel an in-range call will run with the current choice: KHG_KERNEL_SIMT,
 * KHG_KERNEL_TCGEN05, KHG_KERNEL_TCGEN05_F16 or KHG_KERNEL_TCGEN05_F16_GS (AUTO resolved against this model). */
khg_status khg_model_dense_kernel(const khg_model *m, int32_t *kernel);
/* Which kernel the bucketed statistics pass of khg_acc_stats_ali / khg_acc_stats_ali_tids / khg_estep (the per-frame
 * AccumAmDiagGmm::AccumulateForGmm, csrc/mle-am-diag-gmm.cc:41-52, of more than 2048 frames) runs for this model:
 * KHG_KERNEL_TCGEN05_F16 = the tensor-core kernel (dim <= 40, at least half of the Gaussians in pdfs of <= 32 — the items of larger pdfs run on the fp32 kernel —, parameters inside fp16's
 * range after a per-dimension power-of-two scaling; work items whose features or frame weights leave that range
 * are handed, on the device, to the fp32 kernel), KHG_KERNEL_SIMT = the fp32 kernel for everything.  Same
 * statistics either way (tests/test_gpu_stats_tc.py).  Builds the tensor-core operand pack on first use. */
khg_status khg_model_stats_kernel(khg_model *m, int32_t *kernel);
/* Chooses the dense-likelihood kernel (default KHG_KERNEL_AUTO). */
khg_status khg_model_set_kernel(khg_model *m, int32_t kernel);
/* Launches on this CUDA stream (a cudaStream_t; NULL = default stream). */
khg_status khg_model_set_stream(khg_model *m, void *cuda_stream);
/* Synchronises the stream and reports (then clears) a latched NaN/Inf. */
khg_status khg_model_sync(khg_model *m);
void khg_model_destroy(khg_model *m);

/* Stateless DiagGmm::ComputeGconsts (csrc/diag-gmm.cc:103-147), host buffers. */
khg_status khg_compute_gconsts(int32_t nmix, int32_t dim, const float *weights,
                               const float *means_invvars,
                               const float *inv_vars, float *gconsts,
                               int32_t *num_bad);

/* ---------------------------------------------------------- likelihoods --
 * All-pdf log-likelihood block: out(t,p) = scale * LogSumExp_g(loglike(t,g)),
 * g over pdf p.  Replaces, for every (frame, pdf) at once,
 * DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased
 * (csrc/decodable-am-diag-gmm.cc:29-71), DecodableAmDiagGmmScaled::
 * LogLikelihood (csrc/decodable-am-diag-gmm.h:94-98) and AmDiagGmm::
 * LogLikelihood (csrc/am-diag-gmm.cc:121-124).  feats is T x dim. */
khg_status khg_loglikes_all_pdfs(khg_model *m, const float *feats, int64_t T,
                                 int32_t feats_loc, float scale, int32_t layout,
                                 float *out, int64_t ld_out, int32_t out_loc);

/* Same block restricted to the pdfs an utterance's decoding graph can reach (the rows the
 * FST search will ever ask DecodableAmDiagGmmScaled::LogLikelihood for): all pdfs are
 * computed on the device, only the listed ones are gathered and returned, pdf-major:
 * out[i*ld_out + t] = scale * loglike(t, pdf_subset[i]).  pdf_subset is a HOST array.
 * This is what keeps the device->host traffic of a batched gmm-align-compiled small
 * (16.8 KB per frame for all 4200 pdfs vs 4 bytes per listed pdf). */
khg_status khg_loglikes_pdf_subset(khg_model *m, const float *feats, int64_t T,
                                   int32_t feats_loc, const int32_t *pdf_subset,
                                   int32_t n_subset, float scale, float *out,
                                   int64_t ld_out, int32_t out_loc);

/* Per-Gaussian log-likelihoods of ONE pdf: out is T x g_p (frame-major).
 * DiagGmm::LogLikelihoods (csrc/diag-gmm.cc:167-176) for T==1,
 * DiagGmm::LogLikelihoodsMatrix (:177-189) for T>1. */
khg_status khg_pdf_loglikes(khg_model *m, int32_t pdf, const float *feats,
                            int64_t T, int32_t loc, float *out);

/* Posteriors of ONE pdf for T frames: post is T x g_p, loglike is T.
 * DiagGmm::ComponentPosteriors (csrc/diag-gmm.cc:368-392) + Softmax
 * (csrc/eigen.cc:20-32); loglike alone = DiagGmm::LogLikelihood (:150-165).
 * post may be NULL.  NaN/Inf log-like -> KHG_ERR_NONFINITE. */
khg_status khg_pdf_posteriors(khg_model *m, int32_t pdf, const float *feats,
                              int64_t T, int32_t loc, float *post,
                              float *loglike);

/* ---------------------------------------------------------------- stats --
 * Replaces AccumDiagGmm storage (csrc/mle-diag-gmm.h:174-181) for all pdfs and
 * the totals of AccumAmDiagGmm (csrc/mle-am-diag-gmm.h:93-96).  One packed fp64
 * device buffer: [occ G | mean G*D (if m) | var G*D (if v) | tot_like |
 * tot_frames].  flags are augmented v=>m=>w (csrc/model-common.cc:72-84).
 * Lifetime: a khg_stats refers to its khg_model (stream, shapes) for as long as it lives — destroy the statistics
 * before the model they were created for (khg_mle_update returns a NEW model: the old model and its statistics
 * stay valid until the caller destroys them, statistics first).  Device views of the buffer
 * (khg_stats_device_buffer) die with the handle. */
khg_status khg_stats_create(khg_model *m, uint16_t flags, khg_stats **out);
khg_status khg_stats_zero(khg_stats *s);
khg_status khg_stats_flags(const khg_stats *s, uint16_t *flags);
/* Packed device buffer for the multi-GPU sum (NCCL all-reduce through
 * torch.distributed, or AccumAmDiagGmm::Add semantics on one device). */
khg_status khg_stats_device_buffer(khg_stats *s, double **dev_ptr,
                                   int64_t *num_doubles);
/* Host copies; any pointer may be NULL.  totals = {tot_like, tot_frames}. */
khg_status khg_stats_download(khg_stats *s, double *occ, double *mean,
                              double *var, double *totals);
khg_status khg_stats_upload(khg_stats *s, const double *occ, const double *mean,
                            const double *var, const double *totals);
/* dst += scale * src : AccumAmDiagGmm::Add (csrc/mle-am-diag-gmm.cc:119-128);
 * note the reference takes a float scale. */
khg_status khg_stats_add(khg_stats *dst, float scale, const khg_stats *src);
/* AccumAmDiagGmm::Scale (csrc/mle-am-diag-gmm.cc:130-138) */
khg_status khg_stats_scale(khg_stats *s, float scale);
/* Lifetime: a khg_stats keeps a pointer to the model it was created on (its stream,
 * its Gaussian layout).  Destroy the statistics BEFORE their model; a model handle
 * returned by khg_mle_update / split / merge is a different model and needs its own
 * khg_stats.  Views of the device buffer (khg_stats_device_buffer) die with the handle. */
void khg_stats_destroy(khg_stats *s);

/* The gmm-acc-stats-ali inner loop for T frames in one call:
 * for each frame t, AccumAmDiagGmm::AccumulateForGmm(model, feats[t],
 * pdf_ids[t], w[t]) (csrc/mle-am-diag-gmm.cc:41-52 -> csrc/mle-diag-gmm.cc:
 * 145-158 -> csrc/diag-gmm.cc:368-392 -> csrc/mle-diag-gmm.cc:123-143).
 * Frames are bucketed by pdf id on the device (exact integer sort), posteriors
 * and statistics are accumulated per pdf, totals += {sum ll*w, sum w}.
 *   pdf_ids        int32[T]  (same loc as feats)
 *   frame_weights  f32[T] or NULL (=1)         (same loc as feats)
 *   per_frame_loglike f32[T] or NULL: the UNWEIGHTED per-frame log-like
 *                  AccumulateForGmm returns              (same loc as feats)
 *   tot_loglike    host double or NULL: this call's sum ll*w (forces a sync)
 * A pdf id outside [0, num_pdfs) is the reference's KHG_ASSERT
 * (csrc/mle-am-diag-gmm.cc:44): the call — or, for asynchronous device-buffer
 * calls, the next synchronising call on the model — returns KHG_ERR_INVALID.
 * Such frames contribute NOTHING to the statistics (on every internal path); the
 * other frames of the call have been accumulated, so treat the stats handle as
 * poisoned after this error (zero it or destroy it), as the reference's abort does. */
khg_status khg_acc_stats_ali(khg_model *m, khg_stats *s, const float *feats,
                             int64_t T, int32_t loc, const int32_t *pdf_ids,
                             const float *frame_weights,
                             float *per_frame_loglike, double *tot_loglike);

/* Same, from transition-ids: pdf = tid2pdf[tid]
 * (TransitionModel::TransitionIdToPdf, csrc/transition-information.h:71-73) and
 * trans_accs[tid] += 1 (TransitionModel::Accumulate with prob 1,
 * csrc/transition-model.h:183-189) — the whole body of
 * scripts/gmm_acc_stats_ali.py:46-56.  All pointers are HOST pointers.
 *   tid2pdf int32[num_tids+1] (index 0 unused), trans_accs double[num_tids+1]
 *   (updated in place, may be NULL). */
khg_status khg_acc_stats_ali_tids(khg_model *m, khg_stats *s,
                                  const float *feats, int64_t T,
                                  const int32_t *tids, const int32_t *tid2pdf,
                                  int32_t num_tids, double *trans_accs,
                                  double *tot_loglike);

/* AccumDiagGmm::AccumulateFromPosteriors (csrc/mle-diag-gmm.cc:123-143) for T
 * frames of one pdf: post is T x g_p; totals[1] += sum(post)
 * (csrc/mle-am-diag-gmm.cc:78-86).  Host or device buffers per loc. */
khg_status khg_acc_from_posteriors(khg_model *m, khg_stats *s, int32_t pdf,
                                   const float *feats, int64_t T, int32_t loc,
                                   const float *post);

/* Whole E-step over one batch, as BASELINE.json's north_star words it:
 * dense all-pdf log-likelihoods (what gmm-align-compiled consumes) AND the
 * alignment-driven statistics (what gmm-acc-stats-ali produces).
 * loglikes_out: DEVICE buffer of at least num_pdfs * ld_out floats, pdf-major,
 * ld_out >= chunk_frames; it is reused chunk by chunk (the consumer is on the
 * device side of the boundary).  feats/pdf_ids/frame_weights per `loc`; with
 * KHG_HOST the call streams chunks through pinned staging buffers, copy
 * overlapped with compute.  chunk_frames<=0 picks a default.  The statistics
 * pass runs once per group of chunks (up to 16 M frames; with KHG_HOST the
 * frames of two groups are staged on the device: <= 2 x 2.7 GB at dim 40). */
khg_status khg_estep(khg_model *m, khg_stats *s, const float *feats, int64_t T,
                     int32_t loc, const int32_t *pdf_ids,
                     const float *frame_weights, float *loglikes_out,
                     int64_t ld_out, int64_t chunk_frames,
                     double *tot_loglike);

/* ---------------------------------------------------------------- M-step --
 * MleAmDiagGmmUpdate (csrc/mle-am-diag-gmm.cc:153-202) = MleDiagGmmUpdate
 * (csrc/mle-diag-gmm.cc:243-390) for every pdf, on the device, reading the packed
 * statistics in place: weights = occ/sum(occ), mean = x/occ, var = x2/occ - mean^2
 * floored at min_variance, Gaussians with occ <= min_gaussian_occupancy or weight <=
 * min_gaussian_weight removed (never the last one of a pdf) with the weights
 * renormalised, gconsts recomputed, objective change as MlObjective (:479-499).
 * Because removal changes the number of Gaussians, the result is a NEW model handle;
 * the old one stays valid.  update_flags is a GmmUpdateFlags subset of the stats' flags
 * (kGmmTransitions is ignored here). */
typedef struct {
  float min_gaussian_weight;     /* 1e-5  (csrc/mle-diag-gmm.h:23-45) */
  float min_gaussian_occupancy;  /* 10    */
  double min_variance;           /* 0.001 */
  int32_t remove_low_count_gaussians; /* 1 */
} khg_mle_options;

khg_status khg_mle_update(khg_model *m, const khg_stats *s, const khg_mle_options *opts,
                          uint16_t update_flags, khg_model **new_model,
                          float *obj_change, float *count,
                          int32_t *floored_elements, int32_t *floored_gaussians,
                          int32_t *removed_gaussians);

/* Copies the packed model back to the host (any pointer may be NULL):
 * gauss_offsets int32[P+1], weights/gconsts f32[G], means_invvars/inv_vars f32[G*D]. */
khg_status khg_model_download(khg_model *m, int32_t *gauss_offsets, float *weights,
                              float *means_invvars, float *inv_vars, float *gconsts);

/* ------------------------------------------------- batched forced alignment --
 * gmm-align-compiled for a batch of utterances on the device: for every
 * utterance, DecodableAmDiagGmmScaled (csrc/decodable-am-diag-gmm.h:81-109) over the dense
 * all-pdf block + FasterDecoder::Decode / ReachedFinal / GetBestPath
 * (csrc/faster-decoder.cc:125-228, 346-425) as driven by AlignUtteranceWrapper
 * (csrc/decoder-wrappers.cc:16-108): beam, then retry_beam if no final state was reached, with
 * the decoder options the wrapper leaves at their defaults (max_active = int max,
 * min_active = 20, beam_delta = 0.5, csrc/faster-decoder.h:41-43).
 *
 * The device search is the reference's frame-synchronous token passing restated as a pull-style
 * dynamic programme (one CTA per utterance, one token per graph state, costs in double like
 * Token::cost_): a state is expanded at frame t iff its cost < best_t + beam (or inside the
 * min_active cutoff), exactly the reference's rule.  Two things in the reference depend on the ORDER
 * of its token list (csrc/hash-list-inl.h): which token survives an exact cost tie, and — through the
 * running next_weight_cutoff of ProcessEmitting (csrc/faster-decoder.cc:196-216) — which tokens above
 * the frame's final cutoff are kept alive for one more frame.  The device search checks, frame by
 * frame, a certificate that neither can have influenced the result (khg_align.cu); the utterances for
 * which it cannot are re-aligned inside the same call by khg_align_utterance_host below, the exact
 * restatement of the reference's decoder, on the same likelihood block.  So every returned alignment
 * is the reference's, given the likelihoods.  KHG_ALIGN_EXACT=all / none (environment) sends every /
 * no utterance through the host decoder.
 *
 * Graphs: the compiled training graphs (fst::VectorFst<StdArc>, after AddTransitionProbs) as
 * plain arrays, all utterances concatenated.  States and arcs use LOCAL state ids per
 * utterance; arcs are sorted by source state.  `careful` alignment
 * (ModifyGraphForCarefulAlignment) is a graph edit the caller applies before export. */
typedef struct {
  int32_t n_utts;
  const int64_t *frame_offsets;  /* n_utts+1: rows of feats belonging to each utterance        */
  const int32_t *state_offsets;  /* n_utts+1: utterance u owns states [state_offsets[u],
                                    state_offsets[u+1]) of the concatenated state arrays       */
  const int32_t *arc_offsets;    /* total_states+1: CSR over ALL states, absolute arc indices  */
  const int32_t *arc_ilabel;     /* per arc: transition-id, 0 = epsilon                        */
  const int32_t *arc_nextstate;  /* per arc: LOCAL state id within its utterance               */
  const float *arc_weight;       /* per arc: tropical cost                                     */
  const int32_t *start_state;    /* n_utts, local; -1 = empty graph (kNoStateId)               */
  const float *final_cost;       /* per state; +inf = not final                                */
} khg_graph_batch;

enum { KHG_ALIGN_OK = 0, KHG_ALIGN_RETRIED = 1, KHG_ALIGN_FAILED = 2 };

/* feats: sum of frames x dim per `feats_loc`; tid2pdf: HOST int32[n_tids] (index 0 unused,
 * csrc/transition-information.h:71-73); graphs: HOST arrays.
 * Outputs (HOST, any may be NULL):
 *   alignment   int32 per frame: the transition-ids of the best path (0 for failed utterances)
 *   utt_status  int32 per utterance: KHG_ALIGN_OK / _RETRIED / _FAILED
 *   utt_like    float per utterance: -(graph cost + acoustic cost) / acoustic_scale
 *   path_arcs   int32, path_offsets[n_utts+1] (int64): absolute arc ids of the best path
 *               including epsilon arcs, for olabels ("words"); path_capacity entries
 * pdf_ids_dev: optional DEVICE int32 per frame = tid2pdf[alignment] (0 for the frames of failed
 * utterances; give those frames weight 0), so that khg_acc_stats_ali can consume the alignment
 * without leaving the device. */
khg_status khg_align_batch(khg_model *m, const khg_graph_batch *graphs, const float *feats,
                           int32_t feats_loc, const int32_t *tid2pdf, int32_t n_tids,
                           float acoustic_scale, float beam, float retry_beam,
                           int32_t *alignment, int32_t *utt_status, float *utt_like,
                           int32_t *path_arcs, int64_t *path_offsets, int64_t path_capacity,
                           int32_t *pdf_ids_dev);

/* Utterances of the most recent khg_align_batch call (this process) that were re-aligned by the exact
 * host decoder below because the device search could not certify its result (diagnostics). */
int64_t khg_align_last_exact_count(void);
/* Fraction of the (128-frame tile, 240-Gaussian model tile) units of the all-pdf likelihood block that the most recent
 * khg_align_batch call computed: the dense kernel skips model tiles that hold no pdf of the graphs of the frames'
 * utterances (the batched form of the reference's lazy per-pdf evaluation, csrc/decodable-am-diag-gmm.cc:29-71);
 * 1.0 when everything was computed (small batches, KHG_ALIGN_TILE_SUBSET=0, kernels without that mode). */
double khg_align_last_tile_fraction(void);
/* 1 when the most recent khg_align_batch call reused the graph preparation (transposed graphs, their device copy) of
 * the previous call on the same model: graphs, frame offsets and tid2pdf are compared by a 128-bit content hash — the
 * realignment passes of an EM recipe align the same graphs every time (egs/yesno/train.py:165-206).
 * KHG_ALIGN_PREP_CACHE=0 (environment) disables the reuse. */
int32_t khg_align_last_prep_cached(void);

/* The reference's FasterDecoder + AlignUtteranceWrapper on the HOST for ONE utterance of a graph batch,
 * consuming a block of log-likelihoods computed elsewhere (the GPU): csrc/faster-decoder.cc:36-425 with
 * its token container's visiting order (csrc/hash-list-inl.h:26-170: buckets in order of first use,
 * key % hash_size, the hash growing as PossiblyResizeHash does), i.e. including the order-dependent
 * running next_weight_cutoff of ProcessEmitting (:196-216) — the exact reference rule.  khg_align_batch
 * uses it for the utterances whose device search cannot be proven equal to it; no GPU is needed.
 *   loglikes  HOST float, rows x ld: SCALED log-likelihoods, row tid2row[tid] = the row of tid's pdf,
 *             column t = frame t of this utterance (what DecodableAmDiagGmmScaled::LogLikelihood(t, tid)
 *             returns, csrc/decodable-am-diag-gmm.h:94-98)
 *   tid2row   HOST int32[n_tids] (index 0 unused)
 * Outputs (HOST): alignment int32[T]; *status KHG_ALIGN_*; *like = -(graph + acoustic cost) /
 * acoustic_scale; path_arcs (absolute arc ids incl. epsilons, path_capacity entries) / *path_len may be NULL. */
khg_status khg_align_utterance_host(const khg_graph_batch *graphs, int32_t utt, const float *loglikes, int64_t ld,
                                    const int32_t *tid2row, int32_t n_tids, float acoustic_scale, float beam,
                                    float retry_beam, int32_t *alignment, int32_t *status, float *like,
                                    int32_t *path_arcs, int32_t path_capacity, int32_t *path_len);

/* ------------------------------------------------------------------ mix-up --
 * AmDiagGmm::SplitByCount (csrc/am-diag-gmm.cc:72-89) on the packed device model: the per-pdf
 * targets of GetSplitTargets (csrc/model-common.cc:29-70; state_occs = HOST float[num_pdfs],
 * power-law allocation with the min_count rule), then DiagGmm::Split (csrc/diag-gmm.cc:780-851) on
 * every pdf below its target, and ComputeGconsts.  Returns a NEW handle (the Gaussian count changes);
 * the old one stays valid.
 * randn: HOST float[randn_rows x dim], the standard-normal vector of each split in the order the
 * reference draws them (pdf by pdf, split by split) — pass the reference's draws to reproduce its
 * result; NULL = drawn inside from `seed`.  *num_gauss_out = Gaussians of the new model. */
khg_status khg_model_split_by_count(khg_model *m, const float *state_occs, int32_t target_components,
                                    float perturb_factor, float power, float min_count, const float *randn,
                                    int64_t randn_rows, uint64_t seed, khg_model **new_model,
                                    int32_t *num_gauss_out);

/* AmDiagGmm::MergeByCount (csrc/am-diag-gmm.cc:91-108): the same targets (at least 1 per pdf), then
 * DiagGmm::Merge (csrc/diag-gmm.cc:557-746) on every pdf ABOVE its target: greedy merging of the pair
 * of Gaussians whose union loses the least likelihood (target 1: the global mean and variance). */
khg_status khg_model_merge_by_count(khg_model *m, const float *state_occs, int32_t target_components, float power,
                                    float min_count, khg_model **new_model, int32_t *num_gauss_out);

/* -------------------------------------------------------- Gaussian selection --
 * DiagGmm::GaussianSelection for a matrix of frames (csrc/diag-gmm.cc:241-317; the one-frame form
 * :202-239 is T = 1) and DiagGmm::GaussianSelectionPreselect (:319-366) on pdf `pdf` of the model
 * (a UBM is a model with one pdf): per frame the k = min(num_gselect, n) candidates with the largest
 * log-likelihoods, best first — among equal log-likelihoods the larger index first, like the
 * reference's std::greater on (loglike, index) pairs — and the LogAdd chain of their log-likelihoods.
 * preselect (HOST, n_preselect entries, indices into the pdf's Gaussians, duplicates allowed)
 * restricts the candidates; n_preselect = 0 -> all n = NumGauss() of them.
 * Outputs (HOST): out_indices int32[T x k]; out_loglikes float[T x k] or NULL (the selected
 * components' log-likelihoods as the device computed them); frame_loglike float[T] or NULL;
 * *tot_loglike = sum over frames (the matrix form's return value), may be NULL. */
khg_status khg_gaussian_selection(khg_model *m, int32_t pdf, const float *feats, int64_t T, int32_t feats_loc,
                                  const int32_t *preselect, int32_t n_preselect, int32_t num_gselect,
                                  int32_t *out_indices, float *out_loglikes, float *frame_loglike,
                                  double *tot_loglike);

/* Number of kernels this library launched since load (bench.py's
 * gpu_launches). */
int64_t khg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* KHG_B200_H_ */
