// kaldi-hmm-gmm_b200/csrc/host/khg_host.h
//
// C++ host-side mirror of the reference's classes on the E-step hot path, written
// above the C ABI (include/khg_b200.h).  Same class names, method names, argument
// meaning and error behaviour (std::runtime_error, reference csrc/log.h:46-53) as
//   DiagGmm                 csrc/diag-gmm.h
//   AmDiagGmm               csrc/am-diag-gmm.h
//   AccumDiagGmm            csrc/mle-diag-gmm.h
//   AccumAmDiagGmm          csrc/mle-am-diag-gmm.h
//   DecodableInterface      csrc/decodable-itf.h
//   DecodableAmDiagGmm*     csrc/decodable-am-diag-gmm.h
//   GmmUpdateFlags & co     csrc/model-common.h
// of csukuangfj/kaldi-hmm-gmm v1.1.4 (paths relative to kaldi-hmm-gmm/).  Storage is
// host-resident and authoritative (the Python bindings hand out live numpy views,
// python/csrc/diag-gmm.cc:41-45); every likelihood / posterior / accumulation is
// computed by the CUDA kernels behind the C ABI — there is no CPU arithmetic path
// for them.  No Eigen: matrices are row-major std::vector<float|double>.
#ifndef KHG_HOST_H_
#define KHG_HOST_H_

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "khg_b200.h"

namespace khg {

// ---- csrc/model-common.h:18-28 ------------------------------------------------
enum GmmUpdateFlags {
  kGmmMeans = 0x001,
  kGmmVariances = 0x002,
  kGmmWeights = 0x004,
  kGmmTransitions = 0x008,
  kGmmAll = 0x00F
};
typedef uint16_t GmmFlagsType;
GmmFlagsType AugmentGmmFlags(GmmFlagsType flags);         // csrc/model-common.cc:72-84
GmmFlagsType StringToGmmFlags(const std::string &str);    // csrc/model-common.cc:99-124
std::string GmmFlagsToString(GmmFlagsType flags);         // csrc/model-common.cc:126-146

[[noreturn]] void Throw(const std::string &msg);
void Check(khg_status s);  // throws std::runtime_error(khg_last_error()) on failure
#define KHG_HOST_ASSERT(cond)                                             \
  do {                                                                    \
    if (!(cond)) ::khg::Throw(std::string("Assertion failed: ") + #cond); \
  } while (0)

struct FloatMatrix {  // row-major, like the reference's FloatMatrix (csrc/eigen.h:14-16)
  int32_t rows = 0, cols = 0;
  std::vector<float> data;
  FloatMatrix() = default;
  FloatMatrix(int32_t r, int32_t c, float v = 0.f) : rows(r), cols(c), data((size_t)r * c, v) {}
  float &operator()(int32_t r, int32_t c) { return data[(size_t)r * cols + c]; }
  float operator()(int32_t r, int32_t c) const { return data[(size_t)r * cols + c]; }
  const float *row(int32_t r) const { return data.data() + (size_t)r * cols; }
  float *row(int32_t r) { return data.data() + (size_t)r * cols; }
  size_t size() const { return data.size(); }
};
struct DoubleMatrix {
  int32_t rows = 0, cols = 0;
  std::vector<double> data;
  DoubleMatrix() = default;
  DoubleMatrix(int32_t r, int32_t c) : rows(r), cols(c), data((size_t)r * c, 0.0) {}
  double *row(int32_t r) { return data.data() + (size_t)r * cols; }
  const double *row(int32_t r) const { return data.data() + (size_t)r * cols; }
  size_t size() const { return data.size(); }
};
typedef std::vector<float> FloatVector;
typedef std::vector<double> DoubleVector;

struct ModelHandle {  // RAII for khg_model*
  khg_model *h = nullptr;
  ~ModelHandle() { khg_model_destroy(h); }
};
struct StatsHandle {
  khg_stats *h = nullptr;
  ~StatsHandle() { khg_stats_destroy(h); }
};

// ---- csrc/diag-gmm.h ------------------------------------------------------------
class DiagGmm {
 public:
  DiagGmm() = default;
  DiagGmm(const DiagGmm &gmm) { CopyFromDiagGmm(gmm); }
  DiagGmm &operator=(const DiagGmm &gmm) { CopyFromDiagGmm(gmm); return *this; }
  DiagGmm(int32_t nmix, int32_t dim) { Resize(nmix, dim); }
  // csrc/diag-gmm.h:36-55: from parameters (used by pickle); computes gconsts
  DiagGmm(const FloatVector &weights, const FloatMatrix &inv_vars, const FloatMatrix &means_invvars);
  // csrc/diag-gmm.cc:66-101: merge several GMMs
  explicit DiagGmm(const std::vector<std::pair<float, const DiagGmm *>> &gmms);

  void Resize(int32_t nmix, int32_t dim);               // csrc/diag-gmm.cc:31-55
  void CopyFromDiagGmm(const DiagGmm &diaggmm);         // :57-64
  int32_t ComputeGconsts();                             // :103-147, on the device
  float LogLikelihood(const FloatVector &data) const;   // :150-165
  void LogLikelihoods(const FloatVector &data, FloatVector *loglikes) const;         // :167-176
  void LogLikelihoodsMatrix(const FloatMatrix &data, FloatMatrix *loglikes) const;   // :177-189
  void LogLikelihoodsPreselect(const FloatVector &data, const std::vector<int32_t> &indices,
                               FloatVector *loglikes) const;                         // :191-200
  // Gaussian selection on the device (khg_gaussian_selection): top num_gselect per frame
  float GaussianSelection(const FloatVector &data, int32_t num_gselect, std::vector<int32_t> *output) const;  // :202-239
  float GaussianSelection(const FloatMatrix &data, int32_t num_gselect,
                          std::vector<std::vector<int32_t>> *output) const;                                  // :241-317
  float GaussianSelectionPreselect(const FloatVector &data, const std::vector<int32_t> &preselect, int32_t num_gselect,
                                   std::vector<int32_t> *output) const;                                      // :319-366
  float ComponentPosteriors(const FloatVector &data, FloatVector *posterior) const;  // :368-392
  float ComponentLogLikelihood(const FloatVector &data, int32_t comp_id) const;      // :394-409

  void SetWeights(const FloatVector &w);                // :940-944
  void SetMeans(const FloatMatrix &m);                  // :946-952
  void SetInvVars(const FloatMatrix &v);                // :958-966
  void SetInvVarsAndMeans(const FloatMatrix &invvars, const FloatMatrix &means);  // :988-1000
  void SetComponentWeight(int32_t g, float w);          // :970-977
  void SetComponentMean(int32_t g, const FloatVector &v);     // :979-986
  void SetComponentInvVar(int32_t g, const FloatVector &v);   // :1002-1011
  FloatMatrix GetMeans() const;                         // :954-956
  FloatMatrix GetVars() const;                          // :968
  FloatVector GetComponentMean(int32_t gauss) const;    // :1013-1017
  FloatVector GetComponentVariance(int32_t gauss) const;  // :1019-1022
  void RemoveComponent(int32_t gauss, bool renorm_weights);   // :870-937
  void RemoveComponents(const std::vector<int32_t> &gauss, bool renorm_weights);  // :853-868

  int32_t NumGauss() const { return (int32_t)weights_.size(); }
  int32_t Dim() const { return means_invvars_.cols; }
  bool valid_gconsts() const { return valid_gconsts_; }
  const FloatVector &gconsts() const {  // csrc/diag-gmm.h:87-90 asserts validity
    if (!valid_gconsts_) Throw("Must call ComputeGconsts() before computing likelihood");
    return gconsts_;
  }
  const FloatVector &weights() const { return weights_; }
  FloatVector &weights() { return weights_; }  // live, mutable (python/csrc/diag-gmm.cc:41-45)
  const FloatMatrix &means_invvars() const { return means_invvars_; }
  const FloatMatrix &inv_vars() const { return inv_vars_; }

  uint64_t version() const { return version_; }
  // direct parameter replacement used by the M-step (DiagGmmNormal::CopyToDiagGmm)
  void SetParams(const FloatVector *w, const FloatMatrix *inv_vars, const FloatMatrix *means_invvars);
  // all parameters INCLUDING the gconsts the device computed for them (device M-step / mix-up results)
  void SetAllFromDevice(const float *w, const float *inv_vars, const float *means_invvars, const float *gconsts, int32_t nmix,
                        int32_t dim);

 private:
  khg_model *Device() const;  // 1-pdf device pack of this GMM, rebuilt when stale
  void Touch() { valid_gconsts_ = false; Bump(); }
  void Bump();
  FloatVector gconsts_, weights_;
  FloatMatrix inv_vars_, means_invvars_;
  bool valid_gconsts_ = false;
  uint64_t version_ = 0;
  mutable std::shared_ptr<ModelHandle> dev_;
  mutable uint64_t dev_version_ = ~0ull;
};

// ---- csrc/am-diag-gmm.h -----------------------------------------------------------
class AmDiagGmm {
 public:
  AmDiagGmm() = default;
  AmDiagGmm(const AmDiagGmm &) = delete;
  AmDiagGmm &operator=(const AmDiagGmm &) = delete;
  void Init(const DiagGmm &proto, int32_t num_pdfs);    // csrc/am-diag-gmm.cc:17-34
  void AddPdf(const DiagGmm &gmm);                      // :36-44 (deep copy)
  void CopyFromAmDiagGmm(const AmDiagGmm &other);       // :46-56
  int32_t Dim() const { return densities_.empty() ? 0 : densities_[0]->Dim(); }
  int32_t NumPdfs() const { return (int32_t)densities_.size(); }
  int32_t NumGauss() const;                             // :65-70
  int32_t NumGaussInPdf(int32_t pdf_index) const;       // :126-130
  int32_t ComputeGconsts() const;                       // :110-119 (const yet mutating, like the reference)
  float LogLikelihood(int32_t pdf_index, const FloatVector &data) const;  // :121-124
  DiagGmm &GetPdf(int32_t pdf_index);                   // :132-136
  const DiagGmm &GetPdf(int32_t pdf_index) const;       // :138-142
  FloatVector GetGaussianMean(int32_t pdf_index, int32_t gauss) const;
  FloatVector GetGaussianVariance(int32_t pdf_index, int32_t gauss) const;
  void SetGaussianMean(int32_t pdf_index, int32_t gauss_index, const FloatVector &in);
  // Mix-up on the device pack (khg_model_split_by_count), then the host pdfs are rebuilt from it.
  // randn: optional (rows x dim) standard-normal draws in the reference's order; seed otherwise.
  void SplitByCount(const FloatVector &state_occs, int32_t target_components, float perturb_factor, float power,
                    float min_count, const FloatMatrix *randn = nullptr, uint64_t seed = 0);  // csrc/am-diag-gmm.cc:72-89
  void MergeByCount(const FloatVector &state_occs, int32_t target_components, float power, float min_count);  // :91-108
  // Takes ownership of a device model produced FROM this model's pack (M-step, split, merge): the host
  // pdfs are rebuilt from it (all of them, or only those whose Gaussian count changed) and it becomes
  // the current device pack, so the next E-step neither re-uploads nor re-packs anything.
  void AdoptDevice(khg_model *nm, bool all_pdfs);

  // Device pack of the whole model (K4), rebuilt when any pdf changed.  Requires
  // valid gconsts on every pdf (csrc/decodable-am-diag-gmm.cc:49-53).
  khg_model *Device() const { return DeviceShared()->h; }
  std::shared_ptr<ModelHandle> DeviceShared() const;
  std::vector<int32_t> GaussOffsets() const;

 private:
  std::vector<std::unique_ptr<DiagGmm>> densities_;
  mutable std::shared_ptr<ModelHandle> dev_;
  mutable std::vector<std::pair<const DiagGmm *, uint64_t>> dev_sig_;
};

// ---- csrc/mle-diag-gmm.h ------------------------------------------------------------
struct MleDiagGmmOptions {  // csrc/mle-diag-gmm.h:23-45
  float min_gaussian_weight = 1.0e-05f;
  float min_gaussian_occupancy = 10.0f;
  double min_variance = 0.001;
  bool remove_low_count_gaussians = true;
  std::string ToString() const;
};

class AccumDiagGmm {
 public:
  AccumDiagGmm() = default;
  AccumDiagGmm(const DiagGmm &gmm, GmmFlagsType flags) { Resize(gmm.NumGauss(), gmm.Dim(), flags); }
  void Resize(int32_t num_gauss, int32_t dim, GmmFlagsType flags);  // csrc/mle-diag-gmm.cc:43-62
  void Resize(const DiagGmm &gmm, GmmFlagsType flags) { Resize(gmm.NumGauss(), gmm.Dim(), flags); }
  int32_t NumGauss() const { return num_comp_; }
  int32_t Dim() const { return dim_; }
  GmmFlagsType Flags() const { return flags_; }
  void SetZero(GmmFlagsType flags);                     // :64-80
  void Scale(float f, GmmFlagsType flags);              // :82-98
  void AccumulateForComponent(const FloatVector &data, int32_t comp_index, float weight);   // :100-121
  void AccumulateFromPosteriors(const FloatVector &data, const FloatVector &gauss_posteriors);  // :123-143
  float AccumulateFromDiag(const DiagGmm &gmm, const FloatVector &data, float weight);      // :145-158
  void AddStatsForComponent(int32_t g, double occ, const DoubleVector &x_stats, const DoubleVector &x2_stats);  // :160-174
  void Add(float scale, const AccumDiagGmm &acc);       // :176-188
  DoubleVector &occupancy() { return occupancy_; }
  const DoubleVector &occupancy() const { return occupancy_; }
  DoubleMatrix &mean_accumulator() { return mean_accumulator_; }
  const DoubleMatrix &mean_accumulator() const { return mean_accumulator_; }
  DoubleMatrix &variance_accumulator() { return variance_accumulator_; }
  const DoubleMatrix &variance_accumulator() const { return variance_accumulator_; }

 private:
  int32_t dim_ = 0, num_comp_ = 0;
  GmmFlagsType flags_ = 0;
  DoubleVector occupancy_;
  DoubleMatrix mean_accumulator_, variance_accumulator_;
};

// csrc/mle-diag-gmm.cc:243-390, :479-499 — the M-step consumes the stats on the host
// (SURVEY.md §8f row 1 moves it to the device in a later round).
void MleDiagGmmUpdate(const MleDiagGmmOptions &config, const AccumDiagGmm &diag_gmm_acc, GmmFlagsType flags,
                      DiagGmm *gmm, float *obj_change_out, float *count_out, int32_t *floored_elements_out = nullptr,
                      int32_t *floored_gauss_out = nullptr, int32_t *removed_gauss_out = nullptr);
float MlObjective(const DiagGmm &gmm, const AccumDiagGmm &diaggmm_acc);

// ---- csrc/mle-am-diag-gmm.h ----------------------------------------------------------
class AccumAmDiagGmm {
 public:
  AccumAmDiagGmm() = default;
  AccumAmDiagGmm(const AccumAmDiagGmm &) = delete;
  AccumAmDiagGmm &operator=(const AccumAmDiagGmm &) = delete;
  void Init(const AmDiagGmm &model, GmmFlagsType flags);               // csrc/mle-am-diag-gmm.cc:13-21
  void Init(const AmDiagGmm &model, int32_t dim, GmmFlagsType flags);  // :23-33
  void SetZero(GmmFlagsType flags);                                     // :35-39
  float AccumulateForGmm(const AmDiagGmm &model, const FloatVector &data, int32_t gmm_index, float weight);  // :41-52
  float AccumulateForGmmTwofeats(const AmDiagGmm &model, const FloatVector &data1, const FloatVector &data2,
                                 int32_t gmm_index, float weight);                                          // :54-76
  void AccumulateFromPosteriors(const AmDiagGmm &model, const FloatVector &data, int32_t gmm_index,
                                const FloatVector &posteriors);                                              // :78-86
  void AccumulateForGaussian(const AmDiagGmm &am, const FloatVector &data, int32_t gmm_index,
                             int32_t gauss_index, float weight);                                             // :88-97
  int32_t NumAccs() const { return (int32_t)gmm_accumulators_.size(); }
  float TotStatsCount() const;                                          // :99-106
  // csrc/mle-am-diag-gmm.h:72-76 (float!); the device part is read without moving the statistics
  float TotCount() const { double t[2]; DeviceTotals(t); return (float)(total_frames_ + t[1]); }
  float TotLogLike() const { double t[2]; DeviceTotals(t); return (float)(total_log_like_ + t[0]); }
  // per-pdf occupancy (sum over the pdf's Gaussians): what gmm-est hands to SplitByCount / MergeByCount
  // (scripts/gmm_est.py:66-73); only the occupancy vector leaves the device
  FloatVector PdfOccupancies() const;
  bool StatsOnDevice() const { return dev_ && dirty_ && !host_stats_; }
  const AccumDiagGmm &GetAcc(int32_t index) const;                      // :108-117
  AccumDiagGmm &GetAcc(int32_t index);
  void Add(float scale, const AccumAmDiagGmm &other);                   // :119-128
  void Scale(float scale);                                              // :130-138
  int32_t Dim() const { return gmm_accumulators_.empty() || !gmm_accumulators_[0] ? 0 : gmm_accumulators_[0]->Dim(); }

  // ---- batched entry points (new; what the unchanged script functions call) ----
  // Whole utterance(s) in one call: scripts/gmm_acc_stats_ali.py:46-56 without the
  // per-frame Python loop.  Returns sum over frames of the unweighted log-like * w.
  double AccumulateFrames(const AmDiagGmm &model, const float *feats, int64_t num_frames,
                          const int32_t *pdf_ids, const float *frame_weights);
  // From transition-ids; trans_accs (size num_tids+1) is updated in place (may be null).
  double AccumulateAlignment(const AmDiagGmm &model, const std::vector<int32_t> &tid2pdf, const float *feats,
                             int64_t num_frames, const int32_t *tids, double *trans_accs);
  // Device-resident stats of this accumulator (created on demand) — for the NCCL
  // all-reduce; Flush() folds them into the host accumulators.
  khg_stats *DeviceStats(const AmDiagGmm &model);
  void Flush() const;

 private:
  void EnsureDevice(const AmDiagGmm &model) const;
  void DeviceTotals(double tot[2]) const;  // {tot_like, tot_frames} still on the device (0 when clean)
  std::vector<std::unique_ptr<AccumDiagGmm>> gmm_accumulators_;
  mutable double total_frames_ = 0.0, total_log_like_ = 0.0;
  GmmFlagsType flags_ = 0;
  mutable std::shared_ptr<StatsHandle> dev_;
  mutable khg_model *dev_model_ = nullptr;
  mutable std::shared_ptr<ModelHandle> dev_model_keep_;
  mutable bool dirty_ = false;
  // the HOST accumulators may hold statistics (folded in by Flush, accumulated or edited on the host): the
  // device M-step then cannot consume the device buffer alone
  mutable bool host_stats_ = false;
  friend void MleAmDiagGmmUpdate(const MleDiagGmmOptions &, const AccumAmDiagGmm &, GmmFlagsType, AmDiagGmm *,
                                 float *, float *);
};

void MleAmDiagGmmUpdate(const MleDiagGmmOptions &config, const AccumAmDiagGmm &am_diag_gmm_acc, GmmFlagsType flags,
                        AmDiagGmm *am_gmm, float *obj_change_out, float *count_out);  // csrc/mle-am-diag-gmm.cc:153-202

// ---- csrc/decodable-itf.h:65-101 --------------------------------------------------------
class DecodableInterface {
 public:
  virtual ~DecodableInterface() = default;
  virtual float LogLikelihood(int32_t frame, int32_t index) = 0;
  virtual bool IsLastFrame(int32_t frame) const = 0;
  virtual int32_t NumFramesReady() const {
    Throw("NumFramesReady() not implemented for this decodable type.");
  }
  virtual int32_t NumIndices() const = 0;
};

// ---- csrc/decodable-am-diag-gmm.h --------------------------------------------------------
// The reference memoises one frame of per-pdf likelihoods (log_like_cache_, :74-78) and
// computes a GEMV pair per cache miss.  Here the whole (frames x pdfs) block is computed
// once, at construction, by the dense kernel; LogLikelihood is a table lookup, which is
// also what makes the retry pass of AlignUtteranceWrapper (csrc/decoder-wrappers.cc:55-67)
// free.  The model is therefore snapshotted at construction.
class DecodableAmDiagGmmUnmapped : public DecodableInterface {
 public:
  DecodableAmDiagGmmUnmapped(const AmDiagGmm &am, const FloatMatrix &feats, float log_sum_exp_prune = -1.0f);
  // From a precomputed pdf-major block (num_pdfs x num_frames), e.g. one utterance's slice of a
  // batched khg_loglikes_all_pdfs call over many utterances.
  DecodableAmDiagGmmUnmapped(std::vector<float> block, int32_t num_pdfs, int32_t num_frames)
      : num_frames_(num_frames), num_pdfs_(num_pdfs), log_sum_exp_prune_(-1.0f), block_(std::move(block)) {
    KHG_HOST_ASSERT((size_t)num_pdfs * num_frames == block_.size());
  }
  float LogLikelihood(int32_t frame, int32_t state_index) override {  // indices are one-based (:49-53)
    return LogLikelihoodZeroBased(frame, state_index - 1);
  }
  int32_t NumFramesReady() const override { return num_frames_; }
  int32_t NumIndices() const override { return num_pdfs_; }
  bool IsLastFrame(int32_t frame) const override {
    KHG_HOST_ASSERT(frame < NumFramesReady());
    return frame == NumFramesReady() - 1;
  }
  const std::vector<float> &LogLikeBlock() const { return block_; }  // pdf-major: [pdf][frame]

 protected:
  float LogLikelihoodZeroBased(int32_t frame, int32_t state) const;
  int32_t num_frames_ = 0, num_pdfs_ = 0;
  float log_sum_exp_prune_;  // accepted and never used, like the reference (:71)
  std::vector<float> block_;
};

class DecodableAmDiagGmmScaled : public DecodableAmDiagGmmUnmapped {
 public:
  // tid2pdf = TransitionModel::TransitionIdToPdfArray() (csrc/transition-information.h:71-84)
  DecodableAmDiagGmmScaled(const AmDiagGmm &am, std::vector<int32_t> tid2pdf, const FloatMatrix &feats, float scale,
                           float log_sum_exp_prune = -1.0f)
      : DecodableAmDiagGmmUnmapped(am, feats, log_sum_exp_prune), tid2pdf_(std::move(tid2pdf)), scale_(scale) {}
  DecodableAmDiagGmmScaled(std::vector<float> block, int32_t num_pdfs, int32_t num_frames, std::vector<int32_t> tid2pdf,
                           float scale)
      : DecodableAmDiagGmmUnmapped(std::move(block), num_pdfs, num_frames), tid2pdf_(std::move(tid2pdf)), scale_(scale) {}
  float LogLikelihood(int32_t frame, int32_t tid) override {  // :94-98
    KHG_HOST_ASSERT(tid >= 1 && tid < (int32_t)tid2pdf_.size());
    return scale_ * LogLikelihoodZeroBased(frame, tid2pdf_[tid]);
  }
  int32_t NumIndices() const override { return (int32_t)tid2pdf_.size() - 1; }

 private:
  std::vector<int32_t> tid2pdf_;
  float scale_;
};

}  // namespace khg
#endif  // KHG_HOST_H_
