// kaldi-hmm-gmm_b200/csrc/host/khg_host.cc — see khg_host.h.  Host bookkeeping in C++;
// every likelihood / posterior / accumulation goes through the C ABI to the CUDA
// kernels.  Reference citations: kaldi-hmm-gmm/ of csukuangfj/kaldi-hmm-gmm v1.1.4.
#include "khg_host.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <sstream>

namespace khg {

void Throw(const std::string &msg) { throw std::runtime_error(msg); }
void Check(khg_status s) {
  if (s != KHG_OK) Throw(khg_last_error());
}

// ---------------------------------------------------------------- flags --
GmmFlagsType AugmentGmmFlags(GmmFlagsType flags) {
  KHG_HOST_ASSERT((flags & ~kGmmAll) == 0);
  return khg_augment_flags(flags);
}

GmmFlagsType StringToGmmFlags(const std::string &str) {
  GmmFlagsType flags = 0;
  for (char c : str) {
    switch (c) {
      case 'm': flags |= kGmmMeans; break;
      case 'v': flags |= kGmmVariances; break;
      case 'w': flags |= kGmmWeights; break;
      case 't': flags |= kGmmTransitions; break;
      case 'a': flags |= kGmmAll; break;
      default:
        Throw(std::string("Invalid element '") + c + "' of GmmFlagsType option string " + str);
    }
  }
  return flags;
}

std::string GmmFlagsToString(GmmFlagsType flags) {
  std::string ans;
  if (flags & kGmmMeans) ans += "m";
  if (flags & kGmmVariances) ans += "v";
  if (flags & kGmmWeights) ans += "w";
  if (flags & kGmmTransitions) ans += "t";
  return ans;
}

// ---------------------------------------------------------------- DiagGmm --
static std::atomic<uint64_t> g_version{1};
void DiagGmm::Bump() { version_ = g_version.fetch_add(1); }

void DiagGmm::Resize(int32_t nmix, int32_t dim) {
  KHG_HOST_ASSERT(nmix > 0 && dim > 0);
  if ((int32_t)gconsts_.size() != nmix) gconsts_.assign(nmix, 0.f);
  if ((int32_t)weights_.size() != nmix) weights_.assign(nmix, 0.f);
  if (inv_vars_.rows != nmix || inv_vars_.cols != dim) inv_vars_ = FloatMatrix(nmix, dim, 1.0f);
  if (means_invvars_.rows != nmix || means_invvars_.cols != dim) means_invvars_ = FloatMatrix(nmix, dim, 0.f);
  Touch();
}

void DiagGmm::CopyFromDiagGmm(const DiagGmm &o) {
  gconsts_ = o.gconsts_;
  weights_ = o.weights_;
  inv_vars_ = o.inv_vars_;
  means_invvars_ = o.means_invvars_;
  valid_gconsts_ = o.valid_gconsts_;
  Bump();
}

DiagGmm::DiagGmm(const FloatVector &weights, const FloatMatrix &inv_vars, const FloatMatrix &means_invvars) {
  KHG_HOST_ASSERT((int32_t)weights.size() == inv_vars.rows && inv_vars.rows == means_invvars.rows &&
                  inv_vars.cols == means_invvars.cols);
  weights_ = weights;
  inv_vars_ = inv_vars;
  means_invvars_ = means_invvars;
  gconsts_.assign(weights.size(), 0.f);
  Touch();
  ComputeGconsts();
}

DiagGmm::DiagGmm(const std::vector<std::pair<float, const DiagGmm *>> &gmms) {
  if (gmms.empty()) return;
  int32_t num_gauss = 0, dim = gmms[0].second->Dim();
  for (auto &g : gmms) num_gauss += g.second->NumGauss();
  Resize(num_gauss, dim);
  int32_t cur = 0;
  for (auto &pr : gmms) {
    KHG_HOST_ASSERT(pr.first > 0.0);
    const DiagGmm &g = *pr.second;
    KHG_HOST_ASSERT(g.Dim() == dim);
    std::copy(g.means_invvars_.data.begin(), g.means_invvars_.data.end(), means_invvars_.row(cur));
    std::copy(g.inv_vars_.data.begin(), g.inv_vars_.data.end(), inv_vars_.row(cur));
    for (int32_t i = 0; i < g.NumGauss(); ++i) weights_[cur + i] = pr.first * g.weights_[i];
    cur += g.NumGauss();
  }
  ComputeGconsts();
}

int32_t DiagGmm::ComputeGconsts() {
  int32_t nmix = NumGauss(), dim = Dim();
  KHG_HOST_ASSERT(nmix > 0 && dim > 0);
  if ((int32_t)gconsts_.size() != nmix) gconsts_.resize(nmix);
  int32_t num_bad = 0;
  Check(khg_compute_gconsts(nmix, dim, weights_.data(), means_invvars_.data.data(), inv_vars_.data.data(),
                            gconsts_.data(), &num_bad));
  valid_gconsts_ = true;
  Bump();
  return num_bad;
}

khg_model *DiagGmm::Device() const {
  if (!dev_ || dev_version_ != version_) {
    int32_t offs[2] = {0, NumGauss()};
    auto h = std::make_shared<ModelHandle>();
    Check(khg_model_create(Dim(), 1, offs, &h->h));
    Check(khg_model_set_kernel(h->h, KHG_KERNEL_SIMT));
    // gconsts are taken as stored on the host (they may be stale with respect to the
    // weights when the caller edited the live `weights` view: same as the reference).
    Check(khg_model_upload(h->h, weights_.data(), means_invvars_.data.data(), inv_vars_.data.data(),
                           gconsts_.data(), nullptr));
    dev_ = h;
    dev_version_ = version_;
  }
  return dev_->h;
}

void DiagGmm::LogLikelihoods(const FloatVector &data, FloatVector *loglikes) const {
  if ((int32_t)data.size() != Dim()) {
    std::ostringstream os;
    os << "DiagGmm::LogLikelihoods, dimension mismatch " << data.size() << " vs. " << Dim();
    Throw(os.str());
  }
  loglikes->resize(NumGauss());
  Check(khg_pdf_loglikes(Device(), 0, data.data(), 1, KHG_HOST, loglikes->data()));
}

void DiagGmm::LogLikelihoodsMatrix(const FloatMatrix &data, FloatMatrix *loglikes) const {
  KHG_HOST_ASSERT(data.rows != 0);
  if (data.cols != Dim()) {
    std::ostringstream os;
    os << "DiagGmm::LogLikelihoods, dimension mismatch " << data.cols << " vs. " << Dim();
    Throw(os.str());
  }
  *loglikes = FloatMatrix(data.rows, NumGauss());
  Check(khg_pdf_loglikes(Device(), 0, data.data.data(), data.rows, KHG_HOST, loglikes->data.data()));
}

void DiagGmm::LogLikelihoodsPreselect(const FloatVector &data, const std::vector<int32_t> &indices,
                                      FloatVector *loglikes) const {
  KHG_HOST_ASSERT((int32_t)data.size() == Dim());
  FloatVector all;
  LogLikelihoods(data, &all);
  loglikes->resize(indices.size());
  for (size_t i = 0; i < indices.size(); ++i) {
    KHG_HOST_ASSERT(indices[i] >= 0 && indices[i] < NumGauss());
    (*loglikes)[i] = all[indices[i]];
  }
}

float DiagGmm::GaussianSelection(const FloatVector &data, int32_t num_gselect, std::vector<int32_t> *output) const {
  KHG_HOST_ASSERT((int32_t)data.size() == Dim());
  output->assign(std::max(1, std::min(num_gselect, NumGauss())), 0);
  double tot = 0;
  Check(khg_gaussian_selection(Device(), 0, data.data(), 1, KHG_HOST, nullptr, 0, num_gselect, output->data(), nullptr, nullptr,
                               &tot));
  output->resize(std::min(num_gselect, NumGauss()));
  return (float)tot;
}

float DiagGmm::GaussianSelection(const FloatMatrix &data, int32_t num_gselect,
                                 std::vector<std::vector<int32_t>> *output) const {
  KHG_HOST_ASSERT(data.rows != 0);  // csrc/diag-gmm.cc:272
  KHG_HOST_ASSERT(data.cols == Dim());
  const int32_t k = std::min(num_gselect, NumGauss());
  std::vector<int32_t> idx((size_t)data.rows * std::max(k, 1));
  double tot = 0;
  Check(khg_gaussian_selection(Device(), 0, data.data.data(), data.rows, KHG_HOST, nullptr, 0, num_gselect, idx.data(), nullptr,
                               nullptr, &tot));
  output->assign(data.rows, {});
  for (int32_t t = 0; t < data.rows; ++t) (*output)[t].assign(idx.begin() + (size_t)t * k, idx.begin() + (size_t)(t + 1) * k);
  return (float)tot;
}

float DiagGmm::GaussianSelectionPreselect(const FloatVector &data, const std::vector<int32_t> &preselect, int32_t num_gselect,
                                          std::vector<int32_t> *output) const {
  KHG_HOST_ASSERT((int32_t)data.size() == Dim());
  KHG_HOST_ASSERT(!preselect.empty());  // KHG_ASSERT(!output->empty()), csrc/diag-gmm.cc:363
  const int32_t k = std::min<int32_t>(num_gselect, (int32_t)preselect.size());
  output->assign(std::max(k, 1), 0);
  double tot = 0;
  Check(khg_gaussian_selection(Device(), 0, data.data(), 1, KHG_HOST, preselect.data(), (int32_t)preselect.size(), num_gselect,
                               output->data(), nullptr, nullptr, &tot));
  output->resize(k);
  return (float)tot;
}

float DiagGmm::LogLikelihood(const FloatVector &data) const {
  if (!valid_gconsts_) Throw("Must call ComputeGconsts() before computing likelihood");
  if ((int32_t)data.size() != Dim()) {
    std::ostringstream os;
    os << "DiagGmm::LogLikelihoods, dimension mismatch " << data.size() << " vs. " << Dim();
    Throw(os.str());
  }
  float ll = 0.f;
  Check(khg_pdf_posteriors(Device(), 0, data.data(), 1, KHG_HOST, nullptr, &ll));
  return ll;
}

float DiagGmm::ComponentPosteriors(const FloatVector &data, FloatVector *posterior) const {
  if (!valid_gconsts_) Throw("Must call ComputeGconsts() before computing likelihood");
  if (posterior == nullptr) Throw("NULL pointer passed as return argument.");
  if ((int32_t)data.size() != Dim()) {
    std::ostringstream os;
    os << "DiagGmm::LogLikelihoods, dimension mismatch " << data.size() << " vs. " << Dim();
    Throw(os.str());
  }
  posterior->resize(NumGauss());
  float ll = 0.f;
  Check(khg_pdf_posteriors(Device(), 0, data.data(), 1, KHG_HOST, posterior->data(), &ll));
  return ll;
}

float DiagGmm::ComponentLogLikelihood(const FloatVector &data, int32_t comp_id) const {
  if (!valid_gconsts_) Throw("Must call ComputeGconsts() before computing likelihood");
  if ((int32_t)data.size() != Dim()) {
    std::ostringstream os;
    os << "DiagGmm::ComponentLogLikelihood, dimension mismatch " << data.size() << " vs. " << Dim();
    Throw(os.str());
  }
  KHG_HOST_ASSERT(comp_id >= 0 && comp_id < NumGauss());
  FloatVector all;
  LogLikelihoods(data, &all);
  return all[comp_id];
}

void DiagGmm::SetWeights(const FloatVector &w) {
  KHG_HOST_ASSERT(weights_.size() == w.size());
  weights_ = w;
  Touch();
}

void DiagGmm::SetMeans(const FloatMatrix &m) {
  KHG_HOST_ASSERT(means_invvars_.rows == m.rows && means_invvars_.cols == m.cols);
  for (size_t i = 0; i < m.size(); ++i) means_invvars_.data[i] = m.data[i] * inv_vars_.data[i];
  Touch();
}

FloatMatrix DiagGmm::GetMeans() const {
  FloatMatrix m(means_invvars_.rows, means_invvars_.cols);
  for (size_t i = 0; i < m.size(); ++i) m.data[i] = means_invvars_.data[i] / inv_vars_.data[i];
  return m;
}

void DiagGmm::SetInvVars(const FloatMatrix &v) {
  KHG_HOST_ASSERT(inv_vars_.rows == v.rows && inv_vars_.cols == v.cols);
  for (size_t i = 0; i < v.size(); ++i)
    means_invvars_.data[i] = means_invvars_.data[i] / inv_vars_.data[i] * v.data[i];
  inv_vars_ = v;
  Touch();
}

FloatMatrix DiagGmm::GetVars() const {
  FloatMatrix m(inv_vars_.rows, inv_vars_.cols);
  for (size_t i = 0; i < m.size(); ++i) m.data[i] = (float)(1.0 / inv_vars_.data[i]);
  return m;
}

void DiagGmm::SetComponentWeight(int32_t g, float w) {
  KHG_HOST_ASSERT(w > 0.0);
  KHG_HOST_ASSERT(g >= 0 && g < NumGauss());
  weights_[g] = w;
  Touch();
}

void DiagGmm::SetComponentMean(int32_t g, const FloatVector &v) {
  KHG_HOST_ASSERT(g >= 0 && g < NumGauss() && Dim() == (int32_t)v.size());
  for (int32_t d = 0; d < Dim(); ++d) means_invvars_(g, d) = inv_vars_(g, d) * v[d];
  Touch();
}

void DiagGmm::SetInvVarsAndMeans(const FloatMatrix &invvars, const FloatMatrix &means) {
  KHG_HOST_ASSERT(means_invvars_.rows == means.rows && means_invvars_.cols == means.cols &&
                  inv_vars_.rows == invvars.rows && inv_vars_.cols == invvars.cols);
  inv_vars_ = invvars;
  for (size_t i = 0; i < means.size(); ++i) means_invvars_.data[i] = means.data[i] * inv_vars_.data[i];
  Touch();
}

void DiagGmm::SetComponentInvVar(int32_t g, const FloatVector &v) {
  KHG_HOST_ASSERT(g >= 0 && g < NumGauss() && (int32_t)v.size() == Dim());
  for (int32_t d = 0; d < Dim(); ++d) {
    means_invvars_(g, d) = means_invvars_(g, d) / inv_vars_(g, d) * v[d];
    inv_vars_(g, d) = v[d];
  }
  Touch();
}

FloatVector DiagGmm::GetComponentMean(int32_t gauss) const {
  KHG_HOST_ASSERT(gauss >= 0 && gauss < NumGauss());
  FloatVector out(Dim());
  for (int32_t d = 0; d < Dim(); ++d) out[d] = means_invvars_(gauss, d) / inv_vars_(gauss, d);
  return out;
}

FloatVector DiagGmm::GetComponentVariance(int32_t gauss) const {
  KHG_HOST_ASSERT(gauss >= 0 && gauss < NumGauss());
  FloatVector out(Dim());
  for (int32_t d = 0; d < Dim(); ++d) out[d] = 1.0f / inv_vars_(gauss, d);
  return out;
}

void DiagGmm::RemoveComponent(int32_t gauss, bool renorm_weights) {
  KHG_HOST_ASSERT(gauss < NumGauss());
  KHG_HOST_ASSERT(gauss >= 0);
  if (NumGauss() == 1) Throw("Attempting to remove the only remaining component.");
  const int32_t dim = Dim();
  weights_.erase(weights_.begin() + gauss);
  gconsts_.erase(gconsts_.begin() + gauss);
  auto drop_row = [&](FloatMatrix &m) {
    m.data.erase(m.data.begin() + (size_t)gauss * dim, m.data.begin() + (size_t)(gauss + 1) * dim);
    m.rows -= 1;
  };
  drop_row(means_invvars_);
  drop_row(inv_vars_);
  if (renorm_weights) {
    float s = 0.f;
    for (float w : weights_) s += w;
    for (float &w : weights_) w /= s;
    valid_gconsts_ = false;
  }
  Bump();
}

void DiagGmm::RemoveComponents(const std::vector<int32_t> &gauss_in, bool renorm_weights) {
  std::vector<int32_t> gauss(gauss_in);
  std::sort(gauss.begin(), gauss.end());
  for (size_t i = 1; i < gauss.size(); ++i) KHG_HOST_ASSERT(gauss[i] != gauss[i - 1]);
  for (size_t i = 0; i < gauss.size(); ++i) {
    RemoveComponent(gauss[i], renorm_weights);
    for (size_t j = i + 1; j < gauss.size(); ++j) gauss[j]--;
  }
}

void DiagGmm::SetAllFromDevice(const float *w, const float *inv_vars, const float *means_invvars, const float *gconsts,
                               int32_t nmix, int32_t dim) {
  weights_.assign(w, w + nmix);
  gconsts_.assign(gconsts, gconsts + nmix);
  inv_vars_ = FloatMatrix(nmix, dim);
  means_invvars_ = FloatMatrix(nmix, dim);
  std::copy(inv_vars, inv_vars + (size_t)nmix * dim, inv_vars_.data.begin());
  std::copy(means_invvars, means_invvars + (size_t)nmix * dim, means_invvars_.data.begin());
  Bump();
  valid_gconsts_ = true;
}

void DiagGmm::SetParams(const FloatVector *w, const FloatMatrix *inv_vars, const FloatMatrix *means_invvars) {
  if (w) weights_ = *w;
  if (inv_vars) inv_vars_ = *inv_vars;
  if (means_invvars) means_invvars_ = *means_invvars;
  Touch();
}

// -------------------------------------------------------------- AmDiagGmm --
void AmDiagGmm::Init(const DiagGmm &proto, int32_t num_pdfs) {
  densities_.clear();
  dev_.reset();
  if (num_pdfs == 0) return;
  for (int32_t i = 0; i < num_pdfs; ++i) densities_.emplace_back(new DiagGmm(proto));
}

void AmDiagGmm::AddPdf(const DiagGmm &gmm) {
  if (!densities_.empty()) KHG_HOST_ASSERT(gmm.Dim() == this->Dim());
  densities_.emplace_back(new DiagGmm(gmm));
}

void AmDiagGmm::CopyFromAmDiagGmm(const AmDiagGmm &other) {
  densities_.clear();
  dev_.reset();
  for (int32_t i = 0; i < other.NumPdfs(); ++i) densities_.emplace_back(new DiagGmm(*other.densities_[i]));
}

void AmDiagGmm::SplitByCount(const FloatVector &state_occs, int32_t target_components, float perturb_factor, float power,
                             float min_count, const FloatMatrix *randn, uint64_t seed) {
  KHG_HOST_ASSERT((int32_t)state_occs.size() == NumPdfs());
  static uint64_t calls = 0;  // distinct draws per call when no seed is given (the reference uses a global generator)
  khg_model *nm = nullptr;
  Check(khg_model_split_by_count(Device(), state_occs.data(), target_components, perturb_factor, power, min_count,
                                 randn ? randn->data.data() : nullptr, randn ? randn->rows : 0,
                                 seed ? seed : 0x9E3779B97F4A7C15ull + (++calls), &nm, nullptr));
  AdoptDevice(nm, false);
}

void AmDiagGmm::MergeByCount(const FloatVector &state_occs, int32_t target_components, float power, float min_count) {
  KHG_HOST_ASSERT((int32_t)state_occs.size() == NumPdfs());
  khg_model *nm = nullptr;
  Check(khg_model_merge_by_count(Device(), state_occs.data(), target_components, power, min_count, &nm, nullptr));
  AdoptDevice(nm, false);
}

void AmDiagGmm::AdoptDevice(khg_model *nm, bool all_pdfs) {
  auto h = std::make_shared<ModelHandle>();
  h->h = nm;
  int32_t D = 0, P = 0, G = 0;
  Check(khg_model_info(nm, &D, &P, &G));
  KHG_HOST_ASSERT(P == NumPdfs());
  std::vector<int32_t> offs(P + 1);
  std::vector<float> w(G), miv((size_t)G * D), iv((size_t)G * D), gc(G);
  Check(khg_model_download(nm, offs.data(), w.data(), miv.data(), iv.data(), gc.data()));
  for (int32_t p = 0; p < P; ++p) {
    const int32_t g0 = offs[p], n = offs[p + 1] - g0;
    if (!all_pdfs && n == densities_[p]->NumGauss()) continue;  // neither split nor merged
    densities_[p]->SetAllFromDevice(w.data() + g0, iv.data() + (size_t)g0 * D, miv.data() + (size_t)g0 * D, gc.data() + g0, n, D);
  }
  dev_ = h;
  dev_sig_.clear();
  for (auto &d : densities_) dev_sig_.emplace_back(d.get(), d->version());
}

int32_t AmDiagGmm::NumGauss() const {
  int32_t ans = 0;
  for (auto &d : densities_) ans += d->NumGauss();
  return ans;
}

int32_t AmDiagGmm::NumGaussInPdf(int32_t pdf_index) const {
  KHG_HOST_ASSERT(pdf_index >= 0 && (size_t)pdf_index < densities_.size());
  return densities_[pdf_index]->NumGauss();
}

int32_t AmDiagGmm::ComputeGconsts() const {
  int32_t num_bad = 0;
  for (auto &d : densities_) num_bad += d->ComputeGconsts();
  return num_bad;
}

float AmDiagGmm::LogLikelihood(int32_t pdf_index, const FloatVector &data) const {
  KHG_HOST_ASSERT(pdf_index >= 0 && (size_t)pdf_index < densities_.size());
  return densities_[pdf_index]->LogLikelihood(data);
}

DiagGmm &AmDiagGmm::GetPdf(int32_t pdf_index) {
  KHG_HOST_ASSERT(pdf_index >= 0 && (size_t)pdf_index < densities_.size());
  return *densities_[pdf_index];
}
const DiagGmm &AmDiagGmm::GetPdf(int32_t pdf_index) const {
  KHG_HOST_ASSERT(pdf_index >= 0 && (size_t)pdf_index < densities_.size());
  return *densities_[pdf_index];
}

FloatVector AmDiagGmm::GetGaussianMean(int32_t pdf_index, int32_t gauss) const {
  return GetPdf(pdf_index).GetComponentMean(gauss);
}
FloatVector AmDiagGmm::GetGaussianVariance(int32_t pdf_index, int32_t gauss) const {
  return GetPdf(pdf_index).GetComponentVariance(gauss);
}
void AmDiagGmm::SetGaussianMean(int32_t pdf_index, int32_t gauss_index, const FloatVector &in) {
  GetPdf(pdf_index).SetComponentMean(gauss_index, in);
}

std::vector<int32_t> AmDiagGmm::GaussOffsets() const {
  std::vector<int32_t> offs(densities_.size() + 1, 0);
  for (size_t i = 0; i < densities_.size(); ++i) offs[i + 1] = offs[i] + densities_[i]->NumGauss();
  return offs;
}

std::shared_ptr<ModelHandle> AmDiagGmm::DeviceShared() const {
  KHG_HOST_ASSERT(!densities_.empty());
  bool stale = !dev_ || dev_sig_.size() != densities_.size();
  for (size_t i = 0; !stale && i < densities_.size(); ++i)
    stale = dev_sig_[i].first != densities_[i].get() || dev_sig_[i].second != densities_[i]->version();
  if (!stale) return dev_;
  const int32_t dim = Dim();
  std::vector<int32_t> offs = GaussOffsets();
  const size_t G = offs.back();
  std::vector<float> w(G), gc(G), miv(G * dim), iv(G * dim);
  for (size_t p = 0; p < densities_.size(); ++p) {
    const DiagGmm &g = *densities_[p];
    if (g.Dim() != dim) Throw("Dim mismatch between pdfs of the AmDiagGmm");
    if (!g.valid_gconsts()) {
      std::ostringstream os;
      os << "State " << p << ": Must call ComputeGconsts() before computing likelihood.";
      Throw(os.str());  // csrc/decodable-am-diag-gmm.cc:49-53
    }
    std::copy(g.weights().begin(), g.weights().end(), w.begin() + offs[p]);
    std::copy(g.gconsts().begin(), g.gconsts().end(), gc.begin() + offs[p]);
    std::copy(g.means_invvars().data.begin(), g.means_invvars().data.end(), miv.begin() + (size_t)offs[p] * dim);
    std::copy(g.inv_vars().data.begin(), g.inv_vars().data.end(), iv.begin() + (size_t)offs[p] * dim);
  }
  auto h = std::make_shared<ModelHandle>();
  Check(khg_model_create(dim, (int32_t)densities_.size(), offs.data(), &h->h));
  Check(khg_model_upload(h->h, w.data(), miv.data(), iv.data(), gc.data(), nullptr));
  dev_ = h;
  dev_sig_.clear();
  for (auto &d : densities_) dev_sig_.emplace_back(d.get(), d->version());
  return dev_;
}

// ------------------------------------------------------------ AccumDiagGmm --
std::string MleDiagGmmOptions::ToString() const {
  std::ostringstream os;
  os << "MleDiagGmmOptions(min_gaussian_weight=" << min_gaussian_weight
     << ", min_gaussian_occupancy=" << min_gaussian_occupancy << ", min_variance=" << min_variance
     << ", remove_low_count_gaussians=" << (remove_low_count_gaussians ? "True" : "False") << ")";
  return os.str();
}

void AccumDiagGmm::Resize(int32_t num_comp, int32_t dim, GmmFlagsType flags) {
  KHG_HOST_ASSERT(num_comp > 0 && dim > 0);
  num_comp_ = num_comp;
  dim_ = dim;
  flags_ = AugmentGmmFlags(flags);
  occupancy_.assign(num_comp, 0.0);
  mean_accumulator_ = (flags_ & kGmmMeans) ? DoubleMatrix(num_comp, dim) : DoubleMatrix();
  variance_accumulator_ = (flags_ & kGmmVariances) ? DoubleMatrix(num_comp, dim) : DoubleMatrix();
}

void AccumDiagGmm::SetZero(GmmFlagsType flags) {
  if (flags & ~flags_) Throw("Flags in argument do not match the active accumulators");
  if (flags & kGmmWeights) std::fill(occupancy_.begin(), occupancy_.end(), 0.0);
  if (flags & kGmmMeans) std::fill(mean_accumulator_.data.begin(), mean_accumulator_.data.end(), 0.0);
  if (flags & kGmmVariances) std::fill(variance_accumulator_.data.begin(), variance_accumulator_.data.end(), 0.0);
}

void AccumDiagGmm::Scale(float f, GmmFlagsType flags) {
  if (flags & ~flags_) Throw("Flags in argument do not match the active accumulators");
  double d = f;
  if (flags & kGmmWeights) for (double &v : occupancy_) v *= d;
  if (flags & kGmmMeans) for (double &v : mean_accumulator_.data) v *= d;
  if (flags & kGmmVariances) for (double &v : variance_accumulator_.data) v *= d;
}

// A one-pdf device model + stats used by the single-frame AccumDiagGmm methods: the
// arithmetic is the batched kernels' (T = 1), the result is folded into the host buffers.
namespace {
struct TempAcc {
  ModelHandle model;
  StatsHandle stats;
  TempAcc(int32_t nmix, int32_t dim, GmmFlagsType flags, const float *w, const float *miv, const float *iv,
          const float *gc) {
    int32_t offs[2] = {0, nmix};
    Check(khg_model_create(dim, 1, offs, &model.h));
    Check(khg_model_set_kernel(model.h, KHG_KERNEL_SIMT));
    Check(khg_model_upload(model.h, w, miv, iv, gc, nullptr));
    Check(khg_stats_create(model.h, flags, &stats.h));
  }
  void FoldInto(AccumDiagGmm *acc) {
    DoubleVector occ(acc->NumGauss());
    DoubleMatrix mean(acc->mean_accumulator().rows, acc->mean_accumulator().cols);
    DoubleMatrix var(acc->variance_accumulator().rows, acc->variance_accumulator().cols);
    Check(khg_stats_download(stats.h, occ.data(), mean.size() ? mean.data.data() : nullptr,
                             var.size() ? var.data.data() : nullptr, nullptr));
    for (size_t i = 0; i < occ.size(); ++i) acc->occupancy()[i] += occ[i];
    for (size_t i = 0; i < mean.size(); ++i) acc->mean_accumulator().data[i] += mean.data[i];
    for (size_t i = 0; i < var.size(); ++i) acc->variance_accumulator().data[i] += var.data[i];
  }
};
}  // namespace

void AccumDiagGmm::AccumulateForComponent(const FloatVector &data, int32_t comp_index, float weight) {
  if (flags_ & kGmmMeans) KHG_HOST_ASSERT((int32_t)data.size() == Dim());
  KHG_HOST_ASSERT(comp_index >= 0 && comp_index < NumGauss());
  // A one-hot "posterior" vector through the same device path as AccumulateFromPosteriors
  // would round weight*x in fp32; the reference promotes x to double first
  // (csrc/mle-diag-gmm.cc:115), so this single-Gaussian update is plain host bookkeeping.
  double wt = weight;
  occupancy_[comp_index] += wt;
  if (flags_ & kGmmMeans) {
    double *m = mean_accumulator_.row(comp_index);
    for (int32_t d = 0; d < dim_; ++d) m[d] += (double)data[d] * wt;
    if (flags_ & kGmmVariances) {
      double *v = variance_accumulator_.row(comp_index);
      for (int32_t d = 0; d < dim_; ++d) v[d] += (double)((data[d] * data[d]) * (float)wt);  // :117-119
    }
  }
}

void AccumDiagGmm::AccumulateFromPosteriors(const FloatVector &data, const FloatVector &posteriors) {
  if (flags_ & kGmmMeans) KHG_HOST_ASSERT((int32_t)data.size() == Dim());
  KHG_HOST_ASSERT((int32_t)posteriors.size() == NumGauss());
  // model parameters are irrelevant for this entry point; a unit model carries the shape
  std::vector<float> ones((size_t)num_comp_ * dim_, 1.0f), zeros((size_t)num_comp_ * dim_, 0.0f), gc(num_comp_, 0.f);
  FloatVector x = data;
  x.resize(dim_, 0.f);
  TempAcc t(num_comp_, dim_, flags_, nullptr, zeros.data(), ones.data(), gc.data());
  Check(khg_acc_from_posteriors(t.model.h, t.stats.h, 0, x.data(), 1, KHG_HOST, posteriors.data()));
  t.FoldInto(this);
}

float AccumDiagGmm::AccumulateFromDiag(const DiagGmm &gmm, const FloatVector &data, float weight) {
  KHG_HOST_ASSERT(gmm.NumGauss() == NumGauss());
  KHG_HOST_ASSERT(gmm.Dim() == Dim());
  KHG_HOST_ASSERT((int32_t)data.size() == Dim());
  if (!gmm.valid_gconsts()) Throw("Must call ComputeGconsts() before computing likelihood");
  TempAcc t(num_comp_, dim_, flags_, gmm.weights().data(), gmm.means_invvars().data.data(),
            gmm.inv_vars().data.data(), gmm.gconsts().data());
  int32_t pdf = 0;
  float ll = 0.f;
  Check(khg_acc_stats_ali(t.model.h, t.stats.h, data.data(), 1, KHG_HOST, &pdf, &weight, &ll, nullptr));
  t.FoldInto(this);
  return ll;
}

void AccumDiagGmm::AddStatsForComponent(int32_t g, double occ, const DoubleVector &x_stats,
                                        const DoubleVector &x2_stats) {
  KHG_HOST_ASSERT(g >= 0 && g < NumGauss());
  occupancy_[g] += occ;
  if (flags_ & kGmmMeans) {
    KHG_HOST_ASSERT((int32_t)x_stats.size() == dim_);
    for (int32_t d = 0; d < dim_; ++d) mean_accumulator_.row(g)[d] += x_stats[d];
  }
  if (flags_ & kGmmVariances) {
    KHG_HOST_ASSERT((int32_t)x2_stats.size() == dim_);
    for (int32_t d = 0; d < dim_; ++d) variance_accumulator_.row(g)[d] += x2_stats[d];
  }
}

void AccumDiagGmm::Add(float scale, const AccumDiagGmm &acc) {
  KHG_HOST_ASSERT(acc.NumGauss() == NumGauss() && acc.Dim() == Dim());
  for (size_t i = 0; i < occupancy_.size(); ++i) occupancy_[i] += acc.occupancy_[i] * scale;
  if (flags_ & kGmmMeans) {
    KHG_HOST_ASSERT(acc.mean_accumulator_.size() == mean_accumulator_.size());
    for (size_t i = 0; i < mean_accumulator_.size(); ++i) mean_accumulator_.data[i] += acc.mean_accumulator_.data[i] * scale;
  }
  if (flags_ & kGmmVariances) {
    KHG_HOST_ASSERT(acc.variance_accumulator_.size() == variance_accumulator_.size());
    for (size_t i = 0; i < variance_accumulator_.size(); ++i)
      variance_accumulator_.data[i] += acc.variance_accumulator_.data[i] * scale;
  }
}

// ---------------------------------------------------------------- M-step --
float MlObjective(const DiagGmm &gmm, const AccumDiagGmm &acc) {  // csrc/mle-diag-gmm.cc:479-499
  GmmFlagsType f = acc.Flags();
  const FloatVector &gc = gmm.gconsts();
  double o = 0.0;
  for (int32_t g = 0; g < gmm.NumGauss(); ++g) o += acc.occupancy()[g] * (double)gc[g];
  float obj = (float)o;
  if (f & kGmmMeans) {
    double s = 0.0;
    for (size_t i = 0; i < acc.mean_accumulator().size(); ++i)
      s += acc.mean_accumulator().data[i] * (double)gmm.means_invvars().data[i];
    obj = (float)((double)obj + s);
  }
  if (f & kGmmVariances) {
    double s = 0.0;
    for (size_t i = 0; i < acc.variance_accumulator().size(); ++i)
      s += acc.variance_accumulator().data[i] * (double)gmm.inv_vars().data[i];
    obj = (float)((double)obj - 0.5 * s);
  }
  return obj;
}

void MleDiagGmmUpdate(const MleDiagGmmOptions &config, const AccumDiagGmm &acc, GmmFlagsType flags, DiagGmm *gmm,
                      float *obj_change_out, float *count_out, int32_t *floored_elements_out,
                      int32_t *floored_gauss_out, int32_t *removed_gauss_out) {
  KHG_HOST_ASSERT(gmm != nullptr);
  if (flags & ~acc.Flags()) Throw("Flags in argument do not match the active accumulators");
  KHG_HOST_ASSERT(acc.NumGauss() == gmm->NumGauss() && acc.Dim() == gmm->Dim());
  const int32_t G = gmm->NumGauss(), D = gmm->Dim();
  double occ_sum = 0.0;
  for (double v : acc.occupancy()) occ_sum += v;
  int32_t elements_floored = 0, gauss_floored = 0;
  gmm->ComputeGconsts();
  const float obj_old = MlObjective(*gmm, acc);

  // "normal" (mean / variance) double representation of the current model
  // (DiagGmmNormal::CopyFromDiagGmm, csrc/diag-gmm-normal.cc:14-20)
  std::vector<double> nw(G), nvar((size_t)G * D), nmean((size_t)G * D);
  for (int32_t g = 0; g < G; ++g) nw[g] = gmm->weights()[g];
  for (size_t i = 0; i < nvar.size(); ++i) {
    nvar[i] = 1.0 / (double)gmm->inv_vars().data[i];
    nmean[i] = (double)gmm->means_invvars().data[i] * nvar[i];
  }
  const std::vector<double> old_mean_all = nmean, old_var_all = nvar;
  std::vector<int32_t> to_remove;
  for (int32_t i = 0; i < G; ++i) {
    const double occ = acc.occupancy()[i];
    const double prob = occ_sum > 0.0 ? occ / occ_sum : 1.0 / G;
    if (occ > (double)config.min_gaussian_occupancy && prob > (double)config.min_gaussian_weight) {
      nw[i] = prob;
      double *mean = &nmean[(size_t)i * D], *var = &nvar[(size_t)i * D];
      if (acc.Flags() & (kGmmMeans | kGmmVariances))
        for (int32_t d = 0; d < D; ++d) mean[d] = acc.mean_accumulator().row(i)[d] / occ;
      if (acc.Flags() & kGmmVariances) {
        KHG_HOST_ASSERT(acc.Flags() & kGmmMeans);
        int32_t floored = 0;
        for (int32_t d = 0; d < D; ++d) {
          double v = acc.variance_accumulator().row(i)[d] / occ - mean[d] * mean[d];
          if (!(flags & kGmmMeans)) {  // variance-only update: compensate for the mean shift (:300-304)
            const double dm = old_mean_all[(size_t)i * D + d] - mean[d];
            v += dm * dm;
          }
          if (v < config.min_variance) {
            v = config.min_variance;
            ++floored;
          }
          var[d] = v;
        }
        if (floored) {
          elements_floored += floored;
          ++gauss_floored;
        }
      }
    } else if (config.remove_low_count_gaussians && (int32_t)to_remove.size() < G - 1) {
      to_remove.push_back(i);
    } else {
      nw[i] = std::max(prob, (double)config.min_gaussian_weight);
    }
  }
  // back to the exponential form according to `flags` (DiagGmmNormal::CopyToDiagGmm,
  // csrc/diag-gmm-normal.cc:22-48)
  FloatVector w = gmm->weights();
  FloatMatrix iv = gmm->inv_vars(), miv = gmm->means_invvars();
  if (flags & kGmmWeights) for (int32_t g = 0; g < G; ++g) w[g] = (float)nw[g];
  if (flags & kGmmVariances) {
    for (size_t i = 0; i < iv.size(); ++i) iv.data[i] = (float)(1.0 / nvar[i]);
    if (!(flags & kGmmMeans))
      for (size_t i = 0; i < miv.size(); ++i) miv.data[i] = (float)old_mean_all[i] * iv.data[i];
  }
  if (flags & kGmmMeans) for (size_t i = 0; i < miv.size(); ++i) miv.data[i] = (float)nmean[i] * iv.data[i];
  gmm->SetParams(&w, &iv, &miv);
  gmm->ComputeGconsts();
  const float obj_new = MlObjective(*gmm, acc);
  if (obj_change_out) *obj_change_out = obj_new - obj_old;
  if (count_out) *count_out = (float)occ_sum;
  if (floored_elements_out) *floored_elements_out = elements_floored;
  if (floored_gauss_out) *floored_gauss_out = gauss_floored;
  if (!to_remove.empty()) {
    gmm->RemoveComponents(to_remove, true);
    gmm->ComputeGconsts();
  }
  if (removed_gauss_out) *removed_gauss_out = (int32_t)to_remove.size();
}

// ---------------------------------------------------------- AccumAmDiagGmm --
void AccumAmDiagGmm::Init(const AmDiagGmm &model, GmmFlagsType flags) {
  gmm_accumulators_.clear();
  dev_.reset();
  dev_model_keep_.reset();
  dev_model_ = nullptr;
  dirty_ = false;
  host_stats_ = false;
  total_frames_ = total_log_like_ = 0.0;  // fresh object semantics (members start at 0, .h:93-96)
  flags_ = AugmentGmmFlags(flags);
  for (int32_t i = 0; i < model.NumPdfs(); ++i) {
    gmm_accumulators_.emplace_back(new AccumDiagGmm());
    gmm_accumulators_.back()->Resize(model.GetPdf(i), flags);
  }
}

void AccumAmDiagGmm::Init(const AmDiagGmm &model, int32_t dim, GmmFlagsType flags) {
  KHG_HOST_ASSERT(dim > 0);
  gmm_accumulators_.clear();
  dev_.reset();
  dev_model_keep_.reset();
  dev_model_ = nullptr;
  dirty_ = false;
  host_stats_ = false;
  total_frames_ = total_log_like_ = 0.0;
  flags_ = AugmentGmmFlags(flags);
  for (int32_t i = 0; i < model.NumPdfs(); ++i) {
    gmm_accumulators_.emplace_back(new AccumDiagGmm());
    gmm_accumulators_.back()->Resize(model.GetPdf(i).NumGauss(), dim, flags);
  }
}

void AccumAmDiagGmm::SetZero(GmmFlagsType flags) {
  Flush();
  for (auto &a : gmm_accumulators_) a->SetZero(flags);
}

void AccumAmDiagGmm::EnsureDevice(const AmDiagGmm &model) const {
  KHG_HOST_ASSERT(model.NumPdfs() == NumAccs());
  std::shared_ptr<ModelHandle> keep = model.DeviceShared();
  khg_model *m = keep->h;
  if (dev_ && dev_model_ == m) return;
  Flush();  // statistics gathered under the previous pack go to the host first
  dev_.reset();            // stats handle first: it points into its model
  dev_model_keep_ = keep;  // keeps the pack alive for as long as our stats refer to it
  for (int32_t i = 0; i < NumAccs(); ++i)
    KHG_HOST_ASSERT(gmm_accumulators_[i]->NumGauss() == model.GetPdf(i).NumGauss() &&
                    gmm_accumulators_[i]->Dim() == model.Dim());
  auto h = std::make_shared<StatsHandle>();
  Check(khg_stats_create(m, flags_, &h->h));
  dev_ = h;
  dev_model_ = m;
}

khg_stats *AccumAmDiagGmm::DeviceStats(const AmDiagGmm &model) {
  EnsureDevice(model);
  dirty_ = true;
  return dev_->h;
}

void AccumAmDiagGmm::Flush() const {
  if (!dev_ || !dirty_) return;
  int32_t dim = 0, P = 0, G = 0;
  Check(khg_model_info(dev_model_, &dim, &P, &G));
  DoubleVector occ(G), mean, var;
  if (flags_ & kGmmMeans) mean.resize((size_t)G * dim);
  if (flags_ & kGmmVariances) var.resize((size_t)G * dim);
  double tot[2] = {0, 0};
  Check(khg_stats_download(dev_->h, occ.data(), mean.empty() ? nullptr : mean.data(), var.empty() ? nullptr : var.data(), tot));
  size_t g0 = 0;
  for (auto &a : gmm_accumulators_) {
    const size_t ng = a->NumGauss();
    for (size_t i = 0; i < ng; ++i) a->occupancy()[i] += occ[g0 + i];
    if (!mean.empty()) for (size_t i = 0; i < ng * dim; ++i) a->mean_accumulator().data[i] += mean[g0 * dim + i];
    if (!var.empty()) for (size_t i = 0; i < ng * dim; ++i) a->variance_accumulator().data[i] += var[g0 * dim + i];
    g0 += ng;
  }
  total_log_like_ += tot[0];
  total_frames_ += tot[1];
  Check(khg_stats_zero(dev_->h));
  dirty_ = false;
  host_stats_ = true;
}

void AccumAmDiagGmm::DeviceTotals(double tot[2]) const {
  tot[0] = tot[1] = 0.0;
  if (dev_ && dirty_) Check(khg_stats_download(dev_->h, nullptr, nullptr, nullptr, tot));
}

FloatVector AccumAmDiagGmm::PdfOccupancies() const {
  FloatVector out(gmm_accumulators_.size(), 0.f);
  DoubleVector occ;
  if (dev_ && dirty_) {
    int32_t dim = 0, P = 0, G = 0;
    Check(khg_model_info(dev_model_, &dim, &P, &G));
    occ.resize(G);
    Check(khg_stats_download(dev_->h, occ.data(), nullptr, nullptr, nullptr));
  }
  size_t g0 = 0;
  for (size_t p = 0; p < gmm_accumulators_.size(); ++p) {
    double s = 0.0;
    const size_t ng = gmm_accumulators_[p]->NumGauss();
    for (size_t i = 0; i < ng; ++i) s += gmm_accumulators_[p]->occupancy()[i] + (occ.empty() ? 0.0 : occ[g0 + i]);
    out[p] = (float)s;
    g0 += ng;
  }
  return out;
}

float AccumAmDiagGmm::AccumulateForGmm(const AmDiagGmm &model, const FloatVector &data, int32_t gmm_index, float weight) {
  KHG_HOST_ASSERT(gmm_index >= 0 && gmm_index < NumAccs());
  KHG_HOST_ASSERT((int32_t)data.size() == model.Dim());
  EnsureDevice(model);
  float ll = 0.f;
  dirty_ = true;
  Check(khg_acc_stats_ali(dev_model_, dev_->h, data.data(), 1, KHG_HOST, &gmm_index, &weight, &ll, nullptr));
  return ll;
}

float AccumAmDiagGmm::AccumulateForGmmTwofeats(const AmDiagGmm &model, const FloatVector &data1,
                                               const FloatVector &data2, int32_t gmm_index, float weight) {
  KHG_HOST_ASSERT(gmm_index >= 0 && gmm_index < NumAccs());
  EnsureDevice(model);
  const int32_t ng = model.GetPdf(gmm_index).NumGauss();
  FloatVector post(ng);
  float ll = 0.f;
  Check(khg_pdf_posteriors(dev_model_, gmm_index, data1.data(), 1, KHG_HOST, post.data(), &ll));
  for (float &p : post) p *= weight;
  dirty_ = true;
  Check(khg_acc_from_posteriors(dev_model_, dev_->h, gmm_index, data2.data(), 1, KHG_HOST, post.data()));
  // khg_acc_from_posteriors adds sum(post) to the frame total; the reference adds `weight`
  // and log_like * weight here (csrc/mle-am-diag-gmm.cc:71-72): correct on the host side.
  double s = 0.0;
  for (float p : post) s += (double)p;
  total_frames_ += (double)weight - s;
  total_log_like_ += (double)(ll * weight);
  return ll;
}

void AccumAmDiagGmm::AccumulateFromPosteriors(const AmDiagGmm &model, const FloatVector &data, int32_t gmm_index,
                                              const FloatVector &posteriors) {
  KHG_HOST_ASSERT(gmm_index >= 0 && gmm_index < NumAccs());
  KHG_HOST_ASSERT((int32_t)posteriors.size() == model.GetPdf(gmm_index).NumGauss());
  EnsureDevice(model);
  dirty_ = true;
  Check(khg_acc_from_posteriors(dev_model_, dev_->h, gmm_index, data.data(), 1, KHG_HOST, posteriors.data()));
}

void AccumAmDiagGmm::AccumulateForGaussian(const AmDiagGmm &am, const FloatVector &data, int32_t gmm_index,
                                           int32_t gauss_index, float weight) {
  KHG_HOST_ASSERT(gmm_index >= 0 && gmm_index < NumAccs());
  KHG_HOST_ASSERT(gauss_index >= 0 && gauss_index < am.GetPdf(gmm_index).NumGauss());
  host_stats_ = true;
  gmm_accumulators_[gmm_index]->AccumulateForComponent(data, gauss_index, weight);
}

float AccumAmDiagGmm::TotStatsCount() const {
  Flush();
  double ans = 0.0;
  for (auto &a : gmm_accumulators_)
    for (double v : a->occupancy()) ans += v;
  return (float)ans;
}

const AccumDiagGmm &AccumAmDiagGmm::GetAcc(int32_t index) const {
  KHG_HOST_ASSERT(index >= 0 && index < NumAccs());
  Flush();
  return *gmm_accumulators_[index];
}
AccumDiagGmm &AccumAmDiagGmm::GetAcc(int32_t index) {
  KHG_HOST_ASSERT(index >= 0 && index < NumAccs());
  Flush();
  host_stats_ = true;  // a mutable view: the caller may edit it
  return *gmm_accumulators_[index];
}

void AccumAmDiagGmm::Add(float scale, const AccumAmDiagGmm &other) {
  Flush();
  other.Flush();
  host_stats_ = true;
  total_frames_ += scale * other.total_frames_;
  total_log_like_ += scale * other.total_log_like_;
  KHG_HOST_ASSERT(NumAccs() == other.NumAccs());
  for (int32_t i = 0; i < NumAccs(); ++i) gmm_accumulators_[i]->Add(scale, *other.gmm_accumulators_[i]);
}

void AccumAmDiagGmm::Scale(float scale) {
  Flush();
  for (auto &a : gmm_accumulators_) a->Scale(scale, a->Flags());
  total_frames_ *= scale;
  total_log_like_ *= scale;
}

double AccumAmDiagGmm::AccumulateFrames(const AmDiagGmm &model, const float *feats, int64_t num_frames,
                                        const int32_t *pdf_ids, const float *frame_weights) {
  EnsureDevice(model);
  double tot = 0.0;
  dirty_ = true;
  Check(khg_acc_stats_ali(dev_model_, dev_->h, feats, num_frames, KHG_HOST, pdf_ids, frame_weights, nullptr, &tot));
  return tot;
}

double AccumAmDiagGmm::AccumulateAlignment(const AmDiagGmm &model, const std::vector<int32_t> &tid2pdf,
                                           const float *feats, int64_t num_frames, const int32_t *tids,
                                           double *trans_accs) {
  EnsureDevice(model);
  KHG_HOST_ASSERT(tid2pdf.size() >= 2);
  double tot = 0.0;
  dirty_ = true;
  Check(khg_acc_stats_ali_tids(dev_model_, dev_->h, feats, num_frames, tids, tid2pdf.data(),
                               (int32_t)tid2pdf.size() - 1, trans_accs, &tot));
  return tot;
}

void MleAmDiagGmmUpdate(const MleDiagGmmOptions &config, const AccumAmDiagGmm &acc, GmmFlagsType flags,
                        AmDiagGmm *am_gmm, float *obj_change_out, float *count_out) {
  KHG_HOST_ASSERT(am_gmm != nullptr);
  // Device M-step (khg_mle_update): when every statistic of this accumulator is still on the
  // device and was gathered under the model's CURRENT pack, the update runs there — the 26 MB of statistics never
  // leave the GPU — and the host pdfs are rebuilt from the new pack.  Anything else (statistics folded into or
  // edited on the host, a model changed since the E-step) takes the host path below.
  if (acc.dev_ && acc.dirty_ && !acc.host_stats_ && acc.NumAccs() == am_gmm->NumPdfs() && acc.Dim() == am_gmm->Dim() &&
      acc.dev_model_ == am_gmm->Device()) {
    khg_mle_options o;
    o.min_gaussian_weight = config.min_gaussian_weight;
    o.min_gaussian_occupancy = config.min_gaussian_occupancy;
    o.min_variance = config.min_variance;
    o.remove_low_count_gaussians = config.remove_low_count_gaussians ? 1 : 0;
    khg_model *nm = nullptr;
    float obj = 0.f, cnt = 0.f;
    Check(khg_mle_update(acc.dev_model_, acc.dev_->h, &o, flags, &nm, &obj, &cnt, nullptr, nullptr, nullptr));
    am_gmm->AdoptDevice(nm, true);
    if (obj_change_out) *obj_change_out = obj;
    if (count_out) *count_out = cnt;
    return;
  }
  acc.Flush();
  if (acc.Dim() != am_gmm->Dim()) Throw("Dimensions of accumulator and gmm do not match");
  KHG_HOST_ASSERT(acc.NumAccs() == am_gmm->NumPdfs());
  float tot_obj = 0.f, tot_count = 0.f;
  for (int32_t i = 0; i < acc.NumAccs(); ++i) {
    float oc = 0.f, c = 0.f;
    MleDiagGmmUpdate(config, *acc.gmm_accumulators_[i], flags, &am_gmm->GetPdf(i), &oc, &c);
    tot_obj += oc;
    tot_count += c;
  }
  if (obj_change_out) *obj_change_out = tot_obj;
  if (count_out) *count_out = tot_count;
}

// --------------------------------------------------------------- Decodable --
DecodableAmDiagGmmUnmapped::DecodableAmDiagGmmUnmapped(const AmDiagGmm &am, const FloatMatrix &feats,
                                                       float log_sum_exp_prune)
    : log_sum_exp_prune_(log_sum_exp_prune) {
  num_frames_ = feats.rows;
  num_pdfs_ = am.NumPdfs();
  if (num_frames_ == 0) return;
  if (am.Dim() != feats.cols) {
    std::ostringstream os;
    os << "Dim mismatch: data dim = " << feats.cols << " vs. model dim = " << am.Dim();
    Throw(os.str());  // csrc/decodable-am-diag-gmm.cc:44-47
  }
  block_.resize((size_t)num_pdfs_ * num_frames_);
  Check(khg_loglikes_all_pdfs(am.Device(), feats.data.data(), num_frames_, KHG_HOST, 1.0f, KHG_PDF_MAJOR,
                              block_.data(), num_frames_, KHG_HOST));
}

float DecodableAmDiagGmmUnmapped::LogLikelihoodZeroBased(int32_t frame, int32_t state) const {
  KHG_HOST_ASSERT(static_cast<size_t>(frame) < static_cast<size_t>(num_frames_));
  if (!(static_cast<size_t>(state) < static_cast<size_t>(num_pdfs_)))
    Throw("Assertion failed: state < NumIndices(): Likely graph/model mismatch, e.g. using wrong HCLG.fst");
  return block_[(size_t)state * num_frames_ + frame];
}

}  // namespace khg
