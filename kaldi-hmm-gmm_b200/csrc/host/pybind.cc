// kaldi-hmm-gmm_b200/csrc/host/pybind.cc — pybind11 module `_khg_b200`: the hot subset of
// the reference's `_kaldi_hmm_gmm` extension (reference kaldi-hmm-gmm/python/csrc/
// diag-gmm.cc, am-diag-gmm.cc, mle-diag-gmm.cc, mle-am-diag-gmm.cc,
// decodable-am-diag-gmm.cc, decodable-itf.cc, model-common.cc) with the same class
// names, method names, keyword arguments, dtypes and exceptions.  numpy arrays replace
// pybind11/eigen.h (no Eigen here); properties that are live views in the reference
// (python/csrc/diag-gmm.cc:41-45, mle-diag-gmm.cc:75-99) are live numpy views here too.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "khg_host.h"

namespace py = pybind11;
using namespace khg;

using FArr = py::array_t<float, py::array::c_style | py::array::forcecast>;
using DArr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using IArr = py::array_t<int32_t, py::array::c_style | py::array::forcecast>;

static FloatVector ToVec(const FArr &a) {
  if (a.ndim() != 1) throw std::runtime_error("expected a 1-D float array");
  return FloatVector(a.data(), a.data() + a.shape(0));
}
static DoubleVector ToDVec(const DArr &a) {
  if (a.ndim() != 1) throw std::runtime_error("expected a 1-D double array");
  return DoubleVector(a.data(), a.data() + a.shape(0));
}
static FloatMatrix ToMat(const FArr &a) {
  if (a.ndim() != 2) throw std::runtime_error("expected a 2-D float array");
  FloatMatrix m((int32_t)a.shape(0), (int32_t)a.shape(1));
  std::copy(a.data(), a.data() + a.size(), m.data.begin());
  return m;
}
static py::array FromVec(const FloatVector &v) {
  py::array_t<float> a((py::ssize_t)v.size());
  std::copy(v.begin(), v.end(), a.mutable_data());
  return a;
}
static py::array FromMat(const FloatMatrix &m) {
  py::array_t<float> a({(py::ssize_t)m.rows, (py::ssize_t)m.cols});
  std::copy(m.data.begin(), m.data.end(), a.mutable_data());
  return a;
}
// Live view on C++ storage, kept alive by `owner` (reference_internal semantics).
template <class T>
static py::array View1(std::vector<T> &v, py::handle owner, bool writeable) {
  py::array_t<T> a({(py::ssize_t)v.size()}, {(py::ssize_t)sizeof(T)}, v.data(), owner);
  if (!writeable) py::detail::array_proxy(a.ptr())->flags &= ~py::detail::npy_api::NPY_ARRAY_WRITEABLE_;
  return a;
}
template <class T>
static py::array View2(std::vector<T> &v, int32_t rows, int32_t cols, py::handle owner, bool writeable) {
  if (v.empty()) return py::array_t<T>(std::vector<py::ssize_t>{0, 0});  // len(acc.mean_accumulator) == 0
  py::array_t<T> a({(py::ssize_t)rows, (py::ssize_t)cols}, {(py::ssize_t)(sizeof(T) * cols), (py::ssize_t)sizeof(T)}, v.data(), owner);
  if (!writeable) py::detail::array_proxy(a.ptr())->flags &= ~py::detail::npy_api::NPY_ARRAY_WRITEABLE_;
  return a;
}

// tid -> pdf array from either a numpy array or an object with `id2pdf_id`
// (the reference's TransitionModel exposes it: python/csrc/transition-model.cc:76).
static std::vector<int32_t> Tid2Pdf(const py::object &tm) {
  py::object src = py::hasattr(tm, "id2pdf_id") ? tm.attr("id2pdf_id") : tm;
  IArr a = IArr::ensure(src);
  if (!a || a.ndim() != 1) throw std::runtime_error("tm must be a TransitionModel (id2pdf_id) or a 1-D int32 tid->pdf array");
  return std::vector<int32_t>(a.data(), a.data() + a.shape(0));
}

class PyDecodableInterface : public DecodableInterface {  // python/csrc/decodable-itf.cc:16-41
 public:
  using DecodableInterface::DecodableInterface;
  float LogLikelihood(int32_t frame, int32_t index) override {
    PYBIND11_OVERRIDE_PURE_NAME(float, DecodableInterface, "log_likelihood", LogLikelihood, frame, index);
  }
  bool IsLastFrame(int32_t frame) const override {
    PYBIND11_OVERRIDE_PURE_NAME(bool, DecodableInterface, "is_last_frame", IsLastFrame, frame);
  }
  int32_t NumFramesReady() const override {
    PYBIND11_OVERRIDE_NAME(int32_t, DecodableInterface, "num_frames_ready", NumFramesReady);
  }
  int32_t NumIndices() const override {
    PYBIND11_OVERRIDE_PURE_NAME(int32_t, DecodableInterface, "num_indices", NumIndices);
  }
};

PYBIND11_MODULE(_khg_b200, m) {
  m.doc() = "B200-native diag-GMM E-step behind kaldi-hmm-gmm's class API (hot subset)";

  // ---- model-common (python/csrc/model-common.cc:10-26) ----
  py::enum_<GmmUpdateFlags>(m, "GmmUpdateFlags", py::arithmetic())
      .value("kGmmMeans", kGmmMeans)
      .value("kGmmVariances", kGmmVariances)
      .value("kGmmWeights", kGmmWeights)
      .value("kGmmTransitions", kGmmTransitions)
      .value("kGmmAll", kGmmAll)
      .export_values();
  m.def("str_to_gmm_flags", &StringToGmmFlags);
  m.def("gmm_flags_to_str", &GmmFlagsToString);
  m.def("augment_gmm_flags", &AugmentGmmFlags);

  // ---- DiagGmm (python/csrc/diag-gmm.cc:15-168) ----
  py::class_<DiagGmm>(m, "DiagGmm")
      .def(py::init<>())
      .def(py::init<const DiagGmm &>(), py::arg("gmm"))
      .def(py::init<int32_t, int32_t>(), py::arg("nmix"), py::arg("dim"))
      .def(py::init([](const std::vector<std::pair<float, const DiagGmm *>> &gmms) { return new DiagGmm(gmms); }),
           py::arg("gmms"))
      .def("resize", &DiagGmm::Resize, py::arg("nmix"), py::arg("dim"))
      .def("copy_from_diag_gmm", &DiagGmm::CopyFromDiagGmm, py::arg("diaggmm"))
      .def("compute_gconsts", &DiagGmm::ComputeGconsts)
      .def("set_weights", [](DiagGmm &s, const FArr &w) { s.SetWeights(ToVec(w)); }, py::arg("w"))
      .def("set_means", [](DiagGmm &s, const FArr &x) { s.SetMeans(ToMat(x)); }, py::arg("m"))
      .def("set_invvars", [](DiagGmm &s, const FArr &x) { s.SetInvVars(ToMat(x)); }, py::arg("inv_vars"))
      .def("log_likelihood", [](const DiagGmm &s, const FArr &d) { return s.LogLikelihood(ToVec(d)); }, py::arg("data"),
           "Return the total loglikes in a float")
      .def("log_likelihoods",
           [](const DiagGmm &s, const FArr &d) {
             FloatVector ans;
             s.LogLikelihoods(ToVec(d), &ans);
             return FromVec(ans);
           },
           py::arg("data"), "Return the loglike of each component in a 1-D tensor")
      .def_property(
          "weights", [](py::object self) { return View1(self.cast<DiagGmm &>().weights(), self, true); },
          [](DiagGmm &s, const FArr &w) { s.SetWeights(ToVec(w)); })
      .def_property_readonly("means", [](const DiagGmm &s) { return FromMat(s.GetMeans()); })
      .def_property_readonly("vars", [](const DiagGmm &s) { return FromMat(s.GetVars()); })
      .def_property_readonly("num_gauss", &DiagGmm::NumGauss)
      .def_property_readonly("dim", &DiagGmm::Dim)
      .def_property_readonly("valid_gconsts", &DiagGmm::valid_gconsts)
      .def_property_readonly("gconsts",
                             [](py::object self) {
                               auto &g = self.cast<DiagGmm &>();
                               return View1(const_cast<FloatVector &>(g.gconsts()), self, false);
                             })
      .def_property_readonly("means_invvars",
                             [](py::object self) {
                               auto &g = self.cast<DiagGmm &>();
                               auto &mm = const_cast<FloatMatrix &>(g.means_invvars());
                               return View2(mm.data, mm.rows, mm.cols, self, false);
                             })
      .def_property_readonly("inv_vars",
                             [](py::object self) {
                               auto &g = self.cast<DiagGmm &>();
                               auto &mm = const_cast<FloatMatrix &>(g.inv_vars());
                               return View2(mm.data, mm.rows, mm.cols, self, false);
                             })
      .def("set_component_weight", &DiagGmm::SetComponentWeight, py::arg("gauss"), py::arg("weight"))
      .def("set_component_mean", [](DiagGmm &s, int32_t g, const FArr &v) { s.SetComponentMean(g, ToVec(v)); },
           py::arg("gauss"), py::arg("mean"))
      .def("set_invvars_and_means",
           [](DiagGmm &s, const FArr &iv, const FArr &mean) { s.SetInvVarsAndMeans(ToMat(iv), ToMat(mean)); },
           py::arg("inv_vars"), py::arg("means"))
      .def("set_component_inv_var", [](DiagGmm &s, int32_t g, const FArr &v) { s.SetComponentInvVar(g, ToVec(v)); },
           py::arg("gauss"), py::arg("inv_var"))
      .def("get_component_mean", [](const DiagGmm &s, int32_t g) { return FromVec(s.GetComponentMean(g)); }, py::arg("gauss"))
      .def("get_component_variance", [](const DiagGmm &s, int32_t g) { return FromVec(s.GetComponentVariance(g)); },
           py::arg("gauss"))
      .def("remove_component", &DiagGmm::RemoveComponent, py::arg("gauss"), py::arg("renorm_weights"))
      .def("remove_components", &DiagGmm::RemoveComponents, py::arg("gauss"), py::arg("renorm_weights"))
      .def("log_likelihoods_matrix",
           [](const DiagGmm &s, const FArr &d) {
             FloatMatrix ans;
             s.LogLikelihoodsMatrix(ToMat(d), &ans);
             return FromMat(ans);
           },
           py::arg("data"),
           "data is a 2-D tensor of shape (N, dim); it returns a 2-D tensor of shape (N, nmix) containing the "
           "loglike of each component")
      .def("log_likelihoods_preselect",
           [](const DiagGmm &s, const FArr &d, const std::vector<int32_t> &indices) {
             FloatVector ans;
             s.LogLikelihoodsPreselect(ToVec(d), indices, &ans);
             return FromVec(ans);
           },
           py::arg("data"), py::arg("indices"))
      // python/csrc/diag-gmm.cc:108-136: (total log-like, indices) — selection on the device
      .def("gaussian_selection_1d",
           [](const DiagGmm &s, const FArr &d, int32_t num_gselect) {
             std::vector<int32_t> out;
             float f = s.GaussianSelection(ToVec(d), num_gselect, &out);
             return std::make_pair(f, out);
           },
           py::arg("data"), py::arg("num_gselect"))
      .def("gaussian_selection_2d",
           [](const DiagGmm &s, const FArr &d, int32_t num_gselect) {
             std::vector<std::vector<int32_t>> out;
             float f = s.GaussianSelection(ToMat(d), num_gselect, &out);
             return std::make_pair(f, out);
           },
           py::arg("data"), py::arg("num_gselect"))
      .def("gaussian_selection_preselect",
           [](const DiagGmm &s, const FArr &d, const std::vector<int32_t> &preselect, int32_t num_gselect) {
             std::vector<int32_t> out;
             float f = s.GaussianSelectionPreselect(ToVec(d), preselect, num_gselect, &out);
             return std::make_pair(f, out);
           },
           py::arg("data"), py::arg("preselect"), py::arg("num_gselect"))
      .def("component_posteriors",
           [](const DiagGmm &s, const FArr &d) {
             FloatVector post;
             float f = s.ComponentPosteriors(ToVec(d), &post);
             return std::make_pair(f, FromVec(post));
           },
           py::arg("data"))
      .def("component_log_likelihood",
           [](const DiagGmm &s, const FArr &d, int32_t c) { return s.ComponentLogLikelihood(ToVec(d), c); },
           py::arg("data"), py::arg("comp_id"))
      .def(py::pickle(  // (weights, inv_vars, means_invvars): python/csrc/diag-gmm.cc:157-167
          [](const DiagGmm &s) { return py::make_tuple(FromVec(s.weights()), FromMat(s.inv_vars()), FromMat(s.means_invvars())); },
          [](const py::tuple &t) {
            return std::make_unique<DiagGmm>(ToVec(t[0].cast<FArr>()), ToMat(t[1].cast<FArr>()), ToMat(t[2].cast<FArr>()));
          }));

  // ---- AmDiagGmm (python/csrc/am-diag-gmm.cc:13-72) ----
  py::class_<AmDiagGmm>(m, "AmDiagGmm")
      .def(py::init<>())
      .def_property_readonly("dim", &AmDiagGmm::Dim)
      .def_property_readonly("num_pdfs", &AmDiagGmm::NumPdfs)
      .def_property_readonly("num_gauss", &AmDiagGmm::NumGauss)
      .def("num_gauss_in_pdf", &AmDiagGmm::NumGaussInPdf, py::arg("pdf_index"))
      .def("init", &AmDiagGmm::Init, py::arg("proto"), py::arg("num_pdfs"))
      .def("add_pdf", &AmDiagGmm::AddPdf, py::arg("gmm"))
      .def("copy_from_am_diag_gmm", &AmDiagGmm::CopyFromAmDiagGmm, py::arg("other"))
      .def("compute_gconsts", &AmDiagGmm::ComputeGconsts)
      .def("log_likelihood", [](const AmDiagGmm &s, int32_t p, const FArr &d) { return s.LogLikelihood(p, ToVec(d)); },
           py::arg("pdf_index"), py::arg("data"))
      .def("get_pdf", [](AmDiagGmm &s, int32_t p) -> DiagGmm & { return s.GetPdf(p); }, py::arg("pdf_index"),
           py::return_value_policy::reference_internal)
      .def("get_gaussian_mean", [](const AmDiagGmm &s, int32_t p, int32_t g) { return FromVec(s.GetGaussianMean(p, g)); },
           py::arg("pdf_index"), py::arg("gauss"))
      .def("get_gaussian_variance",
           [](const AmDiagGmm &s, int32_t p, int32_t g) { return FromVec(s.GetGaussianVariance(p, g)); },
           py::arg("pdf_index"), py::arg("gauss"))
      .def("set_gaussian_mean",
           [](AmDiagGmm &s, int32_t p, int32_t g, const FArr &in) { s.SetGaussianMean(p, g, ToVec(in)); },
           py::arg("pdf_index"), py::arg("gauss_index"), py::arg("in"))
      // python/csrc/am-diag-gmm.cc:27-30 (+ optional randn / seed: reproducible perturbations)
      .def("split_by_count",
           [](AmDiagGmm &s, const FArr &occs, int32_t target, float perturb, float power, float min_count, py::object randn,
              uint64_t seed) {
             if (randn.is_none()) {
               s.SplitByCount(ToVec(occs), target, perturb, power, min_count, nullptr, seed);
             } else {
               FloatMatrix r = ToMat(randn.cast<FArr>());
               s.SplitByCount(ToVec(occs), target, perturb, power, min_count, &r, seed);
             }
           },
           py::arg("state_occs"), py::arg("target_components"), py::arg("perturb_factor"), py::arg("power"),
           py::arg("min_count"), py::arg("randn") = py::none(), py::arg("seed") = 0)
      .def("merge_by_count",  // python/csrc/am-diag-gmm.cc:31-33
           [](AmDiagGmm &s, const FArr &occs, int32_t target, float power, float min_count) {
             s.MergeByCount(ToVec(occs), target, power, min_count);
           },
           py::arg("state_occs"), py::arg("target_components"), py::arg("power"), py::arg("min_count"))
      // new, batched: (T, num_pdfs) block of per-pdf log-likelihoods computed by the dense kernel
      .def("log_likelihoods_all_pdfs",
           [](const AmDiagGmm &s, const FArr &feats, float scale) {
             if (feats.ndim() != 2 || feats.shape(1) != s.Dim()) throw std::runtime_error("feats must be (T, dim)");
             py::array_t<float> out({(py::ssize_t)feats.shape(0), (py::ssize_t)s.NumPdfs()});
             if (feats.shape(0) > 0) {
               khg_model *dev = s.Device();
               const float *fp = feats.data();
               float *op = out.mutable_data();
               const int64_t T = feats.shape(0);
               const int32_t P = s.NumPdfs();
               khg_status st;
               {
                 py::gil_scoped_release nogil;
                 st = khg_loglikes_all_pdfs(dev, fp, T, KHG_HOST, scale, KHG_FRAME_MAJOR, op, P, KHG_HOST);
               }
               Check(st);
             }
             return out;
           },
           py::arg("feats"), py::arg("scale") = 1.0f)
      // new, batched: gmm-align-compiled for many utterances in one call (khg_align_batch);
      // graphs as the plain arrays of khg_graph_batch (include/khg_b200.h)
      .def("align_batch",
           [](const AmDiagGmm &s, const FArr &feats, const py::array_t<int64_t, py::array::c_style | py::array::forcecast> &frame_offsets,
              const IArr &state_offsets, const IArr &arc_offsets, const IArr &arc_ilabel, const IArr &arc_nextstate,
              const FArr &arc_weight, const IArr &start_state, const FArr &final_cost, const IArr &tid2pdf, float acoustic_scale,
              float beam, float retry_beam, bool want_paths) {
             if (feats.ndim() != 2 || feats.shape(1) != s.Dim()) throw std::runtime_error("feats must be (T, dim)");
             const int32_t U = (int32_t)start_state.shape(0);
             if (frame_offsets.shape(0) != U + 1 || state_offsets.shape(0) != U + 1)
               throw std::runtime_error("frame_offsets / state_offsets must have n_utts + 1 entries");
             const int64_t T = U ? frame_offsets.data()[U] : 0;
             if (T > feats.shape(0)) throw std::runtime_error("frame_offsets exceed the feature matrix");
             khg_graph_batch gb{U, frame_offsets.data(), state_offsets.data(), arc_offsets.data(), arc_ilabel.data(),
                                arc_nextstate.data(), arc_weight.data(), start_state.data(), final_cost.data()};
             py::array_t<int32_t> ali((py::ssize_t)T), status((py::ssize_t)U);
             py::array_t<float> like((py::ssize_t)U);
             py::array_t<int64_t> poff((py::ssize_t)U + 1);
             int64_t cap = want_paths ? T + T / 2 + 16 * (int64_t)U + 1024 : 0;
             py::array_t<int32_t> paths;
             khg_model *dev = s.Device();
             for (;;) {
               paths = py::array_t<int32_t>((py::ssize_t)cap);
               int32_t *pa = want_paths ? paths.mutable_data() : nullptr, *al = ali.mutable_data(), *stp = status.mutable_data();
               float *lk = like.mutable_data();
               int64_t *po = poff.mutable_data();
               const float *fp = feats.data();
               const int32_t *t2p = tid2pdf.data();
               const int32_t n_tids = (int32_t)tid2pdf.shape(0);
               khg_status st;
               {
                 py::gil_scoped_release nogil;
                 st = khg_align_batch(dev, &gb, fp, KHG_HOST, t2p, n_tids, acoustic_scale, beam, retry_beam, al, stp, lk, pa, po, cap,
                                      nullptr);
               }
               if (st != KHG_OK && want_paths && std::string(khg_last_error()).find("path_capacity") != std::string::npos) {
                 cap *= 4;
                 continue;
               }
               Check(st);
               break;
             }
             return py::make_tuple(ali, status, like, poff, paths);
           },
           py::arg("feats"), py::arg("frame_offsets"), py::arg("state_offsets"), py::arg("arc_offsets"), py::arg("arc_ilabel"),
           py::arg("arc_nextstate"), py::arg("arc_weight"), py::arg("start_state"), py::arg("final_cost"), py::arg("tid2pdf"),
           py::arg("acoustic_scale") = 1.0f, py::arg("beam") = 200.0f, py::arg("retry_beam") = 0.0f, py::arg("want_paths") = true)
      .def(py::pickle(  // 3 arrays per pdf: python/csrc/am-diag-gmm.cc:47-71
          [](const AmDiagGmm &s) {
            py::tuple t(s.NumPdfs() * 3);
            for (int32_t i = 0; i < s.NumPdfs(); ++i) {
              const DiagGmm &g = s.GetPdf(i);
              t[3 * i + 0] = FromVec(g.weights());
              t[3 * i + 1] = FromMat(g.inv_vars());
              t[3 * i + 2] = FromMat(g.means_invvars());
            }
            return t;
          },
          [](const py::tuple &t) {
            auto ans = std::make_unique<AmDiagGmm>();
            for (size_t i = 0; i < t.size() / 3; ++i)
              ans->AddPdf(DiagGmm(ToVec(t[3 * i].cast<FArr>()), ToMat(t[3 * i + 1].cast<FArr>()), ToMat(t[3 * i + 2].cast<FArr>())));
            return ans;
          }));

  // ---- MleDiagGmmOptions / AccumDiagGmm (python/csrc/mle-diag-gmm.cc:16-118) ----
  py::class_<MleDiagGmmOptions>(m, "MleDiagGmmOptions")
      .def(py::init([](float w, float occ, double var, bool rm) {
             auto o = std::make_unique<MleDiagGmmOptions>();
             o->min_gaussian_weight = w;
             o->min_gaussian_occupancy = occ;
             o->min_variance = var;
             o->remove_low_count_gaussians = rm;
             return o;
           }),
           py::arg("min_gaussian_weight") = 1.0e-05, py::arg("min_gaussian_occupancy") = 10.0,
           py::arg("min_variance") = 0.001, py::arg("remove_low_count_gaussians") = true)
      .def_readwrite("min_gaussian_weight", &MleDiagGmmOptions::min_gaussian_weight)
      .def_readwrite("min_gaussian_occupancy", &MleDiagGmmOptions::min_gaussian_occupancy)
      .def_readwrite("min_variance", &MleDiagGmmOptions::min_variance)
      .def_readwrite("remove_low_count_gaussians", &MleDiagGmmOptions::remove_low_count_gaussians)
      .def("__str__", &MleDiagGmmOptions::ToString);

  py::class_<AccumDiagGmm>(m, "AccumDiagGmm")
      .def(py::init<>())
      .def(py::init<const DiagGmm &, GmmFlagsType>(), py::arg("gmm"), py::arg("flags"))
      .def(py::init<const AccumDiagGmm &>())
      .def("resize", (void (AccumDiagGmm::*)(int32_t, int32_t, GmmFlagsType))(&AccumDiagGmm::Resize), py::arg("num_gauss"),
           py::arg("dim"), py::arg("flags"))
      .def_property_readonly("num_gauss", &AccumDiagGmm::NumGauss)
      .def_property_readonly("dim", &AccumDiagGmm::Dim)
      .def_property_readonly("flags", &AccumDiagGmm::Flags)
      .def_property(
          "occupancy", [](py::object self) { return View1(self.cast<AccumDiagGmm &>().occupancy(), self, true); },
          [](AccumDiagGmm &s, const DArr &v) { s.occupancy() = ToDVec(v); })
      .def_property(
          "mean_accumulator",
          [](py::object self) {
            auto &mm = self.cast<AccumDiagGmm &>().mean_accumulator();
            return View2(mm.data, mm.rows, mm.cols, self, true);
          },
          [](AccumDiagGmm &s, const DArr &v) {
            if (v.ndim() != 2) throw std::runtime_error("expected a 2-D double array");
            auto &mm = s.mean_accumulator();
            mm.rows = (int32_t)v.shape(0);
            mm.cols = (int32_t)v.shape(1);
            mm.data.assign(v.data(), v.data() + v.size());
          })
      .def_property(
          "variance_accumulator",
          [](py::object self) {
            auto &mm = self.cast<AccumDiagGmm &>().variance_accumulator();
            return View2(mm.data, mm.rows, mm.cols, self, true);
          },
          [](AccumDiagGmm &s, const DArr &v) {
            if (v.ndim() != 2) throw std::runtime_error("expected a 2-D double array");
            auto &mm = s.variance_accumulator();
            mm.rows = (int32_t)v.shape(0);
            mm.cols = (int32_t)v.shape(1);
            mm.data.assign(v.data(), v.data() + v.size());
          })
      .def("set_zero", &AccumDiagGmm::SetZero, py::arg("flags"))
      .def("scale", &AccumDiagGmm::Scale, py::arg("f"), py::arg("flags"))
      .def("accumulate_for_component",
           [](AccumDiagGmm &s, const FArr &d, int32_t c, float w) { s.AccumulateForComponent(ToVec(d), c, w); },
           py::arg("data"), py::arg("comp_index"), py::arg("weight"))
      .def("accumulate_from_posteriors",
           [](AccumDiagGmm &s, const FArr &d, const FArr &p) { s.AccumulateFromPosteriors(ToVec(d), ToVec(p)); },
           py::arg("data"), py::arg("gauss_posteriors"))
      .def("accumulate_from_diag",
           [](AccumDiagGmm &s, const DiagGmm &g, const FArr &d, float w) { return s.AccumulateFromDiag(g, ToVec(d), w); },
           py::arg("gmm"), py::arg("data"), py::arg("weight"))
      .def("add_stats_for_component",
           [](AccumDiagGmm &s, int32_t g, double occ, const DArr &x, const DArr &x2) {
             s.AddStatsForComponent(g, occ, ToDVec(x), ToDVec(x2));
           },
           py::arg("g"), py::arg("occ"), py::arg("x_stats"), py::arg("x2_stats"))
      .def("add", &AccumDiagGmm::Add, py::arg("scale"), py::arg("acc"));

  m.def(
      "mle_diag_gmm_update",
      [](const MleDiagGmmOptions &config, const AccumDiagGmm &acc, GmmFlagsType flags, DiagGmm *gmm) {
        float oc, c;
        int32_t fe, fg, rg;
        MleDiagGmmUpdate(config, acc, flags, gmm, &oc, &c, &fe, &fg, &rg);
        return std::make_tuple(oc, c, fe, fg, rg);
      },
      py::arg("config"), py::arg("diag_gmm_acc"), py::arg("flags"), py::arg("gmm"));
  m.def("ml_objective", &MlObjective, py::arg("gmm"), py::arg("diaggmm_acc"));

  // ---- AccumAmDiagGmm (python/csrc/mle-am-diag-gmm.cc:15-59) ----
  py::class_<AccumAmDiagGmm>(m, "AccumAmDiagGmm")
      .def(py::init<>())
      .def("init", (void (AccumAmDiagGmm::*)(const AmDiagGmm &, GmmFlagsType))(&AccumAmDiagGmm::Init), py::arg("model"),
           py::arg("flags"))
      .def("init", (void (AccumAmDiagGmm::*)(const AmDiagGmm &, int32_t, GmmFlagsType))(&AccumAmDiagGmm::Init),
           py::arg("model"), py::arg("dim"), py::arg("flags"))
      .def("set_zero", &AccumAmDiagGmm::SetZero, py::arg("flags"))
      .def("accumulate_for_gmm",
           [](AccumAmDiagGmm &s, const AmDiagGmm &model, const FArr &d, int32_t i, float w) {
             return s.AccumulateForGmm(model, ToVec(d), i, w);
           },
           py::arg("model"), py::arg("data"), py::arg("gmm_index"), py::arg("weight"))
      .def("accumulate_for_gmm_two_feats",
           [](AccumAmDiagGmm &s, const AmDiagGmm &model, const FArr &d1, const FArr &d2, int32_t i, float w) {
             return s.AccumulateForGmmTwofeats(model, ToVec(d1), ToVec(d2), i, w);
           },
           py::arg("model"), py::arg("data1"), py::arg("data2"), py::arg("gmm_index"), py::arg("weight"))
      .def("accumulate_from_posteriors",  // 4th kwarg is (mis)named `weight` in the reference (:31-33)
           [](AccumAmDiagGmm &s, const AmDiagGmm &model, const FArr &d, int32_t i, const FArr &post) {
             s.AccumulateFromPosteriors(model, ToVec(d), i, ToVec(post));
           },
           py::arg("model"), py::arg("data"), py::arg("gmm_index"), py::arg("weight"))
      .def("accumulate_for_gaussian",
           [](AccumAmDiagGmm &s, const AmDiagGmm &am, const FArr &d, int32_t i, int32_t g, float w) {
             s.AccumulateForGaussian(am, ToVec(d), i, g, w);
           },
           py::arg("am"), py::arg("data"), py::arg("gmm_index"), py::arg("gauss_index"), py::arg("weight"))
      .def_property_readonly("num_accs", &AccumAmDiagGmm::NumAccs)
      .def_property_readonly("tot_stats_count", &AccumAmDiagGmm::TotStatsCount)
      .def_property_readonly("tot_count", &AccumAmDiagGmm::TotCount)
      .def_property_readonly("tot_log_like", &AccumAmDiagGmm::TotLogLike)
      .def("get_acc", [](AccumAmDiagGmm &s, int32_t index) { return AccumDiagGmm(s.GetAcc(index)); })  // a copy (:41-42)
      .def("add", &AccumAmDiagGmm::Add, py::arg("scale"), py::arg("other"))
      .def("scale", &AccumAmDiagGmm::Scale, py::arg("scale"))
      .def_property_readonly("dim", &AccumAmDiagGmm::Dim)
      // ---- new, batched (what the unchanged-signature script functions call) ----
      .def("accumulate_frames",
           [](AccumAmDiagGmm &s, const AmDiagGmm &model, const FArr &feats, const IArr &pdf_ids, py::object weights) {
             if (feats.ndim() != 2 || feats.shape(1) != model.Dim()) throw std::runtime_error("feats must be (T, dim)");
             if (pdf_ids.ndim() != 1 || pdf_ids.shape(0) != feats.shape(0)) throw std::runtime_error("len(pdf_ids) != num frames");
             FArr w;
             const float *wp = nullptr;
             if (!weights.is_none()) {
               w = FArr::ensure(weights);
               if (!w || w.ndim() != 1 || w.shape(0) != feats.shape(0)) throw std::runtime_error("len(weights) != num frames");
               wp = w.data();
             }
             const float *fp = feats.data();
             const int32_t *ip = pdf_ids.data();
             const int64_t T = feats.shape(0);
             py::gil_scoped_release nogil;  // a host thread pool can feed several GPUs / streams from Python
             return s.AccumulateFrames(model, fp, T, ip, wp);
           },
           py::arg("model"), py::arg("feats"), py::arg("pdf_ids"), py::arg("weights") = py::none())
      .def("accumulate_alignment",
           [](AccumAmDiagGmm &s, const AmDiagGmm &model, const py::object &tm, const FArr &feats, const IArr &ali,
              py::object trans_accs) {
             if (feats.ndim() != 2 || feats.shape(1) != model.Dim()) throw std::runtime_error("feats must be (T, dim)");
             if (ali.ndim() != 1 || ali.shape(0) != feats.shape(0)) throw std::runtime_error("len(ali) != num frames");
             std::vector<int32_t> t2p = Tid2Pdf(tm);
             double *tp = nullptr;
             if (!trans_accs.is_none()) {
               auto ta = trans_accs.cast<py::array_t<double>>();  // must be a real float64 array: updated in place
               if (ta.ndim() != 1 || (size_t)ta.shape(0) != t2p.size() || !ta.writeable())
                 throw std::runtime_error("transition_accs must be a writeable float64 array of size num_transition_ids+1");
               tp = ta.mutable_data();
             }
             const float *fp = feats.data();
             const int32_t *ap = ali.data();
             const int64_t T = feats.shape(0);
             py::gil_scoped_release nogil;
             return s.AccumulateAlignment(model, t2p, fp, T, ap, tp);
           },
           py::arg("model"), py::arg("transition_model"), py::arg("feats"), py::arg("ali"),
           py::arg("transition_accs") = py::none())
      .def("flush", &AccumAmDiagGmm::Flush)
      // diagnostics: True while every accumulated statistic is still on the device and none was folded into the host
      // accumulators — the condition under which mle_am_diag_gmm_update runs the device M-step
      .def_property_readonly("stats_on_device", &AccumAmDiagGmm::StatsOnDevice)
      // per-pdf occupancies for mix-up / mix-down without moving the mean / variance statistics
      .def("pdf_occupancies", [](const AccumAmDiagGmm &s) {
        FloatVector v = s.PdfOccupancies();
        py::array_t<float> out((py::ssize_t)v.size());
        std::copy(v.begin(), v.end(), out.mutable_data());
        return out;
      });

  m.def(
      "mle_am_diag_gmm_update",
      [](const MleDiagGmmOptions &config, const AccumAmDiagGmm &acc, GmmFlagsType flags, AmDiagGmm *am_gmm) {
        float oc, c;
        {
          py::gil_scoped_release nogil;
          MleAmDiagGmmUpdate(config, acc, flags, am_gmm, &oc, &c);
        }
        return std::make_pair(oc, c);
      },
      py::arg("config"), py::arg("amdiag_gmm_acc"), py::arg("flags"), py::arg("am_gmm"));

  // ---- AlignConfig (python/csrc/decoder-wrappers.cc:15-23; csrc/decoder-wrappers.h:22-36) ----
  struct AlignConfig {
    float beam, retry_beam;
    bool careful;
  };
  py::class_<AlignConfig>(m, "AlignConfig")
      .def(py::init([](float beam, float retry_beam, bool careful) { return AlignConfig{beam, retry_beam, careful}; }),
           py::arg("beam") = 200.0f, py::arg("retry_beam") = 0.0f, py::arg("careful") = false)
      .def_readwrite("beam", &AlignConfig::beam)
      .def_readwrite("retry_beam", &AlignConfig::retry_beam)
      .def_readwrite("careful", &AlignConfig::careful);

  // ---- decodables (python/csrc/decodable-itf.cc, decodable-am-diag-gmm.cc) ----
  py::class_<DecodableInterface, PyDecodableInterface>(m, "DecodableInterface")
      .def(py::init<>())
      .def("log_likelihood", &DecodableInterface::LogLikelihood, py::arg("frame"), py::arg("index"))
      .def("is_last_frame", &DecodableInterface::IsLastFrame, py::arg("frame"))
      .def("num_frames_ready", &DecodableInterface::NumFramesReady)
      .def("num_indices", &DecodableInterface::NumIndices);

  py::class_<DecodableAmDiagGmmUnmapped, DecodableInterface>(m, "DecodableAmDiagGmmUnmapped")
      .def(py::init([](const AmDiagGmm &am, const FArr &feats, float prune) {
             return new DecodableAmDiagGmmUnmapped(am, ToMat(feats), prune);
           }),
           py::arg("am"), py::arg("feats"), py::arg("log_sum_exp_prune") = -1.0)
      .def_property_readonly("log_like_block", [](py::object self) {
        auto &d = self.cast<DecodableAmDiagGmmUnmapped &>();
        auto &blk = const_cast<std::vector<float> &>(d.LogLikeBlock());
        return View2(blk, d.NumIndices() > 0 && d.NumFramesReady() > 0 ? (int32_t)(blk.size() / d.NumFramesReady()) : 0,
                     d.NumFramesReady(), self, false);
      });

  // the object handed in as `tm` is kept and returned by .transition_model, like the reference's
  // TransModel() (python/csrc/decodable-am-diag-gmm.cc:26, csrc/decodable-am-diag-gmm.h:104)
  struct ScaledWithTm : DecodableAmDiagGmmScaled {
    using DecodableAmDiagGmmScaled::DecodableAmDiagGmmScaled;
    py::object tm;
  };
  py::class_<DecodableAmDiagGmmScaled, DecodableAmDiagGmmUnmapped>(m, "DecodableAmDiagGmmScaled")
      .def(py::init([](const AmDiagGmm &am, const py::object &tm, const FArr &feats, float scale, float prune) {
             auto *d = new ScaledWithTm(am, Tid2Pdf(tm), ToMat(feats), scale, prune);
             d->tm = tm;
             return static_cast<DecodableAmDiagGmmScaled *>(d);
           }),
           py::arg("am"), py::arg("tm"), py::arg("feats"), py::arg("scale"), py::arg("log_sum_exp_prune") = -1.0)
      .def_property_readonly("transition_model", [](DecodableAmDiagGmmScaled &self) -> py::object {
        auto *d = dynamic_cast<ScaledWithTm *>(&self);
        return d ? d->tm : py::none();
      })
      // new: wrap one utterance's (num_pdfs x num_frames) slice of a batched all-pdf block
      .def_static(
          "from_block",
          [](const FArr &block, const py::object &tm, float scale) {
            if (block.ndim() != 2) throw std::runtime_error("block must be (num_pdfs, num_frames)");
            std::vector<float> b(block.data(), block.data() + block.size());
            auto *d = new ScaledWithTm(std::move(b), (int32_t)block.shape(0), (int32_t)block.shape(1), Tid2Pdf(tm), scale);
            d->tm = tm;
            return static_cast<DecodableAmDiagGmmScaled *>(d);
          },
          py::arg("block"), py::arg("tm"), py::arg("scale"));
}
