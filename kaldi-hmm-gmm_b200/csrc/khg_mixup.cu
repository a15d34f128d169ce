// kaldi-hmm-gmm_b200/csrc/khg_mixup.cu — khg_model_split_by_count: AmDiagGmm::SplitByCount on the
// packed device model, so that E-step -> M-step -> mix-up -> next E-step never
// leaves the device.
//
// Reference (paths relative to kaldi-hmm-gmm/csrc/):
//   AmDiagGmm::SplitByCount  am-diag-gmm.cc:72-89   (pdfs below their target are split up to it)
//   GetSplitTargets          model-common.cc:14-70  (power-law allocation by a priority queue)
//   DiagGmm::Split           diag-gmm.cc:780-851    (halve the heaviest weight, perturb the two
//                                                    means_invvars by +-perturb * randn * sqrt(inv_var))
// The allocation is host logic (a priority queue over P floats, same comparator as the reference);
// the splits run on the device, one CTA per pdf (a pdf's splits depend on each other through the
// weights, different pdfs are independent).  The standard-normal draws are an INPUT (one row of
// `dim` values per new Gaussian, consumed in the reference's order: pdf by pdf, split by split) so
// that results can be compared with the reference's arithmetic; NULL = drawn here (mt19937_64).
#include <algorithm>
#include <cmath>
#include <queue>
#include <random>
#include <vector>

#include "khg_internal.h"

namespace khg {

// model-common.cc:14-27
struct CountStats {
  int32_t pdf_index, num_components;
  float occupancy;
  bool operator<(const CountStats &o) const {
    return occupancy / (num_components + 1.0e-10) < o.occupancy / (o.num_components + 1.0e-10);
  }
};

// model-common.cc:29-70
static void get_split_targets(const float *state_occs, int32_t num_pdfs, int32_t target_components, float power, float min_count,
                              std::vector<int32_t> *targets) {
  std::priority_queue<CountStats> q;
  for (int32_t p = 0; p < num_pdfs; ++p) q.push(CountStats{p, 1, (float)pow(state_occs[p], power)});
  for (int32_t num_gauss = num_pdfs; num_gauss < target_components;) {
    CountStats s = q.top();
    if (s.occupancy == 0) break;  // "Could not split up to ... due to min-count (or no counts at all)"
    q.pop();
    const float orig_occ = state_occs[s.pdf_index];
    if ((s.num_components + 1) * min_count >= orig_occ) {
      s.occupancy = 0;  // min-count active: no more splits of this pdf
    } else {
      ++s.num_components;
      ++num_gauss;
    }
    q.push(s);
  }
  targets->assign(num_pdfs, 0);
  while (!q.empty()) {
    (*targets)[q.top().pdf_index] = q.top().num_components;
    q.pop();
  }
}

// One CTA per pdf: copy the pdf's Gaussians to their new place, then DiagGmm::Split.
__global__ void __launch_bounds__(128) split_kernel(int P, int D, const int32_t *__restrict__ old_off, const int32_t *__restrict__ new_off,
                                                    const int32_t *__restrict__ rand_row0, const float *__restrict__ randn,
                                                    float perturb, const float *__restrict__ w_old,
                                                    const float *__restrict__ miv_old, const float *__restrict__ iv_old,
                                                    float *__restrict__ w_new, float *__restrict__ miv_new, float *__restrict__ iv_new) {
  const int p = blockIdx.x;
  const int o0 = old_off[p], n_old = old_off[p + 1] - o0, n0 = new_off[p], n_new = new_off[p + 1] - n0;
  const int tid = threadIdx.x;
  __shared__ int s_max;
  for (int i = tid; i < n_old; i += blockDim.x) w_new[n0 + i] = w_old[o0 + i];
  for (int e = tid; e < n_old * D; e += blockDim.x) {
    miv_new[(size_t)n0 * D + e] = miv_old[(size_t)o0 * D + e];
    iv_new[(size_t)n0 * D + e] = iv_old[(size_t)o0 * D + e];
  }
  __syncthreads();
  float *w = w_new + n0;
  float *miv = miv_new + (size_t)n0 * D, *iv = iv_new + (size_t)n0 * D;
  for (int cur = n_old; cur < n_new; ++cur) {
    if (tid < 32) {  // the heaviest component, the first one among equals (diag-gmm.cc:806-813)
      float bw = -1.f;
      int bi = 0x7fffffff;
      for (int i = tid; i < cur; i += 32) {
        const float v = w[i];
        if (v > bw) { bw = v; bi = i; }
      }
      for (int s = 16; s > 0; s >>= 1) {
        const float ow = __shfl_xor_sync(0xffffffffu, bw, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (ow > bw || (ow == bw && oi < bi)) { bw = ow; bi = oi; }
      }
      if (tid == 0) {
        s_max = bi;
        const float half = w[bi] / 2;  // weights_[max_idx] /= 2; weights_[cur] = weights_[max_idx]
        w[bi] = half;
        w[cur] = half;
      }
    }
    __syncthreads();
    const int mx = s_max;
    const float *rn = randn + (size_t)(rand_row0[p] + (cur - n_old)) * D;
    for (int d = tid; d < D; d += blockDim.x) {
      const float v = iv[(size_t)mx * D + d];
      const float r = __fmul_rn(__fmul_rn(rn[d], sqrtf(v)), perturb);  // rand * sqrt(inv_var) * perturb_factor
      const float m = miv[(size_t)mx * D + d];
      iv[(size_t)cur * D + d] = v;
      miv[(size_t)cur * D + d] = __fadd_rn(m, r);
      miv[(size_t)mx * D + d] = __fsub_rn(m, r);
    }
    __syncthreads();
  }
}

}  // namespace khg

using namespace khg;

extern "C" khg_status khg_model_split_by_count(khg_model *m, const float *state_occs, int32_t target_components,
                                               float perturb_factor, float power, float min_count, const float *randn,
                                               int64_t randn_rows, uint64_t seed, khg_model **new_model,
                                               int32_t *num_gauss_out) {
  KHG_REQUIRE(m && m->uploaded && state_occs && new_model, "bad argument");
  const int P = m->P, D = m->dim;
  std::vector<int32_t> targets;
  get_split_targets(state_occs, P, target_components, power, min_count, &targets);
  std::vector<int32_t> new_off(P + 1, 0), row0(P, 0);
  int64_t rows = 0;
  for (int p = 0; p < P; ++p) {
    const int n_old = m->h_offsets[p + 1] - m->h_offsets[p];
    const int n_new = n_old < targets[p] ? targets[p] : n_old;  // am-diag-gmm.cc:80-83
    new_off[p + 1] = new_off[p] + n_new;
    row0[p] = (int32_t)rows;
    rows += n_new - n_old;
  }
  KHG_REQUIRE(!randn || randn_rows >= rows, "randn has fewer rows than Gaussians to create");
  std::vector<float> drawn;
  if (!randn && rows > 0) {
    std::mt19937_64 gen(seed);
    std::normal_distribution<float> nd(0.f, 1.f);
    drawn.resize((size_t)rows * D);
    for (auto &v : drawn) v = nd(gen);
    randn = drawn.data();
  }
  khg_model *nm = nullptr;
  KHG_TRY(khg_model_create(D, P, new_off.data(), &nm));
  nm->stream = m->stream;
  nm->kernel = m->kernel;
  cudaStream_t st = m->stream;
  DevTmp tmp;
  float *d_rand = nullptr;
  int32_t *d_row0 = nullptr;
  khg_status s = tmp.alloc(&d_rand, std::max<size_t>(1, (size_t)rows * D));
  if (s == KHG_OK) s = tmp.alloc(&d_row0, (size_t)P);
  if (s != KHG_OK) { khg_model_destroy(nm); return s; }
  cudaError_t ce = cudaSuccess;
  if (rows > 0) ce = cudaMemcpyAsync(d_rand, randn, sizeof(float) * (size_t)rows * D, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_row0, row0.data(), sizeof(int32_t) * P, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) {
    split_kernel<<<P, 128, 0, st>>>(P, D, m->d_offsets, nm->d_offsets, d_row0, d_rand, perturb_factor, m->d_weights, m->d_miv,
                                    m->d_iv, nm->d_weights, nm->d_miv, nm->d_iv);
    ++g_launch_count;
    ce = cudaGetLastError();
  }
  if (ce != cudaSuccess) {
    set_error(std::string("split_by_count: ") + cudaGetErrorString(ce));
    khg_model_destroy(nm);
    return KHG_ERR_CUDA;
  }
  s = finish_model_from_device(nm, nullptr);  // ComputeGconsts() at the end of every Split (diag-gmm.cc:850)
  if (s != KHG_OK) { khg_model_destroy(nm); return s; }
  if (num_gauss_out) *num_gauss_out = nm->G;
  *new_model = nm;
  return KHG_OK;
}
