// kaldi-hmm-gmm_b200/csrc/khg_mixup.cu — khg_model_split_by_count / khg_model_merge_by_count:
// AmDiagGmm::SplitByCount and MergeByCount on the packed device model, so that
// E-step -> M-step -> mix-up / mix-down -> next E-step never leaves the device.
//
// Reference (paths relative to kaldi-hmm-gmm/csrc/):
//   AmDiagGmm::SplitByCount  am-diag-gmm.cc:72-89   (pdfs below their target are split up to it)
//   GetSplitTargets          model-common.cc:14-70  (power-law allocation by a priority queue)
//   DiagGmm::Split           diag-gmm.cc:780-851    (halve the heaviest weight, perturb the two
//                                                    means_invvars by +-perturb * randn * sqrt(inv_var))
//   AmDiagGmm::MergeByCount  am-diag-gmm.cc:91-108  (pdfs above their target are merged down to it)
//   DiagGmm::Merge           diag-gmm.cc:557-746    (greedy pairwise merging by the smallest loss of
//                                                    likelihood; target 1 = global mean and variance)
//   MergedComponentsLogdet   diag-gmm.cc:748-767
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

// Sum over the CTA (128 threads) of one float per thread; all threads get the result.
__device__ __forceinline__ float block_sum_128(float v, float *red) {
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return (red[0] + red[1]) + (red[2] + red[3]);
}

struct MergeArgs {
  int P, D;
  const int32_t *old_off, *new_off;
  const int64_t *delta_off;  // per pdf: first float of its n x n delta_like matrix (pdfs that merge)
  const float *w_old, *miv_old, *iv_old;
  float *w_new, *miv_new, *iv_new;
  float *means, *vars;       // scratch, G x D: first / second-order statistics normalised by the weights
  float *wts, *logdet;       // scratch, G
  float *delta;              // scratch
  int32_t *discarded;        // scratch, G: 0 untouched, 1 merged away, 2 survivor of a merge
};

// -0.5 * sum_d log(var of the merged pair), MergedComponentsLogdet (diag-gmm.cc:748-767); one warp.
__device__ __forceinline__ float merged_logdet_warp(float w1, float w2, const float *f1, const float *f2, const float *s1,
                                                    const float *s2, int D) {
  const float w_sum = w1 + w2, r21 = w2 / w1, r1s = w1 / w_sum;
  float acc = 0.f;
  for (int d = threadIdx.x & 31; d < D; d += 32) {
    const float tm = __fmul_rn(__fadd_rn(f1[d], __fmul_rn(f2[d], r21)), r1s);
    const float tv = __fsub_rn(__fmul_rn(__fadd_rn(s1[d], __fmul_rn(s2[d], r21)), r1s), __fmul_rn(tm, tm));
    acc += logf(tv);
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  return -0.5f * acc;
}

// One CTA (128 threads = 4 warps) per pdf: DiagGmm::Merge down to the pdf's new Gaussian count.
__global__ void __launch_bounds__(128) merge_kernel(MergeArgs a) {
  const int p = blockIdx.x, D = a.D, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int o0 = a.old_off[p], n = a.old_off[p + 1] - o0, n0 = a.new_off[p], tgt = a.new_off[p + 1] - n0;
  __shared__ float red[4];
  __shared__ float s_best[4];
  __shared__ int s_bi[4], s_bj[4];
  if (tgt == n) {  // untouched
    for (int i = tid; i < n; i += 128) a.w_new[n0 + i] = a.w_old[o0 + i];
    for (int e = tid; e < n * D; e += 128) {
      a.miv_new[(size_t)n0 * D + e] = a.miv_old[(size_t)o0 * D + e];
      a.iv_new[(size_t)n0 * D + e] = a.iv_old[(size_t)o0 * D + e];
    }
    return;
  }
  const float *w_in = a.w_old + o0, *miv_in = a.miv_old + (size_t)o0 * D, *iv_in = a.iv_old + (size_t)o0 * D;
  float *means = a.means + (size_t)o0 * D, *vars = a.vars + (size_t)o0 * D;
  // vars = 1 / inv_vars; means = means_invvars * vars; vars += means^2  (second-order stats)
  for (int e = tid; e < n * D; e += 128) {
    const float v = 1.0f / iv_in[e], mu = __fmul_rn(miv_in[e], v);
    means[e] = mu;
    vars[e] = __fadd_rn(v, __fmul_rn(mu, mu));
  }
  __syncthreads();
  if (tgt == 1) {  // global mean and variance (diag-gmm.cc:571-607)
    float wsum = 0.f;
    for (int i = 0; i < n; ++i) wsum += w_in[i];
    const bool rescale = !(fabsf(wsum - 1.0f) <= 1e-6f * (fabsf(wsum) + 1.0f));  // !ApproxEqual(sum, 1, 1e-6)
    for (int d = tid; d < D; d += 128) {
      float m1 = 0.f, m2 = 0.f;
      for (int i = 0; i < n; ++i) {
        m1 = fmaf(w_in[i], means[(size_t)i * D + d], m1);
        m2 = fmaf(w_in[i], vars[(size_t)i * D + d], m2);
      }
      if (rescale) {  // "Weights sum to ...: rescaling." — the reference multiplies
        m1 *= wsum;
        m2 *= wsum;
      }
      const float iv = 1.0f / __fsub_rn(m2, __fmul_rn(m1, m1));
      a.iv_new[(size_t)n0 * D + d] = iv;
      a.miv_new[(size_t)n0 * D + d] = __fmul_rn(m1, iv);
    }
    if (tid == 0) a.w_new[n0] = rescale ? 1.0f : wsum;
    return;
  }
  float *wts = a.wts + o0, *logdet = a.logdet + o0, *delta = a.delta + a.delta_off[p];
  int32_t *disc = a.discarded + o0;
  for (int i = tid; i < n; i += 128) { wts[i] = w_in[i]; disc[i] = 0; }
  // logdet[i] = 0.5 * sum_d log(inv_var)
  for (int i = warp; i < n; i += 4) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc += logf(iv_in[(size_t)i * D + d]);
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) logdet[i] = 0.5f * acc;
  }
  // working copies of the parameters of the kept components live in the OLD layout of the new
  // arrays' scratch: use the means/vars scratch plus per-component iv/miv recomputed at the end
  __syncthreads();
  // delta_like(i, j), j < i: the change of likelihood if i and j were merged
  for (int pr = warp; pr < n * (n - 1) / 2; pr += 4) {
    int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)pr)) * 0.5f);
    while (i * (i - 1) / 2 > pr) --i;
    while ((i + 1) * i / 2 <= pr) ++i;
    const int j = pr - i * (i - 1) / 2;
    const float w1 = wts[i], w2 = wts[j];
    const float ml = merged_logdet_warp(w1, w2, means + (size_t)i * D, means + (size_t)j * D, vars + (size_t)i * D, vars + (size_t)j * D, D);
    if (lane == 0) {
      const float v = __fsub_rn(__fsub_rn(__fmul_rn(w1 + w2, ml), __fmul_rn(w1, logdet[i])), __fmul_rn(w2, logdet[j]));
      delta[(size_t)i * n + j] = v;
      delta[(size_t)j * n + i] = v;
    }
  }
  __syncthreads();
  for (int removed = 0; removed < n - tgt; ++removed) {
    // the pair with the largest delta_like; the first one in (i ascending, j ascending) order on ties
    float bv = -3.402823466e+38f;
    int bi = -1, bj = -1;
    for (int i = tid; i < n; i += 128) {
      if (disc[i] == 1) continue;
      for (int j = 0; j < i; ++j) {
        if (disc[j] == 1) continue;
        const float v = delta[(size_t)i * n + j];
        if (v > bv) { bv = v; bi = i; bj = j; }
      }
    }
    for (int s = 16; s > 0; s >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, s);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, s), oj = __shfl_xor_sync(0xffffffffu, bj, s);
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && (oi < bi || (oi == bi && oj < bj))))) { bv = ov; bi = oi; bj = oj; }
    }
    if (lane == 0) { s_best[warp] = bv; s_bi[warp] = bi; s_bj[warp] = bj; }
    __syncthreads();
    bv = s_best[0]; bi = s_bi[0]; bj = s_bj[0];
    for (int w = 1; w < 4; ++w) {
      const float ov = s_best[w];
      const int oi = s_bi[w], oj = s_bj[w];
      if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && (oi < bi || (oi == bi && oj < bj))))) { bv = ov; bi = oi; bj = oj; }
    }
    const int mi = bi, mj = bj;  // KHG_ASSERT(max_i != max_j && max_i != -1 && max_j != -1)
    __syncthreads();
    const float w1 = wts[mi], w2 = wts[mj], w_sum = w1 + w2, r21 = w2 / w1;
    float ld_acc = 0.f;
    for (int d = tid; d < D; d += 128) {
      // means.row(i) = (means.row(i) + w2 / w1 * means.row(j)) * w1 / w_sum;  same for vars
      const float mu = __fmul_rn(__fadd_rn(means[(size_t)mi * D + d], __fmul_rn(r21, means[(size_t)mj * D + d])), w1) / w_sum;
      const float sv = __fmul_rn(__fadd_rn(vars[(size_t)mi * D + d], __fmul_rn(r21, vars[(size_t)mj * D + d])), w1) / w_sum;
      means[(size_t)mi * D + d] = mu;
      vars[(size_t)mi * D + d] = sv;
      ld_acc += logf(1.0f / __fsub_rn(sv, __fmul_rn(mu, mu)));
    }
    const float ld = 0.5f * block_sum_128(ld_acc, red);
    if (tid == 0) {
      wts[mi] = w_sum;
      logdet[mi] = ld;
      disc[mj] = 1;
      disc[mi] = 2;  // alive, parameters to be rebuilt from the merged statistics
    }
    __syncthreads();
    for (int j = warp; j < n; j += 4) {
      if (j == mi || disc[j] == 1) continue;
      const float a1 = wts[mi], a2 = wts[j];
      const float ml = merged_logdet_warp(a1, a2, means + (size_t)mi * D, means + (size_t)j * D, vars + (size_t)mi * D, vars + (size_t)j * D, D);
      if (lane == 0) {
        const float v = __fsub_rn(__fsub_rn(__fmul_rn(a1 + a2, ml), __fmul_rn(a1, logdet[mi])), __fmul_rn(a2, logdet[j]));
        delta[(size_t)mi * n + j] = v;
        delta[(size_t)j * n + mi] = v;
      }
    }
    __syncthreads();
  }
  // the kept components, in order: inv_var = 1 / (second-order - mean^2), means_invvars = mean * inv_var
  // for merged components; untouched components keep their parameters bit for bit
  __shared__ int s_slot;
  if (tid == 0) s_slot = 0;
  __syncthreads();
  for (int i = 0; i < n; ++i) {
    if (disc[i] == 1) continue;
    const int slot = s_slot;
    const bool merged = disc[i] == 2;
    for (int d = tid; d < D; d += 128) {
      float iv = iv_in[(size_t)i * D + d], miv = miv_in[(size_t)i * D + d];
      if (merged) {
        const float mu = means[(size_t)i * D + d];
        iv = 1.0f / __fsub_rn(vars[(size_t)i * D + d], __fmul_rn(mu, mu));
        miv = __fmul_rn(mu, iv);
      }
      a.iv_new[(size_t)(n0 + slot) * D + d] = iv;
      a.miv_new[(size_t)(n0 + slot) * D + d] = miv;
    }
    if (tid == 0) a.w_new[n0 + slot] = wts[i];
    __syncthreads();
    if (tid == 0) s_slot = slot + 1;
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

extern "C" khg_status khg_model_merge_by_count(khg_model *m, const float *state_occs, int32_t target_components, float power,
                                               float min_count, khg_model **new_model, int32_t *num_gauss_out) {
  KHG_REQUIRE(m && m->uploaded && state_occs && new_model, "bad argument");
  const int P = m->P, D = m->dim, G = m->G;
  std::vector<int32_t> targets;
  get_split_targets(state_occs, P, target_components, power, min_count, &targets);
  std::vector<int32_t> new_off(P + 1, 0);
  std::vector<int64_t> delta_off(P, 0);
  int64_t delta_floats = 0;
  for (int p = 0; p < P; ++p) {
    const int n_old = m->h_offsets[p + 1] - m->h_offsets[p];
    const int tgt = std::max(1, targets[p]);                  // "can't merge below 1", am-diag-gmm.cc:98
    const int n_new = n_old > tgt ? tgt : n_old;              // :99-100
    new_off[p + 1] = new_off[p] + n_new;
    delta_off[p] = delta_floats;
    if (n_new != n_old && n_new > 1) delta_floats += (int64_t)n_old * n_old;
  }
  khg_model *nm = nullptr;
  KHG_TRY(khg_model_create(D, P, new_off.data(), &nm));
  nm->stream = m->stream;
  nm->kernel = m->kernel;
  cudaStream_t st = m->stream;
  DevTmp tmp;
  MergeArgs a;
  a.P = P; a.D = D;
  a.old_off = m->d_offsets; a.new_off = nm->d_offsets;
  a.w_old = m->d_weights; a.miv_old = m->d_miv; a.iv_old = m->d_iv;
  a.w_new = nm->d_weights; a.miv_new = nm->d_miv; a.iv_new = nm->d_iv;
  int64_t *d_doff = nullptr;
  khg_status s = tmp.alloc(&a.means, (size_t)G * D);
  if (s == KHG_OK) s = tmp.alloc(&a.vars, (size_t)G * D);
  if (s == KHG_OK) s = tmp.alloc(&a.wts, (size_t)G);
  if (s == KHG_OK) s = tmp.alloc(&a.logdet, (size_t)G);
  if (s == KHG_OK) s = tmp.alloc(&a.discarded, (size_t)G);
  if (s == KHG_OK) s = tmp.alloc(&a.delta, (size_t)std::max<int64_t>(1, delta_floats));
  if (s == KHG_OK) s = tmp.alloc(&d_doff, (size_t)P);
  if (s != KHG_OK) { khg_model_destroy(nm); return s; }
  a.delta_off = d_doff;
  cudaError_t ce = cudaMemcpyAsync(d_doff, delta_off.data(), sizeof(int64_t) * P, cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) {
    merge_kernel<<<P, 128, 0, st>>>(a);
    ++g_launch_count;
    ce = cudaGetLastError();
  }
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) {
    set_error(std::string("merge_by_count: ") + cudaGetErrorString(ce));
    khg_model_destroy(nm);
    return KHG_ERR_CUDA;
  }
  s = finish_model_from_device(nm, nullptr);  // ComputeGconsts() at the end of every Merge (diag-gmm.cc:745)
  if (s != KHG_OK) { khg_model_destroy(nm); return s; }
  if (num_gauss_out) *num_gauss_out = nm->G;
  *new_model = nm;
  return KHG_OK;
}
