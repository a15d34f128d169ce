// kaldi-hmm-gmm_b200/csrc/khg_b200.cu — C ABI (include/khg_b200.h) of the B200-native
// diag-GMM E-step.  Host-side orchestration only; the arithmetic is in
// khg_kernels.cuh (SIMT) and khg_loglikes_tc.cu (tcgen05).  No CPU fallback:
// every compute entry point needs a CUDA device.
#include <cub/device/device_radix_sort.cuh>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "khg_host_pool.h"
#include "khg_internal.h"
#include "khg_kernels.cuh"

namespace khg {

static thread_local std::string t_last_error;
void set_error(const std::string &msg) { t_last_error = msg; }
int64_t g_launch_count = 0;

khg_status Buf::reserve(size_t bytes) {
  if (bytes <= cap) return KHG_OK;
  release();
  size_t want = std::max(bytes, (size_t)256);
  if (pinned)
    KHG_CUDA_TRY(cudaMallocHost(&p, want));
  else
    KHG_CUDA_TRY(cudaMalloc(&p, want));
  cap = want;
  return KHG_OK;
}
void Buf::release() {
  if (p) {
    if (pinned) cudaFreeHost(p); else cudaFree(p);
  }
  p = nullptr;
  cap = 0;
}

// NVTX range around every batch entry point (visible in nsys / ncu --nvtx timelines).
struct NvtxRange {
  explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// Reads and clears the latched device error flag after synchronising.
static khg_status sync_check(khg_model *m) {
  KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  int flag = 0;
  KHG_CUDA_TRY(cudaMemcpy(&flag, m->d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (flag) {
    KHG_CUDA_TRY(cudaMemset(m->d_err, 0, sizeof(int)));
    if (flag & ERR_BAD_INDEX) {
      set_error("index out of range (pdf id / transition id): KHG_ASSERT(gmm_index >= 0 && gmm_index < NumAccs())");
      return KHG_ERR_INVALID;
    }
    set_error("Invalid answer (overflow or invalid variances/features?)");
    return KHG_ERR_NONFINITE;
  }
  return KHG_OK;
}

// Returns a device pointer for a caller buffer: the buffer itself (KHG_DEVICE) or
// a staged copy (KHG_HOST).
// Host -> device copy on the model's stream.  A large PAGEABLE source (a numpy array: what a Python caller hands in)
// goes through two pinned slots filled by the pool's threads — the driver's own staging of pageable memory is one
// thread at 6-12 GB/s — the copy of slice i + 1 into its slot running under the DMA of slice i.  Pinned sources and
// small buffers are one cudaMemcpyAsync.  KHG_STAGE_THREADS=1 keeps the plain copy.
khg_status h2d_copy(khg_model *m, void *dst, const void *src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return KHG_OK;
  if (!stream) stream = m->stream;
  constexpr size_t kSlice = 16u << 20;
  int threads = std::min(8, HostPool::get().workers());
  if (const char *e = getenv("KHG_STAGE_THREADS")) threads = std::max(1, std::min(threads, atoi(e)));
  bool pageable = false;
  if (bytes >= kSlice / 4 && threads > 1) {  // (4 MB: below that the driver's own staging is as fast)
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) != cudaSuccess) {
      (void)cudaGetLastError();
      pageable = true;
    } else {
      pageable = at.type == cudaMemoryTypeUnregistered;
    }
  }
  if (!pageable) {
    KHG_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return KHG_OK;
  }
  for (int i = 0; i < 2; ++i) {
    m->pin_stage[i].pinned = true;
    if (m->pin_stage[i].reserve(kSlice) != KHG_OK) {  // no page-locked memory to be had: the driver's own copy
      (void)cudaGetLastError();
      KHG_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
      return KHG_OK;
    }
    if (!m->ev_stage[i]) KHG_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_stage[i], cudaEventDisableTiming));
  }
  const size_t n_slices = (bytes + kSlice - 1) / kSlice;
  for (size_t i = 0; i < n_slices; ++i) {
    const int b = (int)(i & 1);
    const size_t off = i * kSlice, n = std::min(kSlice, bytes - off);
    if (i >= 2) KHG_CUDA_TRY(cudaEventSynchronize(m->ev_stage[b]));  // the DMA out of slot b (slice i - 2) has finished
    const char *s0 = static_cast<const char *>(src) + off;
    char *p0 = m->pin_stage[b].as<char>();
    parallel_memcpy(p0, s0, n, threads);
    KHG_CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(dst) + off, p0, n, cudaMemcpyHostToDevice, stream));
    KHG_CUDA_TRY(cudaEventRecord(m->ev_stage[b], stream));
  }
  // the slots are reused by the next call: wait for the last two DMAs (the data is then on the device; later work on
  // the stream is ordered behind it anyway)
  for (int b = 0; b < 2; ++b) KHG_CUDA_TRY(cudaEventSynchronize(m->ev_stage[b]));
  return KHG_OK;
}

// Device -> host copy of `rows` rows of `width` bytes (pitches in bytes) on the model's stream, complete on return.
// A large PAGEABLE destination (a fresh numpy array: every page of it still to be faulted in) is filled from two pinned
// slots by the pool's threads while the DMA of the next band of rows runs; pinned destinations and small blocks are one
// cudaMemcpy2DAsync + synchronize.
khg_status d2h_copy_2d(khg_model *m, void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t width, size_t rows) {
  if (width == 0 || rows == 0) return KHG_OK;
  constexpr size_t kSlice = 16u << 20, kPiece = 1u << 20;
  int threads = std::min(8, HostPool::get().workers());
  if (const char *e = getenv("KHG_STAGE_THREADS")) threads = std::max(1, std::min(threads, atoi(e)));
  bool staged = width * rows >= kSlice / 2 && width <= kSlice && threads > 1;
  if (staged) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, dst) != cudaSuccess) (void)cudaGetLastError();
    else staged = at.type == cudaMemoryTypeUnregistered;
  }
  if (!staged) {
    KHG_CUDA_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, rows, cudaMemcpyDeviceToHost, m->stream));
    KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
    return KHG_OK;
  }
  for (int i = 0; i < 2; ++i) {
    m->pin_stage[i].pinned = true;
    if (m->pin_stage[i].reserve(kSlice) != KHG_OK) {  // no page-locked memory to be had: the driver's own copy
      (void)cudaGetLastError();
      KHG_CUDA_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width, rows, cudaMemcpyDeviceToHost, m->stream));
      KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
      return KHG_OK;
    }
    if (!m->ev_stage[i]) KHG_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_stage[i], cudaEventDisableTiming));
  }
  const size_t band = std::max<size_t>(1, kSlice / width), n_bands = (rows + band - 1) / band;
  // rows r0 .. r0 + nr of a band wait densely packed in slot b: out to the destination, one piece per job
  auto drain = [&](int b, size_t r0, size_t nr) {
    const char *p0 = m->pin_stage[b].as<char>();
    char *d0 = static_cast<char *>(dst) + r0 * dst_pitch;
    const size_t per_row = (width + kPiece - 1) / kPiece;            // pieces of a long row
    const size_t rows_per_job = std::max<size_t>(1, kPiece / width);  // short rows: several per job
    const int n_jobs = (int)(per_row > 1 ? nr * per_row : (nr + rows_per_job - 1) / rows_per_job);
    std::atomic<int> next{0};
    HostPool::get().run(std::min(threads, std::max(1, n_jobs)), [&](int) {
      for (int q = next.fetch_add(1); q < n_jobs; q = next.fetch_add(1)) {
        if (per_row > 1) {
          const size_t r = (size_t)q / per_row, c0 = ((size_t)q % per_row) * kPiece;
          std::memcpy(d0 + r * dst_pitch + c0, p0 + r * width + c0, std::min(kPiece, width - c0));
        } else {
          const size_t ra = (size_t)q * rows_per_job, rb = std::min(nr, ra + rows_per_job);
          if (dst_pitch == width) std::memcpy(d0 + ra * width, p0 + ra * width, (rb - ra) * width);
          else for (size_t r = ra; r < rb; ++r) std::memcpy(d0 + r * dst_pitch, p0 + r * width, width);
        }
      }
    });
  };
  for (size_t i = 0; i < n_bands; ++i) {
    const int b = (int)(i & 1);
    const size_t r0 = i * band, nr = std::min(band, rows - r0);
    // (slot b was drained two bands ago, by this thread)
    KHG_CUDA_TRY(cudaMemcpy2DAsync(m->pin_stage[b].p, width, static_cast<const char *>(src) + r0 * src_pitch, src_pitch, width, nr,
                                   cudaMemcpyDeviceToHost, m->stream));
    KHG_CUDA_TRY(cudaEventRecord(m->ev_stage[b], m->stream));
    if (i > 0) {
      KHG_CUDA_TRY(cudaEventSynchronize(m->ev_stage[b ^ 1]));
      drain(b ^ 1, (i - 1) * band, band);
    }
  }
  const size_t last = n_bands - 1;
  KHG_CUDA_TRY(cudaEventSynchronize(m->ev_stage[last & 1]));
  drain((int)(last & 1), last * band, rows - last * band);
  return KHG_OK;
}

template <class T>
static khg_status stage_in(khg_model *m, Buf &buf, const T *src, size_t count, int loc, const T **dev) {
  if (src == nullptr) { *dev = nullptr; return KHG_OK; }
  if (loc == KHG_DEVICE) { *dev = src; return KHG_OK; }
  KHG_TRY(buf.reserve(count * sizeof(T)));
  KHG_TRY(h2d_copy(m, buf.p, src, count * sizeof(T)));
  *dev = buf.as<T>();
  return KHG_OK;
}

}  // namespace khg

using namespace khg;

extern "C" {

const char *khg_last_error(void) { return t_last_error.c_str(); }
int32_t khg_abi_version(void) { return 4; }  // 4: + khg_model_stats_kernel
int64_t khg_launch_count(void) { return g_launch_count; }

khg_status khg_device_count(int32_t *count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    set_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return KHG_ERR_CUDA;
  }
  *count = n;
  return KHG_OK;
}

khg_status khg_set_device(int32_t device) {
  KHG_CUDA_TRY(cudaSetDevice(device));
  return KHG_OK;
}

uint16_t khg_augment_flags(uint16_t flags) {  // csrc/model-common.cc:72-84
  flags &= KHG_GMM_ALL;
  if (flags & KHG_GMM_VARIANCES) flags |= KHG_GMM_MEANS;
  if (flags & KHG_GMM_MEANS) flags |= KHG_GMM_WEIGHTS;
  if (!(flags & KHG_GMM_WEIGHTS)) flags |= KHG_GMM_WEIGHTS;
  return flags;
}

// ------------------------------------------------------------------ model --
khg_status khg_model_create(int32_t dim, int32_t num_pdfs, const int32_t *gauss_offsets,
                            khg_model **out) {
  KHG_REQUIRE(out != nullptr && gauss_offsets != nullptr, "null argument");
  KHG_REQUIRE(dim > 0 && num_pdfs > 0, "nmix > 0 && dim > 0");
  KHG_REQUIRE(gauss_offsets[0] == 0, "gauss_offsets[0] == 0");
  int dev_count = 0;
  KHG_TRY(khg_device_count(&dev_count));
  if (dev_count <= 0) {
    set_error("no CUDA device: libkhg_b200 has no CPU fallback");
    return KHG_ERR_CUDA;
  }
  khg_model *m = new khg_model();
  m->dim = dim;
  m->P = num_pdfs;
  m->h_offsets.assign(gauss_offsets, gauss_offsets + num_pdfs + 1);
  for (int p = 0; p < num_pdfs; ++p) {
    int ng = gauss_offsets[p + 1] - gauss_offsets[p];
    if (ng <= 0) {
      delete m;
      set_error("KHG_ASSERT failed: every pdf needs at least one Gaussian (nmix > 0)");
      return KHG_ERR_INVALID;
    }
    m->max_gp = std::max(m->max_gp, ng);
  }
  m->G = gauss_offsets[num_pdfs];
  m->n_chunks = (m->G + kSimtChunk - 1) / kSimtChunk + 1;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev);
  auto fail = [&](cudaError_t e, const char *what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    khg_model_destroy(m);
    return KHG_ERR_CUDA;
  };
  cudaError_t e;
  size_t gd = (size_t)m->G * dim;
  if ((e = cudaMalloc(&m->d_offsets, sizeof(int32_t) * (num_pdfs + 1))) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_weights, sizeof(float) * m->G)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_miv, sizeof(float) * gd)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_iv, sizeof(float) * gd)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_gconsts, sizeof(float) * m->G)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_packT, sizeof(float) * (size_t)m->n_chunks * 2 * dim * kSimtChunk)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_err, sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_scratch_int, sizeof(int) * 4)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMemset(m->d_err, 0, sizeof(int))) != cudaSuccess) return fail(e, "cudaMemset");
  if ((e = cudaMemcpy(m->d_offsets, gauss_offsets, sizeof(int32_t) * (num_pdfs + 1), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy");
  // groups of 8 Gaussians per pdf for K3
  std::vector<int32_t> grp(num_pdfs + 1, 0);
  for (int p = 0; p < num_pdfs; ++p) grp[p + 1] = grp[p] + (gauss_offsets[p + 1] - gauss_offsets[p] + 7) / 8;
  m->h_grp_start = grp;
  if ((e = cudaMalloc(&m->d_grp_start, sizeof(int32_t) * (num_pdfs + 1))) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMemcpy(m->d_grp_start, grp.data(), sizeof(int32_t) * (num_pdfs + 1), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e, "cudaMemcpy");
  if ((e = cudaMalloc(&m->d_pack8, sizeof(float) * (size_t)grp[num_pdfs] * 2 * dim * 8)) != cudaSuccess) return fail(e, "cudaMalloc");
  if ((e = cudaMalloc(&m->d_gc8, sizeof(float) * (size_t)grp[num_pdfs] * 8)) != cudaSuccess) return fail(e, "cudaMalloc");
  *out = m;
  return KHG_OK;
}

khg_status khg_model_upload(khg_model *m, const float *weights, const float *means_invvars,
                            const float *inv_vars, const float *gconsts, int32_t *num_bad) {
  NvtxRange nvtx_range("khg_model_upload");
  KHG_REQUIRE(m && means_invvars && inv_vars, "null argument");
  KHG_REQUIRE(weights || gconsts, "need weights or gconsts");
  if (m->gsel_shadow) {  // derived from the parameters being replaced
    khg_model_destroy(m->gsel_shadow);
    m->gsel_shadow = nullptr;
  }
  size_t gd = (size_t)m->G * m->dim;
  cudaStream_t st = m->stream;
  KHG_CUDA_TRY(cudaMemcpyAsync(m->d_miv, means_invvars, sizeof(float) * gd, cudaMemcpyHostToDevice, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(m->d_iv, inv_vars, sizeof(float) * gd, cudaMemcpyHostToDevice, st));
  if (weights) KHG_CUDA_TRY(cudaMemcpyAsync(m->d_weights, weights, sizeof(float) * m->G, cudaMemcpyHostToDevice, st));
  if (num_bad) *num_bad = 0;
  if (gconsts) {
    KHG_CUDA_TRY(cudaMemcpyAsync(m->d_gconsts, gconsts, sizeof(float) * m->G, cudaMemcpyHostToDevice, st));
  } else {
    KHG_CUDA_TRY(cudaMemsetAsync(m->d_scratch_int, 0, sizeof(int) * 4, st));
    gconsts_kernel<<<grid_for(m->G, 128), 128, 0, st>>>(m->G, m->dim, m->d_weights, m->d_miv, m->d_iv, m->d_gconsts, m->d_scratch_int);
    ++g_launch_count;
    int flags[2] = {0, 0};
    KHG_CUDA_TRY(cudaMemcpyAsync(flags, m->d_scratch_int, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    KHG_CUDA_TRY(cudaStreamSynchronize(st));
    if (flags[1]) {
      set_error("not a number in gconst computation");  // csrc/diag-gmm.cc:132-135
      return KHG_ERR_NONFINITE;
    }
    if (num_bad) *num_bad = flags[0];
  }
  pack_simt_kernel<<<std::min(1024u, grid_for((int64_t)m->n_chunks * 2 * m->dim * kSimtChunk, 256)), 256, 0, st>>>(
      m->G, m->dim, m->n_chunks, m->d_miv, m->d_iv, m->d_packT);
  pack8_kernel<<<m->P, 128, 0, st>>>(m->P, m->dim, m->d_offsets, m->d_grp_start, m->d_miv, m->d_iv, m->d_gconsts, m->d_pack8, m->d_gc8);
  g_launch_count += 2;
  KHG_CUDA_TRY(cudaGetLastError());
  m->uploaded = true;
  m->tc.ready = false;
  stats_tc_free(m);
  if (m->kernel != KHG_KERNEL_SIMT && tc_supported(m)) {
    khg_status s = tc_pack_build(m);
    if (s != KHG_OK && m->kernel >= KHG_KERNEL_TCGEN05) return s;
  }
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  return KHG_OK;
}

khg_status khg_model_get_gconsts(khg_model *m, float *gconsts) {
  KHG_REQUIRE(m && gconsts && m->uploaded, "model not uploaded");
  KHG_CUDA_TRY(cudaMemcpyAsync(gconsts, m->d_gconsts, sizeof(float) * m->G, cudaMemcpyDeviceToHost, m->stream));
  KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return KHG_OK;
}

khg_status khg_model_info(const khg_model *m, int32_t *dim, int32_t *num_pdfs, int32_t *num_gauss) {
  KHG_REQUIRE(m, "null model");
  if (dim) *dim = m->dim;
  if (num_pdfs) *num_pdfs = m->P;
  if (num_gauss) *num_gauss = m->G;
  return KHG_OK;
}

khg_status khg_model_dense_kernel(const khg_model *m, int32_t *kernel) {
  KHG_REQUIRE(m && kernel, "null argument");
  if (m->kernel == KHG_KERNEL_SIMT || !m->tc.ready) *kernel = KHG_KERNEL_SIMT;
  else if (m->kernel == KHG_KERNEL_TCGEN05_F16_GS && gs_supported(m)) *kernel = KHG_KERNEL_TCGEN05_F16_GS;
  else if (m->tc.f16_ready && m->kernel != KHG_KERNEL_TCGEN05) *kernel = KHG_KERNEL_TCGEN05_F16;
  else if (m->tc.tf32_ready) *kernel = KHG_KERNEL_TCGEN05;
  else *kernel = KHG_KERNEL_SIMT;
  return KHG_OK;
}

// the choice acc_device makes (KHG_STATS_KERNEL=simt forces the fp32 kernel: experiments / A-B runs)
static khg_status stats_use_tc(khg_model *m, bool *use_tc) {
  *use_tc = false;
  const char *e = getenv("KHG_STATS_KERNEL");
  if (e && !strcmp(e, "simt")) return KHG_OK;
  KHG_TRY(stats_tc_build(m));
  *use_tc = m->stk.ready;
  return KHG_OK;
}

khg_status khg_model_stats_kernel(khg_model *m, int32_t *kernel) {
  KHG_REQUIRE(m && kernel && m->uploaded, "model not uploaded");
  bool use_tc = false;
  KHG_TRY(stats_use_tc(m, &use_tc));
  *kernel = use_tc ? KHG_KERNEL_TCGEN05_F16 : KHG_KERNEL_SIMT;
  return KHG_OK;
}

khg_status khg_model_set_kernel(khg_model *m, int32_t kernel) {
  KHG_REQUIRE(m && kernel >= KHG_KERNEL_AUTO && kernel <= KHG_KERNEL_TCGEN05_F16_GS, "bad kernel id");
  if (kernel >= KHG_KERNEL_TCGEN05 && !tc_supported(m)) {
    set_error("tcgen05 kernel does not support this model shape (needs 2*dim+2 <= 160 for the tf32 split / <= 320 for the fp16 split; pdfs of more than 240 Gaussians need the tf32 shape)");
    return KHG_ERR_UNSUPPORTED;
  }
  m->kernel = kernel;
  if (kernel != KHG_KERNEL_SIMT && m->uploaded && !m->tc.ready && tc_supported(m)) return tc_pack_build(m);
  return KHG_OK;
}

khg_status khg_model_set_stream(khg_model *m, void *cuda_stream) {
  KHG_REQUIRE(m, "null model");
  m->stream = static_cast<cudaStream_t>(cuda_stream);
  return KHG_OK;
}

khg_status khg_model_sync(khg_model *m) {
  KHG_REQUIRE(m, "null model");
  return sync_check(m);
}

void khg_model_destroy(khg_model *m) {
  if (!m) return;
  if (m->gsel_shadow) khg_model_destroy(m->gsel_shadow);
  tc_pack_free(m);
  stats_tc_free(m);
  align_cache_free(m);
  cudaFree(m->d_offsets); cudaFree(m->d_weights); cudaFree(m->d_miv); cudaFree(m->d_iv);
  cudaFree(m->d_gconsts); cudaFree(m->d_packT); cudaFree(m->d_err); cudaFree(m->d_scratch_int);
  cudaFree(m->d_grp_start); cudaFree(m->d_pack8); cudaFree(m->d_gc8);
  for (Buf *b : {&m->w_feats, &m->w_ids, &m->w_wts, &m->w_out, &m->w_pf, &m->w_keys, &m->w_vals_in,
                 &m->w_vals_out, &m->w_cub, &m->w_starts, &m->w_item_start, &m->w_tot, &m->w_tid,
                 &m->w_tid2pdf, &m->w_trans, &m->w_keys_out, &m->w_sub, &m->w_full, &m->w_al_graph,
                 &m->w_al_block, &m->w_al_bp, &m->w_al_cost, &m->w_al_ali, &m->w_al_path, &m->w_al_xlist, &m->w_al_xll, &m->w_item_desc, &m->w_fb_items, &m->w_al_tiles, &m->pin_al_tiles})
    b->release();
  for (int i = 0; i < 2; ++i) {
    m->pin_feats[i].release(); m->pin_ids[i].release(); m->pin_wts[i].release();
    m->w_efeats[i].release(); m->w_eids[i].release(); m->w_ewts[i].release();
    m->pin_stage[i].release();
    if (m->ev_stage[i]) cudaEventDestroy(m->ev_stage[i]);
    if (m->ev_copy[i]) cudaEventDestroy(m->ev_copy[i]);
    if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
  }
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  delete m;
}

khg_status khg_compute_gconsts(int32_t nmix, int32_t dim, const float *weights,
                               const float *means_invvars, const float *inv_vars,
                               float *gconsts, int32_t *num_bad) {
  KHG_REQUIRE(nmix > 0 && dim > 0 && weights && means_invvars && inv_vars && gconsts, "bad argument");
  for (int i = 0; i < nmix; ++i) KHG_REQUIRE(weights[i] >= 0, "weights_[mix] >= 0");  // csrc/diag-gmm.cc:116
  int32_t offs[2] = {0, nmix};
  khg_model *m = nullptr;
  KHG_TRY(khg_model_create(dim, 1, offs, &m));
  m->kernel = KHG_KERNEL_SIMT;
  khg_status s = khg_model_upload(m, weights, means_invvars, inv_vars, nullptr, num_bad);
  if (s == KHG_OK) s = khg_model_get_gconsts(m, gconsts);
  khg_model_destroy(m);
  return s;
}

// ------------------------------------------------------------ likelihoods --
// fp32 SIMT dense kernel; out[p*sp + t*stt].  gate (optional): see loglikes_simt_kernel.
static khg_status simt_launch(khg_model *m, const float *d_feats, int64_t T, float scale, float *d_out, int64_t sp,
                              int64_t stt, const unsigned *gate, float gate_limit) {
  const int D = m->dim;
  size_t smem = sizeof(float) * ((size_t)D * kDenseXP + 4 + 2 * (size_t)D * kSimtChunk);
  if (smem > 220 * 1024) {
    set_error("feature dimension too large for the dense SIMT kernel");
    return KHG_ERR_UNSUPPORTED;
  }
  KHG_CUDA_TRY(cudaFuncSetAttribute(loglikes_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));  // (per device: set on every call)
  int64_t n_ft = (T + kDenseFrames - 1) / kDenseFrames;
  // enough CTAs for >= 2 waves when T is small: split the pdf range
  int groups = (int)std::min<int64_t>(std::max<int64_t>(1, (4LL * m->sm_count + n_ft - 1) / n_ft), std::max(1, m->P / 4));
  int ppg = (m->P + groups - 1) / groups;
  groups = (m->P + ppg - 1) / ppg;
  for (int64_t f0 = 0; f0 < n_ft; f0 += 65535 * 32) {  // gridDim.x is huge, but keep launches bounded
    int64_t nf = std::min<int64_t>(n_ft - f0, 65535LL * 32);
    dim3 grid((unsigned)nf, groups);
    loglikes_simt_kernel<<<grid, 128, smem, m->stream>>>(
        d_feats + f0 * kDenseFrames * D, T - f0 * kDenseFrames, D, m->d_packT, m->d_gconsts, m->d_offsets,
        m->P, ppg, scale, d_out + f0 * kDenseFrames * stt, sp, stt, m->d_err, gate, gate_limit);
    ++g_launch_count;
  }
  KHG_CUDA_TRY(cudaGetLastError());
  return KHG_OK;
}

constexpr int kSubsetMinTilesPerSm = 1;  // the tile subset needs n_splits == 1 in tc_launch: at least 2 frame tiles per SM
static khg_status dense_device(khg_model *m, const float *d_feats, int64_t T, float scale,
                               int layout, float *d_out, int64_t ld, const TileSubset *subset = nullptr, bool *subset_used = nullptr) {
  if (subset_used) *subset_used = false;
  const bool use_tc = m->kernel != KHG_KERNEL_SIMT && m->tc.ready;
  const int prec = m->kernel == KHG_KERNEL_TCGEN05 ? 1 : (m->kernel == KHG_KERNEL_TCGEN05_F16 ? 2 : (m->kernel == KHG_KERNEL_TCGEN05_F16_GS ? 3 : 0));
  if (m->kernel >= KHG_KERNEL_TCGEN05 && !m->tc.ready) {
    set_error("tcgen05 kernel requested but its model pack is not built");
    return KHG_ERR_UNSUPPORTED;
  }
  if (use_tc) {
    // The tensor-core kernel writes pdf-major rows of its (virtual) pdfs.  Directly into d_out when that is the layout
    // asked for and no pdf is split; otherwise through a bounded scratch block, a chunk of frames at a time:
    // transposed (frame-major) or merged (pdfs of more than 240 Gaussians run as virtual pdfs of unscaled
    // log-sum-exps, merge_virtual_kernel combines and scales them into either layout).
    const int Pv = m->tc.Pv;
    const bool split = Pv != m->P;
    const unsigned *gate = nullptr;
    float gate_limit = 0.f;
    if (layout == KHG_PDF_MAJOR && !split) {
      // (the tile subset of the batched aligner: honoured by the frame-stationary kernels with whole tiles per CTA,
      // and only when no gated fp32 fall-back is involved; otherwise everything is computed)
      const bool sub = subset && subset->off && prec != 3 && (m->tc.tf32_ready || prec == 2) && T >= 2LL * kSubsetMinTilesPerSm * m->sm_count * 128;
      KHG_TRY(tc_loglikes(m, d_feats, T, scale, d_out, ld, prec, &gate, &gate_limit, sub ? subset : nullptr));
      if (subset_used) *subset_used = sub;
      // shapes the tf32 split cannot take (2*dim+1 > 160): the fp16 split's out-of-range
      // fall-back is the fp32 SIMT kernel, gated on the same device word
      if (gate != nullptr) KHG_TRY(simt_launch(m, d_feats, T, scale, d_out, ld, 1, gate, gate_limit));
      return KHG_OK;
    }
    int64_t cap = std::max<int64_t>(1024, ((int64_t)(512e6 / (4.0 * Pv))) & ~(int64_t)1023);  // frames per scratch block (~512 MB)
    if (const char *e = getenv("KHG_DENSE_SCRATCH_FRAMES")) cap = std::max<int64_t>(128, atoll(e) & ~(int64_t)127);  // (tests: several blocks)
    const int64_t chunk = std::min<int64_t>(cap, (T + 3) & ~(int64_t)3);
    KHG_TRY(m->w_out.reserve(sizeof(float) * (size_t)Pv * chunk));
    float *scr = m->w_out.as<float>();
    for (int64_t t0 = 0; t0 < T; t0 += chunk) {
      const int64_t n = std::min(chunk, T - t0);
      KHG_TRY(tc_loglikes(m, d_feats + t0 * m->dim, n, split ? 1.0f : scale, scr, chunk, prec, &gate, &gate_limit));
      if (gate != nullptr) {
        if (split) {
          set_error("internal: virtual pdfs with a gated fp32 fall-back");
          return KHG_ERR_UNSUPPORTED;
        }
        KHG_TRY(simt_launch(m, d_feats + t0 * m->dim, n, scale, scr, chunk, 1, gate, gate_limit));
      }
      float *dst = layout == KHG_PDF_MAJOR ? d_out + t0 : d_out + t0 * ld;
      if (split) {
        dim3 grid(grid_for(n, 32), grid_for(m->P, 32)), block(32, 8);
        merge_virtual_kernel<<<grid, block, 0, m->stream>>>(scr, chunk, m->tc.d_vfirst, m->P, n, scale, dst, layout == KHG_PDF_MAJOR ? ld : 1,
                                                            layout == KHG_PDF_MAJOR ? 1 : ld, m->d_err);
      } else {
        dim3 grid(grid_for(n, 32), grid_for(m->P, 32)), block(32, 8);
        transpose_kernel<<<grid, block, 0, m->stream>>>(scr, m->P, n, chunk, dst, ld);
      }
      ++g_launch_count;
      KHG_CUDA_TRY(cudaGetLastError());
    }
    return KHG_OK;
  }
  int64_t sp = layout == KHG_PDF_MAJOR ? ld : 1, stt = layout == KHG_PDF_MAJOR ? 1 : ld;
  return simt_launch(m, d_feats, T, scale, d_out, sp, stt, nullptr, 0.f);
}

khg_status khg_loglikes_all_pdfs(khg_model *m, const float *feats, int64_t T, int32_t feats_loc,
                                 float scale, int32_t layout, float *out, int64_t ld_out,
                                 int32_t out_loc) {
  NvtxRange nvtx_range("khg_loglikes_all_pdfs");
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  KHG_REQUIRE(T >= 0 && (T == 0 || (feats && out)), "null buffer");
  KHG_REQUIRE(layout == KHG_FRAME_MAJOR || layout == KHG_PDF_MAJOR, "bad layout");
  KHG_REQUIRE(ld_out >= (layout == KHG_PDF_MAJOR ? T : m->P), "ld_out too small");
  if (T == 0) return KHG_OK;
  const int D = m->dim;
  if (feats_loc == KHG_DEVICE && out_loc == KHG_DEVICE)
    return dense_device(m, feats, T, scale, layout, out, ld_out);
  // host buffers: stream in chunks of frames
  const int64_t chunk = std::max<int64_t>(256, std::min<int64_t>(T, (int64_t)(256e6 / (4.0 * m->P))) & ~(int64_t)255);
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    int64_t n = std::min(chunk, T - t0);
    const float *d_f = nullptr;
    KHG_TRY(stage_in(m, m->w_feats, feats + t0 * D, (size_t)n * D, feats_loc, &d_f));
    float *d_o;
    int64_t ldd;
    if (out_loc == KHG_DEVICE) {
      d_o = layout == KHG_PDF_MAJOR ? out + t0 : out + t0 * ld_out;
      ldd = ld_out;
    } else {
      ldd = layout == KHG_PDF_MAJOR ? ((n + 3) & ~(int64_t)3) : m->P;
      KHG_TRY(m->w_pf.reserve(sizeof(float) * (size_t)(layout == KHG_PDF_MAJOR ? m->P * ldd : n * ldd)));
      d_o = m->w_pf.as<float>();
    }
    KHG_TRY(dense_device(m, d_f, n, scale, layout, d_o, ldd));
    if (out_loc == KHG_HOST) {
      if (layout == KHG_PDF_MAJOR)
        KHG_TRY(d2h_copy_2d(m, out + t0, sizeof(float) * ld_out, d_o, sizeof(float) * ldd, sizeof(float) * n, m->P));
      else
        KHG_TRY(d2h_copy_2d(m, out + t0 * ld_out, sizeof(float) * ld_out, d_o, sizeof(float) * ldd, sizeof(float) * m->P, n));
    }
    KHG_TRY(sync_check(m));
  }
  return KHG_OK;
}

khg_status khg_loglikes_pdf_subset(khg_model *m, const float *feats, int64_t T, int32_t feats_loc, const int32_t *pdf_subset,
                                   int32_t n_subset, float scale, float *out, int64_t ld_out, int32_t out_loc) {
  NvtxRange nvtx_range("khg_loglikes_pdf_subset");
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  KHG_REQUIRE(T >= 0 && n_subset >= 0 && (T == 0 || n_subset == 0 || (feats && out && pdf_subset)), "null buffer");
  KHG_REQUIRE(ld_out >= T, "ld_out too small");
  if (T == 0 || n_subset == 0) return KHG_OK;
  const int D = m->dim;
  cudaStream_t st = m->stream;
  KHG_TRY(m->w_sub.reserve(sizeof(int32_t) * n_subset));
  KHG_CUDA_TRY(cudaMemcpyAsync(m->w_sub.p, pdf_subset, sizeof(int32_t) * n_subset, cudaMemcpyHostToDevice, st));
  const int64_t chunk = std::max<int64_t>(256, std::min<int64_t>(T, (int64_t)(512e6 / (4.0 * m->P))) & ~(int64_t)255);
  for (int64_t t0 = 0; t0 < T; t0 += chunk) {
    const int64_t n = std::min(chunk, T - t0);
    const float *d_f = nullptr;
    KHG_TRY(stage_in(m, m->w_feats, feats + t0 * D, (size_t)n * D, feats_loc, &d_f));
    const int64_t ldf = (n + 3) & ~(int64_t)3;
    KHG_TRY(m->w_full.reserve(sizeof(float) * (size_t)m->P * ldf));
    KHG_TRY(dense_device(m, d_f, n, scale, KHG_PDF_MAJOR, m->w_full.as<float>(), ldf));
    float *d_o = out + t0;
    int64_t ldo = ld_out;
    if (out_loc == KHG_HOST) {
      KHG_TRY(m->w_pf.reserve(sizeof(float) * (size_t)n_subset * ldf));
      d_o = m->w_pf.as<float>();
      ldo = ldf;
    }
    dim3 grid((unsigned)std::min<int64_t>(64, (n + 255) / 256), n_subset);
    gather_rows_kernel<<<grid, 256, 0, st>>>(m->w_full.as<float>(), ldf, m->w_sub.as<int32_t>(), n_subset, n, d_o, ldo, m->P, m->d_err);
    ++g_launch_count;
    KHG_CUDA_TRY(cudaGetLastError());
    if (out_loc == KHG_HOST)
      KHG_TRY(d2h_copy_2d(m, out + t0, sizeof(float) * ld_out, d_o, sizeof(float) * ldo, sizeof(float) * n, n_subset));
    KHG_TRY(sync_check(m));
  }
  return KHG_OK;
}

khg_status khg_pdf_loglikes(khg_model *m, int32_t pdf, const float *feats, int64_t T, int32_t loc, float *out) {
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  KHG_REQUIRE(pdf >= 0 && pdf < m->P, "pdf_index out of range");
  KHG_REQUIRE(T > 0 && feats && out, "data.rows() != 0");  // csrc/diag-gmm.cc:179
  const int g0 = m->h_offsets[pdf], ng = m->h_offsets[pdf + 1] - g0, D = m->dim;
  const float *d_f = nullptr;
  KHG_TRY(stage_in(m, m->w_feats, feats, (size_t)T * D, loc, &d_f));
  float *d_o = out;
  if (loc == KHG_HOST) {
    KHG_TRY(m->w_pf.reserve(sizeof(float) * (size_t)T * ng));
    d_o = m->w_pf.as<float>();
  }
  pdf_loglikes_kernel<<<grid_for(T * ng, 128), 128, 0, m->stream>>>(d_f, T, D, m->d_miv + (size_t)g0 * D, m->d_iv + (size_t)g0 * D, m->d_gconsts + g0, ng, d_o);
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  if (loc == KHG_HOST) {
    KHG_CUDA_TRY(cudaMemcpyAsync(out, d_o, sizeof(float) * (size_t)T * ng, cudaMemcpyDeviceToHost, m->stream));
    KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  }
  return KHG_OK;
}

khg_status khg_pdf_posteriors(khg_model *m, int32_t pdf, const float *feats, int64_t T, int32_t loc,
                              float *post, float *loglike) {
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  KHG_REQUIRE(pdf >= 0 && pdf < m->P, "pdf_index out of range");
  KHG_REQUIRE(T > 0 && feats && (post || loglike), "null buffer");
  const int g0 = m->h_offsets[pdf], ng = m->h_offsets[pdf + 1] - g0, D = m->dim;
  const float *d_f = nullptr;
  KHG_TRY(stage_in(m, m->w_feats, feats, (size_t)T * D, loc, &d_f));
  KHG_TRY(m->w_out.reserve(sizeof(float) * (size_t)T * ng));
  float *d_ll = m->w_out.as<float>();
  float *d_post = post, *d_like = loglike;
  if (loc == KHG_HOST) {
    KHG_TRY(m->w_pf.reserve(sizeof(float) * (size_t)T));
    d_post = post ? d_ll : nullptr;
    d_like = loglike ? m->w_pf.as<float>() : nullptr;
  }
  pdf_loglikes_kernel<<<grid_for(T * ng, 128), 128, 0, m->stream>>>(d_f, T, D, m->d_miv + (size_t)g0 * D, m->d_iv + (size_t)g0 * D, m->d_gconsts + g0, ng, d_ll);
  pdf_softmax_kernel<<<grid_for(T, 128), 128, 0, m->stream>>>(d_ll, T, ng, d_post, d_like, m->d_err);
  g_launch_count += 2;
  KHG_CUDA_TRY(cudaGetLastError());
  if (loc == KHG_HOST) {
    if (post) KHG_CUDA_TRY(cudaMemcpyAsync(post, d_post, sizeof(float) * (size_t)T * ng, cudaMemcpyDeviceToHost, m->stream));
    if (loglike) KHG_CUDA_TRY(cudaMemcpyAsync(loglike, d_like, sizeof(float) * (size_t)T, cudaMemcpyDeviceToHost, m->stream));
    return sync_check(m);
  }
  return KHG_OK;
}

// ------------------------------------------------------------------ stats --
khg_status khg_stats_create(khg_model *m, uint16_t flags, khg_stats **out) {
  KHG_REQUIRE(m && out, "null argument");
  KHG_REQUIRE((flags & ~KHG_GMM_ALL) == 0, "(flags & ~kGmmAll) == 0");  // csrc/model-common.cc:73
  khg_stats *s = new khg_stats();
  s->model = m;
  s->flags = khg_augment_flags(flags);
  int64_t G = m->G, D = m->dim, n = G;
  if (s->flags & KHG_GMM_MEANS) { s->off_mean = n; n += G * D; }
  if (s->flags & KHG_GMM_VARIANCES) { s->off_var = n; n += G * D; }
  s->off_tot = n;
  n += 2;
  s->n = n;
  cudaError_t e = cudaMalloc(&s->buf, sizeof(double) * n);
  if (e == cudaSuccess) e = cudaMemset(s->buf, 0, sizeof(double) * n);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc(stats): ") + cudaGetErrorString(e));
    delete s;
    return KHG_ERR_CUDA;
  }
  *out = s;
  return KHG_OK;
}

khg_status khg_stats_zero(khg_stats *s) {
  KHG_REQUIRE(s, "null stats");
  KHG_CUDA_TRY(cudaMemsetAsync(s->buf, 0, sizeof(double) * s->n, s->model->stream));
  return KHG_OK;
}

khg_status khg_stats_flags(const khg_stats *s, uint16_t *flags) {
  KHG_REQUIRE(s && flags, "null argument");
  *flags = s->flags;
  return KHG_OK;
}

khg_status khg_stats_device_buffer(khg_stats *s, double **dev_ptr, int64_t *num_doubles) {
  KHG_REQUIRE(s && dev_ptr && num_doubles, "null argument");
  *dev_ptr = s->buf;
  *num_doubles = s->n;
  return KHG_OK;
}

khg_status khg_stats_download(khg_stats *s, double *occ, double *mean, double *var, double *totals) {
  KHG_REQUIRE(s, "null stats");
  khg_model *m = s->model;
  size_t G = m->G, GD = (size_t)m->G * m->dim;
  cudaStream_t st = m->stream;
  if (occ) KHG_CUDA_TRY(cudaMemcpyAsync(occ, s->buf, sizeof(double) * G, cudaMemcpyDeviceToHost, st));
  if (mean && s->off_mean >= 0) KHG_CUDA_TRY(cudaMemcpyAsync(mean, s->buf + s->off_mean, sizeof(double) * GD, cudaMemcpyDeviceToHost, st));
  if (var && s->off_var >= 0) KHG_CUDA_TRY(cudaMemcpyAsync(var, s->buf + s->off_var, sizeof(double) * GD, cudaMemcpyDeviceToHost, st));
  if (totals) KHG_CUDA_TRY(cudaMemcpyAsync(totals, s->buf + s->off_tot, sizeof(double) * 2, cudaMemcpyDeviceToHost, st));
  return sync_check(m);
}

khg_status khg_stats_upload(khg_stats *s, const double *occ, const double *mean, const double *var, const double *totals) {
  KHG_REQUIRE(s, "null stats");
  khg_model *m = s->model;
  size_t G = m->G, GD = (size_t)m->G * m->dim;
  cudaStream_t st = m->stream;
  if (occ) KHG_CUDA_TRY(cudaMemcpyAsync(s->buf, occ, sizeof(double) * G, cudaMemcpyHostToDevice, st));
  if (mean && s->off_mean >= 0) KHG_CUDA_TRY(cudaMemcpyAsync(s->buf + s->off_mean, mean, sizeof(double) * GD, cudaMemcpyHostToDevice, st));
  if (var && s->off_var >= 0) KHG_CUDA_TRY(cudaMemcpyAsync(s->buf + s->off_var, var, sizeof(double) * GD, cudaMemcpyHostToDevice, st));
  if (totals) KHG_CUDA_TRY(cudaMemcpyAsync(s->buf + s->off_tot, totals, sizeof(double) * 2, cudaMemcpyHostToDevice, st));
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  return KHG_OK;
}

khg_status khg_stats_add(khg_stats *dst, float scale, const khg_stats *src) {
  KHG_REQUIRE(dst && src, "null stats");
  KHG_REQUIRE(dst->n == src->n && dst->flags == src->flags, "num_accs == other.NumAccs()");
  axpy_f64_kernel<<<std::min(2048u, grid_for(dst->n, 256)), 256, 0, dst->model->stream>>>(dst->buf, src->buf, (double)scale, dst->n);
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  return KHG_OK;
}

khg_status khg_stats_scale(khg_stats *s, float scale) {
  KHG_REQUIRE(s, "null stats");
  scale_f64_kernel<<<std::min(2048u, grid_for(s->n, 256)), 256, 0, s->model->stream>>>(s->buf, (double)scale, s->n);
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  return KHG_OK;
}

void khg_stats_destroy(khg_stats *s) {
  if (!s) return;
  cudaFree(s->buf);
  delete s;
}

constexpr int64_t kDirectMaxFrames = 2048;  // at or below this the one-launch direct kernel is used

// Bucket + accumulate for device-resident inputs; frames processed in slabs so the
// sort workspace stays bounded.
static khg_status acc_device(khg_model *m, khg_stats *s, const float *d_feats, int64_t T,
                             const int32_t *d_ids, const float *d_w, float *d_pf, double *d_call_like) {
  if (m->max_gp > kStatsMaxGp) {
    set_error("a pdf has more Gaussians than the statistics kernel supports (1600)");
    return KHG_ERR_UNSUPPORTED;
  }
  const int D = m->dim, P = m->P;
  cudaStream_t st = m->stream;
  if (T <= kDirectMaxFrames && m->max_gp <= kDirectMaxGp) {
    // small batch (one utterance): one launch, no bucketing
    DirectArgs da;
    da.feats = d_feats; da.ids = d_ids; da.weights = d_w; da.offsets = m->d_offsets;
    da.miv = m->d_miv; da.iv = m->d_iv; da.gconsts = m->d_gconsts;
    da.occ = s->buf;
    da.mean = s->off_mean >= 0 ? s->buf + s->off_mean : nullptr;
    da.var = s->off_var >= 0 ? s->buf + s->off_var : nullptr;
    da.totals = s->buf + s->off_tot;
    da.call_like = d_call_like;
    da.per_frame = d_pf;
    da.err = m->d_err;
    da.T = (int)T; da.P = P; da.D = D;
    const size_t shm = sizeof(float) * kDirectWarps * (size_t)(D + kDirectMaxGp);
    if (shm <= 48 * 1024) {
      stats_direct_kernel<<<grid_for(T, kDirectWarps), 32 * kDirectWarps, shm, st>>>(da);
      ++g_launch_count;
      KHG_CUDA_TRY(cudaGetLastError());
      return KHG_OK;
    }
  }
  const int64_t slab = 1 << 23;
  int end_bit = 1;
  while ((1 << end_bit) <= P) ++end_bit;  // keys 0..P: P is the sentinel bucket of out-of-range ids
  // smem sized for THIS model: posterior tile for the largest pdf, model staging for
  // at most the groups of the largest pdf (bounded) -> several CTAs per SM for C4-like models
  const int max_groups = (m->max_gp + 7) / 8;
  int grp_batch = std::max(1, std::min(std::min(8, max_groups), 20480 / (2 * D * 8 * 4)));
  const int post_cap = std::min(kStatsPostCapMax, kStatsFrames * stats_pitch(m->max_gp));
  size_t smem = sizeof(float) * ((size_t)kStatsFrames * stats_pitch(D) + post_cap + (size_t)grp_batch * 2 * D * 8 + grp_batch * 8 + 128);
  if (smem > 220 * 1024) {
    set_error("feature dimension too large for the statistics kernel");
    return KHG_ERR_UNSUPPORTED;
  }
  KHG_CUDA_TRY(cudaFuncSetAttribute(stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));  // (per device)
  // pdfs of <= 32 Gaussians, dim <= 56, operands inside fp16's range: the tensor-core kernel (khg_stats_tc.cu)
  bool use_tc = false;
  KHG_TRY(stats_use_tc(m, &use_tc));
  for (int64_t t0 = 0; t0 < T; t0 += slab) {
    const int64_t n = std::min(slab, T - t0);
    KHG_TRY(m->w_keys.reserve(sizeof(int32_t) * n));
    KHG_TRY(m->w_keys_out.reserve(sizeof(int32_t) * n));
    KHG_TRY(m->w_vals_in.reserve(sizeof(int32_t) * n));
    KHG_TRY(m->w_vals_out.reserve(sizeof(int32_t) * n));
    KHG_TRY(m->w_starts.reserve(sizeof(int32_t) * (P + 1)));
    KHG_TRY(m->w_item_start.reserve(sizeof(int32_t) * (P + 1)));
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, m->w_keys.as<int32_t>(), m->w_keys_out.as<int32_t>(),
                                    m->w_vals_in.as<int32_t>(), m->w_vals_out.as<int32_t>(), (int)n, 0, end_bit, st);
    KHG_TRY(m->w_cub.reserve(cub_bytes));
    prep_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_ids + t0, n, P, m->w_keys.as<int32_t>(), m->w_vals_in.as<int32_t>(), m->d_err);
    KHG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(m->w_cub.p, cub_bytes, m->w_keys.as<int32_t>(), m->w_keys_out.as<int32_t>(),
                                                 m->w_vals_in.as<int32_t>(), m->w_vals_out.as<int32_t>(), (int)n, 0, end_bit, st));
    bucket_starts_kernel<<<grid_for(n + 1, 256), 256, 0, st>>>(m->w_keys_out.as<int32_t>(), n, P, m->w_starts.as<int32_t>());
    item_scan_kernel<<<1, 1024, 0, st>>>(P, m->d_offsets, m->w_starts.as<int32_t>(), m->w_item_start.as<int32_t>(), post_cap);
    // upper bound on work items: every pdf wastes at most one partial item
    int f_min = stats_frames_for(m->max_gp, post_cap);
    int64_t max_items = n / f_min + P + 1;
    KHG_TRY(m->w_item_desc.reserve(sizeof(int4) * (size_t)max_items));
    item_table_kernel<<<grid_for(P, 4), 128, 0, st>>>(P, m->d_offsets, m->w_starts.as<int32_t>(), m->w_item_start.as<int32_t>(),
                                                      post_cap, m->w_item_desc.as<int4>());
    StatsArgs a;
    a.feats = d_feats + t0 * D;
    a.order = m->w_vals_out.as<int32_t>();
    a.weights = d_w ? d_w + t0 : nullptr;
    a.starts = m->w_starts.as<int32_t>();
    a.item_start = m->w_item_start.as<int32_t>();
    a.item_desc = m->w_item_desc.as<int4>();
    a.offsets = m->d_offsets;
    a.grp_start = m->d_grp_start;
    a.pack8 = m->d_pack8;
    a.gc8 = m->d_gc8;
    a.occ = s->buf;
    a.mean = s->off_mean >= 0 ? s->buf + s->off_mean : nullptr;
    a.var = s->off_var >= 0 ? s->buf + s->off_var : nullptr;
    a.totals = s->buf + s->off_tot;
    a.call_like = d_call_like;
    a.per_frame = d_pf ? d_pf + t0 : nullptr;
    a.err = m->d_err;
    a.P = P;
    a.D = D;
    a.grp_batch = grp_batch;
    a.post_cap = post_cap;
    a.item_list = nullptr;
    a.item_list_n = nullptr;
    if (use_tc) {
      // tensor-core kernel over all items; the items it declines (values outside fp16's range after scaling) go to
      // the fp32 kernel through a device-side list — normally empty, then the second launch is a few idle CTAs
      KHG_TRY(m->w_fb_items.reserve(sizeof(int32_t) * (size_t)max_items));
      StatsTcArgs ta;
      ta.feats = a.feats; ta.order = a.order; ta.weights = a.weights; ta.item_start = a.item_start; ta.item_desc = a.item_desc;
      ta.offsets = a.offsets; ta.occ = a.occ; ta.mean = a.mean; ta.var = a.var; ta.totals = a.totals; ta.call_like = a.call_like;
      ta.per_frame = a.per_frame; ta.err = a.err; ta.fb_items = m->w_fb_items.as<int32_t>(); ta.P = P; ta.D = D; ta.n_frames = (int)n;
      ta.img = nullptr; ta.img_off = nullptr; ta.ascale = nullptr; ta.unscale = nullptr; ta.fb_count = nullptr;
      KHG_TRY(stats_tc_launch(m, ta, st));
      a.item_list = m->w_fb_items.as<int32_t>();
      a.item_list_n = m->stk.fb_count;
      // (a model with pdfs of more than 32 Gaussians hands their items over here: more CTAs than for the normally empty list)
      int list_ctas = m->stk.partial ? 12 : 2;
      if (const char *e = getenv("KHG_STATS_LIST_CTAS_PER_SM")) list_ctas = std::max(1, atoi(e));  // experiments
      stats_kernel<<<(unsigned)std::min<int64_t>(max_items, list_ctas * (int64_t)m->sm_count), 128, smem, st>>>(a);
    } else {
      stats_kernel<<<(unsigned)max_items, 128, smem, st>>>(a);
    }
    g_launch_count += 6 + 3;  // ours + the radix-sort passes (library)
    KHG_CUDA_TRY(cudaGetLastError());
  }
  return KHG_OK;
}

khg_status khg_acc_stats_ali(khg_model *m, khg_stats *s, const float *feats, int64_t T, int32_t loc,
                             const int32_t *pdf_ids, const float *frame_weights,
                             float *per_frame_loglike, double *tot_loglike) {
  NvtxRange nvtx_range("khg_acc_stats_ali");
  KHG_REQUIRE(m && s && s->model == m && m->uploaded, "model/stats mismatch or model not uploaded");
  KHG_REQUIRE(T >= 0, "T >= 0");
  if (T == 0) {
    if (tot_loglike) *tot_loglike = 0.0;
    return KHG_OK;
  }
  KHG_REQUIRE(feats && pdf_ids, "null buffer");
  const int D = m->dim;
  const float *d_f = nullptr, *d_w = nullptr;
  const int32_t *d_i = nullptr;
  KHG_TRY(stage_in(m, m->w_feats, feats, (size_t)T * D, loc, &d_f));
  KHG_TRY(stage_in(m, m->w_ids, pdf_ids, (size_t)T, loc, &d_i));
  KHG_TRY(stage_in(m, m->w_wts, frame_weights, (size_t)T, loc, &d_w));
  float *d_pf = per_frame_loglike;
  if (per_frame_loglike && loc == KHG_HOST) {
    KHG_TRY(m->w_pf.reserve(sizeof(float) * T));
    d_pf = m->w_pf.as<float>();
  }
  double *d_call = nullptr;
  if (tot_loglike) {
    KHG_TRY(m->w_tot.reserve(sizeof(double)));
    d_call = m->w_tot.as<double>();
    KHG_CUDA_TRY(cudaMemsetAsync(d_call, 0, sizeof(double), m->stream));
  }
  KHG_TRY(acc_device(m, s, d_f, T, d_i, d_w, d_pf, d_call));
  const bool need_sync = tot_loglike || loc == KHG_HOST;
  if (per_frame_loglike && loc == KHG_HOST)
    KHG_CUDA_TRY(cudaMemcpyAsync(per_frame_loglike, d_pf, sizeof(float) * T, cudaMemcpyDeviceToHost, m->stream));
  if (tot_loglike)
    KHG_CUDA_TRY(cudaMemcpyAsync(tot_loglike, d_call, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  if (need_sync) return sync_check(m);
  return KHG_OK;
}

khg_status khg_acc_stats_ali_tids(khg_model *m, khg_stats *s, const float *feats, int64_t T,
                                  const int32_t *tids, const int32_t *tid2pdf, int32_t num_tids,
                                  double *trans_accs, double *tot_loglike) {
  NvtxRange nvtx_range("khg_acc_stats_ali_tids");
  KHG_REQUIRE(m && s && s->model == m && m->uploaded, "model/stats mismatch or model not uploaded");
  KHG_REQUIRE(T >= 0 && num_tids > 0 && tid2pdf, "bad argument");
  if (T == 0) {
    if (tot_loglike) *tot_loglike = 0.0;
    return KHG_OK;
  }
  KHG_REQUIRE(feats && tids, "null buffer");
  cudaStream_t st = m->stream;
  const int D = m->dim;
  const float *d_f = nullptr;
  const int32_t *d_t = nullptr, *d_map = nullptr;
  KHG_TRY(stage_in(m, m->w_feats, feats, (size_t)T * D, KHG_HOST, &d_f));
  KHG_TRY(stage_in(m, m->w_tid, tids, (size_t)T, KHG_HOST, &d_t));
  KHG_TRY(stage_in(m, m->w_tid2pdf, tid2pdf, (size_t)num_tids + 1, KHG_HOST, &d_map));
  KHG_TRY(m->w_ids.reserve(sizeof(int32_t) * T));
  unsigned long long *d_cnt = nullptr;
  if (trans_accs) {
    KHG_TRY(m->w_trans.reserve(sizeof(unsigned long long) * ((size_t)num_tids + 1)));
    d_cnt = m->w_trans.as<unsigned long long>();
    KHG_CUDA_TRY(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * ((size_t)num_tids + 1), st));
  }
  map_tids_kernel<<<grid_for(T, 256), 256, 0, st>>>(d_t, T, d_map, num_tids, m->w_ids.as<int32_t>(), d_cnt, m->d_err);
  ++g_launch_count;
  double *d_call = nullptr;
  if (tot_loglike) {
    KHG_TRY(m->w_tot.reserve(sizeof(double)));
    d_call = m->w_tot.as<double>();
    KHG_CUDA_TRY(cudaMemsetAsync(d_call, 0, sizeof(double), st));
  }
  KHG_TRY(acc_device(m, s, d_f, T, m->w_ids.as<int32_t>(), nullptr, nullptr, d_call));
  std::vector<unsigned long long> cnt;
  if (trans_accs) {
    cnt.resize((size_t)num_tids + 1);
    KHG_CUDA_TRY(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(unsigned long long) * cnt.size(), cudaMemcpyDeviceToHost, st));
  }
  if (tot_loglike) KHG_CUDA_TRY(cudaMemcpyAsync(tot_loglike, d_call, sizeof(double), cudaMemcpyDeviceToHost, st));
  KHG_TRY(sync_check(m));
  if (trans_accs)
    for (size_t i = 0; i < cnt.size(); ++i) trans_accs[i] += (double)cnt[i];  // integer counts: exact
  return KHG_OK;
}

khg_status khg_acc_from_posteriors(khg_model *m, khg_stats *s, int32_t pdf, const float *feats,
                                   int64_t T, int32_t loc, const float *post) {
  KHG_REQUIRE(m && s && s->model == m && m->uploaded, "model/stats mismatch or model not uploaded");
  KHG_REQUIRE(pdf >= 0 && pdf < m->P, "gmm_index >= 0 && gmm_index < NumAccs()");
  KHG_REQUIRE(T > 0 && feats && post, "null buffer");
  const int g0 = m->h_offsets[pdf], ng = m->h_offsets[pdf + 1] - g0, D = m->dim;
  const float *d_f = nullptr, *d_p = nullptr;
  KHG_TRY(stage_in(m, m->w_feats, feats, (size_t)T * D, loc, &d_f));
  KHG_TRY(stage_in(m, m->w_wts, post, (size_t)T * ng, loc, &d_p));
  acc_from_post_kernel<<<grid_for(ng * (D + 1), 128), 128, 0, m->stream>>>(
      d_f, T, D, d_p, ng, g0, s->buf, s->off_mean >= 0 ? s->buf + s->off_mean : nullptr,
      s->off_var >= 0 ? s->buf + s->off_var : nullptr, s->buf + s->off_tot);
  ++g_launch_count;
  KHG_CUDA_TRY(cudaGetLastError());
  if (loc == KHG_HOST) KHG_CUDA_TRY(cudaStreamSynchronize(m->stream));
  return KHG_OK;
}

// ------------------------------------------------------------------ M-step --
khg_status khg_model_download(khg_model *m, int32_t *gauss_offsets, float *weights, float *means_invvars,
                              float *inv_vars, float *gconsts) {
  KHG_REQUIRE(m && m->uploaded, "model not uploaded");
  cudaStream_t st = m->stream;
  size_t gd = (size_t)m->G * m->dim;
  if (gauss_offsets) std::memcpy(gauss_offsets, m->h_offsets.data(), sizeof(int32_t) * (m->P + 1));
  if (weights) KHG_CUDA_TRY(cudaMemcpyAsync(weights, m->d_weights, sizeof(float) * m->G, cudaMemcpyDeviceToHost, st));
  if (means_invvars) KHG_CUDA_TRY(cudaMemcpyAsync(means_invvars, m->d_miv, sizeof(float) * gd, cudaMemcpyDeviceToHost, st));
  if (inv_vars) KHG_CUDA_TRY(cudaMemcpyAsync(inv_vars, m->d_iv, sizeof(float) * gd, cudaMemcpyDeviceToHost, st));
  if (gconsts) KHG_CUDA_TRY(cudaMemcpyAsync(gconsts, m->d_gconsts, sizeof(float) * m->G, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  return KHG_OK;
}

extern "C++" {
namespace {
}  // namespace
}  // extern "C++"

khg_status khg_mle_update(khg_model *m, const khg_stats *s, const khg_mle_options *opts, uint16_t update_flags,
                          khg_model **new_model, float *obj_change, float *count, int32_t *floored_elements,
                          int32_t *floored_gaussians, int32_t *removed_gaussians) {
  NvtxRange nvtx_range("khg_mle_update");
  KHG_REQUIRE(m && s && s->model == m && m->uploaded && opts && new_model, "bad argument");
  update_flags &= (KHG_GMM_MEANS | KHG_GMM_VARIANCES | KHG_GMM_WEIGHTS);
  if (update_flags & ~s->flags) {
    set_error("Flags in argument do not match the active accumulators");  // csrc/mle-diag-gmm.cc:253-255
    return KHG_ERR_INVALID;
  }
  const int P = m->P, D = m->dim, G = m->G;
  cudaStream_t st = m->stream;
  DevTmp tmp;
  float *w_new, *miv_new, *iv_new, *gc_old, *gc_new, *obj_old, *obj_new;
  int32_t *remove, *counters, *flags2;
  double *pdf_occ;
  KHG_TRY(tmp.alloc(&w_new, G));
  KHG_TRY(tmp.alloc(&miv_new, (size_t)G * D));
  KHG_TRY(tmp.alloc(&iv_new, (size_t)G * D));
  KHG_TRY(tmp.alloc(&gc_old, G));
  KHG_TRY(tmp.alloc(&gc_new, G));
  KHG_TRY(tmp.alloc(&obj_old, P));
  KHG_TRY(tmp.alloc(&obj_new, P));
  KHG_TRY(tmp.alloc(&remove, G));
  KHG_TRY(tmp.alloc(&counters, 4));
  KHG_TRY(tmp.alloc(&flags2, 4));
  KHG_TRY(tmp.alloc(&pdf_occ, P));
  KHG_CUDA_TRY(cudaMemsetAsync(counters, 0, sizeof(int32_t) * 4, st));
  KHG_CUDA_TRY(cudaMemsetAsync(flags2, 0, sizeof(int32_t) * 4, st));
  const double *occ = s->buf;
  const double *mean = s->off_mean >= 0 ? s->buf + s->off_mean : nullptr;
  const double *var = s->off_var >= 0 ? s->buf + s->off_var : nullptr;
  // gmm->ComputeGconsts(); obj_old = MlObjective(*gmm, acc)   (csrc/mle-diag-gmm.cc:266-267)
  gconsts_kernel<<<grid_for(G, 128), 128, 0, st>>>(G, D, m->d_weights, m->d_miv, m->d_iv, gc_old, flags2);
  mle_objective_kernel<<<P, 128, 0, st>>>(P, D, s->flags, m->d_offsets, occ, mean, var, gc_old, m->d_miv, m->d_iv, obj_old);
  MleArgs a;
  a.P = P; a.D = D; a.acc_flags = s->flags; a.upd_flags = update_flags;
  a.min_w = opts->min_gaussian_weight; a.min_occ = opts->min_gaussian_occupancy; a.min_var = opts->min_variance;
  a.remove_low = opts->remove_low_count_gaussians;
  a.offsets = m->d_offsets; a.occ = occ; a.mean = mean; a.var = var;
  a.w_old = m->d_weights; a.miv_old = m->d_miv; a.iv_old = m->d_iv;
  a.w_new = w_new; a.miv_new = miv_new; a.iv_new = iv_new;
  a.remove = remove; a.counters = counters; a.pdf_occ = pdf_occ;
  mle_update_kernel<<<P, 128, 0, st>>>(a);
  // gmm->ComputeGconsts(); obj_new = MlObjective(*gmm, acc)    (:365-366), before any removal
  gconsts_kernel<<<grid_for(G, 128), 128, 0, st>>>(G, D, w_new, miv_new, iv_new, gc_new, flags2 + 2);
  mle_objective_kernel<<<P, 128, 0, st>>>(P, D, s->flags, m->d_offsets, occ, mean, var, gc_new, miv_new, iv_new, obj_new);
  g_launch_count += 5;
  KHG_CUDA_TRY(cudaGetLastError());
  std::vector<float> h_old(P), h_new(P);
  std::vector<double> h_occ(P);
  std::vector<int32_t> h_remove(G);
  int32_t h_cnt[4], h_fl[4];
  KHG_CUDA_TRY(cudaMemcpyAsync(h_old.data(), obj_old, sizeof(float) * P, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(h_new.data(), obj_new, sizeof(float) * P, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(h_occ.data(), pdf_occ, sizeof(double) * P, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(h_remove.data(), remove, sizeof(int32_t) * G, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(h_cnt, counters, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaMemcpyAsync(h_fl, flags2, sizeof(h_fl), cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  if (h_fl[1] || h_fl[3]) {
    set_error("not a number in gconst computation");
    return KHG_ERR_NONFINITE;
  }
  // per-pdf float results summed in pdf order like MleAmDiagGmmUpdate (csrc/mle-am-diag-gmm.cc:177-190)
  float tot_obj = 0.f, tot_count = 0.f;
  for (int p = 0; p < P; ++p) {
    tot_obj += h_new[p] - h_old[p];
    tot_count += (float)h_occ[p];
  }
  if (obj_change) *obj_change = tot_obj;
  if (count) *count = tot_count;
  if (floored_elements) *floored_elements = h_cnt[0];
  if (floored_gaussians) *floored_gaussians = h_cnt[1];
  if (removed_gaussians) *removed_gaussians = h_cnt[2];
  // new structure
  std::vector<int32_t> new_off(P + 1, 0);
  for (int p = 0; p < P; ++p) {
    int kept = 0;
    for (int g = m->h_offsets[p]; g < m->h_offsets[p + 1]; ++g) kept += h_remove[g] ? 0 : 1;
    new_off[p + 1] = new_off[p] + kept;
  }
  khg_model *nm = nullptr;
  KHG_TRY(khg_model_create(D, P, new_off.data(), &nm));
  nm->stream = st;
  nm->kernel = m->kernel;
  int32_t *d_new_off = nm->d_offsets;
  size_t shm = sizeof(float) * 2 * (size_t)m->max_gp;
  mle_compact_kernel<<<P, 128, shm, st>>>(P, D, m->d_offsets, d_new_off, remove, w_new, miv_new, iv_new,
                                          nm->d_weights, nm->d_miv, nm->d_iv);
  ++g_launch_count;
  khg_status fs = finish_model_from_device(nm, nullptr);
  if (fs != KHG_OK) {
    khg_model_destroy(nm);
    return fs;
  }
  *new_model = nm;
  return KHG_OK;
}

// ------------------------------------------------------------------ E-step --
static khg_status estep_init_streams(khg_model *m) { return khg::ensure_copy_stream(m); }

khg_status khg_estep(khg_model *m, khg_stats *s, const float *feats, int64_t T, int32_t loc,
                     const int32_t *pdf_ids, const float *frame_weights, float *loglikes_out,
                     int64_t ld_out, int64_t chunk_frames, double *tot_loglike) {
  NvtxRange nvtx_range("khg_estep");
  KHG_REQUIRE(m && s && s->model == m && m->uploaded, "model/stats mismatch or model not uploaded");
  KHG_REQUIRE(T >= 0, "T >= 0");
  if (tot_loglike) *tot_loglike = 0.0;
  if (T == 0) return KHG_OK;
  KHG_REQUIRE(feats && pdf_ids && loglikes_out, "null buffer");
  const int D = m->dim;
  if (chunk_frames <= 0) chunk_frames = 128LL * m->sm_count * 4;
  chunk_frames = std::min(chunk_frames, ld_out);
  KHG_REQUIRE(chunk_frames > 0, "ld_out >= chunk_frames > 0");
  double *d_call = nullptr;
  if (tot_loglike) {
    KHG_TRY(m->w_tot.reserve(sizeof(double)));
    d_call = m->w_tot.as<double>();
    KHG_CUDA_TRY(cudaMemsetAsync(d_call, 0, sizeof(double), m->stream));
  }
  // The statistics pass (K2 + K3 / K3t) does not read the dense block, so it runs once per GROUP of chunks: a chunk of
  // 606 208 frames holds ~144 frames per pdf at C4 (303 at C3) — partly filled 128-frame work items and a dozen small
  // launches per chunk (1.1-2.4 G frames/s inside the E-step against 7-8 G for a whole batch).  A group is up to 16 M
  // frames (KHG_ESTEP_STATS_GROUP_FRAMES), a multiple of the chunk.
  int64_t group_frames = 16LL << 20;
  if (const char *e = getenv("KHG_ESTEP_STATS_GROUP_FRAMES")) group_frames = std::max<int64_t>(1, atoll(e));
  const int64_t cpg = std::max<int64_t>(1, group_frames / chunk_frames);  // chunks per group
  if (loc == KHG_DEVICE) {
    int64_t g0 = 0;
    for (int64_t t0 = 0, c = 0; t0 < T; t0 += chunk_frames, ++c) {
      int64_t n = std::min(chunk_frames, T - t0);
      KHG_TRY(dense_device(m, feats + t0 * D, n, 1.0f, KHG_PDF_MAJOR, loglikes_out, ld_out));
      if ((c + 1) % cpg == 0 || t0 + n == T) {
        KHG_TRY(acc_device(m, s, feats + g0 * D, t0 + n - g0, pdf_ids + g0, frame_weights ? frame_weights + g0 : nullptr, nullptr, d_call));
        g0 = t0 + n;
      }
    }
  } else {
    // Host inputs: pinned staging (two chunk slots) -> device staging (two HALVES of one group of chunks each); the
    // H2D copies run on copy_stream ahead of the compute on the model stream: the dense kernel of a chunk starts when
    // its copy has landed, the statistics pass of a group after the group's last dense kernel, and a half is
    // refilled (group g + 2) once the statistics of group g have read it.
    KHG_TRY(estep_init_streams(m));
    // Caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory)
    // are copied from directly: no staging memcpy on the calling thread, which is what limits several
    // ranks feeding their GPUs from one host.
    auto is_pinned = [](const void *p) {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
      }
      return at.type == cudaMemoryTypeHost;
    };
    const bool direct_f = is_pinned(feats), direct_i = is_pinned(pdf_ids), direct_w = frame_weights && is_pinned(frame_weights);
    const int64_t n_chunks = (T + chunk_frames - 1) / chunk_frames;
    const int64_t half_frames = std::min(cpg, n_chunks) * chunk_frames;
    for (int i = 0; i < 2; ++i) {
      if (!direct_f) KHG_TRY(m->pin_feats[i].reserve(sizeof(float) * chunk_frames * D));
      if (!direct_i) KHG_TRY(m->pin_ids[i].reserve(sizeof(int32_t) * chunk_frames));
      if (frame_weights && !direct_w) KHG_TRY(m->pin_wts[i].reserve(sizeof(float) * chunk_frames));
      if (i == 1 && n_chunks <= cpg) break;  // a single group: one half
      KHG_TRY(m->w_efeats[i].reserve(sizeof(float) * half_frames * D));
      KHG_TRY(m->w_eids[i].reserve(sizeof(int32_t) * half_frames));
      if (frame_weights) KHG_TRY(m->w_ewts[i].reserve(sizeof(float) * half_frames));
    }
    bool pin_used[2] = {false, false};
    for (int64_t c = 0; c < n_chunks; ++c) {
      const int b = (int)(c & 1);                      // pinned staging slot
      const int64_t g = c / cpg, k = c % cpg;          // group, chunk inside the group
      const int h = (int)(g & 1);                      // device half
      const int64_t t0 = c * chunk_frames, n = std::min(chunk_frames, T - t0);
      // half h is free once the statistics pass of group g - 2 (recorded on ev_done[h]) has read it
      if (k == 0 && g >= 2) KHG_CUDA_TRY(cudaStreamWaitEvent(m->copy_stream, m->ev_done[h], 0));
      // pinned slot b is free once its previous H2D finished (also bounds how far the host runs ahead)
      if (pin_used[b]) KHG_CUDA_TRY(cudaEventSynchronize(m->ev_copy[b]));
      float *d_f = m->w_efeats[h].as<float>() + k * chunk_frames * D;
      int32_t *d_i = m->w_eids[h].as<int32_t>() + k * chunk_frames;
      const void *src_f = feats + t0 * D, *src_i = pdf_ids + t0;
      if (!direct_f) src_f = parallel_memcpy(m->pin_feats[b].p, src_f, sizeof(float) * n * D, 4);
      if (!direct_i) src_i = std::memcpy(m->pin_ids[b].p, src_i, sizeof(int32_t) * n);
      KHG_CUDA_TRY(cudaMemcpyAsync(d_f, src_f, sizeof(float) * n * D, cudaMemcpyHostToDevice, m->copy_stream));
      KHG_CUDA_TRY(cudaMemcpyAsync(d_i, src_i, sizeof(int32_t) * n, cudaMemcpyHostToDevice, m->copy_stream));
      if (frame_weights) {
        const void *src_w = frame_weights + t0;
        if (!direct_w) src_w = std::memcpy(m->pin_wts[b].p, src_w, sizeof(float) * n);
        KHG_CUDA_TRY(cudaMemcpyAsync(m->w_ewts[h].as<float>() + k * chunk_frames, src_w, sizeof(float) * n, cudaMemcpyHostToDevice, m->copy_stream));
      }
      KHG_CUDA_TRY(cudaEventRecord(m->ev_copy[b], m->copy_stream));
      pin_used[b] = true;
      KHG_CUDA_TRY(cudaStreamWaitEvent(m->stream, m->ev_copy[b], 0));
      KHG_TRY(dense_device(m, d_f, n, 1.0f, KHG_PDF_MAJOR, loglikes_out, ld_out));
      if (k == cpg - 1 || c == n_chunks - 1) {
        KHG_TRY(acc_device(m, s, m->w_efeats[h].as<float>(), k * chunk_frames + n, m->w_eids[h].as<int32_t>(),
                           frame_weights ? m->w_ewts[h].as<float>() : nullptr, nullptr, d_call));
        KHG_CUDA_TRY(cudaEventRecord(m->ev_done[h], m->stream));
      }
    }
  }
  if (tot_loglike) {
    KHG_CUDA_TRY(cudaMemcpyAsync(tot_loglike, d_call, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    return sync_check(m);
  }
  if (loc == KHG_HOST) return sync_check(m);
  return KHG_OK;
}

}  // extern "C"

namespace khg {
// Finishes a handle whose d_weights / d_miv / d_iv were written on the device, exactly like
// khg_model_upload(gconsts = NULL) does from host data: gconsts, the derived packs, the
// tensor-core operand.  num_bad (optional): ComputeGconsts' count of -inf gconsts.
khg_status finish_model_from_device(khg_model *nm, int32_t *num_bad) {
  cudaStream_t st = nm->stream;
  const int D = nm->dim, P = nm->P;
  KHG_CUDA_TRY(cudaMemsetAsync(nm->d_scratch_int, 0, sizeof(int) * 4, st));
  gconsts_kernel<<<grid_for(nm->G, 128), 128, 0, st>>>(nm->G, D, nm->d_weights, nm->d_miv, nm->d_iv, nm->d_gconsts, nm->d_scratch_int);
  pack_simt_kernel<<<std::min(1024u, grid_for((int64_t)nm->n_chunks * 2 * D * kSimtChunk, 256)), 256, 0, st>>>(
      nm->G, D, nm->n_chunks, nm->d_miv, nm->d_iv, nm->d_packT);
  pack8_kernel<<<P, 128, 0, st>>>(P, D, nm->d_offsets, nm->d_grp_start, nm->d_miv, nm->d_iv, nm->d_gconsts, nm->d_pack8, nm->d_gc8);
  g_launch_count += 3;
  KHG_CUDA_TRY(cudaGetLastError());
  int flags[2] = {0, 0};
  KHG_CUDA_TRY(cudaMemcpyAsync(flags, nm->d_scratch_int, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  if (flags[1]) {
    set_error("not a number in gconst computation");  // csrc/diag-gmm.cc:132-135
    return KHG_ERR_NONFINITE;
  }
  if (num_bad) *num_bad = flags[0];
  nm->uploaded = true;
  nm->tc.ready = false;
  stats_tc_free(nm);
  if (nm->kernel != KHG_KERNEL_SIMT && tc_supported(nm)) {
    khg_status ts = tc_pack_build(nm);
    if (ts != KHG_OK && nm->kernel >= KHG_KERNEL_TCGEN05) return ts;
  }
  KHG_CUDA_TRY(cudaStreamSynchronize(st));
  return KHG_OK;
}

khg_status dense_block(khg_model *m, const float *d_feats, int64_t T, float scale, int layout, float *d_out, int64_t ld,
                       const TileSubset *subset, bool *subset_used) {
  return dense_device(m, d_feats, T, scale, layout, d_out, ld, subset, subset_used);
}
khg_status sync_and_check(khg_model *m) { return sync_check(m); }
khg_status ensure_copy_stream(khg_model *m) {
  if (m->copy_stream) return KHG_OK;
  KHG_CUDA_TRY(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    KHG_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_copy[i], cudaEventDisableTiming));
    KHG_CUDA_TRY(cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
    m->pin_feats[i].pinned = m->pin_ids[i].pinned = m->pin_wts[i].pinned = true;
  }
  return KHG_OK;
}
}  // namespace khg
