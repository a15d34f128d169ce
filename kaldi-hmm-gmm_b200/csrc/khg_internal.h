// kaldi-hmm-gmm_b200/csrc/khg_internal.h — internals shared by the .cu files of
// libkhg_b200.so (not part of the ABI; the ABI is include/khg_b200.h).
#ifndef KHG_INTERNAL_H_
#define KHG_INTERNAL_H_

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "khg_b200.h"

namespace khg {

void set_error(const std::string &msg);
extern int64_t g_launch_count;

#define KHG_CUDA_TRY(expr)                                                        \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      ::khg::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));       \
      return KHG_ERR_CUDA;                                                        \
    }                                                                             \
  } while (0)

#define KHG_TRY(expr)                \
  do {                               \
    khg_status _s = (expr);          \
    if (_s != KHG_OK) return _s;     \
  } while (0)

#define KHG_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::khg::set_error(std::string("KHG_ASSERT failed: ") + (msg)); \
      return KHG_ERR_INVALID;                                   \
    }                                                           \
  } while (0)

// Device error-flag bits (latched by kernels, read by synchronising calls).
enum : int { ERR_NONFINITE = 1, ERR_BAD_INDEX = 2 };

// A growable device (or pinned host) scratch buffer.
struct Buf {
  void *p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  khg_status reserve(size_t bytes);
  void release();
  template <class T> T *as() { return static_cast<T *>(p); }
};

// Small RAII bundle of device scratch (M-step, mix-up).
struct DevTmp {
  std::vector<void *> ptrs;
  ~DevTmp() { for (void *p : ptrs) cudaFree(p); }
  template <class T> khg_status alloc(T **out, size_t n) {
    void *p = nullptr;
    KHG_CUDA_TRY(cudaMalloc(&p, (n > 0 ? n : 1) * sizeof(T)));
    ptrs.push_back(p);
    *out = static_cast<T *>(p);
    return KHG_OK;
  }
};

// ---- SIMT model pack ---------------------------------------------------------
// Gaussians are grouped in chunks of 32 (global Gaussian index / 32); chunk c
// holds [which(0=means_invvars,1=inv_vars)][d][g%32], zero padded, so a CTA can
// stage a chunk with one contiguous copy and read 8 consecutive Gaussians of one
// dimension with two broadcast LDS.128.
constexpr int kSimtChunk = 32;

// ---- tcgen05 model pack --------------------------------------------------------
// B operand of the dense contraction: row g = [means_invvars(D) | -0.5*inv_vars(D)
// | gconst | 0-pad] split into TF32 hi and lo parts (3xTF32), K padded to KP.
// Chunks (= TMA stages) of the streamed operand of the tensor-core kernel: see khg_loglikes_tc.cu
constexpr int kMaxStages = 12;
struct StageTab {
  uint32_t e[kMaxStages];  // h0 | nh << 8 | l0 << 12 | nl << 20   (units: K steps)
  int n;
};
struct TcPack {
  bool ready = false;       // tile tables built and at least one operand container usable
  bool tf32_ready = false;  // tf32 operands built (needs 2D+1 <= 160)
  int K = 0;   // 2D+1
  int K8 = 0;  // K rounded up to 8  (UMMA_K for tf32)
  int KP = 0;  // K8 rounded up to 32 (one 128-byte swizzle atom per 32 floats): A row width
  int KPB = 0; // width of the streamed operand rows (tab8.n chunks of 32 floats)
  StageTab tab8, tab16;
  int rows = 0;          // padded row count of the operand
  float *bhi = nullptr;  // rows x KPB: B' (tf32 containers)
  float *blo = nullptr;  // unused (hi and lo live side by side in bhi)
  CUtensorMap map_hi, map_lo;
  // fp16-split variant of the same operand (kind::f16 runs at twice the tf32 rate)
  bool f16_ready = false;
  int K16 = 0, KP16 = 0, KPB16 = 0;  // 2D+2 rounded up to 16; A row width (K16 -> 64); B' row width (K16 + Kc -> 64)
  void *hhi = nullptr, *hlo = nullptr;  // hhi: rows x KPB16 __half B' = [hi | lo]; hlo unused
  float *ascale = nullptr;      // device, 2D floats: power-of-two scale of the [x | x^2] columns
  unsigned *gate = nullptr;     // device word: max |x*ascale| bits of the current call (auto mode)
  CUtensorMap hmap_hi, hmap_lo;
  // N-tiles aligned to pdf boundaries: tile j covers Gaussians
  // [tile_g0[j], tile_g0[j]+kTileN) and owns pdfs [tile_p0[j], tile_p0[j+1]).
  int n_tiles = 0;
  int32_t *tile_g0 = nullptr;  // device, n_tiles
  int32_t *tile_p0 = nullptr;  // device, n_tiles+1
  std::vector<int32_t> h_tile_g0, h_tile_p0;
  // epilogue tables: per (tile, epilogue group) the list of its pdfs (segments), in runs of equal
  // Gaussian count (khg_loglikes_tc.cu)
  void *epi_hdr = nullptr;       // device int2 per (tile, group): {start in seg[], first run | runs << 24}
  uint32_t *runs = nullptr;      // device: length | count << 8
  uint32_t *seg = nullptr;       // device: column | pdf << 8, + 2 sentinels per list
  bool grouped_segs = false;     // some tile has a group of short pdfs read by one load (epi_run_multi)
  bool two_chunk_segs = false;   // some pdf has 17..32 Gaussians: the epilogue's two-load segment form is in use
  bool dead_pdf = false;         // some pdf has only -inf gconsts: every call must fail like the reference
  // pdfs of more than 240 Gaussians are split into virtual pdfs of <= 240 (the kernels' pdf ids are virtual ids);
  // Pv == P when no pdf is split.  The kernels then write Pv rows of unscaled log-sum-exps and merge_virtual_kernel
  // combines them (khg_b200.cu dense_device)
  int Pv = 0;
  std::vector<int32_t> v_off;    // Pv+1 Gaussian offsets of the virtual pdfs
  int32_t *d_vfirst = nullptr;   // device, P+1: first virtual pdf of every pdf
  std::vector<int32_t> h_vfirst; // host copy
  // Gaussian-stationary kernel (khg_loglikes_gs.cu): the split feature operand A' of the current block
  // of frames, a_rows x KPB16 fp16, and its TMA map (box 64 columns x 128 rows)
  Buf a_scr;
  int64_t a_rows = 0;
  CUtensorMap amap;
};


// ---- tensor-core statistics kernel (khg_stats_tc.cu) ---------------------------
// Per-pdf operand images (fp16 hi / lo rows of [means_invvars | -inv_vars/2 | gconst], 128-byte swizzle, ready for
// one bulk copy into shared memory) and the per-dimension power-of-two scaling they were built with.
struct StatsTcPack {
  bool tried = false;      // build attempted for the current parameters
  bool ready = false;      // shape and value range fit (dim <= 40, at least half of the Gaussians in pdfs of <= 32, |operand| <= 3e4)
  int DP = 0;              // dim rounded up to 8
  int np_max = 16;         // operand rows of the largest pdf that has an image (16 or 32)
  bool partial = false;    // some pdfs have more than 32 Gaussians: no image (img_off < 0), their items go to the fp32 kernel
  uint8_t *img = nullptr;
  int32_t *img_off = nullptr;  // P+1, 1024-byte units; -1: the pdf has no image
  float *ascale = nullptr;     // 64 floats: 2^-k_d
  float h_ascale[40] = {};     // host copy (passed to the kernel by value)
  float *unscale = nullptr;    // 128 floats: multiplier of each row of the statistics tile
  int *fb_count = nullptr;     // [0]: items left to the fp32 kernel by the current launch; [1]: pack flag
};
struct StatsTcArgs {
  const float *feats;
  const int32_t *order;
  const float *weights;
  const int32_t *item_start;  // P+1 (item_start[P] = number of items)
  const int4 *item_desc;
  const int32_t *offsets;
  const uint8_t *img;         // per-pdf model images
  const int32_t *img_off;     // P+1, in 1024-byte units
  const float *ascale;        // DP floats: 2^-k_d
  float asc_c[40];            // the same by value (0 beyond dim): constant-bank operands
  const float *unscale;       // 128 floats: what a row of S is multiplied by
  const float *miv, *iv, *gconsts;  // the fp32 parameters (exact re-evaluation of the Gaussians that matter)
  int np_max;                 // 16 or 32: operand rows of the model's largest pdf
  int n_frames;               // entries of `order`
  double *occ, *mean, *var, *totals, *call_like;
  float *per_frame;
  int *err;
  int32_t *fb_items;          // items left to the fp32 kernel
  int *fb_count;
  int P, D;
};

}  // namespace khg

struct khg_model {
  int dim = 0, P = 0, G = 0, max_gp = 0;
  std::vector<int32_t> h_offsets;
  int32_t *d_offsets = nullptr;
  float *d_weights = nullptr;  // G
  float *d_miv = nullptr;      // G x D
  float *d_iv = nullptr;       // G x D
  float *d_gconsts = nullptr;  // G
  float *d_packT = nullptr;    // n_chunks x 2 x D x 32
  int n_chunks = 0;
  std::vector<int32_t> h_grp_start;  // P+1: groups of 8 Gaussians per pdf (K3 pack)
  int32_t *d_grp_start = nullptr;
  float *d_pack8 = nullptr;    // n_groups x 2 x D x 8
  float *d_gc8 = nullptr;      // n_groups x 8 (-inf padded gconsts)
  bool uploaded = false;
  int *d_err = nullptr;
  int *d_scratch_int = nullptr;  // 4 ints: num_bad, has_nan, ...
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D staging for khg_estep(KHG_HOST)
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  int kernel = KHG_KERNEL_AUTO;
  int sm_count = 148;
  khg::TcPack tc;
  khg::StatsTcPack stk;
  khg_model *gsel_shadow = nullptr;  // the Gaussians of pdf gsel_pdf as one-Gaussian pdfs (khg_gselect.cu)
  int gsel_pdf = -1;
  // scratch
  khg::Buf w_feats, w_ids, w_wts, w_out, w_pf;         // device staging of host args
  khg::Buf w_keys, w_keys_out, w_vals_in, w_vals_out, w_cub;       // bucketing (K2)
  khg::Buf w_starts, w_item_start, w_item_desc, w_tot, w_fb_items;              // per-pdf starts, work items, totals
  khg::Buf w_tid, w_tid2pdf, w_trans;                  // tid path
  khg::Buf w_sub, w_full;                              // pdf-subset gather
  khg::Buf w_al_graph, w_al_block, w_al_bp, w_al_cost, w_al_ali, w_al_path;  // khg_align_batch (khg_align.cu)
  void *al_cache = nullptr;                                                   // AlignPrepCache (khg_align.cu)
  khg::Buf pin_al_tiles;  // pinned host image of the tile lists (an upload from pageable memory would wait for the dense kernel running on the stream)
  khg::Buf w_al_tiles;                                                       // ... tile subset lists of the dense kernel
  khg::Buf w_al_xlist, w_al_xll;                                             // ... its exact host pass: flagged list, likelihood rows
  khg::Buf pin_feats[2], pin_ids[2], pin_wts[2];       // pinned staging for estep(HOST)
  khg::Buf w_efeats[2], w_eids[2], w_ewts[2];
  // pinned staging of large pageable host buffers (stage_in): two slots, one event each
  khg::Buf pin_stage[2];
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
};

struct khg_stats {
  khg_model *model = nullptr;
  uint16_t flags = 0;
  int64_t n = 0;      // doubles in buf
  double *buf = nullptr;
  int64_t off_mean = -1, off_var = -1, off_tot = 0;
};

namespace khg {
// khg_loglikes_tc.cu
khg_status tc_pack_build(khg_model *m);
void tc_pack_free(khg_model *m);
bool tc_supported(const khg_model *m);
// Optional restriction of the dense kernel to the model tiles somebody will read (device arrays): frame tiles 2q and
// 2q + 1 compute the tiles tiles[off[q] .. off[q + 1]) only; rows of pdfs in other tiles are left untouched.
struct TileSubset {
  const int32_t *off = nullptr;
  const int32_t *tiles = nullptr;
  int shift = 1;  // list q serves the frame tiles [q << shift, (q + 1) << shift): 1 = per pair of tiles (CTA pairs share the operand stream), 0 = per tile (plain launch)
};
// model tiles [*ja, *jb] that hold Gaussians of pdf p (host tables of the pack), and the number of model tiles
void tc_pdf_tile_range(const khg_model *m, int p, int *ja, int *jb);
int tc_num_tiles(const khg_model *m);
// out is pdf-major: out[p*ld + t]
// *simt_gate (out): non-NULL when the caller must also launch the fp32 SIMT kernel gated on
// that device word (fp16-only shapes whose features may be out of range).
khg_status tc_loglikes(khg_model *m, const float *d_feats, int64_t T, float scale,
                       float *d_out, int64_t ld_out, int precision, const unsigned **simt_gate,
                       float *gate_limit, const TileSubset *subset = nullptr);
// khg_stats_tc.cu: posteriors + statistics of bucketed frames on the tensor cores (lazy pack; launch fills the pack fields of `a`)
khg_status stats_tc_build(khg_model *m);
void stats_tc_free(khg_model *m);
khg_status stats_tc_launch(khg_model *m, StatsTcArgs a, cudaStream_t st);
// khg_loglikes_gs.cu: the Gaussian-stationary form of the fp16-split kernel (model tile resident in shared
// memory, pre-split feature operand streamed)
bool gs_supported(const khg_model *m);
void gs_free(khg_model *m);
khg_status gs_loglikes(khg_model *m, const float *d_feats, int64_t T, float scale, float *d_out, int64_t ld_out);
// khg_align.cu: frees the model's cached graph preparation
void align_cache_free(khg_model *m);
// khg_align_exact.cu: the reference's FasterDecoder on the host for one utterance of a graph batch
khg_status align_exact_host(const khg_graph_batch *gb, int32_t utt, const float *ll, int64_t ld, const int32_t *row_of_tid,
                            const int32_t *row_of_arc, float beam, float retry_beam, int32_t *alignment, int32_t *status,
                            float *cost, std::vector<int32_t> *path);
// khg_b200.cu: the dense all-pdf block of device-resident frames (kernel choice of the model),
// and the synchronising read of the latched device error flag
// (*subset_used, if given, tells whether the tile subset was honoured; when it was not, everything was computed)
khg_status dense_block(khg_model *m, const float *d_feats, int64_t T, float scale, int layout, float *d_out, int64_t ld,
                       const TileSubset *subset = nullptr, bool *subset_used = nullptr);
khg_status sync_and_check(khg_model *m);
// host -> device copy on the model's stream; large pageable sources are staged through pinned slots by several threads
khg_status h2d_copy(khg_model *m, void *dst, const void *src, size_t bytes, cudaStream_t stream = nullptr);  // (nullptr: the model's stream)
// the model's second stream (H2D copies that run under compute) and its events
khg_status ensure_copy_stream(khg_model *m);
khg_status finish_model_from_device(khg_model *nm, int32_t *num_bad);
}  // namespace khg

#endif  // KHG_INTERNAL_H_
