// kaldi-hmm-gmm_b200/csrc/khg_align_exact.cu — the reference's FasterDecoder on the HOST, consuming
// GPU-computed log-likelihood blocks (BASELINE.json north star: "The FST-based Viterbi search stays on
// the host, but it consumes GPU-computed log-likelihood blocks").  Host code only.
//
// khg_align_batch's device search (khg_align.cu) proves, per utterance, that the reference's
// order-dependent pruning cannot have changed the result; the utterances it cannot prove this for are
// re-aligned here, so that every alignment the batch call returns is the reference's.
//
// Restated (paths relative to kaldi-hmm-gmm/csrc/ of the reference):
//   HashList<StateId, Token*>     hash-list-inl.h:26-170   the ORDER of the token list: occupied buckets in
//                                 the order they were first occupied, insertion order inside a bucket,
//                                 bucket = key % hash_size
//   FasterDecoder::InitDecoding   faster-decoder.cc:36-49
//   ::ProcessNonemitting          faster-decoder.cc:51-123  (LIFO queue seeded with the whole list)
//   ::ProcessEmitting             faster-decoder.cc:154-228 (best token first, then the RUNNING
//                                 next_weight_cutoff while walking the list)
//   ::GetCutoff                   faster-decoder.cc:230-320 (min_active 20, beam_delta 0.5; costs rounded to
//                                 float in tmp_array_), ::PossiblyResizeHash :322-329 (hash_ratio 2, from 1000)
//   ::ReachedFinal / GetBestPath  faster-decoder.cc:346-425
//   Token arithmetic              faster-decoder.h:108-141 (double cost_, float arc weight / acoustic cost)
//   AlignUtteranceWrapper         decoder-wrappers.cc:16-108 (beam, then retry_beam on the SAME decoder)
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "khg_internal.h"

namespace khg {

namespace {

struct Tok {
  int32_t arc;   // absolute arc id, -1 for the start token
  int32_t prev;  // index into the token arena, -1 = none
  double cost;
};

struct HashList {
  struct Elem { int32_t key, val, tail; };          // val = token index; tail = next element (-1 = end)
  struct Bucket { int32_t prev_bucket, last_elem; };  // last_elem = -1: empty
  std::vector<Elem> elems;
  std::vector<Bucket> buckets;
  size_t hash_size = 0;
  int32_t list_head = -1, bucket_list_tail = -1;

  void SetSize(size_t size) {  // :26-35
    hash_size = size;
    if (size > buckets.size()) buckets.resize(size, Bucket{0, -1});
  }
  // :37-51: empties the table and returns the old list (element indices stay valid until the next Insert
  // batch is over; the caller copies what it needs)
  int32_t Clear() {
    for (int32_t b = bucket_list_tail; b != -1; b = buckets[b].prev_bucket) buckets[b].last_elem = -1;
    bucket_list_tail = -1;
    const int32_t ans = list_head;
    list_head = -1;
    return ans;
  }
  // :128-170: the existing element of `key`, or a new one holding `val`
  int32_t Insert(int32_t key, int32_t val) {
    const size_t index = (size_t)key % hash_size;
    Bucket &bucket = buckets[index];
    if (bucket.last_elem != -1) {
      const int32_t head = bucket.prev_bucket == -1 ? list_head : elems[buckets[bucket.prev_bucket].last_elem].tail;
      const int32_t tail = elems[bucket.last_elem].tail;
      for (int32_t e = head; e != tail; e = elems[e].tail)
        if (elems[e].key == key) return e;
    }
    const int32_t elem = (int32_t)elems.size();
    elems.push_back(Elem{key, val, -1});
    if (bucket.last_elem == -1) {  // unoccupied bucket: appended to the bucket chain = to the end of the list
      if (bucket_list_tail == -1) list_head = elem;
      else elems[buckets[bucket_list_tail].last_elem].tail = elem;
      bucket.last_elem = elem;
      bucket.prev_bucket = bucket_list_tail;
      bucket_list_tail = (int32_t)index;
    } else {  // occupied bucket: after its last element
      elems[elem].tail = elems[bucket.last_elem].tail;
      elems[bucket.last_elem].tail = elem;
      bucket.last_elem = elem;
    }
    return elem;
  }
};

struct Graph {  // one utterance of a khg_graph_batch
  const int32_t *arc_off;  // S+1, absolute arc ids
  const int32_t *il, *ns;
  const float *w, *fin;
  int32_t S, start;
};

class ExactDecoder {
 public:
  ExactDecoder(const Graph &g, const float *ll, int64_t ld, const int32_t *row_of_tid, const int32_t *row_of_arc, int32_t T)
      : g_(g), ll_(ll), ld_(ld), row_(row_of_tid), arow_(row_of_arc), T_(T) {
    toks_.SetSize(1000);  // faster-decoder.cc:29
  }

  void Decode(float beam) {
    beam_ = beam;
    // InitDecoding
    toks_.Clear();
    toks_.elems.clear();
    arena_.clear();
    arena_.push_back(Tok{-1, -1, 0.0});
    toks_.Insert(g_.start, 0);
    ProcessNonemitting((double)std::numeric_limits<float>::max());
    for (int32_t t = 0; t < T_; ++t) ProcessNonemitting(ProcessEmitting(t));
  }

  bool ReachedFinal() const {
    for (int32_t e = toks_.list_head; e != -1; e = toks_.elems[e].tail)
      if (arena_[toks_.elems[e].val].cost != kInf && g_.fin[toks_.elems[e].key] != std::numeric_limits<float>::infinity()) return true;
    return false;
  }

  // GetBestPath (final states only: the wrapper calls it after ReachedFinal) + GetLinearSymbolSequence:
  // arcs of the best path in forward order; *cost = graph + acoustic cost as LatticeWeight sums them (float)
  bool BestPath(std::vector<int32_t> *arcs, float *cost) const {
    int32_t best_tok = -1, best_state = -1;
    double best = kInf;
    for (int32_t e = toks_.list_head; e != -1; e = toks_.elems[e].tail) {
      const double c = arena_[toks_.elems[e].val].cost + (double)g_.fin[toks_.elems[e].key];
      if (c < best && c != kInf) { best = c; best_tok = toks_.elems[e].val; best_state = toks_.elems[e].key; }
    }
    if (best_tok < 0) return false;
    arcs->clear();
    float graph = 0.f, ac = 0.f;
    for (int32_t t = best_tok; t != -1 && arena_[t].arc >= 0; t = arena_[t].prev) {
      const Tok &k = arena_[t];
      const float tot = (float)(k.cost - (k.prev >= 0 ? arena_[k.prev].cost : 0.0));
      const float gc = g_.w[k.arc];
      arcs->push_back(k.arc);
      graph += gc;
      ac += tot - gc;
    }
    std::reverse(arcs->begin(), arcs->end());
    graph += g_.fin[best_state];
    *cost = graph + ac;
    return true;
  }

 private:
  static constexpr double kInf = std::numeric_limits<double>::infinity();
  static constexpr int kMinActive = 20;
  static constexpr float kBeamDelta = 0.5f, kHashRatio = 2.0f;

  float AcCost(int32_t t, int32_t arc) const { return -1.f * ll_[(int64_t)(arow_ ? arow_[arc] : row_[g_.il[arc]]) * ld_ + t]; }

  void ProcessNonemitting(double cutoff) {
    queue_.clear();
    for (int32_t e = toks_.list_head; e != -1; e = toks_.elems[e].tail) queue_.push_back(e);
    while (!queue_.empty()) {
      const int32_t e = queue_.back();
      queue_.pop_back();
      const int32_t state = toks_.elems[e].key, tok = toks_.elems[e].val;
      const double tc = arena_[tok].cost;
      if (tc > cutoff) continue;
      for (int32_t a = g_.arc_off[state]; a < g_.arc_off[state + 1]; ++a) {
        if (g_.il[a] != 0) continue;
        const double nc = tc + (double)g_.w[a];  // Token(arc, prev): faster-decoder.h:128-137
        if (nc > cutoff) continue;
        const int32_t nt = (int32_t)arena_.size();
        const int32_t found = toks_.Insert(g_.ns[a], nt);
        if (toks_.elems[found].val == nt) {
          arena_.push_back(Tok{a, tok, nc});
          queue_.push_back(found);
        } else if (arena_[toks_.elems[found].val].cost > nc) {  // *(e_found->val) < *new_tok
          arena_.push_back(Tok{a, tok, nc});
          toks_.elems[found].val = nt;
          queue_.push_back(found);
        }
      }
    }
  }

  double GetCutoff(float *adaptive_beam, int32_t *best_elem) {
    double best_cost = kInf;
    tmp_.clear();
    for (size_t i = 0; i < last_.size(); ++i) {
      const double w = arena_[last_[i].second].cost;
      tmp_.push_back((float)w);
      if (w < best_cost) { best_cost = w; *best_elem = (int32_t)i; }
    }
    const double beam_cutoff = best_cost + (double)beam_;
    double min_active_cutoff = kInf;
    if (tmp_.size() > (size_t)kMinActive) {
      std::nth_element(tmp_.begin(), tmp_.begin() + kMinActive, tmp_.end());
      min_active_cutoff = (double)tmp_[kMinActive];
    }
    if (min_active_cutoff > beam_cutoff) {
      *adaptive_beam = (float)(min_active_cutoff - best_cost + (double)kBeamDelta);
      return min_active_cutoff;
    }
    *adaptive_beam = beam_;
    return beam_cutoff;
  }

  double ProcessEmitting(int32_t frame) {
    last_.clear();
    for (int32_t e = toks_.Clear(); e != -1; e = toks_.elems[e].tail) last_.emplace_back(toks_.elems[e].key, toks_.elems[e].val);
    toks_.elems.clear();
    float adaptive_beam = 0.f;
    int32_t best_elem = -1;
    const double weight_cutoff = GetCutoff(&adaptive_beam, &best_elem);
    const size_t new_sz = (size_t)((float)last_.size() * kHashRatio);  // PossiblyResizeHash
    if (new_sz > toks_.hash_size) toks_.SetSize(new_sz);
    double next_weight_cutoff = kInf;
    if (best_elem >= 0) {
      const int32_t state = last_[best_elem].first;
      const double tc = arena_[last_[best_elem].second].cost;
      for (int32_t a = g_.arc_off[state]; a < g_.arc_off[state + 1]; ++a)
        if (g_.il[a] != 0) {
          const float ac_cost = AcCost(frame, a);
          const double new_weight = (double)g_.w[a] + tc + (double)ac_cost;
          if (new_weight + (double)adaptive_beam < next_weight_cutoff) next_weight_cutoff = new_weight + (double)adaptive_beam;
        }
    }
    for (size_t i = 0; i < last_.size(); ++i) {
      const int32_t state = last_[i].first, tok = last_[i].second;
      const double tc = arena_[tok].cost;
      if (!(tc < weight_cutoff)) continue;
      for (int32_t a = g_.arc_off[state]; a < g_.arc_off[state + 1]; ++a) {
        if (g_.il[a] == 0) continue;
        const float ac_cost = AcCost(frame, a);
        const double new_weight = (double)g_.w[a] + tc + (double)ac_cost;
        if (new_weight < next_weight_cutoff) {
          // Token(arc, ac_cost, prev): cost_ = prev->cost_ + arc.weight + ac_cost  (faster-decoder.h:117-126)
          const double nc = tc + (double)g_.w[a] + (double)ac_cost;
          const int32_t nt = (int32_t)arena_.size();
          arena_.push_back(Tok{a, tok, nc});
          const int32_t found = toks_.Insert(g_.ns[a], nt);
          if (new_weight + (double)adaptive_beam < next_weight_cutoff) next_weight_cutoff = new_weight + (double)adaptive_beam;
          if (toks_.elems[found].val != nt) {
            if (arena_[toks_.elems[found].val].cost > nc) toks_.elems[found].val = nt;
            else arena_.pop_back();
          }
        }
      }
    }
    return next_weight_cutoff;
  }

  const Graph &g_;
  const float *ll_;
  int64_t ld_;
  const int32_t *row_, *arow_;
  int32_t T_;
  float beam_ = 0.f;
  HashList toks_;
  std::vector<Tok> arena_;
  std::vector<std::pair<int32_t, int32_t>> last_;  // (state, token) in list order
  std::vector<int32_t> queue_;
  std::vector<float> tmp_;
};

}  // namespace

// One utterance.  ll: rows x ld floats, SCALED log-likelihoods (DecodableAmDiagGmmScaled::LogLikelihood);
// the row of an emitting arc is row_of_arc[absolute arc id] if given, else row_of_tid[its transition-id].
// path: absolute arc ids of the best path, epsilons included; *cost = graph + acoustic cost.
khg_status align_exact_host(const khg_graph_batch *gb, int32_t utt, const float *ll, int64_t ld, const int32_t *row_of_tid,
                            const int32_t *row_of_arc, float beam, float retry_beam, int32_t *alignment, int32_t *status,
                            float *cost, std::vector<int32_t> *path) {
  const int32_t s0 = gb->state_offsets[utt];
  Graph g;
  g.S = gb->state_offsets[utt + 1] - s0;
  g.start = gb->start_state[utt];
  g.arc_off = gb->arc_offsets + s0;
  g.il = gb->arc_ilabel;
  g.ns = gb->arc_nextstate;
  g.w = gb->arc_weight;
  g.fin = gb->final_cost + s0;
  const int32_t T = (int32_t)(gb->frame_offsets[utt + 1] - gb->frame_offsets[utt]);
  *status = KHG_ALIGN_FAILED;
  *cost = 0.f;
  path->clear();
  for (int32_t t = 0; t < T; ++t) alignment[t] = 0;
  if (g.start < 0 || g.S <= 0) return KHG_OK;  // decoder-wrappers.cc:36-42
  ExactDecoder dec(g, ll, ld, row_of_tid, row_of_arc, T);
  dec.Decode(beam);
  bool ans = dec.ReachedFinal();
  int32_t st = KHG_ALIGN_OK;
  if (!ans && retry_beam != 0.f) {
    st = KHG_ALIGN_RETRIED;
    dec.Decode(retry_beam);  // the same decoder: its hash keeps the size it grew to
    ans = dec.ReachedFinal();
  }
  if (!ans || !dec.BestPath(path, cost)) {
    path->clear();
    return KHG_OK;
  }
  int32_t t = 0;
  for (int32_t a : *path)
    if (g.il[a] != 0 && t < T) alignment[t++] = g.il[a];
  *status = st;
  return KHG_OK;
}

}  // namespace khg

using namespace khg;

extern "C" khg_status khg_align_utterance_host(const khg_graph_batch *gb, int32_t utt, const float *loglikes, int64_t ld,
                                               const int32_t *tid2row, int32_t n_tids, float acoustic_scale, float beam,
                                               float retry_beam, int32_t *alignment, int32_t *status, float *like,
                                               int32_t *path_arcs, int32_t path_capacity, int32_t *path_len) {
  KHG_REQUIRE(gb && utt >= 0 && utt < gb->n_utts && loglikes && tid2row && n_tids > 0 && alignment && status, "null / out-of-range argument");
  if ((retry_beam != 0 && retry_beam <= beam) || beam <= 0.0f) {  // decoder-wrappers.cc:29-33
    set_error("Beams do not make sense: beam " + std::to_string(beam) + ", retry-beam " + std::to_string(retry_beam));
    return KHG_ERR_INVALID;
  }
  KHG_REQUIRE(acoustic_scale != 0.f, "acoustic_scale must not be 0");
  const int32_t s0 = gb->state_offsets[utt], s1 = gb->state_offsets[utt + 1];
  for (int32_t a = gb->arc_offsets[s0]; a < gb->arc_offsets[s1]; ++a)
    KHG_REQUIRE(gb->arc_ilabel[a] >= 0 && gb->arc_ilabel[a] < n_tids && gb->arc_nextstate[a] >= 0 && gb->arc_nextstate[a] < s1 - s0,
                "graph: label / state out of range");
  std::vector<int32_t> path;
  float cost = 0.f;
  KHG_TRY(align_exact_host(gb, utt, loglikes, ld, tid2row, nullptr, beam, retry_beam, alignment, status, &cost, &path));
  if (like) *like = *status == KHG_ALIGN_FAILED ? 0.f : -cost / acoustic_scale;  // decoder-wrappers.cc:91
  if (path_len) *path_len = (int32_t)path.size();
  if (path_arcs) {
    KHG_REQUIRE((int64_t)path.size() <= path_capacity, "path_capacity too small");
    std::copy(path.begin(), path.end(), path_arcs);
  }
  return KHG_OK;
}
