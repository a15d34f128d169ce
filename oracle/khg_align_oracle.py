"""CPU oracle for the batched forced aligner — TEST INFRASTRUCTURE ONLY (see oracle/khg_oracle.py).

Restates, in plain Python, the reference's alignment search for one utterance:

* `FasterDecoder` (token passing with beam / min-active pruning): reference
  kaldi-hmm-gmm/csrc/faster-decoder.cc:36-123 (InitDecoding, ProcessNonemitting),
  :125-228 (Decode, AdvanceDecoding, ProcessEmitting), :230-335 (GetCutoff),
  :346-425 (ReachedFinal, GetBestPath); Token arithmetic faster-decoder.h:108-141
  (double cost_, float arc weight and acoustic cost).
* `AlignUtteranceWrapper`: csrc/decoder-wrappers.cc:16-108 (beam, retry beam, like =
  -(graph + acoustic cost) / acoustic_scale).
* `DecodableAmDiagGmmScaled::LogLikelihood(frame, tid)` = scale * loglike(frame, tid2pdf[tid]),
  csrc/decodable-am-diag-gmm.h:94-98, evaluated in float32.

Parity unpinned in the reference: its only use of this code (scripts/test_gmm_align_compiled.py)
asserts nothing and needs OpenFst + data; the oracle is pinned instead by an independent
brute-force Viterbi over all paths (tests/test_oracle_align.py).

A graph is given as plain arrays (what a maintainer exports from the compiled training graph
`fst::VectorFst<StdArc>`): arcs sorted by source state, `arc_offsets[s]..arc_offsets[s+1]`,
`ilabel` (transition-id, 0 = epsilon), `olabel`, `weight` (tropical cost, float32),
`nextstate`; `final[s]` = final cost (inf = not final); `start`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

F32 = np.float32
INF = float("inf")


@dataclass
class Graph:
    arc_offsets: np.ndarray  # int32 [S+1]
    ilabel: np.ndarray       # int32 [A]
    olabel: np.ndarray       # int32 [A]
    weight: np.ndarray       # float32 [A]
    nextstate: np.ndarray    # int32 [A]
    final: np.ndarray        # float32 [S], inf = non-final
    start: int = 0

    @property
    def num_states(self) -> int:
        return int(self.arc_offsets.size - 1)

    @property
    def num_arcs(self) -> int:
        return int(self.ilabel.size)


class _Token:
    __slots__ = ("arc", "prev", "cost")

    def __init__(self, arc: int, prev: Optional["_Token"], cost: float):
        self.arc, self.prev, self.cost = arc, prev, cost


class _Elem:
    __slots__ = ("key", "val", "tail")

    def __init__(self, key, val):
        self.key, self.val, self.tail = key, val, None


class _HashList:
    """csrc/hash-list-inl.h restated: the token container of FasterDecoder.  What matters for parity
    is the ORDER in which ProcessEmitting / ProcessNonemitting / GetCutoff walk the tokens — the
    running `next_weight_cutoff` of ProcessEmitting (faster-decoder.cc:196-216) makes the set of
    surviving tokens depend on it.  The list is the concatenation of the occupied buckets in the
    order they were first occupied; inside a bucket, elements are in insertion order (:128-170).
    bucket = key % hash_size (:133); the size only ever grows (faster-decoder.cc:322-329,
    PossiblyResizeHash) from the constructor's 1000 (:29)."""

    def __init__(self):
        self.hash_size = 0
        self.head = None
        self.bucket_tail = -1
        self.buckets = {}  # index -> [prev_bucket, last_elem]; only occupied buckets are present

    def set_size(self, size: int):  # :26-35
        assert self.head is None and self.bucket_tail == -1
        self.hash_size = size

    def size(self) -> int:
        return self.hash_size

    def clear(self):  # :37-51: hands the list to the caller
        self.buckets = {}
        self.bucket_tail = -1
        head, self.head = self.head, None
        out = []
        while head is not None:
            out.append(head)
            head = head.tail
        return out

    def elems(self):  # GetList()
        e = self.head
        while e is not None:
            yield e
            e = e.tail

    def _bucket_range(self, b):
        prev_bucket, last = b
        head = self.head if prev_bucket == -1 else self.buckets[prev_bucket][1].tail
        return head, last.tail

    def find(self, key):  # :62-82
        b = self.buckets.get(key % self.hash_size)
        if b is None:
            return None
        e, stop = self._bucket_range(b)
        while e is not stop:
            if e.key == key:
                return e
            e = e.tail
        return None

    def insert(self, key, val):  # :128-170: returns the existing element or the new one
        index = key % self.hash_size
        b = self.buckets.get(index)
        if b is not None:
            e, stop = self._bucket_range(b)
            while e is not stop:
                if e.key == key:
                    return e
                e = e.tail
        elem = _Elem(key, val)
        if b is None:  # unoccupied bucket: appended to the chain of buckets = to the end of the list
            if self.bucket_tail == -1:
                self.head = elem
            else:
                self.buckets[self.bucket_tail][1].tail = elem
            elem.tail = None
            self.buckets[index] = [self.bucket_tail, elem]
            self.bucket_tail = index
        else:  # occupied bucket: after its last element, i.e. possibly in the MIDDLE of the list
            elem.tail = b[1].tail
            b[1].tail = elem
            b[1] = elem
        return elem


class FasterDecoderOracle:
    """faster-decoder.cc with the option defaults AlignUtteranceWrapper leaves in place
    (faster-decoder.h:41-43): max_active = int max, min_active = 20, beam_delta = 0.5, hash_ratio = 2."""

    def __init__(self, g: Graph, beam: float, min_active: int = 20, beam_delta: float = 0.5, tight: bool = False,
                 hash_ratio: float = 2.0):
        """tight=False: the reference's rule exactly — the order-dependent running `next_weight_cutoff`, tokens
        visited in the HashList's order (restated above), the hash size growing as PossiblyResizeHash does.
        tight=True: the order-independent rule of the device kernel's fast path — a new token survives iff its
        cost < (lowest new cost of the frame) + adaptive_beam, i.e. the FINAL value of next_weight_cutoff;
        tokens are visited by state id so that ties go to the lowest arc id.  The reference keeps a superset
        of these tokens for one frame; the extra ones are above the next frame's beam cutoff."""
        self.g, self.beam, self.min_active, self.beam_delta = g, float(F32(beam)), min_active, float(F32(beam_delta))
        self.tight = tight
        self.hash_ratio = float(F32(hash_ratio))
        self.hl = _HashList()
        self.hl.set_size(1000)  # faster-decoder.cc:29
        self.frames = 0
        self.max_extra = 0      # diagnostics: most tokens kept above the final cutoff in one frame (tight=False)

    def set_beam(self, beam: float):  # SetOptions, decoder-wrappers.cc:61-63
        self.beam = float(F32(beam))

    @property
    def toks(self):
        """state -> token in the container's order."""
        return {e.key: e.val for e in self.hl.elems()}

    # --- faster-decoder.cc:36-49
    def init_decoding(self):
        self.hl.clear()
        self.hl.insert(self.g.start, _Token(-1, None, 0.0))
        self._process_nonemitting(float(np.finfo(np.float32).max))
        self.frames = 0

    # --- faster-decoder.cc:51-123
    def _process_nonemitting(self, cutoff: float):
        g = self.g
        queue = list(self.hl.elems())
        while queue:
            e = queue.pop()
            state, tok = e.key, e.val
            if tok.cost > cutoff:
                continue
            for a in range(g.arc_offsets[state], g.arc_offsets[state + 1]):
                if g.ilabel[a] != 0:
                    continue
                new = _Token(a, tok, tok.cost + float(g.weight[a]))  # faster-decoder.h:128-137
                if new.cost > cutoff:
                    continue
                found = self.hl.insert(int(g.nextstate[a]), new)
                if found.val is new:
                    queue.append(found)
                elif found.val.cost > new.cost:  # *(e_found->val) < *new_tok
                    found.val = new
                    queue.append(found)

    # --- faster-decoder.cc:230-335 (max_active = int max)
    def _get_cutoff(self, elems) -> Tuple[float, float, Optional[_Elem]]:
        best, best_elem = INF, None
        costs = []
        for e in elems:
            costs.append(e.val.cost)
            if e.val.cost < best:
                best, best_elem = e.val.cost, e
        beam_cutoff = best + self.beam
        min_active_cutoff = INF
        if len(costs) > self.min_active:
            if self.min_active == 0:
                min_active_cutoff = best
            else:
                # tmp_array_ is std::vector<float>: costs are rounded to float there
                arr = np.sort(np.asarray(costs, dtype=np.float32))
                min_active_cutoff = float(arr[self.min_active])
        if min_active_cutoff > beam_cutoff:
            return min_active_cutoff, float(F32(min_active_cutoff - best + self.beam_delta)), best_elem
        return beam_cutoff, self.beam, best_elem

    # --- faster-decoder.cc:154-228
    def _process_emitting(self, loglike) -> float:
        g = self.g
        frame = self.frames
        last = self.hl.clear()
        weight_cutoff, adaptive_beam, best_elem = self._get_cutoff(last)
        new_sz = int(F32(F32(len(last)) * F32(self.hash_ratio)))  # PossiblyResizeHash, :322-329
        if new_sz > self.hl.size():
            self.hl.set_size(new_sz)
        next_cutoff = INF

        def arcs_of(state):
            for a in range(g.arc_offsets[state], g.arc_offsets[state + 1]):
                if g.ilabel[a] != 0:
                    ac = float(F32(-1) * F32(loglike(frame, int(g.ilabel[a]))))
                    yield a, ac

        if self.tight:
            for e in last:
                if e.val.cost < weight_cutoff:
                    for a, ac in arcs_of(e.key):
                        next_cutoff = min(next_cutoff, float(g.weight[a]) + e.val.cost + ac + adaptive_beam)
            for e in sorted(last, key=lambda x: x.key):
                tok = e.val
                if tok.cost < weight_cutoff:
                    for a, ac in arcs_of(e.key):
                        nw = float(g.weight[a]) + tok.cost + ac
                        if nw < next_cutoff:
                            found = self.hl.insert(int(g.nextstate[a]), None)
                            if found.val is None or found.val.cost > nw:
                                found.val = _Token(a, tok, nw)
            self.frames += 1
            return next_cutoff
        if best_elem is not None:  # :177-190: the best token first, for a tight bound
            for a, ac in arcs_of(best_elem.key):
                nw = float(g.weight[a]) + best_elem.val.cost + ac
                if nw + adaptive_beam < next_cutoff:
                    next_cutoff = nw + adaptive_beam
        for e in last:  # :197-225
            tok = e.val
            if tok.cost < weight_cutoff:
                for a, ac in arcs_of(e.key):
                    nw = float(g.weight[a]) + tok.cost + ac
                    if nw < next_cutoff:
                        # Token(arc, ac_cost, prev): cost_ = prev->cost_ + weight + ac_cost
                        new = _Token(a, tok, tok.cost + float(g.weight[a]) + ac)
                        found = self.hl.insert(int(g.nextstate[a]), new)
                        if nw + adaptive_beam < next_cutoff:
                            next_cutoff = nw + adaptive_beam
                        if found.val is not new and found.val.cost > new.cost:
                            found.val = new
        self.max_extra = max(self.max_extra, sum(1 for e in self.hl.elems() if e.val.cost >= next_cutoff))
        self.frames += 1
        return next_cutoff

    def decode(self, num_frames: int, loglike):
        self.init_decoding()
        while self.frames < num_frames:
            cutoff = self._process_emitting(loglike)
            self._process_nonemitting(cutoff)

    # --- faster-decoder.cc:346-354
    def reached_final(self) -> bool:
        return any(e.val.cost != INF and self.g.final[e.key] != INF for e in self.hl.elems())

    # --- faster-decoder.cc:356-425
    def best_path(self):
        """(arc ids of the best path, graph cost, acoustic cost) or None."""
        g = self.g
        best_tok, best_state = None, None
        is_final = self.reached_final()
        if not is_final:
            for e in self.hl.elems():
                if best_tok is None or best_tok.cost > e.val.cost:
                    best_tok, best_state = e.val, e.key
        else:
            best = INF
            for e in self.hl.elems():
                c = e.val.cost + float(g.final[e.key])
                if c < best and c != INF:
                    best, best_tok, best_state = c, e.val, e.key
        if best_tok is None:
            return None
        arcs, graph, ac = [], F32(0), F32(0)
        tok = best_tok
        while tok is not None and tok.arc >= 0:
            tot = F32(tok.cost - (tok.prev.cost if tok.prev else 0.0))
            gc = F32(g.weight[tok.arc])
            arcs.append(tok.arc)
            graph, ac = F32(graph + gc), F32(ac + F32(tot - gc))
            tok = tok.prev
        arcs.reverse()
        if is_final:
            graph = F32(graph + F32(g.final[best_state]))
        return arcs, float(graph), float(ac)


def align_utterance(g: Graph, loglikes_pdf_major: np.ndarray, tid2pdf: np.ndarray, acoustic_scale: float,
                    beam: float = 200.0, retry_beam: float = 0.0, tight: bool = False):
    """AlignUtteranceWrapper (csrc/decoder-wrappers.cc:16-108) for one utterance.
    loglikes_pdf_major: (P, T) float32 UNSCALED all-pdf log-likelihoods.
    Returns dict(status 0 ok / 1 ok after retry / 2 failed, alignment, words, like, path)."""
    T = int(loglikes_pdf_major.shape[1])
    scale = F32(acoustic_scale)

    def loglike(frame: int, tid: int) -> np.float32:  # decodable-am-diag-gmm.h:94-98
        return F32(scale * F32(loglikes_pdf_major[int(tid2pdf[tid]), frame]))

    if (retry_beam != 0 and retry_beam <= beam) or beam <= 0:
        raise RuntimeError(f"Beams do not make sense: beam {beam}, retry-beam {retry_beam}")
    if g.start < 0:  # decoder-wrappers.cc:36-42: empty graph -> num_error
        return dict(status=2, alignment=[], words=[], like=0.0, path=[])
    dec = FasterDecoderOracle(g, beam, tight=tight)
    dec.decode(T, loglike)
    status = 0
    ok = dec.reached_final()
    if not ok and retry_beam != 0:
        status = 1
        dec.set_beam(retry_beam)  # the same decoder object: its hash keeps the size it grew to
        dec.decode(T, loglike)
        ok = dec.reached_final()
    if not ok:
        return dict(status=2, alignment=[], words=[], like=0.0, path=[])
    arcs, graph, ac = dec.best_path()
    ali = [int(g.ilabel[a]) for a in arcs if g.ilabel[a] != 0]
    words = [int(g.olabel[a]) for a in arcs if g.olabel[a] != 0]
    like = -(graph + ac) / float(acoustic_scale)
    return dict(status=status, alignment=ali, words=words, like=like, path=arcs, max_extra=dec.max_extra)


def brute_force_best(g: Graph, loglikes_pdf_major: np.ndarray, tid2pdf: np.ndarray, acoustic_scale: float):
    """Exact Viterbi by dynamic programming over (frame, state) in float64 — the independent
    check that pins the token-passing restatement (no pruning, so compare with a wide beam)."""
    S, T = g.num_states, int(loglikes_pdf_major.shape[1])
    src = np.repeat(np.arange(S), np.diff(g.arc_offsets))
    cost = np.full(S, INF)
    cost[g.start] = 0.0

    def closure(c):
        changed = True
        while changed:
            changed = False
            for a in range(g.num_arcs):
                if g.ilabel[a] == 0 and c[src[a]] + float(g.weight[a]) < c[g.nextstate[a]]:
                    c[g.nextstate[a]] = c[src[a]] + float(g.weight[a])
                    changed = True
        return c

    cost = closure(cost)
    for t in range(T):
        new = np.full(S, INF)
        for a in range(g.num_arcs):
            if g.ilabel[a] != 0 and cost[src[a]] < INF:
                ac = -float(F32(F32(acoustic_scale) * F32(loglikes_pdf_major[int(tid2pdf[g.ilabel[a]]), t])))
                c = cost[src[a]] + float(g.weight[a]) + ac
                if c < new[g.nextstate[a]]:
                    new[g.nextstate[a]] = c
        cost = closure(new)
    tot = cost + g.final.astype(np.float64)
    return float(tot.min()) if np.isfinite(tot).any() else INF


def make_training_graph(rng: np.random.Generator, phones: List[int], states_per_phone: int = 3,
                        optional_sil: bool = True, alt_prob: float = 0.3) -> Tuple[Graph, int]:
    """A synthetic compiled training graph: a left-to-right chain of HMM states with self loops
    (transition-ids 2k+1 = self loop, 2k+2 = forward of HMM-state k), an optional epsilon skip
    over "silence" phones and occasional alternative pronunciations (a parallel branch).
    Returns (graph, number of transition ids used + 1)."""
    arcs = []  # (src, ilabel, olabel, weight, dst)
    n_states = 1
    cur = 0
    max_tid = 0

    def add_phone(src: int, ph: int, word: int) -> int:
        nonlocal n_states, max_tid
        s = src
        for k in range(states_per_phone):
            hs = ph * states_per_phone + k
            t_self, t_fwd = 2 * hs + 1, 2 * hs + 2
            max_tid = max(max_tid, t_fwd)
            nxt = n_states
            n_states += 1
            # enter the HMM state with its forward tid, loop with the self-loop tid
            arcs.append((s, t_fwd, word if k == 0 else 0, float(rng.uniform(0.1, 1.5)), nxt))
            arcs.append((nxt, t_self, 0, float(rng.uniform(0.1, 1.5)), nxt))
            s = nxt
        return s

    for i, ph in enumerate(phones):
        end = add_phone(cur, ph, i + 1)
        if rng.random() < alt_prob:  # alternative pronunciation in parallel, joined by an epsilon
            alt_end = add_phone(cur, (ph + 7) % max(1, max(phones) + 1), i + 1)
            arcs.append((alt_end, 0, 0, float(rng.uniform(0.0, 0.7)), end))
        if optional_sil and ph == 0:  # optional silence: epsilon skip
            arcs.append((cur, 0, 0, float(rng.uniform(0.0, 0.7)), end))
        cur = end
    arcs.sort(key=lambda a: a[0])
    src = np.array([a[0] for a in arcs], np.int32)
    offs = np.zeros(n_states + 1, np.int32)
    np.add.at(offs, src + 1, 1)
    offs = np.cumsum(offs).astype(np.int32)
    final = np.full(n_states, np.inf, np.float32)
    final[cur] = np.float32(rng.uniform(0.0, 1.0))
    g = Graph(offs, np.array([a[1] for a in arcs], np.int32), np.array([a[2] for a in arcs], np.int32),
              np.array([a[3] for a in arcs], np.float32), np.array([a[4] for a in arcs], np.int32), final, 0)
    return g, max_tid + 1


def make_tid2pdf(n_tids: int, num_pdfs: int) -> np.ndarray:
    """tid -> pdf for make_training_graph's transition ids (2k+1 / 2k+2 belong to HMM-state k);
    index 0 unused (csrc/transition-information.h:71-73)."""
    t = np.arange(n_tids, dtype=np.int64)
    out = (((t - 1) // 2) % num_pdfs).astype(np.int32)
    out[0] = 0
    return out


def sample_utterance(rng: np.random.Generator, g: Graph, tid2pdf: np.ndarray, model, means, vars_,
                     max_dur: int = 5, noise: float = 1.0):
    """Frames that follow one path through `g`: walks the graph from the start taking, per state,
    a random number of self loops, emitting for every transition id one frame drawn from a random
    Gaussian of its pdf.  Returns (feats float32 [T, D], the walked transition ids)."""
    offs = model.offsets
    s, tids = g.start, []
    while not np.isfinite(g.final[s]) or g.arc_offsets[s + 1] > g.arc_offsets[s]:
        arcs = list(range(g.arc_offsets[s], g.arc_offsets[s + 1]))
        loops = [a for a in arcs if g.nextstate[a] == s and g.ilabel[a] != 0]
        fwd = [a for a in arcs if g.nextstate[a] != s]
        for a in loops[:1]:
            tids += [int(g.ilabel[a])] * int(rng.integers(0, max_dur))
        if not fwd:
            break
        a = fwd[int(rng.integers(0, len(fwd)))]
        if g.ilabel[a] != 0:
            tids.append(int(g.ilabel[a]))
        s = int(g.nextstate[a])
    D = means.shape[1]
    feats = np.empty((len(tids), D), np.float32)
    for i, t in enumerate(tids):
        p = int(tid2pdf[t])
        k = int(rng.integers(offs[p], offs[p + 1]))
        feats[i] = means[k] + noise * np.sqrt(vars_[k]) * rng.standard_normal(D)
    return feats, tids


def make_bushy_graph(rng: np.random.Generator, chain_lengths: List[int], n_hmm_states: int) -> Tuple[Graph, int]:
    """Start state fanning out into len(chain_lengths) parallel left-to-right chains (chain c has
    chain_lengths[c] HMM states with self loops, its last state is final): more than min_active
    tokens are alive from the first frame on, so GetCutoff's min_active branch decides what
    survives under a narrow beam.  Returns (graph, number of transition ids + 1)."""
    arcs, finals, n_states = [], {}, 1
    for L in chain_lengths:
        s = 0
        for _ in range(L):
            hs = int(rng.integers(0, n_hmm_states))
            nxt = n_states
            n_states += 1
            arcs.append((s, 2 * hs + 2, 0, float(rng.uniform(0.1, 1.5)), nxt))
            arcs.append((nxt, 2 * hs + 1, 0, float(rng.uniform(0.1, 1.5)), nxt))
            s = nxt
        finals[s] = float(rng.uniform(0.0, 1.0))
    arcs.sort(key=lambda a: a[0])
    src = np.array([a[0] for a in arcs], np.int32)
    offs = np.zeros(n_states + 1, np.int32)
    np.add.at(offs, src + 1, 1)
    final = np.full(n_states, np.inf, np.float32)
    for s, f in finals.items():
        final[s] = np.float32(f)
    g = Graph(np.cumsum(offs).astype(np.int32), np.array([a[1] for a in arcs], np.int32), np.array([a[2] for a in arcs], np.int32),
              np.array([a[3] for a in arcs], np.float32), np.array([a[4] for a in arcs], np.int32), final, 0)
    return g, 2 * n_hmm_states + 1
