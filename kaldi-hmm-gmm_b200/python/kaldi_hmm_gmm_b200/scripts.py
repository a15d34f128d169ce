"""The reference's "binaries" on the hot path, with their signatures unchanged
(reference scripts/gmm_acc_stats_ali.py:9-58, scripts/gmm_est.py:8-96) but with the
per-frame Python loop replaced by ONE batched call into the CUDA path.

`transition_model` may be the reference's khg.TransitionModel (anything exposing
`id2pdf_id`, reference python/csrc/transition-model.cc:76) or a plain int32 numpy array
tid -> pdf (index 0 unused, csrc/transition-information.h:71-84)."""
from typing import List, Optional, Tuple

import numpy as np

from . import _khg_b200 as _ext


def _np(x, dtype):
    if hasattr(x, "detach"):  # torch tensor
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype)


def _tid2pdf(transition_model) -> np.ndarray:
    src = transition_model.id2pdf_id if hasattr(transition_model, "id2pdf_id") else transition_model
    return np.ascontiguousarray(src, np.int32)


def gmm_acc_stats_ali(am_gmm, gmm_accs, transition_model, feats, ali: List[int],
                      transition_accs: Optional[np.ndarray] = None) -> Tuple[float, np.ndarray]:
    """Same contract as reference scripts/gmm_acc_stats_ali.py:9-58: returns
    (total log-like of the frames, transition_accs); gmm_accs is changed in place."""
    feats = _np(feats, np.float32)
    assert feats.ndim == 2, feats.shape
    assert len(ali) == feats.shape[0], (len(ali), feats.shape[0])
    t2p = _tid2pdf(transition_model)
    given = transition_accs
    if transition_accs is None:
        # TransitionModel::InitStats: zeros(num_transition_ids + 1), csrc/transition-model.h:176-180
        transition_accs = np.zeros(t2p.size, np.float64)
    elif hasattr(transition_accs, "detach"):
        # a torch tensor: a float64 CPU tensor is updated IN PLACE through its numpy view (shared memory); anything
        # else is accumulated in a float64 copy and written back into the caller's tensor below
        t = transition_accs.detach()
        if t.device.type == "cpu" and t.dtype.is_floating_point and t.element_size() == 8 and t.is_contiguous():
            transition_accs = t.numpy()
        else:
            transition_accs = np.ascontiguousarray(t.cpu().numpy(), np.float64)
    else:
        transition_accs = _np(transition_accs, np.float64)
    log_like = gmm_accs.accumulate_alignment(model=am_gmm, transition_model=t2p, feats=feats,
                                             ali=_np(ali, np.int32), transition_accs=transition_accs)
    if hasattr(given, "detach"):
        if not np.shares_memory(transition_accs, given.detach().cpu().numpy() if given.device.type != "cpu" else given.detach().numpy()):
            import torch

            given.detach().copy_(torch.from_numpy(transition_accs).to(given.dtype))
        return log_like, given
    return log_like, transition_accs


def gmm_est(am_gmm, gmm_accs, transition_model=None, transition_accs=None, tcfg=None, gmm_opts=None, mixup: int = 0,
            mixdown: int = 0, perturb_factor: float = 0.01, power: float = 0.2, min_count: float = 20.0,
            update_flags: str = "mvwt", verbose: bool = False, randn=None, seed: int = 0):
    """Reference scripts/gmm_est.py:8-96 with its argument names and defaults: transition update
    (delegated to transition_model.mle_update when the object provides it), MleAmDiagGmmUpdate, then
    mix-down (`merge_by_count`) and mix-up (`split_by_count`) by the per-pdf occupancies, both on the
    device.  Returns (objf_impr, count, avg_like_per_frame) (the reference prints them).
    randn / seed (not in the reference, whose Split draws from a global generator): the standard-normal
    vectors of the mix-up splits, (rows, dim) in the reference's order, or the seed they are drawn from."""
    flags = _ext.str_to_gmm_flags(update_flags)
    if flags & int(_ext.GmmUpdateFlags.kGmmTransitions) and hasattr(transition_model, "mle_update"):
        transition_model.mle_update(transition_accs, tcfg)
    tot_like, tot_t = gmm_accs.tot_log_like, gmm_accs.tot_count
    objf_impr, count = _ext.mle_am_diag_gmm_update(
        config=gmm_opts if gmm_opts is not None else _ext.MleDiagGmmOptions(), amdiag_gmm_acc=gmm_accs,
        flags=flags & ~int(_ext.GmmUpdateFlags.kGmmTransitions), am_gmm=am_gmm)
    if verbose:
        print("GMM update: Overall", objf_impr / count, "objective function improvement per frame over", count, "frames")
        print("GMM update: Overall avg like per frame =", tot_like / tot_t, "over", tot_t, "frames.")
    if mixup != 0 or mixdown != 0:
        # the reference sums get_acc(i).occupancy per pdf (scripts/gmm_est.py:66-69); here only the occupancy vector
        # leaves the device
        pdf_occs = np.asarray(gmm_accs.pdf_occupancies(), np.float32)
        if mixdown != 0:
            am_gmm.merge_by_count(state_occs=pdf_occs, target_components=mixdown, power=power, min_count=min_count)
        if mixup != 0:
            am_gmm.split_by_count(state_occs=pdf_occs, target_components=mixup, perturb_factor=perturb_factor, power=power,
                                  min_count=min_count, randn=randn, seed=seed)
    return objf_impr, count, (tot_like / tot_t if tot_t else float("nan"))


def make_decodable(am_gmm, transition_model, feats, acoustic_scale: float = 1.0):
    """The decodable reference scripts/gmm_align_compiled.py:43-48 builds; its (frames x pdfs)
    likelihood block is computed once on the GPU at construction."""
    return _ext.DecodableAmDiagGmmScaled(am=am_gmm, tm=_tid2pdf(transition_model), feats=_np(feats, np.float32),
                                         scale=acoustic_scale)


def make_decodables(am_gmm, transition_model, feats_list, acoustic_scale: float = 1.0):
    """Decodables for MANY utterances from ONE dense GPU call: the utterances are concatenated,
    the (frames x pdfs) block is computed once by the tensor-core kernel and sliced per
    utterance (the feed of a batched gmm-align-compiled)."""
    feats_list = [_np(f, np.float32) for f in feats_list]
    lens = [f.shape[0] for f in feats_list]
    if not lens:
        return []
    block = am_gmm.log_likelihoods_all_pdfs(np.concatenate(feats_list, axis=0))  # (sum T, P)
    t2p = _tid2pdf(transition_model)
    out, t0 = [], 0
    for n in lens:
        out.append(_ext.DecodableAmDiagGmmScaled.from_block(np.ascontiguousarray(block[t0:t0 + n].T), t2p, acoustic_scale))
        t0 += n
    return out


class TrainingGraph:
    """One compiled training graph as plain arrays — what `fst::VectorFst<StdArc>` holds after
    `add_transition_probs` (reference scripts/gmm_align_compiled.py:36-41): arcs sorted by source
    state, `arc_offsets[s]..arc_offsets[s+1]`; `final[s]` = final cost, inf = not final."""

    def __init__(self, arc_offsets, ilabel, olabel, weight, nextstate, final, start: int = 0):
        self.arc_offsets = np.ascontiguousarray(arc_offsets, np.int32)
        self.ilabel, self.olabel = np.ascontiguousarray(ilabel, np.int32), np.ascontiguousarray(olabel, np.int32)
        self.weight, self.nextstate = np.ascontiguousarray(weight, np.float32), np.ascontiguousarray(nextstate, np.int32)
        self.final, self.start = np.ascontiguousarray(final, np.float32), int(start)

    @classmethod
    def from_fst(cls, fst) -> "TrainingGraph":
        """From anything with the OpenFst-style Python surface kaldifst.StdVectorFst exposes
        (`start`, `num_states`, `final(s)`, and `kaldifst.ArcIterator(fst, s)` yielding arcs with
        ilabel / olabel / weight / nextstate).  Not exercised in this repository's tests (kaldifst is
        not installed here); the array form above is the tested boundary."""
        import kaldifst

        offs, il, ol, w, ns, fin = [0], [], [], [], [], []
        for s in range(fst.num_states):
            for arc in kaldifst.ArcIterator(fst, s):
                il.append(arc.ilabel), ol.append(arc.olabel), ns.append(arc.nextstate)
                w.append(float(getattr(arc.weight, "value", arc.weight)))
            offs.append(len(il))
            f = fst.final(s)
            fin.append(float(getattr(f, "value", f)))
        return cls(offs, il, ol, w, ns, fin, fst.start)


def _careful(g: TrainingGraph) -> TrainingGraph:
    """ModifyGraphForCarefulAlignment (reference csrc/decoder-wrappers.cc:111-144): Concat(fst, rhs)
    where rhs = a copy of fst without final costs, entered through a new final pre-initial state.
    OpenFst's Concat adds, for every final state s of the first operand, an epsilon arc with the
    final cost to the second operand's start, and clears s's final cost."""
    S, A = g.final.size, g.ilabel.size
    src = np.repeat(np.arange(S, dtype=np.int32), np.diff(g.arc_offsets))
    pre = 2 * S                                     # pre_initial of rhs, after offsetting rhs states by S
    fin_states = np.flatnonzero(np.isfinite(g.final)).astype(np.int32)
    a_src = np.concatenate([src, fin_states, src + S, [pre]]).astype(np.int32)
    a_il = np.concatenate([g.ilabel, np.zeros(fin_states.size, np.int32), g.ilabel, [0]]).astype(np.int32)
    a_ol = np.concatenate([g.olabel, np.zeros(fin_states.size, np.int32), g.olabel, [0]]).astype(np.int32)
    a_w = np.concatenate([g.weight, g.final[fin_states], g.weight, [0.0]]).astype(np.float32)
    a_ns = np.concatenate([g.nextstate, np.full(fin_states.size, pre, np.int32), g.nextstate + S, [g.start + S]]).astype(np.int32)
    order = np.argsort(a_src, kind="stable")
    offs = np.zeros(2 * S + 2, np.int32)
    np.add.at(offs, a_src + 1, 1)
    final = np.full(2 * S + 1, np.inf, np.float32)
    final[pre] = 0.0
    return TrainingGraph(np.cumsum(offs), a_il[order], a_ol[order], a_w[order], a_ns[order], final, g.start)


def gmm_align_compiled_batch(am_gmm, transition_model, utts: List[str], fsts, feats_list, align_config,
                             acoustic_scale: float = 1.0, num_done: int = 0, num_error: int = 0, num_retried: int = 0,
                             tot_like: float = 0, frame_count: int = 0):
    """gmm-align-compiled for MANY utterances in one device call (khg_align_batch): the counters,
    `alignment` and `words` of reference scripts/gmm_align_compiled.py:10-79 per utterance.
    `fsts`: TrainingGraph objects (or kaldifst FSTs) with the transition probabilities already
    added (`khg.add_transition_probs`, FST-side work outside this package)."""
    graphs = [g if isinstance(g, TrainingGraph) else TrainingGraph.from_fst(g) for g in fsts]
    if align_config.careful:
        graphs = [_careful(g) if g.start >= 0 and g.final.size else g for g in graphs]
    feats_list = [_np(f, np.float32) for f in feats_list]
    lens = np.asarray([f.shape[0] for f in feats_list], np.int64)
    t2p = _tid2pdf(transition_model)
    cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if len(xs) else np.zeros(0, dt), dt)  # noqa: E731
    narcs = np.cumsum([0] + [g.ilabel.size for g in graphs])
    arc_offsets = cat([np.zeros(1, np.int64)] + [g.arc_offsets[1:].astype(np.int64) + narcs[i] for i, g in enumerate(graphs)], np.int32)
    olabel = cat([g.olabel for g in graphs], np.int32)
    ali, status, like, poff, paths = am_gmm.align_batch(
        feats=np.concatenate(feats_list, 0) if feats_list else np.zeros((0, am_gmm.dim), np.float32),
        frame_offsets=np.concatenate([[0], np.cumsum(lens)]).astype(np.int64),
        state_offsets=np.concatenate([[0], np.cumsum([g.final.size for g in graphs])]).astype(np.int32),
        arc_offsets=arc_offsets, arc_ilabel=cat([g.ilabel for g in graphs], np.int32),
        arc_nextstate=cat([g.nextstate for g in graphs], np.int32), arc_weight=cat([g.weight for g in graphs], np.float32),
        start_state=np.asarray([g.start for g in graphs], np.int32), final_cost=cat([g.final for g in graphs], np.float32),
        tid2pdf=t2p, acoustic_scale=acoustic_scale, beam=align_config.beam, retry_beam=align_config.retry_beam)
    fo = np.concatenate([[0], np.cumsum(lens)])
    alignments, words = [], []
    for u in range(len(graphs)):
        # AlignUtteranceWrapper's counters (reference csrc/decoder-wrappers.cc:36-107)
        if graphs[u].start >= 0 and status[u] != 0 and align_config.retry_beam != 0:
            num_retried += 1
        if status[u] == 2:
            num_error += 1
            alignments.append([]), words.append([])
            continue
        num_done += 1
        tot_like += float(like[u])
        frame_count += int(lens[u])
        alignments.append(ali[fo[u]:fo[u + 1]].tolist())
        ol = olabel[paths[poff[u]:poff[u + 1]]]
        words.append(ol[ol != 0].tolist())
    return {"num_done": num_done, "num_error": num_error, "num_retried": num_retried, "tot_like": tot_like,
            "frame_count": frame_count, "alignment": alignments, "words": words, "status": status.tolist()}


def gmm_align_compiled(am_gmm, transition_model, utt: str, fst, feats, align_config, acoustic_scale: float = 1.0,
                       transition_scale: float = 1.0, self_loop_scale: float = 1.0, num_done: int = 0, num_error: int = 0,
                       num_retried: int = 0, tot_like: float = 0, frame_count: int = 0):
    """Same contract as reference scripts/gmm_align_compiled.py:10-79 for one utterance.  When `fst`
    is a kaldifst FST and the reference package is importable its `add_transition_probs` is applied
    first (like the reference script); a TrainingGraph is taken as already carrying them."""
    if not isinstance(fst, TrainingGraph):
        import kaldi_hmm_gmm as khg  # the reference's FST-side helper (unchanged)

        khg.add_transition_probs(trans_model=transition_model, transition_scale=transition_scale,
                                 self_loop_scale=self_loop_scale, fst=fst)
    r = gmm_align_compiled_batch(am_gmm, transition_model, [utt], [fst], [feats], align_config, acoustic_scale,
                                 num_done, num_error, num_retried, tot_like, frame_count)
    r["alignment"], r["words"] = r["alignment"][0], r["words"][0]
    del r["status"]
    return r
