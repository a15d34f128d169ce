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
    if transition_accs is None:
        # TransitionModel::InitStats: zeros(num_transition_ids + 1), csrc/transition-model.h:176-180
        transition_accs = np.zeros(t2p.size, np.float64)
    else:
        transition_accs = _np(transition_accs, np.float64)
    log_like = gmm_accs.accumulate_alignment(model=am_gmm, transition_model=t2p, feats=feats,
                                             ali=_np(ali, np.int32), transition_accs=transition_accs)
    return log_like, transition_accs


def gmm_est(am_gmm, gmm_accs, transition_model=None, transition_accs=None, tcfg=None, gmm_opts=None,
            update_flags: str = "mvwt", verbose: bool = False):
    """GMM part of reference scripts/gmm_est.py:8-73 (the transition update is delegated to
    transition_model.mle_update when the object provides it; mix-up/mix-down is model
    surgery outside this package's scope). Returns (objf_impr, count, avg_like_per_frame)."""
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
