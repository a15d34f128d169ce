"""oracle/khg_oracle.py — ctypes loader for the C oracle + an independent numpy
restatement of the same reference lines.

TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module; the
product package (kaldi_hmm_gmm_b200) never does.

Reference = csukuangfj/kaldi-hmm-gmm v1.1.4; `csrc/` = kaldi-hmm-gmm/csrc/.
The reference cannot be built or imported in this image (Eigen 3.4.0 is an
un-vendored, network-fetched dependency: cmake/eigen.cmake:4-6), so both
restatements are pinned by the reference's own known answers and closed-form
tests (tests/test_oracle.py, tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

kGmmMeans, kGmmVariances, kGmmWeights, kGmmTransitions, kGmmAll = 1, 2, 4, 8, 15

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> None:
    """Compile libkhg_oracle.so / libkhg_oracle_fast.so with oracle/Makefile."""
    targets = ["libkhg_oracle.so", "libkhg_oracle_fast.so"]
    if force or not all(os.path.exists(os.path.join(_HERE, t)) for t in targets):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


def _fp(a, ctype=_f32p):
    return None if a is None else a.ctypes.data_as(ctype)


class MleOpts(C.Structure):
    # csrc/mle-diag-gmm.h:23-45 (MleDiagGmmOptions defaults)
    _fields_ = [
        ("min_gaussian_weight", C.c_float),
        ("min_gaussian_occupancy", C.c_float),
        ("min_variance", C.c_double),
        ("remove_low_count_gaussians", C.c_int32),
    ]


class Oracle:
    """Thin ctypes wrapper; `fast=True` loads the -O3 AVX2 build (CPU baseline)."""

    def __init__(self, fast: bool = False):
        name = "libkhg_oracle_fast.so" if fast else "libkhg_oracle.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.khg_oracle_logsumexp.restype = C.c_float
        L.khg_oracle_logsumexp.argtypes = [_f32p, C.c_int32]
        L.khg_oracle_softmax.restype = None
        L.khg_oracle_softmax.argtypes = [_f32p, C.c_int32, _f32p, _f32p]
        L.khg_oracle_augment_flags.restype = C.c_uint16
        L.khg_oracle_augment_flags.argtypes = [C.c_uint16]
        L.khg_oracle_compute_gconsts.restype = C.c_int32
        L.khg_oracle_compute_gconsts.argtypes = [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, _f32p]
        L.khg_oracle_loglikes.restype = None
        L.khg_oracle_loglikes.argtypes = [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.khg_oracle_loglikes_matrix.restype = None
        L.khg_oracle_loglikes_matrix.argtypes = [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, _f32p, C.c_int64, _f32p]
        L.khg_oracle_log_likelihood.restype = C.c_int32
        L.khg_oracle_log_likelihood.argtypes = [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.khg_oracle_component_posteriors.restype = C.c_int32
        L.khg_oracle_component_posteriors.argtypes = [C.c_int32, C.c_int32, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p]
        L.khg_oracle_acc_from_posteriors.restype = None
        L.khg_oracle_acc_from_posteriors.argtypes = [C.c_int32, C.c_int32, C.c_uint16, _f32p, _f32p, _f64p, _f64p, _f64p]
        L.khg_oracle_acc_for_component.restype = None
        L.khg_oracle_acc_for_component.argtypes = [C.c_int32, C.c_int32, C.c_uint16, _f32p, C.c_int32, C.c_float, _f64p, _f64p, _f64p]
        L.khg_oracle_acc_from_diag.restype = C.c_int32
        L.khg_oracle_acc_from_diag.argtypes = [C.c_int32, C.c_int32, C.c_uint16, _f32p, _f32p, _f32p, _f32p, C.c_float, _f64p, _f64p, _f64p, _f32p]
        L.khg_oracle_acc_stats_ali.restype = C.c_int64
        L.khg_oracle_acc_stats_ali.argtypes = [C.c_int32, C.c_int32, _i32p, _f32p, _f32p, _f32p, C.c_uint16, _f32p, C.c_int64, _i32p, _f32p, _f64p, _f64p, _f64p, _f64p, _f32p]
        L.khg_oracle_acc_stats_ali_mt.restype = C.c_int64
        L.khg_oracle_acc_stats_ali_mt.argtypes = [C.c_int32, C.c_int32, _i32p, _f32p, _f32p, _f32p, C.c_uint16, _f32p, C.c_int64, _i32p, _f32p, _f64p, _f64p, _f64p, _f64p, C.c_int32]
        L.khg_oracle_loglikes_all_pdfs.restype = C.c_int64
        L.khg_oracle_loglikes_all_pdfs.argtypes = [C.c_int32, C.c_int32, _i32p, _f32p, _f32p, _f32p, _f32p, C.c_int64, C.c_float, C.c_int32, _f32p, C.c_int32]
        L.khg_oracle_loglikes_all_pdfs_blocked.restype = C.c_int64
        L.khg_oracle_loglikes_all_pdfs_blocked.argtypes = L.khg_oracle_loglikes_all_pdfs.argtypes
        L.khg_oracle_ml_objective.restype = C.c_float
        L.khg_oracle_ml_objective.argtypes = [C.c_int32, C.c_int32, C.c_uint16, _f32p, _f32p, _f32p, _f64p, _f64p, _f64p]
        L.khg_oracle_mle_update.restype = C.c_int32
        L.khg_oracle_mle_update.argtypes = [C.POINTER(MleOpts), C.c_int32, C.c_int32, C.c_uint16, C.c_uint16, _f64p, _f64p, _f64p, _f32p, _f32p, _f32p, _f32p, _i32p, _f32p, _f32p, _i32p, _i32p, _i32p]

    # -- scalar helpers ---------------------------------------------------
    def logsumexp(self, v):
        v = np.ascontiguousarray(v, np.float32)
        return float(self.lib.khg_oracle_logsumexp(_fp(v), v.size))

    def softmax(self, v):
        v = np.ascontiguousarray(v, np.float32)
        out = np.empty_like(v)
        lse = C.c_float()
        self.lib.khg_oracle_softmax(_fp(v), v.size, _fp(out), C.byref(lse))
        return out, float(lse.value)

    def compute_gconsts(self, weights, miv, iv):
        miv = np.ascontiguousarray(miv, np.float32)
        iv = np.ascontiguousarray(iv, np.float32)
        weights = np.ascontiguousarray(weights, np.float32)
        nmix, dim = miv.shape
        gc = np.empty(nmix, np.float32)
        nb = self.lib.khg_oracle_compute_gconsts(nmix, dim, _fp(weights), _fp(miv), _fp(iv), _fp(gc))
        if nb < 0:
            raise RuntimeError("not a number in gconst computation")
        return gc, nb

    def loglikes(self, gc, miv, iv, x):
        nmix, dim = miv.shape
        out = np.empty(nmix, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        self.lib.khg_oracle_loglikes(nmix, dim, _fp(gc), _fp(miv), _fp(iv), _fp(x), _fp(out))
        return out

    def loglikes_matrix(self, gc, miv, iv, feats):
        nmix, dim = miv.shape
        feats = np.ascontiguousarray(feats, np.float32)
        out = np.empty((feats.shape[0], nmix), np.float32)
        self.lib.khg_oracle_loglikes_matrix(nmix, dim, _fp(gc), _fp(miv), _fp(iv), _fp(feats), feats.shape[0], _fp(out))
        return out

    def log_likelihood(self, gc, miv, iv, x):
        nmix, dim = miv.shape
        x = np.ascontiguousarray(x, np.float32)
        ll = C.c_float()
        bad = self.lib.khg_oracle_log_likelihood(nmix, dim, _fp(gc), _fp(miv), _fp(iv), _fp(x), C.byref(ll))
        if bad:
            raise RuntimeError("Invalid answer (overflow or invalid variances/features?)")
        return float(ll.value)

    def component_posteriors(self, gc, miv, iv, x):
        nmix, dim = miv.shape
        x = np.ascontiguousarray(x, np.float32)
        post = np.empty(nmix, np.float32)
        ll = C.c_float()
        bad = self.lib.khg_oracle_component_posteriors(nmix, dim, _fp(gc), _fp(miv), _fp(iv), _fp(x), _fp(post), C.byref(ll))
        if bad:
            raise RuntimeError("Invalid answer (overflow or invalid variances/features?)")
        return float(ll.value), post

    def acc_from_posteriors(self, flags, x, post, occ, mean_acc, var_acc):
        nmix = occ.shape[0]
        x = np.ascontiguousarray(x, np.float32)
        post = np.ascontiguousarray(post, np.float32)
        self.lib.khg_oracle_acc_from_posteriors(nmix, x.size, flags, _fp(x), _fp(post), _fp(occ, _f64p), _fp(mean_acc, _f64p), _fp(var_acc, _f64p))

    def acc_for_component(self, flags, x, comp, weight, occ, mean_acc, var_acc):
        x = np.ascontiguousarray(x, np.float32)
        self.lib.khg_oracle_acc_for_component(occ.shape[0], x.size, flags, _fp(x), comp, weight, _fp(occ, _f64p), _fp(mean_acc, _f64p), _fp(var_acc, _f64p))

    def acc_from_diag(self, flags, gc, miv, iv, x, weight, occ, mean_acc, var_acc):
        nmix, dim = miv.shape
        x = np.ascontiguousarray(x, np.float32)
        ll = C.c_float()
        bad = self.lib.khg_oracle_acc_from_diag(nmix, dim, flags, _fp(gc), _fp(miv), _fp(iv), _fp(x), weight, _fp(occ, _f64p), _fp(mean_acc, _f64p), _fp(var_acc, _f64p), C.byref(ll))
        if bad:
            raise RuntimeError("Invalid answer (overflow or invalid variances/features?)")
        return float(ll.value)

    # -- packed-model (AmDiagGmm) paths -------------------------------------
    def acc_stats_ali(self, model: "PackedModel", feats, pdf_ids, frame_weights=None,
                      flags=kGmmAll, threads: int = 1, want_per_frame=True):
        m = model
        feats = np.ascontiguousarray(feats, np.float32)
        pdf_ids = np.ascontiguousarray(pdf_ids, np.int32)
        T = feats.shape[0]
        aflags = int(self.lib.khg_oracle_augment_flags(flags))
        G, D = m.means_invvars.shape
        occ = np.zeros(G, np.float64)
        mean = np.zeros((G, D), np.float64) if aflags & kGmmMeans else None
        var = np.zeros((G, D), np.float64) if aflags & kGmmVariances else None
        totals = np.zeros(2, np.float64)
        fw = None if frame_weights is None else np.ascontiguousarray(frame_weights, np.float32)
        if threads <= 1:
            pf = np.empty(T, np.float32) if want_per_frame else None
            bad = self.lib.khg_oracle_acc_stats_ali(D, m.num_pdfs, _fp(m.offsets, _i32p), _fp(m.gconsts), _fp(m.means_invvars), _fp(m.inv_vars), flags, _fp(feats), T, _fp(pdf_ids, _i32p), _fp(fw), _fp(occ, _f64p), _fp(mean, _f64p), _fp(var, _f64p), _fp(totals, _f64p), _fp(pf))
        else:
            pf = None
            bad = self.lib.khg_oracle_acc_stats_ali_mt(D, m.num_pdfs, _fp(m.offsets, _i32p), _fp(m.gconsts), _fp(m.means_invvars), _fp(m.inv_vars), flags, _fp(feats), T, _fp(pdf_ids, _i32p), _fp(fw), _fp(occ, _f64p), _fp(mean, _f64p), _fp(var, _f64p), _fp(totals, _f64p), threads)
        return dict(occ=occ, mean=mean, var=var, tot_like=float(totals[0]), tot_frames=float(totals[1]), per_frame=pf, bad=int(bad))

    def loglikes_all_pdfs(self, model: "PackedModel", feats, scale=1.0, pdf_major=False, threads=1, blocked=False):
        """blocked=True: the frame-blocked matrix form (csrc/diag-gmm.cc:177-189), the CPU baseline."""
        m = model
        feats = np.ascontiguousarray(feats, np.float32)
        T = feats.shape[0]
        out = np.empty((m.num_pdfs, T) if pdf_major else (T, m.num_pdfs), np.float32)
        fn = self.lib.khg_oracle_loglikes_all_pdfs_blocked if blocked else self.lib.khg_oracle_loglikes_all_pdfs
        bad = fn(m.dim, m.num_pdfs, _fp(m.offsets, _i32p), _fp(m.gconsts), _fp(m.means_invvars), _fp(m.inv_vars), _fp(feats), T, scale, int(pdf_major), _fp(out), threads)
        return out, int(bad)

    def ml_objective(self, acc_flags, gc, miv, iv, occ, mean, var):
        nmix, dim = miv.shape
        return float(self.lib.khg_oracle_ml_objective(nmix, dim, acc_flags, _fp(gc), _fp(miv), _fp(iv), _fp(occ, _f64p), _fp(mean, _f64p), _fp(var, _f64p)))

    def mle_update(self, weights, miv, iv, occ, mean, var, acc_flags=kGmmAll, update_flags=kGmmAll,
                   min_gaussian_weight=1e-5, min_gaussian_occupancy=10.0, min_variance=0.001,
                   remove_low_count_gaussians=True):
        """MleDiagGmmUpdate for one pdf (csrc/mle-diag-gmm.cc:243-390). Returns a dict."""
        w = np.array(weights, np.float32, copy=True)
        miv = np.array(miv, np.float32, copy=True)
        iv = np.array(iv, np.float32, copy=True)
        nmix, dim = miv.shape
        gc = np.empty(nmix, np.float32)
        opts = MleOpts(min_gaussian_weight, min_gaussian_occupancy, min_variance, int(remove_low_count_gaussians))
        nout = C.c_int32()
        objc = C.c_float()
        cnt = C.c_float()
        fe, fg, rg = C.c_int32(), C.c_int32(), C.c_int32()
        rc = self.lib.khg_oracle_mle_update(C.byref(opts), nmix, dim, acc_flags, update_flags, _fp(occ, _f64p), _fp(mean, _f64p), _fp(var, _f64p), _fp(w), _fp(miv), _fp(iv), _fp(gc), C.byref(nout), C.byref(objc), C.byref(cnt), C.byref(fe), C.byref(fg), C.byref(rg))
        if rc != 0:
            raise RuntimeError(f"mle_update failed rc={rc}")
        n = nout.value
        return dict(weights=w[:n].copy(), means_invvars=miv[:n].copy(), inv_vars=iv[:n].copy(), gconsts=gc[:n].copy(),
                    obj_change=float(objc.value), count=float(cnt.value), floored_elements=fe.value,
                    floored_gaussians=fg.value, removed_gaussians=rg.value)


@dataclass
class PackedModel:
    """An AmDiagGmm (csrc/am-diag-gmm.h:96) flattened: pdf p owns Gaussians
    [offsets[p], offsets[p+1])."""
    offsets: np.ndarray        # int32 [P+1]
    weights: np.ndarray        # f32 [G]
    means_invvars: np.ndarray  # f32 [G, D]
    inv_vars: np.ndarray       # f32 [G, D]
    gconsts: np.ndarray        # f32 [G]

    @property
    def num_pdfs(self):
        return self.offsets.size - 1

    @property
    def dim(self):
        return self.means_invvars.shape[1]

    @property
    def num_gauss(self):
        return int(self.offsets[-1])


def make_synthetic_model(D: int, P: int, G: int, seed: int = 20230414, oracle: Optional[Oracle] = None) -> Tuple[PackedModel, np.ndarray, np.ndarray]:
    """SURVEY.md §8(d) synthetic model: g_p = G//P (+1 for the first G%P pdfs);
    mean ~ 3*N(0,1), var ~ U(0.5,2), weights = softmax(N(0,1)) within each pdf.
    Returns (PackedModel, means, vars)."""
    rng = np.random.default_rng(seed)
    gp = np.full(P, G // P, np.int32)
    gp[: G % P] += 1
    offsets = np.zeros(P + 1, np.int32)
    np.cumsum(gp, out=offsets[1:])
    means = (3.0 * rng.standard_normal((G, D))).astype(np.float32)
    vars_ = rng.uniform(0.5, 2.0, (G, D)).astype(np.float32)
    logits = rng.standard_normal(G).astype(np.float32)
    weights = np.empty(G, np.float32)
    for p in range(P):
        s = slice(offsets[p], offsets[p + 1])
        e = np.exp(logits[s] - logits[s].max())
        weights[s] = e / e.sum()
    iv = (1.0 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np_compute_gconsts(weights, miv, iv)[0] if oracle is None else np.concatenate(
        [oracle.compute_gconsts(weights[offsets[p]:offsets[p + 1]], miv[offsets[p]:offsets[p + 1]], iv[offsets[p]:offsets[p + 1]])[0] for p in range(P)])
    return PackedModel(offsets, weights, miv, iv, gc.astype(np.float32)), means, vars_


def make_synthetic_frames(model: PackedModel, means, vars_, T: int, seed: int = 20230615):
    """each frame is a sample from a random Gaussian of a random
    pdf; the alignment is that pdf. Returns (feats f32 [T,D], pdf_ids int32 [T])."""
    rng = np.random.default_rng(seed)
    P = model.num_pdfs
    pdf = rng.integers(0, P, T).astype(np.int32)
    gp = (model.offsets[1:] - model.offsets[:-1])
    g = model.offsets[pdf] + (rng.random(T) * gp[pdf]).astype(np.int32)
    x = means[g] + np.sqrt(vars_[g]) * rng.standard_normal((T, model.dim)).astype(np.float32)
    return x.astype(np.float32), pdf


# ---------------------------------------------------------------------------
# Independent numpy restatement (fp32 arithmetic through numpy ufuncs).  It
# shares no code with the C file; tests require the two to agree.
# ---------------------------------------------------------------------------
_LOG_2PI = 1.8378770664093454835606594728112  # csrc/kaldi-math.h:25


def np_logsumexp(v):
    """csrc/eigen.cc:14-18."""
    v = np.asarray(v, np.float32)
    m = v.max()
    return np.float32(np.log(np.exp(v - m, dtype=np.float32).sum(dtype=np.float32), dtype=np.float32) + m)


def np_softmax(v):
    """csrc/eigen.cc:20-32. Returns (softmax, log_sum_exp)."""
    v = np.asarray(v, np.float32)
    m = v.max()
    e = np.exp(v - m, dtype=np.float32)
    s = e.sum(dtype=np.float32)
    return (e / s).astype(np.float32), np.float32(np.log(s, dtype=np.float32) + m)


def np_augment_flags(flags):
    """csrc/model-common.cc:72-84."""
    if flags & kGmmVariances:
        flags |= kGmmMeans
    if flags & kGmmMeans:
        flags |= kGmmWeights
    if not flags & kGmmWeights:
        flags |= kGmmWeights
    return flags


def np_compute_gconsts(weights, miv, iv):
    """csrc/diag-gmm.cc:103-147. Returns (gconsts f32, num_bad)."""
    weights = np.asarray(weights, np.float32)
    miv = np.asarray(miv, np.float32)
    iv = np.asarray(iv, np.float32)
    nmix, dim = miv.shape
    offset = np.float32(-0.5 * _LOG_2PI * dim)
    with np.errstate(divide="ignore", invalid="ignore"):
        gc = (np.log(weights, dtype=np.float32) + offset).astype(np.float32)
        rhs = 0.5 * np.log(iv, dtype=np.float32).astype(np.float64) - 0.5 * miv.astype(np.float64) * miv.astype(np.float64) / iv.astype(np.float64)
        for d in range(dim):  # gc rounded back to float after each += (:122-125)
            gc = (gc.astype(np.float64) + rhs[:, d]).astype(np.float32)
    if np.isnan(gc).any():
        raise RuntimeError("not a number in gconst computation")
    bad = np.isinf(gc)
    gc = np.where(bad & (gc > 0), -gc, gc).astype(np.float32)
    return gc, int(bad.sum())


def np_loglikes_matrix(gc, miv, iv, feats):
    """csrc/diag-gmm.cc:177-189 (row t = csrc/diag-gmm.cc:167-176)."""
    feats = np.atleast_2d(np.asarray(feats, np.float32))
    a = feats @ miv.T.astype(np.float32)
    b = np.square(feats, dtype=np.float32) @ iv.T.astype(np.float32)
    return ((gc[None, :] + a) - np.float32(0.5) * b).astype(np.float32)


def np_acc_stats_ali(model: PackedModel, feats, pdf_ids, frame_weights=None, flags=kGmmAll):
    """scripts/gmm_acc_stats_ali.py:46-56 -> csrc/mle-am-diag-gmm.cc:41-52 ->
    csrc/mle-diag-gmm.cc:145-158, :123-143, frame by frame (small inputs only)."""
    flags = np_augment_flags(flags)
    G, D = model.means_invvars.shape
    occ = np.zeros(G)
    mean = np.zeros((G, D)) if flags & kGmmMeans else None
    var = np.zeros((G, D)) if flags & kGmmVariances else None
    tot_like = 0.0
    tot_frames = 0.0
    per_frame = np.empty(len(pdf_ids), np.float32)
    for t, p in enumerate(pdf_ids):
        s = slice(int(model.offsets[p]), int(model.offsets[p + 1]))
        x = np.asarray(feats[t], np.float32)
        w = np.float32(1.0 if frame_weights is None else frame_weights[t])
        ll = np_loglikes_matrix(model.gconsts[s], model.means_invvars[s], model.inv_vars[s], x)[0]
        post, lse = np_softmax(ll)
        post = (post * w).astype(np.float32)
        occ[s] += post.astype(np.float64)
        if mean is not None:
            mean[s] += np.outer(post, x).astype(np.float32).astype(np.float64)
            if var is not None:
                var[s] += np.outer(post, np.square(x, dtype=np.float32)).astype(np.float32).astype(np.float64)
        tot_like += float(np.float32(lse * w))
        tot_frames += float(w)
        per_frame[t] = lse
    return dict(occ=occ, mean=mean, var=var, tot_like=tot_like, tot_frames=tot_frames, per_frame=per_frame)


def np_loglikes_all_pdfs(model: PackedModel, feats, scale=1.0):
    """csrc/decodable-am-diag-gmm.cc:29-71 for all (frame, pdf): (T, P)."""
    ll = np_loglikes_matrix(model.gconsts, model.means_invvars, model.inv_vars, feats)
    out = np.empty((ll.shape[0], model.num_pdfs), np.float32)
    for p in range(model.num_pdfs):
        seg = ll[:, model.offsets[p]:model.offsets[p + 1]]
        m = seg.max(axis=1, keepdims=True)
        out[:, p] = (np.log(np.exp(seg - m, dtype=np.float32).sum(axis=1, dtype=np.float32), dtype=np.float32) + m[:, 0])
    return (np.float32(scale) * out).astype(np.float32)


def np_log_add(x: np.float32, y: np.float32) -> np.float32:
    """LogAdd(float, float), reference csrc/kaldi-math.h:60-78 (kMinLogDiffFloat = log(FLT_EPSILON))."""
    x, y = np.float32(x), np.float32(y)
    if x < y:
        diff, x = np.float32(x - y), y
    else:
        diff = np.float32(y - x)
    if diff >= np.float32(np.log(np.finfo(np.float32).eps)):
        return np.float32(x + np.float32(np.log1p(np.exp(diff, dtype=np.float32), dtype=np.float32)))
    return x


def np_gaussian_selection(loglikes, num_gselect: int, labels=None):
    """DiagGmm::GaussianSelection / GaussianSelectionPreselect for ONE frame given the candidates'
    log-likelihoods (reference csrc/diag-gmm.cc:202-239, 319-366): threshold = the (n - k)-th order
    statistic (std::nth_element), every candidate >= threshold as a (loglike, label) pair, sorted
    with std::greater, the first k kept; tot = LogAdd chain in that order.  labels = the preselect
    list (defaults to 0..n-1).  Returns (tot float32, list of labels)."""
    ll = np.asarray(loglikes, np.float32)
    n = ll.size
    labels = np.arange(n) if labels is None else np.asarray(labels)
    k = min(int(num_gselect), n)
    # (the preselect form always takes the order statistic, the plain form uses -inf when k == n:
    # the same set either way)
    thresh = np.partition(ll, n - k)[n - k] if k < n else -np.inf
    pairs = sorted(((float(ll[p]), int(labels[p])) for p in range(n) if ll[p] >= thresh), reverse=True)
    tot = np.float32(-np.inf)
    out = []
    for v, lab in pairs[:k]:
        out.append(lab)
        tot = np_log_add(tot, np.float32(v))
    return tot, out


def np_get_split_targets(state_occs, target_components: int, power: float, min_count: float):
    """GetSplitTargets, reference csrc/model-common.cc:14-70: every pdf starts with one Gaussian; the pdf with
    the largest pow(occ, power) / num_components gets the next one unless (n + 1) * min_count >= occ,
    which retires it.  (The reference's priority queue breaks exact ties by its heap order; this
    restatement takes the first maximum — use distinct occupancies when comparing.)"""
    occs = np.asarray(state_occs, np.float32)
    P = occs.size
    occ_pow = np.power(occs.astype(np.float64), np.float64(np.float32(power))).astype(np.float32)  # float occ = pow(...)
    n = np.ones(P, np.int64)
    live = occ_pow.astype(np.float64).copy()
    num_gauss = P
    while num_gauss < target_components:
        key = live / (n + 1.0e-10)
        p = int(np.argmax(key))
        if live[p] == 0:
            break
        if (n[p] + 1) * np.float32(min_count) >= occs[p]:
            live[p] = 0.0
        else:
            n[p] += 1
            num_gauss += 1
    return n.astype(np.int32)


def np_split_by_count(model: "PackedModel", state_occs, target_components: int, perturb_factor: float, power: float,
                      min_count: float, randn):
    """AmDiagGmm::SplitByCount (csrc/am-diag-gmm.cc:72-89) + DiagGmm::Split (csrc/diag-gmm.cc:780-851) +
    ComputeGconsts, float32 like the reference.  randn: (rows, dim) standard-normal vectors, one per
    split, consumed pdf by pdf.  Returns a new PackedModel."""
    targets = np_get_split_targets(state_occs, target_components, power, min_count)
    D = model.dim
    randn = np.asarray(randn, np.float32).reshape(-1, D)
    W, MIV, IV, GC, offs = [], [], [], [], [0]
    row = 0
    pf = np.float32(perturb_factor)
    for p in range(model.num_pdfs):
        s = slice(model.offsets[p], model.offsets[p + 1])
        w, miv, iv = model.weights[s].copy(), model.means_invvars[s].copy(), model.inv_vars[s].copy()
        tgt = int(targets[p])
        if w.size < tgt:
            cur = w.size
            w = np.concatenate([w, np.zeros(tgt - cur, np.float32)])
            miv = np.concatenate([miv, np.zeros((tgt - cur, D), np.float32)])
            iv = np.concatenate([iv, np.zeros((tgt - cur, D), np.float32)])
            while cur < tgt:
                mx = int(np.argmax(w[:cur]))  # first maximum (strict > in the reference's scan)
                w[mx] = np.float32(w[mx] / np.float32(2))
                w[cur] = w[mx]
                r = (randn[row] * np.sqrt(iv[mx])).astype(np.float32)
                row += 1
                iv[cur] = iv[mx]
                step = (r * pf).astype(np.float32)
                miv[cur] = miv[mx] + step
                miv[mx] = miv[mx] - step
                cur += 1
        gc, _ = np_compute_gconsts(w, miv, iv)
        W.append(w), MIV.append(miv), IV.append(iv), GC.append(gc)
        offs.append(offs[-1] + w.size)
    return PackedModel(np.asarray(offs, np.int32), np.concatenate(W), np.concatenate(MIV), np.concatenate(IV),
                       np.concatenate(GC))


def _np_merged_logdet(w1, w2, f1, f2, s1, s2):
    """DiagGmm::MergedComponentsLogdet, reference csrc/diag-gmm.cc:748-767 (float32)."""
    w_sum = np.float32(w1 + w2)
    tm = (f1 + f2 * np.float32(w2 / w1)) * np.float32(w1 / w_sum)
    tv = (s1 + s2 * np.float32(w2 / w1)) * np.float32(w1 / w_sum) - tm * tm
    return np.float32(-0.5 * np.log(tv, dtype=np.float32).sum(dtype=np.float32))


def np_diag_gmm_merge(w, miv, iv, target: int):
    """DiagGmm::Merge, reference csrc/diag-gmm.cc:557-746, float32.  Returns (weights, means_invvars, inv_vars)."""
    w, miv, iv = np.array(w, np.float32), np.array(miv, np.float32), np.array(iv, np.float32)
    n = w.size
    assert 0 < target <= n
    if target == n:
        return w, miv, iv
    vars_ = (np.float32(1) / iv).astype(np.float32)
    means = (miv * vars_).astype(np.float32)
    vars_ = (vars_ + means * means).astype(np.float32)
    if target == 1:
        m1 = (w[None, :] @ means).astype(np.float32)[0]
        m2 = (w[None, :] @ vars_).astype(np.float32)[0]
        wsum = np.float32(w.sum(dtype=np.float32))
        if not (abs(wsum - 1.0) <= 1e-6 * (abs(wsum) + 1.0)):
            m1, m2, wsum = (m1 * wsum).astype(np.float32), (m2 * wsum).astype(np.float32), np.float32(1)
        niv = (np.float32(1) / (m2 - m1 * m1)).astype(np.float32)
        return np.array([wsum], np.float32), (m1 * niv)[None, :].astype(np.float32), niv[None, :]
    disc = np.zeros(n, bool)
    logdet = (np.float32(0.5) * np.log(iv, dtype=np.float32).sum(1, dtype=np.float32)).astype(np.float32)
    delta = np.zeros((n, n), np.float32)
    for i in range(n):
        for j in range(i):
            ml = _np_merged_logdet(w[i], w[j], means[i], means[j], vars_[i], vars_[j])
            delta[i, j] = np.float32(np.float32(np.float32(w[i] + w[j]) * ml) - np.float32(w[i] * logdet[i])) - np.float32(w[j] * logdet[j])
    for _ in range(n - target):
        best, bi, bj = -np.finfo(np.float32).max, -1, -1
        for i in range(n):
            if disc[i]:
                continue
            for j in range(i):
                if not disc[j] and delta[i, j] > best:
                    best, bi, bj = delta[i, j], i, j
        w1, w2 = w[bi], w[bj]
        w_sum = np.float32(w1 + w2)
        r21 = np.float32(w2 / w1)
        means[bi] = ((means[bi] + r21 * means[bj]) * w1 / w_sum).astype(np.float32)
        vars_[bi] = ((vars_[bi] + r21 * vars_[bj]) * w1 / w_sum).astype(np.float32)
        w[bi] = w_sum
        iv[bi] = (np.float32(1) / (vars_[bi] - means[bi] * means[bi])).astype(np.float32)
        miv[bi] = (means[bi] * iv[bi]).astype(np.float32)
        logdet[bi] = np.float32(0.5) * np.log(iv[bi], dtype=np.float32).sum(dtype=np.float32)
        disc[bj] = True
        for j in range(n):
            if j == bi or disc[j]:
                continue
            ml = _np_merged_logdet(w[bi], w[j], means[bi], means[j], vars_[bi], vars_[j])
            t = np.float32(np.float32(np.float32(w[bi] + w[j]) * ml) - np.float32(w[bi] * logdet[bi])) - np.float32(w[j] * logdet[j])
            delta[bi, j] = delta[j, bi] = t
    keep = ~disc
    return w[keep], miv[keep], iv[keep]


def np_merge_by_count(model: "PackedModel", state_occs, target_components: int, power: float, min_count: float):
    """AmDiagGmm::MergeByCount, reference csrc/am-diag-gmm.cc:91-108."""
    targets = np.maximum(np_get_split_targets(state_occs, target_components, power, min_count), 1)
    W, MIV, IV, GC, offs = [], [], [], [], [0]
    for p in range(model.num_pdfs):
        s = slice(model.offsets[p], model.offsets[p + 1])
        w, miv, iv = model.weights[s], model.means_invvars[s], model.inv_vars[s]
        if w.size > targets[p]:
            w, miv, iv = np_diag_gmm_merge(w, miv, iv, int(targets[p]))
        gc, _ = np_compute_gconsts(w, miv, iv)
        W.append(w), MIV.append(miv), IV.append(iv), GC.append(gc)
        offs.append(offs[-1] + w.size)
    return PackedModel(np.asarray(offs, np.int32), np.concatenate(W), np.concatenate(MIV), np.concatenate(IV),
                       np.concatenate(GC))
