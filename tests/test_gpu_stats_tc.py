"""K3t — the tensor-core statistics kernel (kaldi-hmm-gmm_b200/csrc/khg_stats_tc.cu) — through the C ABI
(khg_acc_stats_ali, include/khg_b200.h) against the CPU oracle of AccumAmDiagGmm::AccumulateForGmm
(csrc/mle-am-diag-gmm.cc:41-52 -> csrc/mle-diag-gmm.cc:123-158 -> csrc/diag-gmm.cc:368-392).

Tolerances are BASELINE.json's, as in test_gpu_parity.py: per-frame log-likelihoods 1e-3 absolute / 1e-4
relative, statistics 1e-4 relative (entries that cancel below 1e-6 of the array's scale are compared
against that scale), frame counts exact.  Every case asserts which kernel ran (khg_model_stats_kernel).
"""
import os

import numpy as np
import pytest

from oracle import khg_oracle as ko

pytestmark = pytest.mark.gpu

TC, SIMT = 3, 1


def _assert_ll(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    assert err.max() <= 1e-3, f"abs err {err.max()}"
    big = np.abs(ref) > 10.0
    if big.any():
        assert (err[big] / np.abs(ref[big])).max() <= 1e-4


def _assert_stats(got, ref):
    for k in ("occ", "mean", "var"):
        scale = np.abs(ref[k]).max() + 1e-30
        np.testing.assert_allclose(got[k], ref[k], rtol=1e-4, atol=1e-6 * scale, err_msg=k)
    assert abs(got["tot_frames"] - ref["tot_frames"]) <= 1e-6 * max(1.0, abs(ref["tot_frames"]))
    assert abs(got["tot_like"] - ref["tot_like"]) <= 1e-4 * abs(ref["tot_like"]) + 1e-6


def _run(model, feats, pdf, weights=None, expect=TC):
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats

    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    assert dm.stats_kernel() == expect
    st = DeviceStats(dm)
    pf = np.full(feats.shape[0], np.nan, np.float32)
    tot = st.acc_stats_ali(feats, pdf, weights, pf)
    got = st.download()
    return got, tot, pf


def _ragged_model(oracle, D, sizes, seed):
    rng = np.random.default_rng(seed)
    sizes = np.asarray(sizes, np.int32)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(offsets[-1])
    means = (2.0 * rng.standard_normal((G, D))).astype(np.float32)
    vars_ = rng.uniform(0.5, 2, (G, D)).astype(np.float32)
    w = np.concatenate([rng.dirichlet(np.ones(s)) for s in sizes]).astype(np.float32)
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    return ko.PackedModel(offsets, w, miv, iv, gc), means, vars_


@pytest.mark.parametrize("D,P,G,T", [(40, 120, 1100, 30000), (39, 60, 300, 20000), (13, 7, 40, 9000), (40, 50, 1000, 25000), (8, 3, 3, 5000)])
def test_tc_stats_vs_oracle(oracle, D, P, G, T):
    """BASELINE-shaped models at oracle-friendly sizes: dim 40 (16-byte rows), 39 (4-byte aligned rows), 13, 8;
    pdfs of 1, ~5, ~9 and 20 Gaussians (the 32-row operand form); unweighted and weighted."""
    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    fw = (np.random.default_rng(5).random(T) * 2).astype(np.float32)
    for weights in (None, fw):
        got, tot, pf = _run(model, feats, pdf, weights)
        ref = oracle.acc_stats_ali(model, feats, pdf, weights)
        _assert_ll(pf, ref["per_frame"])
        _assert_stats(got, ref)
        assert abs(tot - ref["tot_like"]) <= 1e-4 * abs(ref["tot_like"])


def test_tc_stats_overlapping_gaussians(oracle):
    """Gaussians of a pdf that overlap (posteriors spread over several components — speech-like, unlike the
    well-separated synthetic model): the exact re-evaluation covers every component that matters."""
    rng = np.random.default_rng(21)
    D, P, per = 40, 30, 12
    offsets = (np.arange(P + 1) * per).astype(np.int32)
    G = P * per
    centres = 3.0 * rng.standard_normal((P, D))
    means = (np.repeat(centres, per, 0) + 0.7 * rng.standard_normal((G, D))).astype(np.float32)
    vars_ = rng.uniform(0.6, 1.6, (G, D)).astype(np.float32)
    w = np.concatenate([rng.dirichlet(np.ones(per)) for _ in range(P)]).astype(np.float32)
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 20000)
    got, tot, pf = _run(model, feats, pdf)
    ref = oracle.acc_stats_ali(model, feats, pdf)
    # several components carry weight in most frames
    assert np.median((np.asarray(ref["occ"]) > 0).sum()) > 0
    _assert_ll(pf, ref["per_frame"])
    _assert_stats(got, ref)


def test_tc_stats_many_tiles_per_pdf_and_ragged(oracle):
    """Few pdfs, many frames each: the statistics tile accumulates in TMEM over up to 16 work items, is flushed,
    and starts again (kStkMaxTilesPerFlush); ragged pdf sizes incl. 1, 16, 17 and 32 Gaussians; empty pdfs."""
    model, means, vars_ = _ragged_model(oracle, 24, [1, 16, 17, 32, 5, 9], 31)
    T = 60000
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    pdf[pdf == 4] = 1  # pdf 4 receives no frames
    got, tot, pf = _run(model, feats, pdf)
    ref = oracle.acc_stats_ali(model, feats, pdf)
    _assert_ll(pf, ref["per_frame"])
    _assert_stats(got, ref)
    assert float(np.abs(got["occ"][model.offsets[4]:model.offsets[5]]).max()) == 0.0
    # occupancies of a pdf sum to the number of frames aligned to it (exact property)
    for p in range(model.num_pdfs):
        assert abs(got["occ"][model.offsets[p]:model.offsets[p + 1]].sum() - float((pdf == p).sum())) <= 1e-4 * max(1, (pdf == p).sum())


def test_tc_stats_out_of_range_items_take_the_fp32_kernel(oracle):
    """Frames whose scaled features leave fp16's range (|x 2^-k| > 128) and frame weights above 8 make their work
    items fall through, on the device, to the fp32 kernel; tiny features (fp16 subnormals) stay on the tensor
    cores.  Statistics of the whole batch against the oracle."""
    model, means, vars_ = ko.make_synthetic_model(40, 40, 360, oracle=oracle)
    T = 24000
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    rng = np.random.default_rng(9)
    big = rng.choice(T, 40, replace=False)
    feats[big, rng.integers(0, 40, 40)] = 3000.0           # far outliers (finite): log-likes ~ -1e6, still finite
    tiny = rng.choice(T, 400, replace=False)
    feats[tiny] *= 1e-4                                     # x^2 2^-2k ~ 1e-9: below fp16's normal range
    fw = np.ones(T, np.float32)
    fw[rng.choice(T, 25, replace=False)] = 20.0
    fw[rng.choice(T, 25, replace=False)] = 0.0
    got, tot, pf = _run(model, feats, pdf, fw)
    ref = oracle.acc_stats_ali(model, feats, pdf, fw)
    assert ref["bad"] == 0
    err = np.abs(pf.astype(np.float64) - ref["per_frame"])
    assert (err <= 1e-3 + 1e-4 * np.abs(ref["per_frame"])).all()
    _assert_stats(got, ref)


def test_tc_and_fp32_kernels_agree_and_shapes_outside_fall_back(oracle):
    """KHG_STATS_KERNEL=simt forces the fp32 kernel; models outside the tensor-core kernel's shape (dim > 40, a pdf
    of more than 32 Gaussians) report and use the fp32 kernel."""
    model, means, vars_ = ko.make_synthetic_model(40, 25, 200, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 12000)
    got_tc, _, pf_tc = _run(model, feats, pdf)
    os.environ["KHG_STATS_KERNEL"] = "simt"
    try:
        got_f, _, pf_f = _run(model, feats, pdf, expect=SIMT)
    finally:
        del os.environ["KHG_STATS_KERNEL"]
    assert np.abs(pf_tc - pf_f).max() <= 2e-5
    for k in ("occ", "mean", "var"):
        scale = np.abs(got_f[k]).max()
        np.testing.assert_allclose(got_tc[k], got_f[k], rtol=2e-5, atol=1e-6 * scale)
    for D, sizes in ((48, [3, 4]), (20, [40, 2])):
        m2, mu, va = _ragged_model(oracle, D, sizes, 3)
        f2, p2 = ko.make_synthetic_frames(m2, mu, va, 6000)
        got, _, pf = _run(m2, f2, p2, expect=SIMT)
        _assert_stats(got, oracle.acc_stats_ali(m2, f2, p2))


def test_tc_stats_models_with_some_large_pdfs(oracle):
    """A model after mix-up has pdfs of very different sizes.  Pdfs of more than 32 Gaussians have no tensor-core image:
    their work items are declined one by one and go to the fp32 kernel through the device list, the others stay on the
    tensor-core kernel (at least half of the Gaussians must be in pdfs of at most 32 for the model to take this path)."""
    sizes = [40, 8, 12, 33, 5, 16, 20, 64, 9, 30, 7, 11, 3, 25]
    for D, T, seed in ((40, 50000, 5), (39, 30000, 6), (13, 20000, 7)):
        model, means, vars_ = _ragged_model(oracle, D, sizes, seed)
        feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
        w = np.random.default_rng(seed).uniform(0.5, 1.5, T).astype(np.float32)
        got, tot, pf = _run(model, feats, pdf, w)
        ref = oracle.acc_stats_ali(model, feats, pdf, frame_weights=w)
        _assert_ll(pf, ref["per_frame"])
        _assert_stats(got, ref)
        for p_ in range(model.num_pdfs):  # every pdf, large or small, received its frames
            sel = pdf == p_
            assert abs(got["occ"][model.offsets[p_]:model.offsets[p_ + 1]].sum() - float(w[sel].astype(np.float64).sum())) <= 1e-4 * max(1.0, sel.sum())


def test_tc_stats_nonfinite_features_raise_like_the_reference(oracle):
    """A NaN feature: the item goes to the fp32 kernel, which latches the reference's "Invalid answer"
    (csrc/diag-gmm.cc:158-163) — reported by the synchronising call."""
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats

    model, means, vars_ = ko.make_synthetic_model(40, 10, 60, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 5000)
    feats[1234, 7] = np.nan
    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    assert dm.stats_kernel() == TC
    st = DeviceStats(dm)
    with pytest.raises(Exception):
        st.acc_stats_ali(feats, pdf)
        st.download()


@pytest.mark.parametrize("seed", list(range(8)))
def test_tc_stats_random_shapes(oracle, seed):
    """Random model shapes inside the tensor-core kernel's range (dim 1..40, ragged pdf sizes 1..32, empty pdfs, batches
    just above the bucketing threshold, single-frame items): statistics, totals and per-frame values against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    D = int(rng.choice([1, 2, 3, 7, 8, 9, 13, 16, 20, 24, 31, 32, 33, 39, 40]))
    P = int(rng.integers(1, 60))
    sizes = rng.integers(1, 33, P)
    if seed % 2 == 0:
        sizes = np.minimum(sizes, 16)  # the 16-row instantiation
    model, means, vars_ = _ragged_model(oracle, D, sizes, 7 + seed)
    T = int(rng.choice([2049, 2100, 4097, 9000, 20000]))
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T, seed=seed)
    if P > 3:  # some pdfs get no frames, one gets exactly one
        pdf[pdf == 1] = 0
        pdf[pdf == 2] = 0
        pdf[0] = 2
    weights = None if seed % 3 else (rng.random(T) * 3).astype(np.float32)
    got, tot, pf = _run(model, feats, pdf, weights)
    ref = oracle.acc_stats_ali(model, feats, pdf, weights)
    _assert_ll(pf, ref["per_frame"])
    _assert_stats(got, ref)
