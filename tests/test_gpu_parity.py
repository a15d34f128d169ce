"""GPU parity tests proper: the CUDA path, called through the C ABI
(include/khg_b200.h), against the CPU oracle on the same seeded inputs, the
committed golden fixtures, and size-independent properties at larger sizes.

Tolerances are BASELINE.json's: per-frame log-likelihoods 1e-4 relative /
1e-3 absolute; accumulated statistics 1e-4 relative; integer work (bucketing,
transition counts, frame counts) bit-exact.
"""
import json
import os

import numpy as np
import pytest

from oracle import khg_oracle as ko

pytestmark = pytest.mark.gpu

LL_ATOL, LL_RTOL, STATS_RTOL = 1e-3, 1e-4, 1e-4


def _assert_ll(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    assert err.max() <= LL_ATOL, f"abs err {err.max()}"
    big = np.abs(ref) > 10.0
    if big.any():
        assert (err[big] / np.abs(ref[big])).max() <= LL_RTOL


def _assert_stats(got, ref, T):
    for k in ("occ", "mean", "var"):
        if ref[k] is None:
            assert got[k] is None
            continue
        scale = np.abs(ref[k]).max() + 1e-30
        # 1e-4 relative; entries that cancel to something tiny next to the array's
        # scale are compared against 1e-6 of that scale (fp32 posterior resolution)
        np.testing.assert_allclose(got[k], ref[k], rtol=STATS_RTOL, atol=1e-6 * scale)
    assert abs(got["tot_frames"] - ref["tot_frames"]) <= 1e-9 * max(1.0, abs(ref["tot_frames"]))
    assert abs(got["tot_like"] - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"]) + 1e-6


def _device_model(model, kernel=None):
    from kaldi_hmm_gmm_b200 import DeviceModel

    dm = DeviceModel(model.dim, model.offsets)
    if kernel is not None:
        dm.set_kernel(kernel)
    nbad = dm.upload(model.weights, model.means_invvars, model.inv_vars)
    return dm, nbad


@pytest.fixture(scope="module")
def small(oracle):
    model, means, vars_ = ko.make_synthetic_model(39, 13, 100, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 1000)
    return model, feats, pdf


def test_gconsts_on_device(oracle, small, golden_dir):
    from kaldi_hmm_gmm_b200 import _cabi
    import ctypes as C

    model, _, _ = small
    dm, nbad = _device_model(model)
    assert nbad == 0
    np.testing.assert_allclose(dm.gconsts(), model.gconsts, rtol=1e-5, atol=1e-4)
    # golden closed forms (python/tests/test_diag_gmm.py:45-51)
    for c in json.load(open(os.path.join(golden_dir, "diag_gmm_closed_form.json")))["cases"]:
        w = np.array(c["weights"], np.float32)
        var = np.array(c["vars"], np.float32)
        iv = (1.0 / var).astype(np.float32)
        miv = (np.array(c["means"], np.float32) * iv).astype(np.float32)
        gc = np.empty(c["nmix"], np.float32)
        nb = C.c_int32()
        _cabi.check(_cabi.lib().khg_compute_gconsts(c["nmix"], c["dim"], w.ctypes.data, miv.ctypes.data, iv.ctypes.data, gc.ctypes.data, C.byref(nb)))
        assert nb.value == 0
        np.testing.assert_allclose(gc, c["gconsts"], rtol=2e-5, atol=1e-5)
    # zero weight -> -inf counted bad; NaN -> error (csrc/diag-gmm.cc:132-141)
    miv = np.ones((2, 3), np.float32)
    iv = np.ones((2, 3), np.float32)
    gc = np.empty(2, np.float32)
    nb = C.c_int32()
    w = np.array([0.0, 1.0], np.float32)
    _cabi.check(_cabi.lib().khg_compute_gconsts(2, 3, w.ctypes.data, miv.ctypes.data, iv.ctypes.data, gc.ctypes.data, C.byref(nb)))
    assert nb.value == 1 and np.isneginf(gc[0])
    iv[1, 1] = -1.0
    w[:] = 0.5
    with pytest.raises(RuntimeError, match="not a number"):
        _cabi.check(_cabi.lib().khg_compute_gconsts(2, 3, w.ctypes.data, miv.ctypes.data, iv.ctypes.data, gc.ctypes.data, C.byref(nb)))


def test_golden_closed_form_likelihoods_and_posteriors(golden_dir):
    """python/tests/test_diag_gmm.py:327-403, 529-553 against the CUDA path."""
    from kaldi_hmm_gmm_b200 import DeviceModel

    for c in json.load(open(os.path.join(golden_dir, "diag_gmm_closed_form.json")))["cases"]:
        w = np.array(c["weights"], np.float32)
        var = np.array(c["vars"], np.float32)
        iv = (1.0 / var).astype(np.float32)
        miv = (np.array(c["means"], np.float32) * iv).astype(np.float32)
        x = np.array(c["x"], np.float32)
        dm = DeviceModel(c["dim"], np.array([0, c["nmix"]], np.int32))
        dm.upload(w, miv, iv)
        mat = dm.pdf_loglikes(0, x)
        assert mat.shape == (x.shape[0], c["nmix"])
        np.testing.assert_allclose(mat, c["component_loglikes"], rtol=1e-4, atol=1e-4)
        ll, post = dm.pdf_posteriors(0, x)
        np.testing.assert_allclose(ll, c["loglike"], atol=1e-4)
        np.testing.assert_allclose(post, c["posteriors"], rtol=1e-4, atol=1e-6)
        dense = dm.loglikes_all_pdfs(x)
        np.testing.assert_allclose(dense[:, 0], c["loglike"], atol=1e-4)


@pytest.mark.parametrize("layout", [0, 1])
def test_dense_loglikes_vs_oracle(oracle, small, layout):
    model, feats, _ = small
    dm, _ = _device_model(model, kernel=1)
    got = dm.loglikes_all_pdfs(feats, scale=0.7, layout=layout)
    ref, bad = oracle.loglikes_all_pdfs(model, feats, scale=0.7, pdf_major=bool(layout))
    assert bad == 0
    _assert_ll(got, ref)


@pytest.mark.parametrize("D,P,G,T", [(40, 50, 500, 2049), (5, 3, 3, 1), (13, 7, 300, 255), (80, 11, 64, 600)])
def test_dense_loglikes_shapes(oracle, D, P, G, T):
    import torch

    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    dm, _ = _device_model(model, kernel=1)
    ref, _ = oracle.loglikes_all_pdfs(model, feats)
    _assert_ll(dm.loglikes_all_pdfs(feats), ref)
    # device-resident buffers, pdf-major with padding in the leading dimension
    dfe = torch.from_numpy(feats).cuda()
    out = torch.full((P, T + 5), float("nan"), device="cuda")
    dm.loglikes_all_pdfs(dfe, layout=1, out=out)
    dm.sync()
    _assert_ll(out[:, :T].T.cpu().numpy(), ref)


def test_acc_stats_ali_vs_oracle(oracle, small):
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, feats, pdf = small
    dm, _ = _device_model(model)
    fw = np.random.default_rng(3).random(feats.shape[0]).astype(np.float32)
    for weights in (None, fw):
        st = DeviceStats(dm)
        pf = np.empty(feats.shape[0], np.float32)
        tot = st.acc_stats_ali(feats, pdf, weights, pf)
        ref = oracle.acc_stats_ali(model, feats, pdf, weights)
        _assert_ll(pf, ref["per_frame"])
        got = st.download()
        _assert_stats(got, ref, feats.shape[0])
        assert abs(tot - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"])
        # occupancies of a pdf sum to the weight mass aligned to it (exact property)
        for p in range(model.num_pdfs):
            mass = float((np.ones_like(fw) if weights is None else fw)[pdf == p].astype(np.float64).sum())
            assert abs(got["occ"][model.offsets[p]:model.offsets[p + 1]].sum() - mass) <= 1e-4 * max(1.0, mass)


@pytest.mark.parametrize("flags,has_mean,has_var", [(4, False, False), (1, True, False), (2, True, True), (0, False, False)])
def test_flags_control_buffers(oracle, small, flags, has_mean, has_var):
    # csrc/mle-diag-gmm.cc:43-62 + csrc/model-common.cc:72-84
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, feats, pdf = small
    dm, _ = _device_model(model)
    st = DeviceStats(dm, flags)
    st.acc_stats_ali(feats, pdf)
    got = st.download()
    assert (got["mean"] is not None) == has_mean and (got["var"] is not None) == has_var
    ref = oracle.acc_stats_ali(model, feats, pdf, flags=flags)
    _assert_stats(got, ref, feats.shape[0])


def test_ragged_big_pdf_and_accumulation_across_calls(oracle):
    from kaldi_hmm_gmm_b200 import DeviceStats

    rng = np.random.default_rng(11)
    D = 20
    sizes = np.array([1, 300, 2, 17, 64, 9, 1, 33], np.int32)  # ragged, one big pdf
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(offsets[-1])
    means = rng.standard_normal((G, D)).astype(np.float32) * 2
    vars_ = rng.uniform(0.5, 2, (G, D)).astype(np.float32)
    w = np.concatenate([rng.dirichlet(np.ones(s)) for s in sizes]).astype(np.float32)
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 1500)
    dm, _ = _device_model(model, kernel=1)
    _assert_ll(dm.loglikes_all_pdfs(feats), oracle.loglikes_all_pdfs(model, feats)[0])
    st = DeviceStats(dm)
    tot = 0.0
    for a, b in [(0, 1), (1, 700), (700, 700), (700, 1500)]:  # incl. single-frame and empty calls
        tot += st.acc_stats_ali(feats[a:b], pdf[a:b])
    ref = oracle.acc_stats_ali(model, feats, pdf)
    _assert_stats(st.download(), ref, 1500)
    assert abs(tot - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"])


def test_tids_path_bit_exact_integers(oracle, small):
    """tid->pdf mapping and transition counts are integer work: bit-exact.
    scripts/gmm_acc_stats_ali.py:46-56; test there asserts sum(transition_accs)==frames
    (scripts/test_gmm_acc_stats_ali.py:106)."""
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, feats, _ = small
    rng = np.random.default_rng(5)
    num_tids = 4 * model.num_pdfs
    tid2pdf = np.concatenate([[0], rng.integers(0, model.num_pdfs, num_tids)]).astype(np.int32)
    tids = rng.integers(1, num_tids + 1, feats.shape[0]).astype(np.int32)
    dm, _ = _device_model(model)
    st = DeviceStats(dm)
    trans = np.zeros(num_tids + 1, np.float64)
    trans[3] = 2.0
    tot = st.acc_stats_ali_tids(feats, tids, tid2pdf, trans)
    expect = np.bincount(tids, minlength=num_tids + 1).astype(np.float64)
    expect[3] += 2.0
    assert np.array_equal(trans, expect)
    assert trans.sum() - 2.0 == feats.shape[0]
    ref = oracle.acc_stats_ali(model, feats, tid2pdf[tids])
    _assert_stats(st.download(), ref, feats.shape[0])
    assert abs(tot - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"])
    bad = tids.copy()
    bad[7] = num_tids + 1
    with pytest.raises(RuntimeError, match="out of range"):
        st.acc_stats_ali_tids(feats, bad, tid2pdf)


def test_bad_pdf_id_and_nonfinite_raise(oracle, small):
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, feats, pdf = small
    dm, _ = _device_model(model, kernel=1)
    st = DeviceStats(dm)
    bad = pdf.copy()
    bad[3] = model.num_pdfs
    with pytest.raises(RuntimeError, match="out of range"):
        st.acc_stats_ali(feats, bad)
    f2 = feats.copy()
    f2[5, 2] = np.nan
    with pytest.raises(RuntimeError, match="Invalid answer"):  # csrc/diag-gmm.cc:385-387
        DeviceStats(dm).acc_stats_ali(f2, pdf)
    with pytest.raises(RuntimeError, match="Invalid answer"):  # csrc/decodable-am-diag-gmm.cc:63-65
        dm.loglikes_all_pdfs(f2)
    # a pdf whose Gaussians all have zero weight: LogSumExp is NaN -> error
    w = model.weights.copy()
    w[model.offsets[2]:model.offsets[3]] = 0.0
    from kaldi_hmm_gmm_b200 import DeviceModel

    dm2 = DeviceModel(model.dim, model.offsets)
    assert dm2.upload(w, model.means_invvars, model.inv_vars) == model.offsets[3] - model.offsets[2]
    with pytest.raises(RuntimeError, match="Invalid answer"):
        dm2.loglikes_all_pdfs(feats[:10])


def test_stats_add_scale_and_posteriors(oracle, small):
    """AccumAmDiagGmm::Add/Scale (csrc/mle-am-diag-gmm.cc:119-138) and
    AccumulateFromPosteriors (csrc/mle-diag-gmm.cc:123-143)."""
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, feats, pdf = small
    dm, _ = _device_model(model)
    a, b = DeviceStats(dm), DeviceStats(dm)
    a.acc_stats_ali(feats[:400], pdf[:400])
    b.acc_stats_ali(feats[400:], pdf[400:])
    a.add(1.0, b)
    ref = oracle.acc_stats_ali(model, feats, pdf)
    _assert_stats(a.download(), ref, feats.shape[0])
    a.scale(0.5)
    got = a.download()
    np.testing.assert_allclose(got["occ"], 0.5 * ref["occ"], rtol=1e-4, atol=1e-7)
    assert abs(got["tot_frames"] - 0.5 * ref["tot_frames"]) < 1e-6
    c = DeviceStats(dm)
    p = 4
    ng = model.offsets[p + 1] - model.offsets[p]
    post = np.random.default_rng(2).random((3, ng)).astype(np.float32)
    c.acc_from_posteriors(p, feats[:3], post)
    got = c.download()
    occ = np.zeros(model.num_gauss)
    mean = np.zeros((model.num_gauss, model.dim))
    var = np.zeros_like(mean)
    s = slice(model.offsets[p], model.offsets[p + 1])
    for t in range(3):
        oracle.acc_from_posteriors(ko.kGmmAll, feats[t], post[t], occ[s], mean[s], var[s])
    np.testing.assert_allclose(got["occ"], occ, rtol=1e-6)
    np.testing.assert_allclose(got["mean"], mean, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["var"], var, rtol=1e-6, atol=1e-9)
    assert abs(got["tot_frames"] - post.sum()) < 1e-5 and got["tot_like"] == 0.0


def test_estep_device_and_host_paths(oracle):
    import torch
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, means, vars_ = ko.make_synthetic_model(40, 64, 600, oracle=oracle)
    T = 5000
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    dm, _ = _device_model(model)
    ref = oracle.acc_stats_ali(model, feats, pdf)
    dense_ref, _ = oracle.loglikes_all_pdfs(model, feats)
    block = torch.empty((model.num_pdfs, 2048), device="cuda")
    # device-resident inputs
    st = DeviceStats(dm)
    tot = st.estep(torch.from_numpy(feats).cuda(), torch.from_numpy(pdf).cuda(), block, chunk_frames=2048, want_total=True)
    _assert_stats(st.download(), ref, T)
    assert abs(tot - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"])
    last = T - (T // 2048) * 2048  # the block holds the last chunk
    _assert_ll(block[:, :last].T.cpu().numpy(), dense_ref[T - last:])
    # host inputs streamed through pinned staging
    st2 = DeviceStats(dm)
    tot2 = st2.estep(feats, pdf, block, chunk_frames=2048, want_total=True)
    _assert_stats(st2.download(), ref, T)
    assert abs(tot2 - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"])


def test_estep_statistics_groups(oracle, monkeypatch):
    """khg_estep runs the statistics pass once per GROUP of chunks (two device halves, refilled while the other one
    computes): several groups with a ragged tail, device / pageable / pinned inputs, frame weights; the answer is the
    oracle's and does not depend on the grouping."""
    import torch
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, means, vars_ = ko.make_synthetic_model(40, 50, 400, oracle=oracle)
    T = 23_000
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    w = np.random.default_rng(3).uniform(0.25, 2.0, T).astype(np.float32)
    dm, _ = _device_model(model)
    ref = oracle.acc_stats_ali(model, feats, pdf, frame_weights=w)
    block = torch.empty((model.num_pdfs, 1024), device="cuda")
    hf, hp, hw = (torch.from_numpy(a).pin_memory() for a in (feats, pdf, w))
    dfe, dpd, dw = (torch.from_numpy(a).cuda() for a in (feats, pdf, w))
    for group in ("1", "3000", "5000", "100000"):  # 1, 2, 4 chunks per group (7 groups, the last one ragged); one group
        monkeypatch.setenv("KHG_ESTEP_STATS_GROUP_FRAMES", group)
        for args in ((dfe, dpd, dw), (feats, pdf, w), (hf.numpy(), hp.numpy(), hw.numpy())):
            st = DeviceStats(dm)
            tot = st.estep(args[0], args[1], block, chunk_frames=1024, frame_weights=args[2], want_total=True)
            got = st.download()
            for k in ("occ", "mean", "var"):
                np.testing.assert_allclose(got[k], ref[k], rtol=STATS_RTOL, atol=STATS_RTOL * np.abs(ref[k]).max())
            assert abs(got["tot_frames"] - float(w.astype(np.float64).sum())) < 1e-3
            assert abs(tot - ref["tot_like"]) <= STATS_RTOL * abs(ref["tot_like"])


def test_large_pageable_host_buffers_are_staged_by_the_pool(oracle, monkeypatch):
    """A pageable host buffer of 32 MB or more (a numpy array) reaches the device through two pinned slots filled by
    several host threads (h2d_copy); a ragged last slice, the driver's own pageable copy (KHG_STAGE_THREADS=1) and
    device-resident inputs give the same statistics."""
    import torch
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, means, vars_ = ko.make_synthetic_model(40, 30, 240, oracle=oracle)
    T = 330_001  # 52.8 MB of features: three full slices and a ragged one
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    dm, _ = _device_model(model)
    ref_st = DeviceStats(dm)
    ref_tot = ref_st.acc_stats_ali(torch.from_numpy(feats).cuda(), torch.from_numpy(pdf).cuda(), want_total=True)
    ref = ref_st.download()
    for threads in (None, "1", "3"):
        if threads is None:
            monkeypatch.delenv("KHG_STAGE_THREADS", raising=False)
        else:
            monkeypatch.setenv("KHG_STAGE_THREADS", threads)
        st = DeviceStats(dm)
        pf = np.empty(T, np.float32)
        tot = st.acc_stats_ali(feats, pdf, per_frame=pf, want_total=True)
        got = st.download()
        for k in ("occ", "mean", "var"):
            np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-9 * np.abs(ref[k]).max())
        assert got["tot_frames"] == T and abs(tot - ref_tot) <= 1e-9 * abs(ref_tot)
        assert np.isfinite(pf).all()
    sample = slice(T - 3000, T)  # the tail of the last slice against the oracle
    o = oracle.acc_stats_ali(model, feats[sample], pdf[sample])
    np.testing.assert_allclose(pf[sample], o["per_frame"], rtol=LL_RTOL, atol=LL_ATOL)


def test_dense_block_to_large_pageable_host_outputs(oracle, monkeypatch):
    """khg_loglikes_all_pdfs into a pageable host array of several staging bands (d2h_copy_2d: DMA into two pinned slots,
    the pool's threads move the rows out): both layouts, an output wider than the block (ld_out > row length), against
    the device-resident output of the same call bit for bit; the driver's own copy (KHG_STAGE_THREADS=1) as well."""
    import torch
    from kaldi_hmm_gmm_b200 import _cabi as A

    model, means, vars_ = ko.make_synthetic_model(24, 600, 3000, oracle=oracle)
    T = 21_003
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    dm, _ = _device_model(model)
    dfe = torch.from_numpy(feats).cuda()
    ref_fm = dm.loglikes_all_pdfs(dfe, layout=A.KHG_FRAME_MAJOR).cpu().numpy()
    ref_pm = dm.loglikes_all_pdfs(dfe, layout=A.KHG_PDF_MAJOR).cpu().numpy()
    assert np.array_equal(ref_fm, ref_pm.T)
    P = model.num_pdfs
    for threads in (None, "1"):
        if threads is None:
            monkeypatch.delenv("KHG_STAGE_THREADS", raising=False)
        else:
            monkeypatch.setenv("KHG_STAGE_THREADS", threads)
        assert np.array_equal(dm.loglikes_all_pdfs(feats, layout=A.KHG_FRAME_MAJOR), ref_fm)
        assert np.array_equal(dm.loglikes_all_pdfs(feats, layout=A.KHG_PDF_MAJOR), ref_pm)
        wide = np.full((T, P + 5), -7.0, np.float32)
        dm.loglikes_all_pdfs(feats, layout=A.KHG_FRAME_MAJOR, out=wide)
        assert np.array_equal(wide[:, :P], ref_fm) and (wide[:, P:] == -7.0).all()
        wide = np.full((P, T + 9), -7.0, np.float32)
        dm.loglikes_all_pdfs(feats, layout=A.KHG_PDF_MAJOR, out=wide)
        assert np.array_equal(wide[:, :T], ref_pm) and (wide[:, T:] == -7.0).all()
    sample = slice(T - 500, T)
    o, _ = oracle.loglikes_all_pdfs(model, feats[sample])
    _assert_ll(ref_fm[sample], o)


def test_large_properties(oracle):
    """Size-independent properties at a size the oracle cannot check frame by frame:
    (1) sum of occupancies == number of frames, per pdf, exactly the bucket sizes;
    (2) linearity: stats(A)+stats(B) == stats(A u B); (3) permutation invariance;
    (4) the dense block's aligned-pdf entry equals the stats path's per-frame log-like."""
    import torch
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, means, vars_ = ko.make_synthetic_model(39, 130, 1000, oracle=oracle)
    T = 400_000
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    dm, _ = _device_model(model)
    dfe, dpdf = torch.from_numpy(feats).cuda(), torch.from_numpy(pdf).cuda()
    st = DeviceStats(dm)
    pf = torch.empty(T, device="cuda")
    st.acc_stats_ali(dfe, dpdf, per_frame=pf, want_total=False)
    full = st.download()
    counts = np.bincount(pdf, minlength=model.num_pdfs)
    pdf_occ = np.add.reduceat(full["occ"], model.offsets[:-1])
    np.testing.assert_allclose(pdf_occ, counts, rtol=2e-6)
    assert full["tot_frames"] == T
    half = T // 2
    a, b = DeviceStats(dm), DeviceStats(dm)
    a.acc_stats_ali(dfe[:half], dpdf[:half], want_total=False)
    b.acc_stats_ali(dfe[half:], dpdf[half:], want_total=False)
    a.add(1.0, b)
    got = a.download()
    for k in ("occ", "mean", "var"):
        np.testing.assert_allclose(got[k], full[k], rtol=1e-6, atol=1e-6 * np.abs(full[k]).max())
    perm = torch.randperm(T, device="cuda")
    c = DeviceStats(dm)
    c.acc_stats_ali(dfe[perm].contiguous(), dpdf[perm].contiguous(), want_total=False)
    got = c.download()
    for k in ("occ", "mean", "var"):
        np.testing.assert_allclose(got[k], full[k], rtol=1e-6, atol=1e-6 * np.abs(full[k]).max())
    n = 20000
    dense = dm.loglikes_all_pdfs(dfe[:n], layout=1)
    dm.sync()
    aligned = dense[dpdf[:n].long(), torch.arange(n, device="cuda")]
    assert (aligned - pf[:n]).abs().max().item() < 1e-3
    # oracle spot check on a slice
    ref = oracle.acc_stats_ali(model, feats[:5000], pdf[:5000])
    _assert_ll(pf[:5000].cpu().numpy(), ref["per_frame"])


def _nccl_worker(rank, world, port, out_dir):
    import os
    import sys

    import torch
    import torch.distributed as dist

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "kaldi-hmm-gmm_b200", "python")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats, _cabi
    from kaldi_hmm_gmm_b200 import parallel as par

    _cabi.check(_cabi.lib().khg_set_device(rank))
    ora = ko.Oracle()
    model, means, vars_ = ko.make_synthetic_model(40, 64, 600, oracle=ora)
    T = 20001
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    a, b = par.shard_frames(T, rank, world)
    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    st = DeviceStats(dm)
    st.acc_stats_ali(torch.from_numpy(feats[a:b]).cuda(), torch.from_numpy(pdf[a:b]).cuda(), want_total=False)
    par.allreduce_stats(st)
    torch.cuda.synchronize()
    got = st.download()
    ref = ora.acc_stats_ali(model, feats, pdf)
    _assert_stats(got, ref, T)
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


def test_two_gpu_nccl_allreduce_of_stats(tmp_path):
    """Frames sharded over 2 GPUs, packed stats summed by one NCCL all-reduce == oracle on all frames."""
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(2))


def test_direct_small_batch_path_matches_bucketed_path(oracle):
    """T <= 2048 takes the one-launch direct kernel, larger calls the bucketed kernels: the
    same frames through either path give the same statistics; both match the oracle."""
    from kaldi_hmm_gmm_b200 import DeviceStats

    model, means, vars_ = ko.make_synthetic_model(40, 64, 600, oracle=oracle)
    T = 4000
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    fw = np.random.default_rng(9).random(T).astype(np.float32)
    dm, _ = _device_model(model)
    big = DeviceStats(dm)
    pf_big = np.empty(T, np.float32)
    tot_big = big.acc_stats_ali(feats, pdf, fw, pf_big)            # bucketed
    small = DeviceStats(dm)
    pf_small = np.empty(T, np.float32)
    tot_small = 0.0
    for a in range(0, T, 500):                                      # direct, utterance-sized calls
        tot_small += small.acc_stats_ali(feats[a:a + 500], pdf[a:a + 500], fw[a:a + 500], pf_small[a:a + 500])
    ref = oracle.acc_stats_ali(model, feats, pdf, fw)
    gb, gs = big.download(), small.download()
    _assert_stats(gb, ref, T)
    _assert_stats(gs, ref, T)
    _assert_ll(pf_small, ref["per_frame"])
    np.testing.assert_allclose(pf_small, pf_big, rtol=1e-5, atol=1e-4)
    for k in ("occ", "mean", "var"):
        np.testing.assert_allclose(gs[k], gb[k], rtol=1e-5, atol=1e-6 * np.abs(gb[k]).max())
    assert abs(tot_small - tot_big) <= 1e-6 * abs(tot_big)


def test_pdf_subset_block(oracle, small):
    """khg_loglikes_pdf_subset: only the pdfs on an utterance's graph travel to the host."""
    import torch

    model, feats, _ = small
    dm, _ = _device_model(model)
    ref, _ = oracle.loglikes_all_pdfs(model, feats, scale=0.5)
    sub = np.array([7, 0, 12, 3, 3], np.int32)
    got = dm.loglikes_pdf_subset(feats, sub, scale=0.5)
    assert got.shape == (5, feats.shape[0])
    _assert_ll(got, ref[:, sub].T)
    got_d = dm.loglikes_pdf_subset(torch.from_numpy(feats).cuda(), sub, scale=0.5)
    _assert_ll(got_d.cpu().numpy(), ref[:, sub].T)
    with pytest.raises(RuntimeError, match="out of range"):
        dm.loglikes_pdf_subset(feats, np.array([1, model.num_pdfs], np.int32))


def test_packed_model_pickle_round_trip(oracle):
    """pickle / torch.save of the device-resident packed model: same parameters, same likelihoods."""
    import io
    import pickle

    import torch

    model, means, vars_ = ko.make_synthetic_model(13, 9, 70, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 200)
    dm, _ = _device_model(model)
    ref = dm.loglikes_all_pdfs(feats)
    dm2 = pickle.loads(pickle.dumps(dm))
    buf = io.BytesIO()
    torch.save(dm, buf)
    buf.seek(0)
    dm3 = torch.load(buf, weights_only=False)
    for other in (dm2, dm3):
        a, b = dm.download(), other.download()
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_array_equal(other.loglikes_all_pdfs(feats), ref)
