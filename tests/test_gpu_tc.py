"""tcgen05 (3xTF32, TMA, TMEM) dense log-likelihood kernel vs the CPU oracle and vs the
fp32 SIMT kernel.  Tolerance: BASELINE.json — 1e-3 absolute / 1e-4 relative."""
import numpy as np
import pytest

from oracle import khg_oracle as ko

pytestmark = pytest.mark.gpu

TC, SIMT, TC_F16, AUTO, TC_F16_GS = 2, 1, 3, 0, 4


def _models(model):
    from kaldi_hmm_gmm_b200 import DeviceModel

    out = []
    for k in (TC, SIMT):
        dm = DeviceModel(model.dim, model.offsets)
        dm.set_kernel(k)
        dm.upload(model.weights, model.means_invvars, model.inv_vars)
        out.append(dm)
    return out


def _one(model, kernel):
    from kaldi_hmm_gmm_b200 import DeviceModel

    dm = DeviceModel(model.dim, model.offsets)
    dm.set_kernel(kernel)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    return dm


def _check(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    assert np.isfinite(got).all()
    assert err.max() <= 1e-3, f"abs err {err.max()} at {np.unravel_index(err.argmax(), err.shape)}"
    big = np.abs(ref) > 10
    assert (err[big] / np.abs(ref[big])).max() <= 1e-4


@pytest.mark.parametrize("D,P,G,T", [
    (40, 37, 350, 3000),     # several N tiles, T not a multiple of 128
    (39, 13, 100, 1000),     # one N tile, K8 = 80
    (40, 420, 4000, 700),    # many tiles, few frames -> N range split over CTAs
    (13, 5, 17, 129),        # tiny K (one swizzle atom), single-Gaussian-ish pdfs
    (60, 9, 200, 513),       # K = 121 -> 4 chunks
    (5, 3, 3, 1),            # single frame
])
def test_tc_vs_oracle(oracle, D, P, G, T):
    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    tc, simt = _models(model)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    got = tc.loglikes_all_pdfs(feats)
    _check(got, ref)
    _check(got, simt.loglikes_all_pdfs(feats))
    got_pm = tc.loglikes_all_pdfs(feats, scale=0.1, layout=1)
    _check(got_pm.T * 10.0, ref)


def test_tc_ragged_pdfs_and_zero_weight(oracle):
    rng = np.random.default_rng(11)
    D = 40
    sizes = np.array([1, 240, 2, 17, 64, 9, 1, 33, 100, 139, 1, 1, 1, 230, 11], np.int32)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(offsets[-1])
    means = rng.standard_normal((G, D)).astype(np.float32) * 2
    vars_ = rng.uniform(0.5, 2, (G, D)).astype(np.float32)
    w = np.concatenate([rng.dirichlet(np.ones(s)) for s in sizes]).astype(np.float32)
    w[offsets[3] + 2] = 0.0  # zero-weight Gaussian inside a pdf: gconst = -inf, allowed
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 777)
    tc, _ = _models(model)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    _check(tc.loglikes_all_pdfs(feats), ref)


def test_tc_large_device_resident_vs_simt(oracle):
    import torch

    model, means, vars_ = ko.make_synthetic_model(40, 420, 4000, oracle=oracle)
    T = 148 * 128 * 2 + 77
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    tc, simt = _models(model)
    dfe = torch.from_numpy(feats).cuda()
    a = tc.loglikes_all_pdfs(dfe, layout=1)
    b = simt.loglikes_all_pdfs(dfe, layout=1)
    tc.sync()
    simt.sync()
    assert torch.isfinite(a).all()
    assert (a - b).abs().max().item() < 1e-3
    ref, _ = oracle.loglikes_all_pdfs(model, feats[-300:])
    _check(a[:, -300:].T.cpu().numpy(), ref)


def test_tc_nonfinite_and_all_zero_weight_pdf_raise(oracle):
    from kaldi_hmm_gmm_b200 import DeviceModel

    model, means, vars_ = ko.make_synthetic_model(40, 13, 100, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 300)
    tc, _ = _models(model)
    f2 = feats.copy()
    f2[5, 2] = np.nan
    with pytest.raises(RuntimeError, match="Invalid answer"):
        tc.loglikes_all_pdfs(f2)
    w = model.weights.copy()
    w[model.offsets[2]:model.offsets[3]] = 0.0
    dm = DeviceModel(model.dim, model.offsets)
    dm.set_kernel(TC)
    dm.upload(w, model.means_invvars, model.inv_vars)
    with pytest.raises(RuntimeError, match="Invalid answer"):
        dm.loglikes_all_pdfs(feats[:10])


@pytest.mark.parametrize("kernel", [TC, TC_F16, AUTO], ids=["tf32", "f16", "auto"])
@pytest.mark.parametrize("name,D,P,G,T", [
    ("C2-monophone", 39, 130, 1000, 3000),
    ("C3-tri1", 39, 2000, 10000, 1500),
    ("C4-lda-mllt", 40, 4200, 40000, 1000),
    ("C5-sat", 40, 5000, 100000, 600),
])
def test_baseline_config_shapes(oracle, name, D, P, G, T, kernel):
    """The model shapes BASELINE.json names (configs[1..4]) on EVERY tensor-core kernel choice — the
    3xTF32 split, the 3xFP16 split and the device-gated AUTO that bench.py runs: dense block and
    alignment statistics against the oracle on a slice of frames."""
    import os

    from kaldi_hmm_gmm_b200 import DeviceStats

    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    tc = _one(model, kernel)
    assert tc.dense_kernel() == (TC if kernel == TC else TC_F16)  # AUTO resolves to the fp16 split on these models
    ref, bad = oracle.loglikes_all_pdfs(model, feats, threads=os.cpu_count() or 1)
    assert bad == 0
    _check(tc.loglikes_all_pdfs(feats), ref)
    st = DeviceStats(tc)
    pf = np.empty(T, np.float32)
    tot = st.acc_stats_ali(feats, pdf, per_frame=pf)
    r = oracle.acc_stats_ali(model, feats, pdf)
    got = st.download()
    assert np.abs(pf - r["per_frame"]).max() < 1e-3
    for k in ("occ", "mean", "var"):
        np.testing.assert_allclose(got[k], r[k], rtol=1e-4, atol=1e-6 * np.abs(r[k]).max())
    assert abs(tot - r["tot_like"]) < 1e-4 * abs(r["tot_like"])
    # the aligned-pdf column of the dense block is the per-frame log-like of the stats path
    assert np.abs(ref[np.arange(T), pdf] - r["per_frame"]).max() < 1e-4


@pytest.mark.parametrize("kernel", [TC_F16, AUTO, TC], ids=["f16", "auto", "tf32"])
@pytest.mark.parametrize("name,D,P,G", [("C4-lda-mllt", 40, 4200, 40000), ("C5-sat", 40, 5000, 100000)])
def test_baseline_config_shapes_many_frame_tiles(oracle, name, D, P, G, kernel):
    """C4 / C5 with more than two frame tiles per SM and a ragged tail, device-resident like bench.py and
    khg_align_batch feed them: the launch form used at full size (every N tile of the model, every
    CTA working through many frame tiles).  Oracle on the first and last frames; every entry finite."""
    import os

    import torch

    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    T = 2 * 148 * 128 + 128 * 5 + 77
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    dm = _one(model, kernel)
    out = dm.loglikes_all_pdfs(torch.from_numpy(feats).cuda(), layout=1)
    dm.sync()
    assert out.shape == (P, T) and bool(torch.isfinite(out).all())
    n = 250
    for sl in (slice(0, n), slice(T - n, T), slice(148 * 128 - n // 2, 148 * 128 + n // 2)):
        ref, bad = oracle.loglikes_all_pdfs(model, feats[sl], threads=os.cpu_count() or 1)
        assert bad == 0
        _check(out[:, sl].T.cpu().numpy(), ref)


@pytest.mark.parametrize("D,P,G,T", [(40, 37, 350, 3000), (39, 13, 100, 1000), (40, 420, 4000, 700), (13, 5, 17, 129),
                                     (60, 9, 200, 513), (5, 3, 3, 1)])
def test_tc_f16_split_vs_oracle(oracle, D, P, G, T):
    """3xFP16 split (kind::f16, per-dimension power-of-two scaling): same tolerance as 3xTF32."""
    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    for kernel in (TC_F16, AUTO):
        dm = _one(model, kernel)
        _check(dm.loglikes_all_pdfs(feats), ref)
        _check(dm.loglikes_all_pdfs(feats, scale=0.1, layout=1).T * 10.0, ref)


def test_tc_f16_large_magnitude_model_is_rescaled(oracle):
    """MFCC-like raw magnitudes (means ~ +-80, variances 50..2000): the per-dimension scaling
    keeps the fp16 operands in range; accuracy stays within the tolerance."""
    rng = np.random.default_rng(5)
    D, P, G, T = 39, 20, 200, 600
    offsets = np.arange(0, G + 1, G // P).astype(np.int32)
    means = (80.0 * rng.standard_normal((G, D))).astype(np.float32)
    vars_ = rng.uniform(50, 2000, (G, D)).astype(np.float32)
    w = np.full(G, P / G, np.float32)
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    _check(_one(model, TC_F16).loglikes_all_pdfs(feats), ref)


def test_tc_auto_falls_back_to_tf32_when_out_of_fp16_range(oracle):
    import torch

    model, means, vars_ = ko.make_synthetic_model(40, 37, 350, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 500)
    # (1) features far outside the model's range: x^2 would overflow fp16 -> the device-side
    # gate must route the call to the tf32 kernel, results still within tolerance
    big = feats.copy()
    big[7] *= 300.0
    ref, _ = oracle.loglikes_all_pdfs(model, big)
    auto = _one(model, AUTO)
    got = auto.loglikes_all_pdfs(big)
    err = np.abs(got.astype(np.float64) - ref)
    rel = err / np.maximum(np.abs(ref), 1.0)
    assert np.isfinite(got).all() and rel.max() < 1e-4
    tc = _one(model, TC)
    # host buffers are processed in 256-frame calls here: the call holding frame 7 must have
    # run the tf32 kernel (bit-identical to the forced-tf32 model), the other one the fp16 kernel
    np.testing.assert_array_equal(got[:256], tc.loglikes_all_pdfs(big)[:256])
    # device-resident call stays asynchronous and correct
    out = auto.loglikes_all_pdfs(torch.from_numpy(feats).cuda(), layout=1)
    auto.sync()
    _check(out.T.cpu().numpy(), oracle.loglikes_all_pdfs(model, feats)[0])
    # (2) a model whose parameters do not fit fp16 (tiny variances -> huge means*inv_vars)
    iv = model.inv_vars * 1e5
    miv = model.means_invvars * 1e5
    from kaldi_hmm_gmm_b200 import DeviceModel

    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, miv, iv)  # AUTO: silently stays on tf32
    small = (feats * 0.01).astype(np.float32)
    gcs = dm.gconsts()
    m2 = ko.PackedModel(model.offsets, model.weights, miv.astype(np.float32), iv.astype(np.float32), gcs)
    ref2, _ = oracle.loglikes_all_pdfs(m2, small)
    got2 = dm.loglikes_all_pdfs(small)
    assert (np.abs(got2 - ref2) / np.maximum(np.abs(ref2), 1.0)).max() < 1e-4
    dm.set_kernel(TC_F16)
    with pytest.raises(RuntimeError, match="fp16"):
        dm.loglikes_all_pdfs(small)
    # (3) zero-weight Gaussians (gconst = -inf) keep the model on the tf32 path too
    w = model.weights.copy()
    w[model.offsets[3] + 1] = 0.0
    dz = DeviceModel(model.dim, model.offsets)
    dz.upload(w, model.means_invvars, model.inv_vars)
    gz = dz.gconsts()
    mz = ko.PackedModel(model.offsets, w, model.means_invvars, model.inv_vars, gz)
    _check(dz.loglikes_all_pdfs(feats), oracle.loglikes_all_pdfs(mz, feats)[0])


def test_tc_length_classes_and_kernel_query(oracle):
    """Every segment length 1..40 plus 240-Gaussian pdfs in one model: exercises each
    compile-time length class of the epilogue (1..16 one load, 17..32 two loads merged as two
    (max, sum) pairs), the >32 two-pass path, every kind of run boundary, and tile packing."""
    from kaldi_hmm_gmm_b200 import DeviceModel

    rng = np.random.default_rng(21)
    D = 40
    sizes = np.array(list(range(1, 41)) * 3 + [240, 1, 16, 17, 239, 2, 32, 33, 31], np.int32)
    rng.shuffle(sizes)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(offsets[-1])
    means = rng.standard_normal((G, D)).astype(np.float32) * 2
    vars_ = rng.uniform(0.5, 2, (G, D)).astype(np.float32)
    w = np.concatenate([rng.dirichlet(np.ones(s)) for s in sizes]).astype(np.float32)
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 300)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    for kernel, expect in ((AUTO, 3), (TC, 2), (TC_F16, 3), (SIMT, 1)):
        dm = _one(model, kernel)
        assert dm.dense_kernel() == expect
        _check(dm.loglikes_all_pdfs(feats), ref)
    # a pdf with 241 Gaussians does not fit an accumulator tile: it runs as two virtual pdfs (121 + 120) on the
    # tensor cores; with features too wide for the tf32 split (no device-side fall-back for that form) it is fp32
    off2 = np.array([0, 241, 244], np.int32)
    dm = DeviceModel(D, off2)
    dm.upload(np.full(244, 1 / 122, np.float32), miv[:244], iv[:244])
    assert dm.dense_kernel() == 3
    dm.set_kernel(TC)
    assert dm.dense_kernel() == 2
    Dw = 100
    rng2 = np.random.default_rng(2)
    dmw = DeviceModel(Dw, off2)
    dmw.upload(np.full(244, 1 / 122, np.float32), rng2.standard_normal((244, Dw)).astype(np.float32), np.ones((244, Dw), np.float32))
    assert dmw.dense_kernel() == 1
    with pytest.raises(RuntimeError, match="does not support"):
        dmw.set_kernel(TC)


@pytest.mark.parametrize("sizes", [
    [1] * 130,                                                       # freshly initialised monophone model: 1 Gaussian per pdf
    [1] * 40 + [2] * 17 + [3] * 11 + [5] * 7 + [8] * 5 + [4] * 9 + [7] * 3 + [6] * 4 + [9] * 3 + [20] * 2 + [1] * 3,
    [5] * 100 + [10] * 30 + [5] * 2 + [240] + [3] * 6,               # groups, singles, a full-tile pdf, group leftovers
])
def test_tc_grouped_short_segments(oracle, sizes):
    """Runs of column-adjacent pdfs with the same small Gaussian count are read 16 / len at a time
    by one TMEM load (epi_run_multi); leftovers and other lengths take the single-segment forms."""
    rng = np.random.default_rng(len(sizes))
    D = 39
    sizes = np.asarray(sizes, np.int32)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(offsets[-1])
    means = rng.standard_normal((G, D)).astype(np.float32) * 2
    vars_ = rng.uniform(0.5, 2, (G, D)).astype(np.float32)
    w = np.concatenate([rng.dirichlet(np.ones(s)) for s in sizes]).astype(np.float32)
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 333)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    for kernel in (AUTO, TC, TC_F16):
        _check(_one(model, kernel).loglikes_all_pdfs(feats), ref)


def test_tc_yesno_shape_dim80_fp16_only(oracle):
    """BASELINE configs[0]: the yesno recipe feeds 80-dim fbank (reference
    egs/yesno/local/compute_fbank_yesno.py:32) into 11 pdfs growing to ~1000 Gaussians.
    2*80+2 columns do not fit the tf32 operand tile, they do fit the fp16 one; out-of-range
    calls fall back to the fp32 SIMT kernel through the same device gate."""
    model, means, vars_ = ko.make_synthetic_model(80, 11, 1000, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 700)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    auto = _one(model, AUTO)
    assert auto.dense_kernel() == TC_F16
    _check(auto.loglikes_all_pdfs(feats), ref)
    _check(_one(model, TC_F16).loglikes_all_pdfs(feats, layout=1).T, ref)
    from kaldi_hmm_gmm_b200 import DeviceModel

    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    with pytest.raises(RuntimeError, match="tf32"):
        dm.set_kernel(TC)
        dm.loglikes_all_pdfs(feats[:10])
    big = feats[:300].copy()
    big[3] *= 500.0
    ref_big, _ = oracle.loglikes_all_pdfs(model, big)
    got = auto.loglikes_all_pdfs(big)
    assert np.isfinite(got).all()
    assert (np.abs(got - ref_big) / np.maximum(np.abs(ref_big), 1.0)).max() < 1e-4
    np.testing.assert_array_equal(got[:256], _one(model, SIMT).loglikes_all_pdfs(big)[:256])


@pytest.mark.parametrize("T", [128 * 300, 128 * 301 + 77])
def test_tc_cta_pair_multicast_path(oracle, T, monkeypatch):
    """Enough frame tiles (>= 2 per SM) switch the tensor-core kernel to CTA pairs that share the
    streamed operand by TMA multicast (even and odd tile counts: the odd one has a phantom tile).
    Same results as the plain launch (KHG_TC_CLUSTER=0) bit for bit, and as the oracle within tolerance."""
    import torch

    model, means, vars_ = ko.make_synthetic_model(40, 300, 2900, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    dfe = torch.from_numpy(feats).cuda()
    for kernel in (TC_F16, TC):
        dm = _one(model, kernel)
        pair = dm.loglikes_all_pdfs(dfe, layout=1)
        dm.sync()
        monkeypatch.setenv("KHG_TC_CLUSTER", "0")
        plain = dm.loglikes_all_pdfs(dfe, layout=1)
        dm.sync()
        monkeypatch.delenv("KHG_TC_CLUSTER")
        assert torch.equal(pair, plain)
        n = 3000
        for sl in (slice(0, n), slice(T - n, T)):
            ref, _ = oracle.loglikes_all_pdfs(model, feats[sl])
            _check(pair[:, sl].T.cpu().numpy(), ref)


@pytest.mark.parametrize("D,P,G,T", [(40, 37, 350, 3000), (39, 13, 100, 1000), (40, 420, 4000, 700), (13, 5, 17, 129),
                                     (60, 9, 200, 513), (5, 3, 3, 1), (39, 130, 1000, 40000)])
def test_gaussian_stationary_f16_kernel_vs_oracle(oracle, D, P, G, T):
    """KHG_KERNEL_TCGEN05_F16_GS (khg_loglikes_gs.cu): the model tile stays in shared memory, the pre-split
    feature operand A' is streamed; same 3xFP16 arithmetic, same tolerance.  Work splits: whole rounds of model
    tiles in lock step, leftover tiles dealt out by frame range, fewer units than SMs."""
    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    dm = _one(model, TC_F16_GS)
    assert dm.dense_kernel() == TC_F16_GS
    sl = slice(0, T) if T <= 3000 else slice(T - 1500, T)
    ref, bad = oracle.loglikes_all_pdfs(model, feats[sl])
    assert bad == 0
    _check(dm.loglikes_all_pdfs(feats)[sl], ref)
    _check(dm.loglikes_all_pdfs(feats, scale=0.1, layout=1).T[sl] * 10.0, ref)


@pytest.mark.parametrize("name,D,P,G", [("C4-lda-mllt", 40, 4200, 40000), ("C5-sat", 40, 5000, 100000)])
def test_gaussian_stationary_baseline_shapes_and_length_classes(oracle, name, D, P, G):
    """C4 (167 model tiles: one whole round + 19 leftover tiles) and C5 (two-load segments) device-resident with a
    ragged tail and more frames than one A' sub-block; bit-identical to nothing, but within tolerance of the oracle and
    of the frame-stationary kernel everywhere."""
    import os

    import torch

    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    T = 148 * 128 * 8 + 128 * 3 + 5
    feats, _ = ko.make_synthetic_frames(model, means, vars_, T)
    dfe = torch.from_numpy(feats).cuda()
    gs, fs = _one(model, TC_F16_GS), _one(model, TC_F16)
    a = gs.loglikes_all_pdfs(dfe, layout=1)
    b = fs.loglikes_all_pdfs(dfe, layout=1)
    gs.sync()
    fs.sync()
    assert bool(torch.isfinite(a).all()) and (a - b).abs().max().item() < 1e-3
    for sl in (slice(0, 200), slice(T - 200, T), slice(148 * 128 * 8 - 100, 148 * 128 * 8 + 100)):
        ref, bad = oracle.loglikes_all_pdfs(model, feats[sl], threads=os.cpu_count() or 1)
        assert bad == 0
        _check(a[:, sl].T.cpu().numpy(), ref)


def test_gaussian_stationary_rejects_wide_features_and_raises_on_nonfinite(oracle):
    from kaldi_hmm_gmm_b200 import DeviceModel

    model, means, vars_ = ko.make_synthetic_model(80, 11, 100, oracle=oracle)
    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    dm.set_kernel(TC_F16_GS)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 50)
    with pytest.raises(RuntimeError, match="Gaussian-stationary"):
        dm.loglikes_all_pdfs(feats)
    model, means, vars_ = ko.make_synthetic_model(40, 13, 100, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 300)
    feats[5, 2] = np.nan
    with pytest.raises(RuntimeError, match="Invalid answer"):
        _one(model, TC_F16_GS).loglikes_all_pdfs(feats)


def _ragged(oracle, D, sizes, seed=11):
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


@pytest.mark.parametrize("kernel", [AUTO, TC, TC_F16])
def test_pdfs_of_more_than_240_gaussians_on_the_tensor_cores(oracle, kernel):
    """DiagGmm-sized mixtures (csrc/diag-gmm.cc:241-317 builds UBMs of up to 2048 components): a pdf of more than 240
    Gaussians runs as virtual sub-pdfs whose log-sum-exps merge_virtual_kernel combines — both layouts, device and
    host buffers, next to small pdfs; the kernel that ran is asserted."""
    import torch

    model, means, vars_ = _ragged(oracle, 20, [1, 300, 2, 17, 64, 9, 1, 33, 700, 241, 240])
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 1500)
    dm = _one(model, kernel)
    assert dm.dense_kernel() == (TC if kernel == TC else TC_F16)
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    _check(dm.loglikes_all_pdfs(feats), ref)
    _check(dm.loglikes_all_pdfs(feats, scale=0.1, layout=1).T, 0.1 * ref.astype(np.float64))
    dfeats = torch.from_numpy(feats).cuda()
    _check(dm.loglikes_all_pdfs(dfeats).cpu().numpy(), ref)
    _check(dm.loglikes_all_pdfs(dfeats, layout=1).cpu().numpy().T, ref)


def test_ubm_sized_single_pdf_on_the_tensor_cores(oracle):
    model, means, vars_ = _ragged(oracle, 40, [2048], seed=5)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 700)
    dm = _one(model, AUTO)
    assert dm.dense_kernel() == TC_F16
    ref, bad = oracle.loglikes_all_pdfs(model, feats)
    assert bad == 0
    _check(dm.loglikes_all_pdfs(feats), ref)


def test_frame_major_device_output_in_bounded_scratch_blocks(oracle, monkeypatch):
    """Frame-major output of device-resident frames goes through a bounded scratch block, a chunk of frames at a time
    (KHG_DENSE_SCRATCH_FRAMES shrinks the block so that a small batch takes several)."""
    import torch

    monkeypatch.setenv("KHG_DENSE_SCRATCH_FRAMES", "256")
    model, means, vars_ = ko.make_synthetic_model(40, 37, 350, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 1111)
    dm = _one(model, AUTO)
    ref, _ = oracle.loglikes_all_pdfs(model, feats)
    got = dm.loglikes_all_pdfs(torch.from_numpy(feats).cuda())
    _check(got.cpu().numpy(), ref)
    big, mu, va = _ragged(oracle, 20, [300, 5, 260])
    f2, _ = ko.make_synthetic_frames(big, mu, va, 900)
    d2 = _one(big, AUTO)
    r2, _ = oracle.loglikes_all_pdfs(big, f2)
    _check(d2.loglikes_all_pdfs(torch.from_numpy(f2).cuda(), layout=1).cpu().numpy().T, r2)
    _check(d2.loglikes_all_pdfs(torch.from_numpy(f2).cuda()).cpu().numpy(), r2)
