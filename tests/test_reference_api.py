"""The reference's own Python unit tests for the hot path, re-expressed against this
package's pybind11 mirror (`kaldi_hmm_gmm_b200`), so that a user of kaldi_hmm_gmm finds
the same classes, methods, kwargs, dtypes and exceptions.  Sources (kaldi-hmm-gmm/python/
tests/): test_diag_gmm.py, test_am_diag_gmm.py, test_mle_diag_gmm.py,
test_mle_am_diag_gmm.py, test_gmm_update_flags.py; plus the script-level contract of
scripts/test_gmm_acc_stats_ali.py.  Everything that computes a likelihood runs on the GPU."""
import math
import pickle

import numpy as np
import pytest

khg = pytest.importorskip("kaldi_hmm_gmm_b200")
if not khg.HAVE_EXTENSION:  # pragma: no cover
    pytest.skip(f"pybind11 extension not built: {khg._EXT_ERROR}", allow_module_level=True)

from oracle import khg_oracle as ko  # noqa: E402


def _rand_gmm(rng, nmix, dim):
    g = khg.DiagGmm(nmix=nmix, dim=dim)
    w = rng.random(nmix).astype(np.float32)
    w /= w.sum()
    mean = rng.random((nmix, dim)).astype(np.float32)
    var = (rng.random((nmix, dim)) * 0.9 + 0.1).astype(np.float32)
    g.set_weights(w)
    g.set_means(mean)
    g.set_invvars(1 / var)
    return g, w, mean, var


# ------------------------------------------------------------------ host-only (CPU ok) --
def test_gmm_update_flags():  # test_gmm_update_flags.py
    assert int(khg.GmmUpdateFlags.kGmmMeans) == 1 and int(khg.GmmUpdateFlags.kGmmVariances) == 2
    assert int(khg.GmmUpdateFlags.kGmmWeights) == 4 and int(khg.GmmUpdateFlags.kGmmTransitions) == 8
    assert int(khg.GmmUpdateFlags.kGmmAll) == 15 and khg.kGmmAll == khg.GmmUpdateFlags.kGmmAll
    assert khg.str_to_gmm_flags("mvwt") == 15 and khg.str_to_gmm_flags("a") == 15 and khg.str_to_gmm_flags("mw") == 5
    assert khg.gmm_flags_to_str(15) == "mvwt" and khg.gmm_flags_to_str(6) == "vw"
    with pytest.raises(RuntimeError):
        khg.str_to_gmm_flags("x")


def test_mle_options_and_accumulator_shapes():  # test_mle_diag_gmm.py:13-90
    o = khg.MleDiagGmmOptions()
    assert abs(o.min_gaussian_weight - 1e-5) < 1e-9 and o.min_gaussian_occupancy == 10 and abs(o.min_variance - 0.001) < 1e-12
    assert o.remove_low_count_gaussians is True and "MleDiagGmmOptions(" in str(o)
    o = khg.MleDiagGmmOptions(min_gaussian_weight=1, min_gaussian_occupancy=2, min_variance=3, remove_low_count_gaussians=False)
    assert (o.min_gaussian_weight, o.min_gaussian_occupancy, o.min_variance, o.remove_low_count_gaussians) == (1, 2, 3, False)
    acc = khg.AccumDiagGmm()
    acc.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmAll)
    assert acc.flags == khg.GmmUpdateFlags.kGmmAll and acc.num_gauss == 3 and acc.dim == 5
    assert acc.occupancy.shape == (3,) and acc.mean_accumulator.shape == (3, 5) and acc.variance_accumulator.shape == (3, 5)
    assert acc.occupancy.dtype == np.float64 and acc.mean_accumulator.dtype == np.float64
    acc.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmWeights)
    assert acc.flags == khg.GmmUpdateFlags.kGmmWeights
    assert len(acc.mean_accumulator) == 0 and len(acc.variance_accumulator) == 0
    acc.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmMeans)
    assert acc.mean_accumulator.shape == (3, 5) and len(acc.variance_accumulator) == 0
    acc.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmVariances)
    assert acc.mean_accumulator.shape == (3, 5) and acc.variance_accumulator.shape == (3, 5)


def test_accumulate_for_component_scale_zero_add_stats():  # test_mle_diag_gmm.py:92-163, 254-285
    rng = np.random.default_rng(20230615)
    acc = khg.AccumDiagGmm()
    acc.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmAll)
    d = rng.random(5).astype(np.float32)
    acc.accumulate_for_component(data=d, comp_index=1, weight=0.25)
    np.testing.assert_allclose(acc.occupancy, [0, 0.25, 0])
    np.testing.assert_allclose(acc.mean_accumulator[1], d.astype(np.float64) * 0.25)
    np.testing.assert_allclose(acc.variance_accumulator[1], np.square(d).astype(np.float64) * 0.25, rtol=1e-6)
    acc.scale(f=0.1, flags=khg.GmmUpdateFlags.kGmmAll)
    np.testing.assert_allclose(acc.occupancy, [0, 0.025, 0], rtol=1e-6)
    occ = acc.occupancy  # live view (python/csrc/mle-diag-gmm.cc:75-79)
    occ[2] = 7.0
    assert acc.occupancy[2] == 7.0
    x, x2 = rng.random(5), rng.random(5)
    before = acc.mean_accumulator.copy()
    acc.add_stats_for_component(g=0, occ=0.3, x_stats=x, x2_stats=x2)
    np.testing.assert_allclose(acc.mean_accumulator[0], before[0] + x)
    assert acc.occupancy[0] == 0.3
    acc.set_zero(khg.GmmUpdateFlags.kGmmAll)
    assert acc.occupancy.sum() == 0 and np.abs(acc.mean_accumulator).sum() == 0
    with pytest.raises(RuntimeError):
        w_only = khg.AccumDiagGmm()
        w_only.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmWeights)
        w_only.set_zero(khg.GmmUpdateFlags.kGmmAll)


def test_diag_gmm_get_set_without_gpu():  # test_diag_gmm.py:16-43 minus compute_gconsts
    rng = np.random.default_rng(1)
    g, w, mean, var = _rand_gmm(rng, 10, 8)
    np.testing.assert_allclose(g.weights, w)
    np.testing.assert_allclose(g.means, mean, rtol=1e-5)
    np.testing.assert_allclose(g.vars, var, rtol=1e-5)
    assert g.num_gauss == 10 and g.dim == 8 and g.valid_gconsts is False
    with pytest.raises(RuntimeError):  # csrc/diag-gmm.h:87-90
        g.gconsts
    assert not g.means_invvars.flags.writeable and not g.inv_vars.flags.writeable


# ------------------------------------------------------------------------------ GPU --
@pytest.mark.gpu
def test_diag_gmm_gconsts_loglikes_posteriors():  # test_diag_gmm.py:45-58, 327-403, 529-576
    rng = np.random.default_rng(20230414)
    g, w, mean, var = _rand_gmm(rng, 10, 8)
    assert g.compute_gconsts() == 0 and g.valid_gconsts is True
    expected_gc = np.log(w) - 0.5 * (8 * math.log(2 * math.pi) + np.log(var).sum(1) + (mean ** 2 / var).sum(1))
    np.testing.assert_allclose(g.gconsts, expected_gc, rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(g.means_invvars, mean / var, rtol=1e-5)
    np.testing.assert_allclose(g.inv_vars, 1 / var, rtol=1e-5)
    x = rng.random(8).astype(np.float32)
    comp = np.log(w) + (-(x - mean) ** 2 / (2 * var)).sum(1) - 0.5 * np.log(2 * math.pi * var).sum(1)
    assert abs(g.log_likelihood(x) - np.log(np.exp(comp).sum())) < 1e-4
    np.testing.assert_allclose(g.log_likelihoods(x), comp, rtol=1e-4, atol=1e-4)
    X = rng.random((3, 8)).astype(np.float32)
    mat = g.log_likelihoods_matrix(X)
    assert mat.shape == (3, 10)
    for i in range(3):
        np.testing.assert_allclose(mat[i], g.log_likelihoods(X[i]), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(g.log_likelihoods_preselect(x, [3, 0, 7]), comp[[3, 0, 7]], rtol=1e-4, atol=1e-4)
    log_like, post = g.component_posteriors(x)
    sm = np.exp(comp - comp.max())
    np.testing.assert_allclose(post, sm / sm.sum(), rtol=1e-4, atol=1e-6)
    assert abs(log_like - np.log(np.exp(comp).sum())) < 1e-4
    for i in range(10):
        assert abs(g.component_log_likelihood(x, i) - comp[i]) < 1e-4
    with pytest.raises(RuntimeError, match="mismatch"):
        g.log_likelihoods(np.zeros(5, np.float32))
    g2 = khg.DiagGmm(nmix=2, dim=8)
    with pytest.raises(RuntimeError, match="ComputeGconsts"):
        g2.log_likelihood(x)
    # pickle = (weights, inv_vars, means_invvars), test_diag_gmm.py:819-848
    h = pickle.loads(pickle.dumps(g))
    np.testing.assert_array_equal(h.weights, g.weights)
    np.testing.assert_array_equal(h.inv_vars, g.inv_vars)
    np.testing.assert_allclose(h.gconsts, g.gconsts, rtol=1e-6)
    # remove_component + merge constructor
    h.remove_component(0, True)
    assert h.num_gauss == 9 and abs(h.weights.sum() - 1) < 1e-6 and h.valid_gconsts is False
    m2 = khg.DiagGmm([(0.4, g), (0.6, h)]) if h.compute_gconsts() == 0 else None
    assert m2.num_gauss == 19 and abs(m2.weights.sum() - 1) < 1e-5 and m2.valid_gconsts


@pytest.mark.gpu
def test_am_diag_gmm_container_semantics():  # test_am_diag_gmm.py:16-70
    rng = np.random.default_rng(3)
    g, w, _, _ = _rand_gmm(rng, 4, 6)
    g.compute_gconsts()
    am = khg.AmDiagGmm()
    am.add_pdf(g)
    am.add_pdf(g)
    assert am.num_pdfs == 2 and am.num_gauss == 8 and am.dim == 6 and am.num_gauss_in_pdf(1) == 4
    pdf0 = am.get_pdf(0)
    wv = pdf0.weights
    wv[0] = 0.125  # in-place mutation through the live view reaches the model (test_am_diag_gmm.py:44-47)
    assert am.get_pdf(0).weights[0] == np.float32(0.125) and g.weights[0] != np.float32(0.125)  # add_pdf deep-copies
    x = rng.random(6).astype(np.float32)
    assert abs(am.log_likelihood(1, x) - g.log_likelihood(x)) < 1e-6
    am2 = pickle.loads(pickle.dumps(am))
    assert am2.num_pdfs == 2
    np.testing.assert_array_equal(am2.get_pdf(1).inv_vars, g.inv_vars)
    block = am.log_likelihoods_all_pdfs(rng.random((5, 6)).astype(np.float32))
    assert block.shape == (5, 2)


@pytest.mark.gpu
def test_accumulate_from_posteriors_and_from_diag():  # test_mle_diag_gmm.py:165-252
    rng = np.random.default_rng(20230615)
    acc = khg.AccumDiagGmm()
    acc.resize(num_gauss=3, dim=5, flags=khg.GmmUpdateFlags.kGmmAll)
    acc.accumulate_for_component(data=rng.random(5).astype(np.float32), comp_index=1, weight=0.25)
    occ, ma, va = acc.occupancy.copy(), acc.mean_accumulator.copy(), acc.variance_accumulator.copy()
    data = rng.random(5).astype(np.float32)
    post = rng.random(3).astype(np.float32)
    acc.accumulate_from_posteriors(data=data, gauss_posteriors=post)
    np.testing.assert_allclose(acc.occupancy, occ + post)
    np.testing.assert_allclose(acc.mean_accumulator, ma + post[:, None] * data, rtol=1e-6)
    np.testing.assert_allclose(acc.variance_accumulator, va + post[:, None] * np.square(data), rtol=1e-6)
    occ, ma, va = acc.occupancy.copy(), acc.mean_accumulator.copy(), acc.variance_accumulator.copy()
    g, _, _, _ = _rand_gmm(rng, 3, 5)
    g.compute_gconsts()
    log_like = acc.accumulate_from_diag(gmm=g, data=data, weight=0.2)
    expected_ll, p = g.component_posteriors(data)
    assert abs(log_like - expected_ll) < 1e-5
    p = p * np.float32(0.2)
    np.testing.assert_allclose(acc.occupancy, occ + p, rtol=1e-5)
    np.testing.assert_allclose(acc.mean_accumulator, ma + p[:, None] * data, rtol=1e-5)
    np.testing.assert_allclose(acc.variance_accumulator, va + p[:, None] * np.square(data), rtol=1e-5)


def _am_from_packed(model):
    am = khg.AmDiagGmm()
    for p in range(model.num_pdfs):
        s = slice(model.offsets[p], model.offsets[p + 1])
        g = khg.DiagGmm(nmix=s.stop - s.start, dim=model.dim)
        g.set_weights(model.weights[s])
        g.set_invvars_and_means(model.inv_vars[s], model.means_invvars[s] / model.inv_vars[s])
        am.add_pdf(g)
    assert am.compute_gconsts() == 0
    return am


@pytest.mark.gpu
def test_accum_am_per_frame_api_and_batched_script(oracle):
    """AccumAmDiagGmm::AccumulateForGmm has no asserting test in the reference
    (test_mle_am_diag_gmm.py:29-43); pinned here through the oracle.  The per-frame loop of
    scripts/gmm_acc_stats_ali.py:46-56 and the batched script function must agree, and
    sum(transition_accs) == frames (scripts/test_gmm_acc_stats_ali.py:106)."""
    model, means, vars_ = ko.make_synthetic_model(13, 6, 20, oracle=oracle)
    T = 60
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    am = _am_from_packed(model)
    rng = np.random.default_rng(0)
    tid2pdf = np.concatenate([[0], np.repeat(np.arange(6), 2)]).astype(np.int32)  # 12 tids
    ali = (1 + 2 * pdf + rng.integers(0, 2, T)).astype(np.int32)
    ref = oracle.acc_stats_ali(model, feats, pdf)
    # (1) the unchanged per-frame binding
    accs = khg.AccumAmDiagGmm()
    accs.init(model=am, flags=khg.GmmUpdateFlags.kGmmAll)
    assert accs.num_accs == 6 and accs.dim == 13
    tot = 0.0
    for i in range(T):
        ll = accs.accumulate_for_gmm(model=am, data=feats[i], gmm_index=int(tid2pdf[ali[i]]), weight=1)
        assert abs(ll - ref["per_frame"][i]) < 1e-3
        tot += ll
    assert abs(accs.tot_log_like - ref["tot_like"]) < 1e-4 * abs(ref["tot_like"]) and accs.tot_count == T
    assert abs(accs.tot_stats_count - T) < 1e-3
    # (2) the batched script function, same signature as the reference's
    accs2 = khg.AccumAmDiagGmm()
    accs2.init(model=am, flags=khg.GmmUpdateFlags.kGmmAll)
    log_like, trans = khg.gmm_acc_stats_ali(am_gmm=am, gmm_accs=accs2, transition_model=tid2pdf, feats=feats, ali=ali.tolist())
    assert trans.dtype == np.float64 and trans.sum() == T and np.array_equal(trans, np.bincount(ali, minlength=13))
    assert abs(log_like - tot) < 1e-3 and abs(log_like - ref["tot_like"]) < 1e-4 * abs(ref["tot_like"])
    for p in range(6):
        s = slice(model.offsets[p], model.offsets[p + 1])
        a1, a2 = accs.get_acc(p), accs2.get_acc(p)
        for got in (a1, a2):
            np.testing.assert_allclose(got.occupancy, ref["occ"][s], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(got.mean_accumulator, ref["mean"][s], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(got.variance_accumulator, ref["var"][s], rtol=1e-4, atol=1e-5)
    # get_acc returns a copy (python/csrc/mle-am-diag-gmm.cc:41-42)
    c = accs.get_acc(0)
    c.occupancy[0] = -1.0
    assert accs.get_acc(0).occupancy[0] != -1.0
    # Add / Scale (csrc/mle-am-diag-gmm.cc:119-138)
    accs.add(1.0, accs2)
    assert abs(accs.tot_count - 2 * T) < 1e-6
    accs.scale(0.5)
    np.testing.assert_allclose(accs.get_acc(2).occupancy, ref["occ"][model.offsets[2]:model.offsets[3]], rtol=1e-4, atol=1e-7)
    # gmm-est consumes the stats: re-estimated parameters vs the oracle's M-step (1e-4 rel)
    opts = khg.MleDiagGmmOptions(min_gaussian_occupancy=0.5)
    khg.gmm_est(am_gmm=am, gmm_accs=accs2, gmm_opts=opts, update_flags="mvw")
    for p in range(6):
        s = slice(model.offsets[p], model.offsets[p + 1])
        upd = oracle.mle_update(model.weights[s], model.means_invvars[s], model.inv_vars[s], ref["occ"][s], ref["mean"][s],
                                ref["var"][s], update_flags=7, min_gaussian_occupancy=0.5)
        g = am.get_pdf(p)
        assert g.num_gauss == upd["weights"].size
        np.testing.assert_allclose(g.weights, upd["weights"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(g.inv_vars, upd["inv_vars"], rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(g.means, upd["means_invvars"] / upd["inv_vars"], rtol=1e-4, atol=1e-4)
    # the reference's mix-up arguments (scripts/gmm_est.py:14-18, 75-96): allocation by the per-pdf occupancies
    n0 = am.num_gauss
    accs3 = khg.AccumAmDiagGmm()
    accs3.init(model=am, flags=khg.GmmUpdateFlags.kGmmAll)
    khg.gmm_acc_stats_ali(am_gmm=am, gmm_accs=accs3, transition_model=tid2pdf, feats=feats, ali=ali.tolist())
    pdf_occs = np.asarray([accs3.get_acc(i).occupancy.sum() for i in range(6)], np.float32)
    khg.gmm_est(am_gmm=am, gmm_accs=accs3, gmm_opts=opts, mixup=n0 + 7, perturb_factor=0.01, power=0.2, min_count=2.0,
                update_flags="mvw")
    t = ko.np_get_split_targets(pdf_occs, n0 + 7, 0.2, 2.0)
    # every pdf reaches its target; pdfs already above it keep their Gaussians (csrc/am-diag-gmm.cc:80-83)
    assert n0 < t.sum() <= n0 + 7 and am.num_gauss >= t.sum()
    assert all(am.num_gauss_in_pdf(p) >= t[p] for p in range(6))


@pytest.mark.gpu
def test_decodables(oracle):
    """DecodableAmDiagGmmUnmapped/Scaled (csrc/decodable-am-diag-gmm.h:30-109): one-based
    indices, scale, NumIndices, IsLastFrame; un-asserted in the reference, pinned via the oracle."""
    model, means, vars_ = ko.make_synthetic_model(13, 6, 20, oracle=oracle)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 40)
    am = _am_from_packed(model)
    ref, _ = oracle.loglikes_all_pdfs(model, feats)
    d = khg.DecodableAmDiagGmmUnmapped(am=am, feats=feats)
    assert d.num_frames_ready() == 40 and d.num_indices() == 6
    assert d.is_last_frame(39) and not d.is_last_frame(0)
    for t, p in [(0, 0), (17, 3), (39, 5)]:
        assert abs(d.log_likelihood(t, p + 1) - ref[t, p]) < 1e-3
    with pytest.raises(RuntimeError):
        d.log_likelihood(0, 7)
    tid2pdf = np.array([0, 2, 2, 5, 0, 1], np.int32)
    s = khg.DecodableAmDiagGmmScaled(am=am, tm=tid2pdf, feats=feats, scale=0.1)
    assert s.num_indices() == 5 and isinstance(s, khg.DecodableInterface)
    assert abs(s.log_likelihood(4, 3) - 0.1 * ref[4, 5]) < 1e-4
    # a second pass over the frames (AlignUtteranceWrapper retry) reads the same block
    assert s.log_likelihood(0, 1) == s.log_likelihood(0, 2)
    with pytest.raises(RuntimeError, match="Dim mismatch"):
        khg.DecodableAmDiagGmmUnmapped(am=am, feats=np.zeros((3, 5), np.float32))

    # many utterances, one dense call (batched feed of gmm-align-compiled)
    cuts = [feats[:7], feats[7:30], feats[30:]]
    decs = khg.make_decodables(am, tid2pdf, cuts, acoustic_scale=0.1)
    assert [x.num_frames_ready() for x in decs] == [7, 23, 10] and decs[1].num_indices() == 5
    assert abs(decs[1].log_likelihood(3, 3) - 0.1 * ref[10, 5]) < 1e-4
    assert abs(decs[2].log_likelihood(9, 5) - 0.1 * ref[39, 1]) < 1e-4

    class Mine(khg.DecodableInterface):  # Python-overridable trampoline (python/csrc/decodable-itf.cc:16-41)
        def log_likelihood(self, frame, index):
            return -1.5

        def is_last_frame(self, frame):
            return frame == 0

        def num_indices(self):
            return 3

    m = Mine()
    assert m.log_likelihood(0, 1) == -1.5 and m.num_indices() == 3 and m.is_last_frame(0)


def test_training_graph_from_fst_with_a_kaldifst_shaped_object(monkeypatch):
    """TrainingGraph.from_fst walks an FST through the surface kaldifst.StdVectorFst exposes (`start`,
    `num_states`, `final(s)`, `kaldifst.ArcIterator(fst, s)` with ilabel / olabel / weight / nextstate,
    weights as objects with `.value` like kaldifst's TropicalWeight, or plain floats).  kaldifst is not
    installed here, so a stand-in module with that surface is put in its place; the reference's own use is
    scripts/gmm_align_compiled.py:36-41 / egs/yesno/train.py:176-186.  Host-only."""
    import sys
    import types
    from collections import namedtuple

    import kaldi_hmm_gmm_b200 as khg

    Arc = namedtuple("Arc", "ilabel olabel weight nextstate")
    W = namedtuple("W", "value")

    class FakeFst:
        start = 0
        num_states = 4
        _arcs = {0: [Arc(2, 7, W(0.5), 1), Arc(0, 0, W(0.25), 2)], 1: [Arc(1, 0, 0.75, 1), Arc(4, 0, W(1.5), 2)],
                 2: [Arc(3, 0, W(0.125), 2), Arc(6, 9, W(2.0), 3)], 3: []}
        _final = {3: W(0.375)}

        def final(self, s):
            return self._final.get(s, W(float("inf")))

    fake = types.ModuleType("kaldifst")
    fake.ArcIterator = lambda fst, s: iter(fst._arcs[s])
    monkeypatch.setitem(sys.modules, "kaldifst", fake)
    g = khg.TrainingGraph.from_fst(FakeFst())
    assert g.start == 0 and g.arc_offsets.tolist() == [0, 2, 4, 6, 6]
    assert g.ilabel.tolist() == [2, 0, 1, 4, 3, 6] and g.olabel.tolist() == [7, 0, 0, 0, 0, 9]
    assert g.nextstate.tolist() == [1, 2, 1, 2, 2, 3]
    np.testing.assert_array_equal(g.weight, np.asarray([0.5, 0.25, 0.75, 1.5, 0.125, 2.0], np.float32))
    assert np.isinf(g.final[:3]).all() and g.final[3] == np.float32(0.375)
    assert g.ilabel.dtype == np.int32 and g.weight.dtype == np.float32


def test_decodable_scaled_keeps_its_transition_model_object():
    """DecodableAmDiagGmmScaled.transition_model (reference python/csrc/decodable-am-diag-gmm.cc:26): the object
    handed in as `tm`.  from_block needs no device, so this runs on the CPU."""
    import kaldi_hmm_gmm_b200 as khg

    class Tm:
        id2pdf_id = np.asarray([0, 0, 1, 1], np.int32)

    tm = Tm()
    d = khg.DecodableAmDiagGmmScaled.from_block(np.zeros((2, 3), np.float32), tm, 0.5)
    assert d.transition_model is tm and d.num_indices() == 3 and d.num_frames_ready() == 3


def test_gmm_acc_stats_ali_updates_a_torch_transition_accs_in_place():
    """scripts.gmm_acc_stats_ali: a float64 CPU torch tensor handed in as transition_accs is updated in place
    (and returned); other dtypes are written back into the caller's tensor.  Uses a stand-in accumulator
    (the batched call's contract: accumulate_alignment adds one count per frame), host-only."""
    import torch

    import kaldi_hmm_gmm_b200 as khg

    class FakeAccs:
        def accumulate_alignment(self, model, transition_model, feats, ali, transition_accs):
            np.add.at(transition_accs, ali, 1.0)
            return -1.5 * len(ali)

    t2p = np.asarray([0, 0, 0, 1, 1], np.int32)
    feats = np.zeros((4, 3), np.float32)
    for dtype in (torch.float64, torch.float32):
        acc = torch.zeros(5, dtype=dtype)
        ll, out = khg.gmm_acc_stats_ali(am_gmm=None, gmm_accs=FakeAccs(), transition_model=t2p, feats=feats, ali=[1, 1, 3, 4],
                                        transition_accs=acc)
        assert out is acc and acc.tolist() == [0, 2, 0, 1, 1] and ll == -6.0
    ll, out = khg.gmm_acc_stats_ali(am_gmm=None, gmm_accs=FakeAccs(), transition_model=t2p, feats=feats, ali=[2, 2, 2, 2])
    assert out.dtype == np.float64 and out.tolist() == [0, 0, 4, 0, 0]
