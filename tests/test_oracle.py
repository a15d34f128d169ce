"""Pins the CPU oracle (oracle/) against the reference's own known answers and
closed-form unit tests, and the C and numpy restatements against each other.
CPU only.

Reference tests re-expressed here (kaldi-hmm-gmm/...):
  csrc/eigen-test.cc:460-474, 641-654            LogSumExp / Softmax known answers
  python/tests/test_diag_gmm.py:45-51            gconsts closed form
  python/tests/test_diag_gmm.py:327-403          log-likelihoods vs explicit Gaussian
  python/tests/test_diag_gmm.py:529-576          posteriors == softmax, component ll
  python/tests/test_mle_diag_gmm.py:48-285       accumulator shapes / accumulate_*
"""
import json
import os

import numpy as np
import pytest

from oracle import khg_oracle as ko


def test_known_answers_logsumexp_softmax(oracle, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "eigen_known_answers.json")))
    for case in g["logsumexp"]:
        assert abs(oracle.logsumexp(case["v"]) - case["expected"]) < case["tol"]
        assert abs(float(ko.np_logsumexp(case["v"])) - case["expected"]) < case["tol"]
    sm = g["softmax"]
    out, _ = oracle.softmax(sm["v"])
    np.testing.assert_allclose(out, sm["expected"], atol=sm["tol"])
    np.testing.assert_allclose(ko.np_softmax(sm["v"])[0], sm["expected"], atol=sm["tol"])


def _closed_form_cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, "diag_gmm_closed_form.json")))["cases"]


def test_closed_form_diag_gmm(oracle, golden_dir):
    for c in _closed_form_cases(golden_dir):
        w = np.array(c["weights"], np.float32)
        mean = np.array(c["means"], np.float32)
        var = np.array(c["vars"], np.float32)
        iv = (1.0 / var).astype(np.float32)
        miv = (mean * iv).astype(np.float32)
        gc, nbad = oracle.compute_gconsts(w, miv, iv)
        assert nbad == 0
        # test_diag_gmm.py:45-51 uses torch.allclose defaults (rtol 1e-5, atol 1e-8)
        np.testing.assert_allclose(gc, c["gconsts"], rtol=2e-5, atol=1e-5)
        np.testing.assert_allclose(ko.np_compute_gconsts(w, miv, iv)[0], gc, rtol=1e-6, atol=1e-6)
        x = np.array(c["x"], np.float32)
        mat = oracle.loglikes_matrix(gc, miv, iv, x)
        assert mat.shape == (x.shape[0], c["nmix"])  # test_diag_gmm.py:391
        np.testing.assert_allclose(mat, c["component_loglikes"], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(ko.np_loglikes_matrix(gc, miv, iv, x), mat, rtol=1e-5, atol=1e-4)
        for i in range(x.shape[0]):
            ll = oracle.log_likelihood(gc, miv, iv, x[i])
            assert abs(ll - c["loglike"][i]) < 1e-4  # test_diag_gmm.py:349
            np.testing.assert_allclose(oracle.loglikes(gc, miv, iv, x[i]), mat[i], rtol=0, atol=0)
            ll2, post = oracle.component_posteriors(gc, miv, iv, x[i])
            assert abs(ll2 - ll) < 1e-5
            np.testing.assert_allclose(post, c["posteriors"][i], rtol=1e-4, atol=1e-6)  # :551
        # test_mle_diag_gmm.py:200-252
        occ = np.zeros(c["nmix"])
        ma = np.zeros((c["nmix"], c["dim"]))
        va = np.zeros((c["nmix"], c["dim"]))
        ll = oracle.acc_from_diag(ko.kGmmAll, gc, miv, iv, x[0], c["acc_weight"], occ, ma, va)
        assert abs(ll - c["loglike"][0]) < 1e-4
        np.testing.assert_allclose(occ, c["occ"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(ma, c["mean_acc"], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(va, c["var_acc"], rtol=1e-4, atol=1e-7)


def test_flag_augmentation(oracle):
    # csrc/model-common.cc:72-84; python/tests/test_mle_diag_gmm.py:66-90
    aug = oracle.lib.khg_oracle_augment_flags
    assert aug(ko.kGmmAll & 7) == 7
    assert aug(ko.kGmmWeights) == ko.kGmmWeights
    assert aug(ko.kGmmMeans) == ko.kGmmMeans | ko.kGmmWeights
    assert aug(ko.kGmmVariances) == 7
    assert aug(0) == ko.kGmmWeights
    for f in range(8):
        assert ko.np_augment_flags(f) == aug(f)


def test_acc_for_component_and_posteriors(oracle):
    # python/tests/test_mle_diag_gmm.py:92-198
    rng = np.random.default_rng(20230615)
    ng, dim = 3, 5
    occ, ma, va = np.zeros(ng), np.zeros((ng, dim)), np.zeros((ng, dim))
    d = rng.random(dim).astype(np.float32)
    oracle.acc_for_component(ko.kGmmAll, d, 1, 0.25, occ, ma, va)
    np.testing.assert_allclose(occ, [0, 0.25, 0])
    np.testing.assert_allclose(ma[1], d.astype(np.float64) * 0.25)
    np.testing.assert_allclose(va[1], np.square(d).astype(np.float64) * 0.25, rtol=1e-6)
    occ0, ma0, va0 = occ.copy(), ma.copy(), va.copy()
    data = rng.random(dim).astype(np.float32)
    post = rng.random(ng).astype(np.float32)
    oracle.acc_from_posteriors(ko.kGmmAll, data, post, occ, ma, va)
    np.testing.assert_allclose(occ, occ0 + post)
    np.testing.assert_allclose(ma, ma0 + post[:, None] * data, rtol=1e-6)
    np.testing.assert_allclose(va, va0 + post[:, None] * np.square(data), rtol=1e-6)
    # weights-only flags: mean/var untouched (csrc/mle-diag-gmm.cc:133)
    occ2 = np.zeros(ng)
    oracle.acc_from_posteriors(ko.kGmmWeights, data, post, occ2, None, None)
    np.testing.assert_allclose(occ2, post)


def test_gconsts_zero_weight_and_nan(oracle):
    # csrc/diag-gmm.cc:132-141: zero weight -> -inf (counted bad), NaN -> error
    miv = np.ones((2, 3), np.float32)
    iv = np.ones((2, 3), np.float32)
    gc, nbad = oracle.compute_gconsts(np.array([0.0, 1.0], np.float32), miv, iv)
    assert nbad == 1 and np.isneginf(gc[0]) and np.isfinite(gc[1])
    iv_bad = iv.copy()
    iv_bad[1, 1] = -1.0
    with pytest.raises(RuntimeError):
        oracle.compute_gconsts(np.array([0.5, 0.5], np.float32), miv, iv_bad)
    with pytest.raises(RuntimeError):
        ko.np_compute_gconsts(np.array([0.5, 0.5], np.float32), miv, iv_bad)
    # an all -inf pdf makes LogSumExp NaN -> error (eigen.cc:14-18, diag-gmm.cc:160-162)
    gc2, _ = oracle.compute_gconsts(np.array([0.0, 0.0], np.float32), miv, iv)
    with pytest.raises(RuntimeError):
        oracle.log_likelihood(gc2, miv, iv, np.zeros(3, np.float32))


@pytest.mark.parametrize("D,P,G,T", [(39, 13, 100, 400), (40, 7, 33, 257), (5, 3, 3, 64)])
def test_c_vs_numpy_packed_paths(oracle, D, P, G, T):
    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    fw = np.random.default_rng(1).random(T).astype(np.float32)
    a = oracle.acc_stats_ali(model, feats, pdf, fw)
    b = ko.np_acc_stats_ali(model, feats, pdf, fw)
    assert a["bad"] == 0
    np.testing.assert_allclose(a["per_frame"], b["per_frame"], rtol=1e-5, atol=1e-4)
    for k in ("occ", "mean", "var"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-4, atol=1e-5)
    assert abs(a["tot_like"] - b["tot_like"]) < 1e-5 * abs(b["tot_like"]) + 1e-3
    assert abs(a["tot_frames"] - b["tot_frames"]) < 1e-9 * T + 1e-9
    # multi-thread sharding + Add merge == serial (csrc/mle-am-diag-gmm.cc:119-128)
    c = oracle.acc_stats_ali(model, feats, pdf, fw, threads=3)
    for k in ("occ", "mean", "var"):
        np.testing.assert_allclose(c[k], a[k], rtol=1e-12, atol=1e-12)
    # all-pdf likelihoods, both layouts
    d, bad = oracle.loglikes_all_pdfs(model, feats, scale=0.5)
    assert bad == 0
    np.testing.assert_allclose(d, ko.np_loglikes_all_pdfs(model, feats, 0.5), rtol=1e-5, atol=1e-4)
    e, _ = oracle.loglikes_all_pdfs(model, feats, scale=0.5, pdf_major=True, threads=2)
    np.testing.assert_array_equal(e.T, d)
    # aligned-pdf entry of the dense block == per-frame log-like of the stats path
    np.testing.assert_allclose(2 * d[np.arange(T), pdf], a["per_frame"], rtol=1e-6, atol=1e-5)


def test_weights_only_flags_leave_mean_var_empty(oracle):
    model, means, vars_ = ko.make_synthetic_model(6, 4, 9, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 50)
    r = oracle.acc_stats_ali(model, feats, pdf, flags=ko.kGmmWeights)
    assert r["mean"] is None and r["var"] is None and abs(r["occ"].sum() - 50) < 1e-4
    r = oracle.acc_stats_ali(model, feats, pdf, flags=ko.kGmmMeans)
    assert r["mean"] is not None and r["var"] is None


def test_mle_update_recovers_generating_model(oracle):
    # Downstream check (csrc/mle-diag-gmm.cc:243-390): ML re-estimation from the
    # stats of samples drawn from a single Gaussian recovers its mean/variance.
    rng = np.random.default_rng(7)
    D, T = 6, 20000
    mean = rng.standard_normal(D).astype(np.float32)
    var = rng.uniform(0.5, 2, D).astype(np.float32)
    x = (mean + np.sqrt(var) * rng.standard_normal((T, D))).astype(np.float32)
    w = np.array([1.0], np.float32)
    iv = np.ones((1, D), np.float32)
    miv = np.zeros((1, D), np.float32)
    gc, _ = oracle.compute_gconsts(w, miv, iv)
    model = ko.PackedModel(np.array([0, 1], np.int32), w, miv, iv, gc)
    st = oracle.acc_stats_ali(model, x, np.zeros(T, np.int32))
    upd = oracle.mle_update(w, miv, iv, st["occ"], st["mean"], st["var"])
    new_var = 1.0 / upd["inv_vars"][0]
    new_mean = upd["means_invvars"][0] * new_var
    np.testing.assert_allclose(new_mean, x.astype(np.float64).mean(0), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(new_var, x.astype(np.float64).var(0), rtol=1e-3)
    assert upd["obj_change"] > 0 and upd["removed_gaussians"] == 0 and abs(upd["count"] - T) < 1e-2


def test_mle_update_removes_low_count_gaussian(oracle):
    # csrc/mle-diag-gmm.cc:337-350, 378-382 and RemoveComponents renormalisation
    D = 3
    w = np.array([0.5, 0.3, 0.2], np.float32)
    iv = np.ones((3, D), np.float32)
    miv = np.zeros((3, D), np.float32)
    occ = np.array([100.0, 2.0, 50.0])
    mean = np.outer(occ, np.ones(D)) * 0.5
    var = np.outer(occ, np.ones(D)) * 1.25
    upd = oracle.mle_update(w, miv, iv, occ, mean, var)
    assert upd["removed_gaussians"] == 1 and upd["weights"].shape == (2,)
    np.testing.assert_allclose(upd["weights"].sum(), 1.0, rtol=1e-6)
    np.testing.assert_allclose(upd["weights"], np.array([100, 50]) / 150.0, rtol=1e-5)
    np.testing.assert_allclose(1.0 / upd["inv_vars"], 1.0, rtol=1e-5)


@pytest.mark.parametrize("fast", [False, True])
def test_blocked_matrix_form_agrees_with_per_frame_form(fast):
    """The frame-blocked LogLikelihoodsMatrix form (csrc/diag-gmm.cc:177-189; bench.py's CPU baseline)
    against the per-frame parity form (csrc/decodable-am-diag-gmm.cc:29-71): same numbers to fp32
    rounding, both layouts, ragged tail block, several threads."""
    ora = ko.Oracle(fast=fast)
    model, means, vars_ = ko.make_synthetic_model(39, 57, 500, oracle=ora)
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 64 * 3 + 17)
    a, bad_a = ora.loglikes_all_pdfs(model, feats, scale=0.5)
    for threads in (1, 3):
        b, bad_b = ora.loglikes_all_pdfs(model, feats, scale=0.5, blocked=True, threads=threads)
        assert bad_a == bad_b == 0
        assert np.abs(a - b).max() < 1e-3 and (np.abs(a - b) / np.maximum(np.abs(a), 10.0)).max() < 1e-4
        c, _ = ora.loglikes_all_pdfs(model, feats, scale=0.5, blocked=True, threads=threads, pdf_major=True)
        np.testing.assert_array_equal(c.T, b)
