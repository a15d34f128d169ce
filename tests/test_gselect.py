"""Gaussian selection: DiagGmm::GaussianSelection / GaussianSelectionPreselect,
reference csrc/diag-gmm.cc:202-366.

CPU: the numpy restatement (oracle.khg_oracle.np_gaussian_selection) against the closed form the
reference's own tests assert (python/tests/test_diag_gmm.py:436-527: indices == descending sort,
log-like == logsumexp of the top k within 1e-4).
GPU: khg_gaussian_selection through the C ABI and through the reference's Python method names."""
import numpy as np
import pytest

from oracle import khg_oracle as ko


def _lse(v):
    v = np.asarray(v, np.float64)
    return float(v.max() + np.log(np.exp(v - v.max()).sum()))


@pytest.mark.parametrize("seed", range(6))
def test_oracle_selection_closed_form(seed):
    rng = np.random.default_rng(seed)
    n, k = 10, 3
    ll = rng.standard_normal(n).astype(np.float32) * 5
    tot, idx = ko.np_gaussian_selection(ll, k)
    order = np.argsort(-ll, kind="stable")
    assert idx == order[:k].tolist()                               # test_diag_gmm.py:459-461
    assert abs(float(tot) - _lse(ll[order[:k]])) < 1e-4            # :462
    # k >= n keeps everything, best first
    tot_all, idx_all = ko.np_gaussian_selection(ll, 25)
    assert idx_all == order.tolist() and abs(float(tot_all) - _lse(ll)) < 1e-4
    # preselect: labels need not be sorted or unique (test_diag_gmm.py:511-527)
    pre = [0, 1, 3, 8, 7, 8, 3, 2]
    tot_p, idx_p = ko.np_gaussian_selection(ll[pre], k, labels=pre)
    s2u = np.argsort(-ll[pre], kind="stable")
    assert sorted(ll[idx_p].tolist(), reverse=True) == ll[pre][s2u[:k]].tolist()
    assert abs(float(tot_p) - _lse(ll[pre][s2u[:k]])) < 1e-4


def test_oracle_selection_ties_prefer_larger_index():
    ll = np.array([1.0, 3.0, 3.0, 0.5, 3.0], np.float32)
    _, idx = ko.np_gaussian_selection(ll, 2)
    assert idx == [4, 2]                    # std::greater on (loglike, index) pairs
    _, idx = ko.np_gaussian_selection(ll, 4)
    assert idx == [4, 2, 1, 0]


def _ubm(seed, ng, D, dup=False):
    rng = np.random.default_rng(seed)
    means = rng.standard_normal((ng, D)).astype(np.float32) * 2
    vars_ = rng.uniform(0.5, 2.0, (ng, D)).astype(np.float32)
    w = rng.dirichlet(np.ones(ng)).astype(np.float32)
    if dup:  # exact ties: identical components
        means[5], vars_[5], w[5] = means[2], vars_[2], w[2]
        means[ng - 1], vars_[ng - 1], w[ng - 1] = means[2], vars_[2], w[2]
    iv = (1 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    gc = ko.np_compute_gconsts(w, miv, iv)[0]
    model = ko.PackedModel(np.array([0, ng], np.int32), w, miv, iv, gc)
    feats = (means[rng.integers(0, ng, 300)] + rng.standard_normal((300, D)) * 1.5).astype(np.float32)
    return model, feats


def _check_selection(model, feats, k, idx, sel_ll, frame_ll, tot, labels=None):
    ref_ll = ko.np_loglikes_matrix(model.gconsts, model.means_invvars, model.inv_vars, feats)  # (T, ng)
    lab = np.arange(model.num_gauss) if labels is None else np.asarray(labels)
    cand = ref_ll[:, lab]
    kk = min(k, lab.size)
    assert idx.shape == (feats.shape[0], kk)
    # the device's own log-likes of what it selected agree with the oracle's (1e-3 / 1e-4)
    got = np.take_along_axis(ref_ll, idx, axis=1)
    assert np.abs(sel_ll - got).max() <= 1e-3
    for t in range(feats.shape[0]):
        v = sel_ll[t]
        # best first; among equal log-likes the larger index first
        assert all(v[j] > v[j + 1] or (v[j] == v[j + 1] and idx[t, j] >= idx[t, j + 1]) for j in range(kk - 1))
        assert set(idx[t]) <= set(lab.tolist())
        # nothing left out beats the k-th selected by more than the likelihood tolerance
        chosen = np.zeros(lab.size, bool)
        for i in idx[t]:
            chosen[np.flatnonzero((lab == i) & ~chosen)[0]] = True
        if (~chosen).any():
            assert cand[t][~chosen].max() <= v[-1] + 2e-3
        # the LogAdd chain over the device's values, in order
        ref_tot, _ = ko.np_gaussian_selection(v, kk)
        assert abs(frame_ll[t] - ref_tot) <= 1e-5 * max(1.0, abs(ref_tot))
        # and the reference test's closed form on the oracle's values
        assert abs(frame_ll[t] - _lse(np.sort(cand[t])[::-1][:kk])) <= 2e-3
    assert abs(tot - frame_ll.astype(np.float64).sum()) <= 1e-6 * abs(tot) + 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("ng,D,k", [(10, 8, 3), (64, 13, 5), (400, 40, 20), (2048, 40, 50), (7, 5, 30)])
def test_gpu_gaussian_selection_matches_oracle(ng, D, k):
    from kaldi_hmm_gmm_b200 import DeviceModel

    model, feats = _ubm(ng + k, ng, D)
    dm = DeviceModel(D, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    tot, idx, fl, sl = dm.gaussian_selection(0, feats, k, want_loglikes=True)
    _check_selection(model, feats, k, idx, sl, fl, tot)
    # exact agreement with the oracle's selection where the margin around the cut is clear
    ref_ll = ko.np_loglikes_matrix(model.gconsts, model.means_invvars, model.inv_vars, feats)
    n_exact = 0
    for t in range(feats.shape[0]):
        s = np.sort(ref_ll[t])[::-1]
        if np.min(np.abs(np.diff(s[: min(k, ng) + 1]))) > 5e-3:
            assert idx[t].tolist() == ko.np_gaussian_selection(ref_ll[t], k)[1]
            n_exact += 1
    assert n_exact > 0


@pytest.mark.gpu
def test_gpu_gaussian_selection_in_several_chunks(monkeypatch):
    """Frames are processed in chunks (bounded per-Gaussian block): same result as one chunk."""
    from kaldi_hmm_gmm_b200 import DeviceModel

    model, feats = _ubm(9, 64, 13)
    dm = DeviceModel(13, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    one = dm.gaussian_selection(0, feats, 7, want_loglikes=True)
    monkeypatch.setenv("KHG_GSEL_CHUNK_FRAMES", "77")
    many = dm.gaussian_selection(0, feats, 7, want_loglikes=True)
    monkeypatch.delenv("KHG_GSEL_CHUNK_FRAMES")
    assert abs(one[0] - many[0]) <= 1e-9 * abs(one[0])
    for a, b in zip(one[1:], many[1:]):
        np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
def test_gpu_gaussian_selection_ties_preselect_and_device_feats():
    import torch
    from kaldi_hmm_gmm_b200 import DeviceModel

    model, feats = _ubm(3, 40, 10, dup=True)
    dm = DeviceModel(10, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    tot, idx, fl, sl = dm.gaussian_selection(0, torch.from_numpy(feats).cuda(), 40, want_loglikes=True)
    _check_selection(model, feats, 40, idx, sl, fl, tot)
    # identical components 2, 5, 39 tie exactly: larger index first
    for t in range(feats.shape[0]):
        pos = {int(g): j for j, g in enumerate(idx[t])}
        assert pos[39] + 1 == pos[5] and pos[5] + 1 == pos[2]
    pre = [0, 1, 3, 8, 7, 8, 3, 2, 39, 5]
    tot, idx, fl, sl = dm.gaussian_selection(0, feats, 4, preselect=pre, want_loglikes=True)
    _check_selection(model, feats, 4, idx, sl, fl, tot, labels=pre)
    # a pdf inside a larger model
    big, _, _ = ko.make_synthetic_model(10, 6, 90)
    dmb = DeviceModel(10, big.offsets)
    dmb.upload(big.weights, big.means_invvars, big.inv_vars)
    s = slice(big.offsets[4], big.offsets[5])
    sub = ko.PackedModel(np.array([0, s.stop - s.start], np.int32), big.weights[s], big.means_invvars[s], big.inv_vars[s], big.gconsts[s])
    tot, idx, fl, sl = dmb.gaussian_selection(4, feats, 6, want_loglikes=True)
    _check_selection(sub, feats, 6, idx, sl, fl, tot)
    with pytest.raises(RuntimeError):
        dmb.gaussian_selection(4, feats, 0)


@pytest.mark.gpu
def test_gpu_reference_method_names():
    """python/tests/test_diag_gmm.py:436-527 re-expressed against this package."""
    import kaldi_hmm_gmm_b200 as khg

    rng = np.random.default_rng(0)
    nmix, dim = 10, 8
    g = khg.DiagGmm(nmix=nmix, dim=dim)
    w = rng.random(nmix).astype(np.float32)
    g.set_weights(w / w.sum())
    g.set_means(rng.random((nmix, dim)).astype(np.float32))
    g.set_invvars((1 / (rng.random((nmix, dim)) + 0.1)).astype(np.float32))
    g.compute_gconsts()
    x = rng.random(dim).astype(np.float32)
    log_like, indexes = g.gaussian_selection_1d(x, 3)
    ll = g.log_likelihoods(x)
    order = np.argsort(-ll, kind="stable")
    assert indexes == order[:3].tolist() and abs(log_like - _lse(ll[order[:3]])) < 1e-4
    X = rng.random((5, dim)).astype(np.float32)
    tot, lists = g.gaussian_selection_2d(X, 3)
    acc = 0.0
    for i in range(5):
        li, ind = g.gaussian_selection_1d(X[i], 3)
        assert ind == lists[i]
        acc += li
    assert abs(tot - acc) < 1e-4
    pre = [0, 1, 3, 8, 7, 8, 3, 2]
    lp, sel = g.gaussian_selection_preselect(x, preselect=pre, num_gselect=3)
    llp = g.log_likelihoods_preselect(x, pre)
    s2u = np.argsort(-llp, kind="stable")
    assert [ll[i] for i in sel] == llp[s2u[:3]].tolist()
    assert abs(lp - _lse(llp[s2u[:3]])) < 1e-4
