"""Mix-up on the packed device model: AmDiagGmm::SplitByCount, reference
csrc/am-diag-gmm.cc:72-89, csrc/model-common.cc:14-70, csrc/diag-gmm.cc:780-851.
The reference has no asserting test for SplitByCount; the restatement is pinned by the properties
its code guarantees (allocation invariants, weight conservation, mean preservation)."""
import numpy as np
import pytest

from oracle import khg_oracle as ko


def test_oracle_split_targets_properties():
    rng = np.random.default_rng(0)
    occs = (rng.random(50) * 5000).astype(np.float32)
    occs[[3, 17]] = 0.0                       # no counts: never split
    occs[5] = 30.0                            # min_count 20: (1 + 1) * 20 >= 30 -> stays at 1
    t = ko.np_get_split_targets(occs, 400, 0.2, 20.0)
    assert t.sum() == 400 and t.min() == 1 and t[3] == 1 and t[17] == 1 and t[5] == 1
    assert np.all(t * 20.0 < np.maximum(occs, 21.0))            # min_count rule: n * min_count < occ
    # power-law: allocations ordered like the occupancies, ratios ~ occ^power
    order = np.argsort(occs)
    assert np.all(np.diff(t[order]) >= 0)
    big = occs > 1000
    ratio = t[big] / occs[big] ** 0.2
    assert ratio.max() / ratio.min() < 1.35
    # a target that min_count makes unreachable: stops early instead of looping
    t2 = ko.np_get_split_targets(np.array([50.0, 45.0], np.float32), 100, 0.2, 20.0)
    assert t2.tolist() == [2, 2]


def test_oracle_split_preserves_mass_and_means():
    model, means, vars_ = ko.make_synthetic_model(13, 6, 18)
    occs = np.array([900, 50, 4000, 0, 2500, 700], np.float32)
    randn = np.random.default_rng(1).standard_normal((200, 13)).astype(np.float32)
    new = ko.np_split_by_count(model, occs, 40, 0.01, 0.2, 20.0, randn)
    t = ko.np_get_split_targets(occs, 40, 0.2, 20.0)
    assert t.sum() == 40 and new.num_pdfs == 6
    # pdfs already above their target keep their Gaussians (am-diag-gmm.cc:80-83)
    assert new.num_gauss == int(np.maximum(t, 3).sum()) == 43
    for p in range(6):
        a, b = slice(model.offsets[p], model.offsets[p + 1]), slice(new.offsets[p], new.offsets[p + 1])
        assert abs(new.weights[b].sum() - model.weights[a].sum()) < 1e-6            # halving conserves the mass
        # every split is symmetric: the weighted mean of means_invvars is unchanged
        np.testing.assert_allclose((new.weights[b, None] * new.means_invvars[b]).sum(0),
                                   (model.weights[a, None] * model.means_invvars[a]).sum(0), rtol=1e-4, atol=1e-4)
        assert b.stop - b.start >= a.stop - a.start
    assert new.offsets[4] - new.offsets[3] == 3  # the zero-count pdf keeps its 3 Gaussians


@pytest.mark.gpu
@pytest.mark.parametrize("P,G,target", [(6, 18, 40), (37, 150, 400), (130, 300, 1000)])
def test_gpu_split_by_count_matches_oracle(P, G, target):
    from kaldi_hmm_gmm_b200 import DeviceModel

    rng = np.random.default_rng(P)
    model, means, vars_ = ko.make_synthetic_model(13, P, G)
    occs = (rng.random(P) * 6000).astype(np.float32)
    occs[rng.integers(0, P, 2)] = 0.0
    randn = rng.standard_normal((target, 13)).astype(np.float32)
    ref = ko.np_split_by_count(model, occs, target, 0.01, 0.2, 20.0, randn)
    dm = DeviceModel(13, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    new = dm.split_by_count(occs, target, 0.01, 0.2, 20.0, randn=randn)
    got = new.download()
    np.testing.assert_array_equal(got["offsets"], ref.offsets)                        # allocation: exact
    np.testing.assert_array_equal(got["weights"], ref.weights)                        # halvings are exact
    np.testing.assert_allclose(got["inv_vars"], ref.inv_vars, rtol=0, atol=0)
    np.testing.assert_allclose(got["means_invvars"], ref.means_invvars, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(got["gconsts"], ref.gconsts, rtol=1e-5, atol=1e-4)
    assert new.num_gauss == ref.num_gauss
    # the new handle is a full model: likelihoods agree with the oracle's on the split model
    feats, _ = ko.make_synthetic_frames(model, means, vars_, 200)
    np.testing.assert_allclose(new.loglikes_all_pdfs(feats), ko.np_loglikes_all_pdfs(ref, feats), rtol=1e-4, atol=1e-3)
    # drawn inside: same structure, different perturbations, reproducible from the seed
    a = dm.split_by_count(occs, target, seed=7).download()
    b = dm.split_by_count(occs, target, seed=7).download()
    np.testing.assert_array_equal(a["offsets"], ref.offsets)
    np.testing.assert_array_equal(a["means_invvars"], b["means_invvars"])
    assert not np.array_equal(a["means_invvars"], got["means_invvars"])


@pytest.mark.gpu
def test_gpu_em_iteration_stays_on_the_device():
    """E-step -> device M-step -> device mix-up -> E-step on the mixed-up model, all through handles."""
    import torch
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats

    model, means, vars_ = ko.make_synthetic_model(13, 9, 27)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, 6000)
    dm = DeviceModel(13, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    dfe, dpdf = torch.from_numpy(feats).cuda(), torch.from_numpy(pdf).cuda()
    st = DeviceStats(dm)
    like0 = st.acc_stats_ali(dfe, dpdf)
    occ = st.download()["occ"]
    pdf_occs = np.add.reduceat(occ, model.offsets[:-1]).astype(np.float32)
    dm1, info = st.mle_update()
    dm2 = dm1.split_by_count(pdf_occs, 45, seed=3)
    t = ko.np_get_split_targets(pdf_occs, 45, 0.2, 20.0)
    n1 = np.diff(dm1.offsets)
    assert dm2.num_pdfs == 9 and dm2.num_gauss == int(np.maximum(t, n1).sum()) and dm2.num_gauss > dm1.num_gauss
    st2 = DeviceStats(dm2)
    like2 = st2.acc_stats_ali(dfe, dpdf)
    assert like2 > like0 - 1e-3 * abs(like0)      # re-estimated + mixed-up model explains its data no worse
    assert abs(st2.download()["occ"].sum() - 6000) < 1e-6 * 6000


@pytest.mark.gpu
def test_gpu_am_diag_gmm_split_by_count_method():
    """The reference's method name and kwargs (python/csrc/am-diag-gmm.cc:27-30; scripts/gmm_est.py:89-96)
    on the host mirror: pdfs are rebuilt from the device result, gconsts valid, pickle still works."""
    import pickle

    import kaldi_hmm_gmm_b200 as khg

    model, means, vars_ = ko.make_synthetic_model(13, 6, 18)
    am = khg.AmDiagGmm()
    for p in range(6):
        s = slice(model.offsets[p], model.offsets[p + 1])
        g = khg.DiagGmm(nmix=s.stop - s.start, dim=13)
        g.set_weights(model.weights[s])
        g.set_invvars_and_means(model.inv_vars[s], model.means_invvars[s] / model.inv_vars[s])
        am.add_pdf(g)
    am.compute_gconsts()
    occs = np.array([900, 50, 4000, 0, 2500, 700], np.float32)
    randn = np.random.default_rng(1).standard_normal((200, 13)).astype(np.float32)
    ref = ko.np_split_by_count(model, occs, 40, 0.01, 0.2, 20.0, randn)
    am.split_by_count(state_occs=occs, target_components=40, perturb_factor=0.01, power=0.2, min_count=20.0, randn=randn)
    assert am.num_gauss == ref.num_gauss == 43
    for p in range(6):
        s = slice(ref.offsets[p], ref.offsets[p + 1])
        g = am.get_pdf(p)
        np.testing.assert_array_equal(g.weights, ref.weights[s])
        np.testing.assert_allclose(g.means_invvars, ref.means_invvars[s], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(g.gconsts, ref.gconsts[s], rtol=1e-5, atol=1e-4)
    x = means[0].astype(np.float32)
    assert abs(am.log_likelihood(2, x) - ko.np_loglikes_all_pdfs(ref, x[None])[0, 2]) < 1e-3
    am2 = pickle.loads(pickle.dumps(am))
    assert am2.num_gauss == 43
    am.split_by_count(state_occs=occs, target_components=60, perturb_factor=0.01, power=0.2, min_count=20.0)  # internal draws
    assert am.num_gauss > 43
    # mix-down with the reference's method name (scripts/gmm_est.py:81-87)
    before = am.num_gauss
    am.merge_by_count(state_occs=occs, target_components=20, power=0.2, min_count=20.0)
    assert am.num_gauss < before and am.get_pdf(2).valid_gconsts
    assert np.isfinite(am.log_likelihood(2, x))


def _moments(w, miv, iv):
    """(total weight, mixture mean, mixture second moment) of a diagonal GMM in exponential form."""
    var = 1.0 / iv.astype(np.float64)
    mu = miv * var
    return w.sum(), (w[:, None] * mu).sum(0), (w[:, None] * (var + mu * mu)).sum(0)


def test_oracle_merge_conserves_moments():
    """Merging two Gaussians keeps the mixture's weight, mean and second moment (what
    DiagGmm::Merge's statistics-domain arithmetic guarantees); target 1 = the global Gaussian."""
    model, means, vars_ = ko.make_synthetic_model(13, 1, 12)
    w, miv, iv = model.weights, model.means_invvars, model.inv_vars
    ref = _moments(w, miv, iv)
    for tgt in (7, 3, 1):
        w2, miv2, iv2 = ko.np_diag_gmm_merge(w, miv, iv, tgt)
        assert w2.size == tgt and (iv2 > 0).all()
        got = _moments(w2, miv2, iv2)
        assert abs(got[0] - ref[0]) < 1e-5
        np.testing.assert_allclose(got[1], ref[1], rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(got[2], ref[2], rtol=1e-4, atol=1e-4)
    # merging the two closest of three well separated pairs: the far component survives untouched
    w = np.array([0.3, 0.3, 0.4], np.float32)
    mu = np.array([[0.0, 0.0], [0.1, 0.0], [9.0, 9.0]], np.float32)
    iv = np.ones((3, 2), np.float32)
    w2, miv2, iv2 = ko.np_diag_gmm_merge(w, mu * iv, iv, 2)
    assert w2.tolist() == [np.float32(0.3) + np.float32(0.3), np.float32(0.4)] and miv2[1].tolist() == [9.0, 9.0]


@pytest.mark.gpu
@pytest.mark.parametrize("P,G,target", [(6, 60, 25), (37, 400, 150), (5, 40, 5), (3, 90, 12)])
def test_gpu_merge_by_count_matches_oracle(P, G, target):
    from kaldi_hmm_gmm_b200 import DeviceModel

    rng = np.random.default_rng(P + G)
    model, means, vars_ = ko.make_synthetic_model(13, P, G)
    occs = (rng.random(P) * 6000 + 100).astype(np.float32)
    ref = ko.np_merge_by_count(model, occs, target, 0.2, 20.0)
    dm = DeviceModel(13, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    new = dm.merge_by_count(occs, target, 0.2, 20.0)
    got = new.download()
    np.testing.assert_array_equal(got["offsets"], ref.offsets)                  # allocation: exact
    assert new.num_gauss < model.num_gauss
    np.testing.assert_allclose(got["weights"], ref.weights, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(got["inv_vars"], ref.inv_vars, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(got["means_invvars"], ref.means_invvars, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["gconsts"], ref.gconsts, rtol=1e-4, atol=1e-3)
    for p in range(P):  # mixture moments of every pdf survive the merge
        a, b = slice(model.offsets[p], model.offsets[p + 1]), slice(got["offsets"][p], got["offsets"][p + 1])
        m0 = _moments(model.weights[a], model.means_invvars[a], model.inv_vars[a])
        m1 = _moments(got["weights"][b], got["means_invvars"][b], got["inv_vars"][b])
        assert abs(m0[0] - m1[0]) < 1e-5
        np.testing.assert_allclose(m1[1], m0[1], rtol=1e-3, atol=1e-3)
        np.testing.assert_allclose(m1[2], m0[2], rtol=1e-3, atol=1e-3)
