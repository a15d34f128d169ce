"""Device M-step (khg_mle_update = MleAmDiagGmmUpdate, reference csrc/mle-am-diag-gmm.cc:
153-202 / csrc/mle-diag-gmm.cc:243-390) against the oracle's per-pdf restatement.
BASELINE.json: re-estimated parameters within 1e-4 relative."""
import numpy as np
import pytest

from oracle import khg_oracle as ko

pytestmark = pytest.mark.gpu


def _setup(oracle, D, P, G, T, seed=0):
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats

    model, means, vars_ = ko.make_synthetic_model(D, P, G, oracle=oracle)
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T, seed=20230615 + seed)
    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    st = DeviceStats(dm)
    st.acc_stats_ali(feats, pdf)
    return model, feats, pdf, dm, st


def _oracle_update(oracle, model, stats, **kw):
    out = []
    tot_obj, tot_cnt = np.float32(0), np.float32(0)
    fe = fg = rg = 0
    for p in range(model.num_pdfs):
        s = slice(model.offsets[p], model.offsets[p + 1])
        u = oracle.mle_update(model.weights[s], model.means_invvars[s], model.inv_vars[s], stats["occ"][s],
                              None if stats["mean"] is None else stats["mean"][s],
                              None if stats["var"] is None else stats["var"][s], **kw)
        out.append(u)
        tot_obj = np.float32(tot_obj + np.float32(u["obj_change"]))
        tot_cnt = np.float32(tot_cnt + np.float32(u["count"]))
        fe += u["floored_elements"]
        fg += u["floored_gaussians"]
        rg += u["removed_gaussians"]
    return out, float(tot_obj), float(tot_cnt), fe, fg, rg


def _compare(new_dm, ref_pdfs):
    got = new_dm.download()
    sizes = [u["weights"].size for u in ref_pdfs]
    assert np.array_equal(np.diff(got["offsets"]), sizes)
    w = np.concatenate([u["weights"] for u in ref_pdfs])
    miv = np.concatenate([u["means_invvars"] for u in ref_pdfs])
    iv = np.concatenate([u["inv_vars"] for u in ref_pdfs])
    gc = np.concatenate([u["gconsts"] for u in ref_pdfs])
    np.testing.assert_allclose(got["weights"], w, rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(got["inv_vars"], iv, rtol=1e-4)
    # means (not means*inv_vars) are the re-estimated parameter; compare in that form
    np.testing.assert_allclose(got["means_invvars"] / got["inv_vars"], miv / iv, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(got["gconsts"], gc, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("update_flags", [15, 7, 5, 4, 6])
def test_mle_update_vs_oracle(oracle, update_flags):
    model, feats, pdf, dm, st = _setup(oracle, 13, 9, 40, 6000)
    stats = oracle.acc_stats_ali(model, feats, pdf)
    kw = dict(update_flags=update_flags & 7, min_gaussian_occupancy=3.0)
    ref, obj, cnt, fe, fg, rg = _oracle_update(oracle, model, stats, **kw)
    new_dm, info = st.mle_update(update_flags=update_flags, min_gaussian_occupancy=3.0)
    _compare(new_dm, ref)
    assert info["removed_gaussians"] == rg and info["floored_elements"] == fe and info["floored_gaussians"] == fg
    assert abs(info["count"] - cnt) <= 1e-4 * cnt
    # the objective change is a float difference of two large floats in the reference
    assert abs(info["obj_change"] - obj) <= 2e-3 * max(1.0, abs(obj)) + 0.5
    if update_flags & 1:
        assert info["obj_change"] > 0
    # the new model is usable: likelihood of the training frames went up (EM property)
    from kaldi_hmm_gmm_b200 import DeviceStats

    st2 = DeviceStats(new_dm)
    tot_new = st2.acc_stats_ali(feats, pdf)
    if update_flags == 15 or update_flags == 7:
        assert tot_new > stats["tot_like"]


def test_mle_update_removal_flooring_and_last_gaussian(oracle):
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats

    rng = np.random.default_rng(4)
    D = 6
    sizes = np.array([3, 1, 4, 2], np.int32)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    G = int(offsets[-1])
    w = np.concatenate([rng.dirichlet(np.ones(s)) for s in sizes]).astype(np.float32)
    iv = rng.uniform(0.5, 2, (G, D)).astype(np.float32)
    miv = (rng.standard_normal((G, D)) * iv).astype(np.float32)
    occ = np.array([100.0, 2.0, 50.0,   1.0,   0.5, 0.2, 0.1, 0.3,   40.0, 60.0])
    #                pdf0: middle one removed | pdf1: single gaussian kept | pdf2: all low -> last kept | pdf3 fine
    mean = occ[:, None] * rng.standard_normal((G, D))
    var = occ[:, None] * (rng.uniform(0.5, 2, (G, D)) + (mean / occ[:, None]) ** 2)
    var[8, 2] = occ[8] * (mean[8, 2] / occ[8]) ** 2  # zero variance -> floored
    dm = DeviceModel(D, offsets)
    dm.upload(w, miv, iv)
    st = DeviceStats(dm)
    st.upload(occ, mean, var, np.array([0.0, occ.sum()]))
    new_dm, info = st.mle_update()
    gc = np.concatenate([oracle.compute_gconsts(w[a:b], miv[a:b], iv[a:b])[0] for a, b in zip(offsets[:-1], offsets[1:])])
    model = ko.PackedModel(offsets, w, miv, iv, gc)
    ref, obj, cnt, fe, fg, rg = _oracle_update(oracle, model, dict(occ=occ, mean=mean, var=var))
    assert [u["weights"].size for u in ref] == [2, 1, 1, 2]
    _compare(new_dm, ref)
    assert info["removed_gaussians"] == rg == 4 and info["floored_elements"] == fe >= 1 and info["floored_gaussians"] == fg
    assert abs(info["count"] - cnt) < 1e-3
    with pytest.raises(RuntimeError, match="do not match"):
        DeviceStats(dm, 4).mle_update(update_flags=7)
