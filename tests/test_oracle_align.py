"""CPU tests of the forced-alignment oracle (oracle/khg_align_oracle.py), which restates
FasterDecoder + AlignUtteranceWrapper (reference csrc/faster-decoder.cc, csrc/decoder-wrappers.cc).
The reference has no asserting test for this code ("parity unpinned"), so the restatement is
pinned by an independent exact Viterbi (dynamic programming in float64, no pruning)."""
import numpy as np
import pytest

from oracle import khg_align_oracle as ao


def _case(seed, n_phones=6, P=23, T_extra=0, eps=True):
    rng = np.random.default_rng(seed)
    phones = [int(x) for x in rng.integers(0, 9, n_phones)]
    g, n_tids = ao.make_training_graph(rng, phones, optional_sil=eps, alt_prob=0.3 if eps else 0.0)
    t2p = ao.make_tid2pdf(n_tids, P)
    T = 3 * n_phones + int(rng.integers(0, 20)) + T_extra
    ll = (-8.0 * rng.random((P, T)) - 1.0).astype(np.float32)
    return g, t2p, ll


def _path_cost(g, arcs, ll, t2p, scale):
    c, t = 0.0, 0
    for a in arcs:
        c += float(g.weight[a])
        if g.ilabel[a] != 0:
            c += -float(np.float32(np.float32(scale) * ll[t2p[g.ilabel[a]], t]))
            t += 1
    return c, t


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("tight", [False, True])
def test_wide_beam_equals_exact_viterbi(seed, tight):
    g, t2p, ll = _case(seed)
    scale = 0.7
    r = ao.align_utterance(g, ll, t2p, scale, beam=1e4, tight=tight)
    best = ao.brute_force_best(g, ll, t2p, scale)
    assert r["status"] == 0 and len(r["alignment"]) == ll.shape[1]
    c, t = _path_cost(g, r["path"], ll, t2p, scale)
    assert t == ll.shape[1]
    # the returned path is a real path: consecutive arcs connect, it starts at start and ends final
    src = np.repeat(np.arange(g.num_states), np.diff(g.arc_offsets))
    assert src[r["path"][0]] == g.start and np.isfinite(g.final[g.nextstate[r["path"][-1]]])
    assert all(g.nextstate[a] == src[b] for a, b in zip(r["path"][:-1], r["path"][1:]))
    tot = c + float(g.final[g.nextstate[r["path"][-1]]])
    assert abs(tot - best) <= 1e-9 * abs(best)
    assert abs(r["like"] - (-tot / scale)) <= 2e-5 * abs(tot / scale)


def _realistic_case(seed, n_phones=10, P=23, D=13):
    """Frames drawn along a path of the graph from a synthetic model, likelihoods from the numpy
    restatement of the dense block: the regime alignment runs in (one dominant path)."""
    from oracle import khg_oracle as ko

    rng = np.random.default_rng(seed)
    model, means, vars_ = ko.make_synthetic_model(D, P, 3 * P)
    phones = [int(x) for x in rng.integers(0, 9, n_phones)]
    g, n_tids = ao.make_training_graph(rng, phones)
    t2p = ao.make_tid2pdf(n_tids, P)
    feats, _ = ao.sample_utterance(rng, g, t2p, model, means, vars_)
    ll = np.ascontiguousarray(ko.np_loglikes_all_pdfs(model, feats).T.astype(np.float32))
    return g, t2p, ll


@pytest.mark.parametrize("seed", range(12))
def test_tight_pruning_gives_the_reference_alignment(seed):
    """The device kernel's order-independent pruning (tight=True) against the reference's running
    cutoff with the recipe's beams (egs/yesno/train.py:165-167: beam 6, retry 40).  The two can
    differ only through tokens ABOVE the beam that the reference's visiting order happens to keep
    for one frame (they can push the token count over min_active); on data with a dominant path
    they give the same alignment."""
    g, t2p, ll = _realistic_case(100 + seed)
    a = ao.align_utterance(g, ll, t2p, 1.0, beam=6.0, retry_beam=40.0, tight=False)
    b = ao.align_utterance(g, ll, t2p, 1.0, beam=6.0, retry_beam=40.0, tight=True)
    assert a["status"] == b["status"]
    assert a["alignment"] == b["alignment"] and a["words"] == b["words"]


def test_retry_and_failure():
    g, t2p, ll = _case(7, n_phones=8)
    # too few frames to traverse the graph: never final, with or without retry
    short = ll[:, :5]
    assert ao.align_utterance(g, short, t2p, 1.0, beam=10.0, retry_beam=40.0)["status"] == 2
    # a beam so narrow that the first pass dies, the retry succeeds
    rng = np.random.default_rng(3)
    ll2 = (-60.0 * rng.random(ll.shape) - 1.0).astype(np.float32)
    st = [ao.align_utterance(g, ll2, t2p, 1.0, beam=0.05, retry_beam=1e4, tight=t)["status"] for t in (False, True)]
    assert st[0] == st[1] and st[0] in (0, 1)
    with pytest.raises(RuntimeError):
        ao.align_utterance(g, ll, t2p, 1.0, beam=10.0, retry_beam=5.0)
    with pytest.raises(RuntimeError):
        ao.align_utterance(g, ll, t2p, 1.0, beam=0.0)


def test_zero_frames_and_empty_graph():
    g, t2p, ll = _case(1)
    r = ao.align_utterance(g, ll[:, :0], t2p, 1.0)
    assert r["status"] == 2  # start is not final in these graphs
    g.start = -1
    assert ao.align_utterance(g, ll, t2p, 1.0)["status"] == 2


# ---------------------------------------------------------------------------------------------
# The C++ restatement of the same decoder inside libkhg_b200.so (khg_align_utterance_host: the
# exact host path of khg_align_batch) against the Python oracle with the reference's rule
# (tight=False).  Host code only: runs without a GPU.
def _host_align(g, ll, t2p, scale, beam, retry):
    from kaldi_hmm_gmm_b200 import GraphBatch, align_utterance_host

    gb = GraphBatch([g], [ll.shape[1]])
    scaled = (np.float32(scale) * ll).astype(np.float32)  # decodable-am-diag-gmm.h:94-98, fp32 product
    return align_utterance_host(gb, 0, scaled, t2p, scale, beam, retry)


@pytest.mark.parametrize("seed", range(24))
def test_host_exact_decoder_equals_reference_rule_oracle(seed):
    """Recipe beams (egs/yesno/train.py:165-167: 6 / 40, and 10 / 40) on data with a dominant path, plus
    flat random likelihoods where many tokens compete: alignment, status and best path bit-exact, like
    to float rounding — including the utterances where the reference keeps tokens above the final cutoff."""
    if seed % 2 == 0:
        g, t2p, ll = _realistic_case(300 + seed, n_phones=12)
    else:
        g, t2p, ll = _case(400 + seed, n_phones=9, T_extra=15)
    for scale, beam, retry in ((1.0, 6.0, 40.0), (0.7, 10.0, 40.0), (1.0, 2.0, 0.0), (1.0, 1e4, 0.0)):
        ref = ao.align_utterance(g, ll, t2p, scale, beam=beam, retry_beam=retry, tight=False)
        got = _host_align(g, ll, t2p, scale, beam, retry)
        assert got["status"] == ref["status"], (seed, beam)
        if ref["status"] == 2:
            assert not got["alignment"].any() and got["like"] == 0.0 and got["path"].size == 0
        else:
            assert got["alignment"].tolist() == ref["alignment"], (seed, beam)
            assert got["path"].tolist() == ref["path"], (seed, beam)
            assert abs(got["like"] - ref["like"]) <= 1e-5 * abs(ref["like"]) + 1e-4


def test_host_exact_decoder_hash_collisions_change_the_visiting_order():
    """More than 1000 states: state ids collide modulo the initial hash size (faster-decoder.cc:29), so the
    token list is no longer in insertion order (hash-list-inl.h:160-168).  Still bit-exact with the oracle's
    restatement of the container — and the container order is exercised (some bucket holds two states)."""
    rng = np.random.default_rng(77)
    phones = [int(x) for x in rng.integers(1, 9, 360)]  # ~1100+ states, no silence skips
    g, n_tids = ao.make_training_graph(rng, phones, alt_prob=0.1)
    assert g.num_states > 1100
    P = 23
    t2p = ao.make_tid2pdf(n_tids, P)
    T = 5 * len(phones)
    ll = (-6.0 * rng.random((P, T)) - 1.0).astype(np.float32)
    statuses = []
    for beam, retry in ((4.0, 60.0), (12.0, 400.0), (30.0, 0.0)):
        ref = ao.align_utterance(g, ll, t2p, 1.0, beam=beam, retry_beam=retry, tight=False)
        got = _host_align(g, ll, t2p, 1.0, beam, retry)
        assert got["status"] == ref["status"]
        statuses.append(ref["status"])
        if ref["status"] != 2:
            assert got["alignment"].tolist() == ref["alignment"] and got["path"].tolist() == ref["path"]
    assert min(statuses) < 2, statuses


def test_host_exact_decoder_errors():
    g, t2p, ll = _case(3)
    with pytest.raises(RuntimeError, match="Beams do not make sense"):
        _host_align(g, ll, t2p, 1.0, 10.0, 5.0)
    g.start = -1
    assert _host_align(g, ll, t2p, 1.0, 10.0, 40.0)["status"] == 2
