"""Synthetic workloads of SURVEY.md §8(d) for bench.py and tools/: random-init models of the
named sizes, MFCC-shaped frames drawn from the model, and compiled-training-graph-shaped
utterance graphs for the batched aligner.  No datasets, no oracle imports.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Sequence, Tuple

import numpy as np


def host_model(D: int, P: int, G: int, seed: int = 20230414, size_range=None) -> dict:
    """g_p = G//P (+1 for the first G%P pdfs), or uniform in size_range = (lo, hi) — a model after mix-up has pdfs of
    very different sizes; G is then what the sizes sum to; mean ~ 3*N(0,1), var ~ U(0.5,2), weights =
    softmax(N(0,1)) within each pdf."""
    rng = np.random.default_rng(seed)
    if size_range is None:
        gp = np.full(P, G // P, np.int32)
        gp[: G % P] += 1
    else:
        gp = rng.integers(size_range[0], size_range[1] + 1, P).astype(np.int32)
        G = int(gp.sum())
    offsets = np.zeros(P + 1, np.int32)
    np.cumsum(gp, out=offsets[1:])
    means = (3.0 * rng.standard_normal((G, D))).astype(np.float32)
    vars_ = rng.uniform(0.5, 2.0, (G, D)).astype(np.float32)
    logits = rng.standard_normal(G).astype(np.float64)
    e = np.exp(logits)
    denom = np.add.reduceat(e, offsets[:-1])
    weights = (e / np.repeat(denom, gp)).astype(np.float32)
    iv = (1.0 / vars_).astype(np.float32)
    miv = (means * iv).astype(np.float32)
    return dict(offsets=offsets, gp=gp, means=means, vars=vars_, weights=weights, iv=iv, miv=miv)


def device_frames(hm: dict, T: int, seed: int, device, pdf_seq=None):
    """Each frame = a sample from a random Gaussian of a random pdf (or of pdf_seq[t]); alignment = that pdf."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    D = hm["means"].shape[1]
    P = hm["offsets"].size - 1
    means = torch.from_numpy(hm["means"]).to(device)
    std = torch.from_numpy(np.sqrt(hm["vars"])).to(device)
    offs = torch.from_numpy(hm["offsets"][:-1].astype(np.int64)).to(device)
    gp = torch.from_numpy(hm["gp"].astype(np.int64)).to(device)
    feats = torch.empty((T, D), dtype=torch.float32, device=device)
    pdf = torch.empty(T, dtype=torch.int32, device=device)
    given = None if pdf_seq is None else torch.as_tensor(np.asarray(pdf_seq, np.int64), device=device)
    step = 4_000_000
    for t0 in range(0, T, step):
        n = min(step, T - t0)
        p = torch.randint(0, P, (n,), generator=gen, device=device) if given is None else given[t0:t0 + n]
        g = offs[p] + (torch.rand(n, generator=gen, device=device) * gp[p]).long().minimum(gp[p] - 1)
        feats[t0:t0 + n] = means[g] + std[g] * torch.randn((n, D), generator=gen, device=device)
        pdf[t0:t0 + n] = p.int()
    return feats, pdf


def host_frames(hm: dict, T: int, seed: int):
    rng = np.random.default_rng(seed)
    P = hm["offsets"].size - 1
    D = hm["means"].shape[1]
    p = rng.integers(0, P, T).astype(np.int32)
    g = hm["offsets"][p] + np.minimum((rng.random(T) * hm["gp"][p]).astype(np.int32), hm["gp"][p] - 1)
    x = hm["means"][g] + np.sqrt(hm["vars"][g]) * rng.standard_normal((T, D)).astype(np.float32)
    return x.astype(np.float32), p


# ---------------------------------------------------------------- utterance graphs --
def chain_graph(rng: np.random.Generator, phones: Sequence[int], n_phone_ids: int, states_per_phone: int = 3,
                alt_prob: float = 0.2, sil_phone: int = 0, max_loops: int = 5) -> Tuple[SimpleNamespace, np.ndarray]:
    """A graph shaped like a compiled training graph after AddSelfLoops: a left-to-right chain of HMM
    states (HMM-state k = phone * states_per_phone + position; transition-ids 2k+1 = self loop, 2k+2 =
    forward), an optional alternative pronunciation in parallel (joined by an epsilon arc), an epsilon
    skip over the silence phone.  Returns (graph with arc_offsets / ilabel / olabel / weight / nextstate
    / final / start, the transition-id sequence of a random walk along the main chain)."""
    src: List[int] = []
    ilab: List[int] = []
    olab: List[int] = []
    dst: List[int] = []
    n_states = 1
    walk: List[int] = []

    def phone_chain(s: int, ph: int, word: int, record: bool) -> int:
        nonlocal n_states
        for k in range(states_per_phone):
            hs = ph * states_per_phone + k
            nxt = n_states
            n_states += 1
            src.extend((s, nxt)); ilab.extend((2 * hs + 2, 2 * hs + 1)); olab.extend((word if k == 0 else 0, 0)); dst.extend((nxt, nxt))
            if record:
                walk.append(2 * hs + 2)
                walk.extend([2 * hs + 1] * int(rng.integers(0, max_loops + 1)))
            s = nxt
        return s

    cur = 0
    for i, ph in enumerate(phones):
        end = phone_chain(cur, int(ph), i + 1, True)
        if rng.random() < alt_prob:
            alt_end = phone_chain(cur, (int(ph) + 7) % n_phone_ids, i + 1, False)
            src.append(alt_end); ilab.append(0); olab.append(0); dst.append(end)
        if ph == sil_phone:
            src.append(cur); ilab.append(0); olab.append(0); dst.append(end)
        cur = end
    src_a = np.asarray(src, np.int32)
    order = np.argsort(src_a, kind="stable")
    n_arcs = src_a.size
    is_eps = np.asarray(ilab, np.int32)[order] == 0
    weight = np.where(is_eps, rng.uniform(0.0, 0.7, n_arcs), rng.uniform(0.1, 1.5, n_arcs)).astype(np.float32)
    offs = np.zeros(n_states + 1, np.int32)
    np.cumsum(np.bincount(src_a, minlength=n_states), out=offs[1:])
    final = np.full(n_states, np.inf, np.float32)
    final[cur] = np.float32(rng.uniform(0.0, 1.0))
    g = SimpleNamespace(arc_offsets=offs, ilabel=np.asarray(ilab, np.int32)[order], olabel=np.asarray(olab, np.int32)[order],
                        weight=weight, nextstate=np.asarray(dst, np.int32)[order], final=final, start=0)
    return g, np.asarray(walk, np.int32)


def tid2pdf_table(n_tids: int, num_pdfs: int) -> np.ndarray:
    """tid -> pdf for chain_graph's transition ids (2k+1 / 2k+2 belong to HMM-state k); index 0 unused
    (reference csrc/transition-information.h:71-73)."""
    t = np.arange(n_tids, dtype=np.int64)
    out = (((t - 1) // 2) % num_pdfs).astype(np.int32)
    out[0] = 0
    return out


def alignment_workload(hm: dict, n_utts: int, phones_per_utt: int = 48, seed: int = 20230615):
    """C5-shaped batch for khg_align_batch: n_utts graphs (one pdf per HMM state, ~500 frames each) and
    the transition-id sequence every utterance's frames are drawn along.
    Returns (graphs, lens, tid2pdf, tid_seq)."""
    rng = np.random.default_rng(seed)
    P = hm["offsets"].size - 1
    n_phone_ids = max(1, P // 3)
    graphs, walks = [], []
    for _ in range(n_utts):
        g, w = chain_graph(rng, rng.integers(0, n_phone_ids, phones_per_utt), n_phone_ids)
        graphs.append(g)
        walks.append(w)
    n_tids = 2 * 3 * n_phone_ids + 3
    return graphs, [int(w.size) for w in walks], tid2pdf_table(n_tids, P), np.concatenate(walks)
