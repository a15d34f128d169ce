"""Batched forced alignment at BASELINE config C5 (40-dim, 5000 pdfs / 100k Gaussians; 2000
utterances x ~500 frames): khg_align_batch = dense all-pdf likelihoods (K1) + device Viterbi.
Prints one JSON line; KHG_ALIGN_TIMING=1 adds the library's own breakdown on stderr.
  python tools/bench_align.py [--utts 2000] [--phones 48] [--pdfs 5000] [--gauss 100000] [--check 8]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=2000)
    ap.add_argument("--phones", type=int, default=48)
    ap.add_argument("--dim", type=int, default=40)
    ap.add_argument("--pdfs", type=int, default=5000)
    ap.add_argument("--gauss", type=int, default=100000)
    ap.add_argument("--beam", type=float, default=10.0)
    ap.add_argument("--retry", type=float, default=40.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", type=int, default=8, help="utterances re-aligned by the CPU oracle")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    from kaldi_hmm_gmm_b200 import DeviceModel, GraphBatch, align_batch
    from kaldi_hmm_gmm_b200 import _cabi as A
    from oracle import khg_align_oracle as ao
    from oracle import khg_oracle as ko

    rng = np.random.default_rng(20230615)
    model, means, vars_ = ko.make_synthetic_model(a.dim, a.pdfs, a.gauss)
    n_hmm_states = a.pdfs  # one pdf per HMM state
    graphs, tids_all, n_tids = [], [], 1
    for _ in range(a.utts):
        phones = [int(x) for x in rng.integers(0, n_hmm_states // 3, a.phones)]
        g, nt = ao.make_training_graph(rng, phones, alt_prob=0.2)
        graphs.append(g)
        n_tids = max(n_tids, nt)
    t2p = ao.make_tid2pdf(n_tids, a.pdfs)
    lens = []
    for g in graphs:  # walk the main path; frames are drawn in one vectorised step below
        s, tids = g.start, []
        while g.arc_offsets[s + 1] > g.arc_offsets[s]:
            arcs = range(g.arc_offsets[s], g.arc_offsets[s + 1])
            loops = [x for x in arcs if g.nextstate[x] == s and g.ilabel[x] != 0]
            fwd = [x for x in arcs if g.nextstate[x] != s and g.ilabel[x] != 0]
            if loops:
                tids += [int(g.ilabel[loops[0]])] * int(rng.integers(0, 6))
            if not fwd:
                break
            tids.append(int(g.ilabel[fwd[0]]))
            s = int(g.nextstate[fwd[0]])
        tids_all.append(np.asarray(tids, np.int32))
        lens.append(len(tids))
    tid_seq = np.concatenate(tids_all)
    pdf_seq = t2p[tid_seq]
    k = (model.offsets[pdf_seq] + (rng.random(pdf_seq.size) * (model.offsets[pdf_seq + 1] - model.offsets[pdf_seq])).astype(np.int64))
    feats = (means[k] + np.sqrt(vars_[k]) * rng.standard_normal((pdf_seq.size, a.dim))).astype(np.float32)
    T = feats.shape[0]
    gb = GraphBatch(graphs, lens)
    dm = DeviceModel(model.dim, model.offsets)
    dm.upload(model.weights, model.means_invvars, model.inv_vars)
    dfe = torch.from_numpy(feats).cuda()
    pdf_ids = torch.zeros(T, dtype=torch.int32, device="cuda")

    def run(host):
        return align_batch(dm, gb, feats if host else dfe, t2p, 1.0, a.beam, a.retry, want_paths=False, pdf_ids_out=pdf_ids)

    out = run(False)
    times = {}
    for host in (False, True):
        ts = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = run(host)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        times["host" if host else "device"] = min(ts)
    # dense kernel alone on the same frames (pdf-major block kept on the device)
    block = torch.empty((a.pdfs, (T + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dm.loglikes_all_pdfs(dfe, layout=A.KHG_PDF_MAJOR, out=block)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(a.reps):
        dm.loglikes_all_pdfs(dfe, layout=A.KHG_PDF_MAJOR, out=block)
    ev1.record()
    torch.cuda.synchronize()
    dense_s = ev0.elapsed_time(ev1) / 1e3 / a.reps
    # parity on a sample: the CPU oracle on the device's own likelihood block
    fo = gb.frame_offsets
    n_ok = 0
    t_cpu = 0.0
    for u in range(min(a.check, a.utts)):
        ll = block[:, fo[u]:fo[u + 1]].cpu().numpy()
        t0 = time.perf_counter()
        ref = ao.align_utterance(graphs[u], np.ascontiguousarray(ll), t2p, 1.0, beam=a.beam, retry_beam=a.retry, tight=False)
        t_cpu += time.perf_counter() - t0
        assert out["status"][u] == ref["status"]
        assert out["alignment"][fo[u]:fo[u + 1]].tolist() == ref["alignment"], u
        n_ok += 1
    correct = float((out["alignment"] == tid_seq).mean())
    rec = {"metric": "frames/sec (all-pdf likelihoods + batched forced alignment)", "config": {
        "workload": f"C5: D={a.dim} P={a.pdfs} G={a.gauss}, {a.utts} utterances, {T} frames, beam {a.beam}/{a.retry}",
        "states": int(gb.state_offsets[-1]), "arcs": int(gb.arc_ilabel.size)},
        "value_device_feats": T / times["device"], "e2e_host_feats": T / times["host"], "unit": "frames/s",
        "ms_call_device": times["device"] * 1e3, "ms_call_host": times["host"] * 1e3, "ms_dense_only": dense_s * 1e3,
        "dense_frames_per_s": T / dense_s, "status_counts": np.bincount(out["status"], minlength=3).tolist(),
        "oracle_checked_utts": n_ok, "frames_equal_to_generating_path": correct,
        "python_oracle_search_frames_per_s": float(sum(lens[: n_ok]) / t_cpu) if n_ok else None}
    print(json.dumps(rec))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
