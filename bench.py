#!/usr/bin/env python3
"""bench.py — E-step throughput of the B200-native diag-GMM path (frames/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Headline workload (BASELINE.json `metric`, configs[3], "C4"): LDA+MLLT-scale model, D=40,
4200 pdfs / 40k Gaussians, 100M synthetic frames, sharded over N GPUs (strong scaling: the
total is fixed) with an NCCL all-reduce of the packed fp64 statistics.  One step = one E-step
pass over the whole batch = dense all-pdf log-likelihoods (tcgen05 split-precision kernel; the
block gmm-align-compiled consumes, kept on the device) + alignment-driven statistics
(bucketing + posteriors + fp64 stats; what gmm-acc-stats-ali produces).  Synthetic data per
SURVEY.md §8(d).

The JSON line also carries
  roofline      dominant kernel (dense log-likelihoods, tensor-bound; algorithmic flops =
                2*G*(2D+1) per frame, counted once — not x3 for the split), timed live with CUDA
                events inside the timed region
  parity_check  after the timed region: the dense block of the LAST timed chunk (the bench's own
                data, the kernel that was timed) and one chunk's statistics against the CPU oracle
                at BASELINE.json's tolerances; the run fails if it is not ok
  e2e           the same metric through the C ABI with HOST (pinned) buffers at every N: H2D of
                the step's inputs, the all-reduce and the D2H of the statistics inside the timed region
  workloads     (N=1) the other configurations, each with its own CUDA-event time, roofline and
                clock sample: W-aligned C4 (HBM-bound), the literal 3xTF32 kernel at C4, the E-step at
                C2 and C3, the dense block and the batched aligner at C5
  cpu_baseline / cpu_matrix   the oracle port on the box's host cores (bounded samples): 1 thread and
                all cores, W-aligned and W-dense (frame-blocked LogLikelihoodsMatrix form), C2..C5
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

from kaldi_hmm_gmm_b200.synth import alignment_workload, device_frames, host_frames, host_model  # noqa: E402,F401

CONFIGS = {
    # name: (D, P, G, T_total)
    "c2": (39, 130, 1000, 1_000_000),
    "c3": (39, 2000, 10_000, 10_000_000),
    "c4": (40, 4200, 40_000, 100_000_000),
    "c5": (40, 5000, 100_000, 1_000_000),  # dense block feeding the batched aligner
}
METRIC = "frames/sec (loglike+E-step stats, 1/2/4/8 B200); % tensor-core peak"
KERNELS = {"auto": 0, "simt": 1, "tcgen05": 2, "tcgen05_f16": 3}
KERNEL_NAMES = {1: "loglikes_simt_kernel (fp32 FMA)", 2: "tcgen05 3xTF32 split", 3: "tcgen05 3xFP16 split (device-gated fallback to 3xTF32)"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="override total frames (debug)")
    ap.add_argument("--kernel", default="auto", choices=sorted(KERNELS))
    ap.add_argument("--chunk", type=int, default=148 * 128 * 32, help="frames per dense block (10 GB at C4; twice that measures the same, half of it 1 % less: profiles/r3l_estep_chunk_size.txt)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-workloads", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks sampler --
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_ms=200):
        self.index = index
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.period_ms), "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm --
def oracle_model(ora, hm):
    from oracle import khg_oracle as ko

    gc = np.concatenate([ora.compute_gconsts(hm["weights"][a:b], hm["miv"][a:b], hm["iv"][a:b])[0]
                         for a, b in zip(hm["offsets"][:-1], hm["offsets"][1:])])
    return ko.PackedModel(hm["offsets"], hm["weights"], hm["miv"], hm["iv"], gc)


def _timed_rate(run, n0, target_s, n_max):
    """frames/s of run(n) on a bounded sample: grows n until one call takes ~target_s."""
    n = n0
    run(min(n0, 256))  # spins up the OpenMP team, faults the pages in
    dt = run(n)
    while dt < 0.25 * target_s and n < n_max:
        n = int(min(n_max, max(2 * n, n * 0.8 * target_s / max(dt, 1e-6))))
        dt = run(n)
    return n / dt, n, dt


def cpu_rates(ora, model, x, p, threads, what, target_s):
    """what = 'aligned' (per-frame aligned-pdf posteriors + stats: what AccumulateForGmm executes,
    csrc/mle-am-diag-gmm.cc:41-52), 'dense' (frame-blocked all-pdf LogLikelihoodsMatrix form,
    csrc/diag-gmm.cc:177-189) or 'estep' (both) on the first n frames of (x, p).
    Returns (frames/s, sample frames, seconds)."""
    cap = x.shape[0] if what == "aligned" else min(x.shape[0], 131_072)

    def run(n):
        n = min(n, cap)
        t0 = time.perf_counter()
        if what in ("dense", "estep"):
            ora.loglikes_all_pdfs(model, x[:n], pdf_major=True, threads=threads, blocked=True)
        if what in ("aligned", "estep"):
            ora.acc_stats_ali(model, x[:n], p[:n], threads=threads, want_per_frame=False)
        return time.perf_counter() - t0

    if what == "aligned" and threads > 1:
        # per-thread accumulators + their merge (AccumAmDiagGmm::Add, Kaldi's gmm-sum-accs) are a fixed
        # cost per call; a long job amortises it, so report the marginal rate between two sample sizes
        n = cap // 2
        run(1024)
        t1, t2 = run(n), run(2 * n)
        return (n / (t2 - t1) if t2 > t1 * 1.05 else 2 * n / t2), 2 * n, t2
    return _timed_rate(run, 64 * max(1, threads), target_s, cap)


def cpu_matrix(threads, target_s=0.7, configs=("c2", "c3", "c4", "c5")):
    """BASELINE.md §3's CPU matrix: 1 thread (what the reference's recipe executes per job) and all
    host cores, W-aligned and W-dense, per config — the oracle port, -O3 AVX2 build."""
    from oracle import khg_oracle as ko

    ko.build()
    ora = ko.Oracle(fast=True)
    out = {}
    for name in configs:
        D, P, G, _ = CONFIGS[name]
        hm = host_model(D, P, G)
        model = oracle_model(ora, hm)
        x, p = host_frames(hm, 1_000_000, 20230615)
        row = {}
        for what in ("aligned", "dense"):
            for th in sorted({1, threads}):
                r, n, dt = cpu_rates(ora, model, x, p, th, what, target_s)
                row[f"w_{what}_{'1thread' if th == 1 else 'allcores'}"] = {
                    "value": r, "unit": "frames/s", "cores": th,
                    "sample": f"{n} frames in {dt:.2f} s" + (" (marginal rate between n/2 and n frames: the merge of the "
                                                             "per-thread accumulators is a fixed cost per call)" if what == "aligned" and th > 1 else "")}
        out[name] = row
    return out


_CPU_SAMPLE = {}


def cpu_estep_rate(hm, threads, target_s):
    from oracle import khg_oracle as ko

    if "ora" not in _CPU_SAMPLE:
        ko.build()
        ora = ko.Oracle(fast=True)
        _CPU_SAMPLE.update(ora=ora, model=oracle_model(ora, hm), xp=host_frames(hm, 131_072, 20230615))
    x, p = _CPU_SAMPLE["xp"]
    return cpu_rates(_CPU_SAMPLE["ora"], _CPU_SAMPLE["model"], x, p, threads, "estep", target_s)


CPU_KIND_NOTE = ("oracle/khg_oracle.c -O3 AVX2 (port of the reference's algorithm; its Eigen build is unbuildable here: Eigen 3.4.0 is "
                 "network-fetched): dense all-pdf log-likelihoods in the frame-blocked LogLikelihoodsMatrix form + alignment statistics")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, P, G, T_total = CONFIGS[args.config]
    hm = host_model(D, P, G)
    threads = os.cpu_count() or 1
    rates = []
    for i in range(args.warmup + args.steps):
        r, T, dt = cpu_estep_rate(hm, threads, target_s=4.0 if i >= args.warmup else 1.0)
        if i >= args.warmup:
            rates.append((r, T, dt))
    value = float(np.mean([r for r, _, _ in rates]))
    T = rates[-1][1]
    sample = f"{T} frames/step of the {T_total}-frame workload, {threads} OpenMP threads; {CPU_KIND_NOTE}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([dt for _, _, dt in rates])),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (f64 stats)", "data": "synthetic",
        "config": {"workload": f"{args.config}: D={D} P={P} G={G} T={T_total} E-step (dense loglikes + stats)",
                   "sample_frames_per_step": T},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_cpu:
        line["cpu_matrix"] = cpu_matrix(threads, configs=(args.config,))
    emit_line(line)


# ------------------------------------------------------------------ GPU arm --
def measure_tf32_peak(torch):
    n = 8192
    a = torch.randn((n, n), device="cuda")
    b = torch.randn((n, n), device="cuda")
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    best = 0.0
    for i in range(13):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return best


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def tensor_roofline(ran, D, G, fps, peaks, tf32_peak):
    """Roofline record of the dense kernel at `fps` frames/s (algorithmic flops counted once)."""
    K = 2 * D + 1
    achieved = 2.0 * G * K * fps / 1e12
    uk = 16 if ran == 3 else 8
    phys = ((2 * D + 2 + uk - 1) // uk * uk + 2 * ((2 * D + uk - 1) // uk * uk)) / K if ran != 1 else 1.0
    if ran == 3:
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 TFLOP/s sustained bf16 (B200_PROFILING.md)"
    else:
        peak, src = tf32_peak, "TF32 dense measured in this run (torch.matmul fp32 8192^3, allow_tf32)"
    return {"kernel": KERNEL_NAMES[ran], "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "peak_source": src,
            "frac_of_tf32_peak_in_run": achieved / tf32_peak if tf32_peak else None,
            "physical_per_algorithmic": phys, "physical_frac_of_peak": achieved * phys / peak if peak else None}


def timed(torch, fn, reps, gpu_index, min_s=0.0):
    """CUDA-event time of `reps` calls of fn (after one warm call), clocks sampled meanwhile."""
    fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(gpu_index, 100).start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.perf_counter()
    e0.record()
    while n < reps or time.perf_counter() - t0 < min_s:
        fn()
        n += 1
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, sampler.stop(), n


def parity_check(torch, ora_mod, hm, dm, feats, pdf, block, a, b, D):
    """The timed kernel's own output on the bench's own data against the CPU oracle."""
    from kaldi_hmm_gmm_b200 import DeviceStats

    ora_mod.build()
    ora = ora_mod.Oracle()
    model = oracle_model(ora, hm)
    threads = os.cpu_count() or 1
    n_d = min(2048, b - a)
    x = feats[a:a + n_d].cpu().numpy()
    ref, bad = ora.loglikes_all_pdfs(model, x, threads=threads)
    got = block[:, :n_d].T.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref)
    max_abs = float(err.max())
    big = np.abs(ref) > 10
    max_rel = float((err[big] / np.abs(ref[big])).max()) if big.any() else 0.0
    # one chunk's statistics
    xs = feats[a:b].cpu().numpy()
    ps = pdf[a:b].cpu().numpy()
    stp = DeviceStats(dm)
    stp.acc_stats_ali(feats[a:b], pdf[a:b], want_total=False)
    g = stp.download()
    r = ora.acc_stats_ali(model, xs, ps, threads=min(threads, 16), want_per_frame=False)
    srel = 0.0
    for k in ("occ", "mean", "var"):
        floor = 1e-2 * np.abs(r[k]).max()
        srel = max(srel, float((np.abs(g[k] - r[k]) / np.maximum(np.abs(r[k]), floor)).max()))
    counts = np.bincount(ps, minlength=hm["offsets"].size - 1).astype(np.float64)
    occ_pdf = np.add.reduceat(g["occ"], hm["offsets"][:-1])
    occ_err = float(np.abs(occ_pdf - counts).max())
    ok = bool(bad == 0 and np.isfinite(got).all() and max_abs <= 1e-3 and max_rel <= 1e-4 and srel <= 1e-4
              and occ_err <= 1e-3 * max(1.0, counts.max()) and abs(g["tot_frames"] - (b - a)) < 0.5)
    return {"ok": ok, "max_abs": max_abs, "max_rel": max_rel, "stats_max_rel": srel, "occ_vs_bucket_counts_max_abs": occ_err,
            "dense_frames_checked": int(n_d), "stats_frames_checked": int(b - a), "tolerance": "1e-3 abs and 1e-4 rel (|ll|>10) "
            "log-likelihoods; 1e-4 rel statistics (entries below 1e-2 of the array maximum are compared against that floor)",
            "what": "dense block of the last timed chunk (kernel that was timed) and acc_stats_ali of that chunk vs oracle/khg_oracle.c"}


def run_workloads(torch, args, dev, local, peaks, tf32_peak, dm4, feats4, pdf4, block4):
    """The other configurations, N=1: each with its own CUDA-event time, roofline and clock sample."""
    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats, GraphBatch, _cabi, align_batch

    out = {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    D4, P4, G4, _ = CONFIGS["c4"]

    # -- W-aligned at C4: K2 + K3 only (what gmm-acc-stats-ali computes), HBM-bound
    n = min(feats4.shape[0], 20_000_000)
    st = DeviceStats(dm4)
    ms, ck, reps = timed(torch, lambda: st.acc_stats_ali(feats4[:n], pdf4[:n], want_total=False), 5, local, 1.0)
    fps = n / (ms * 1e-3)
    out["w_aligned_c4"] = {"value": fps, "unit": "frames/s", "ms_per_call": ms, "frames_per_call": n, "reps": reps, "clocks": ck,
                           "stats_kernel": int(dm4.stats_kernel()) if hasattr(dm4, "stats_kernel") else None,
                           "roofline": {"bound": "hbm", "achieved": fps * (4 * D4 + 4) / 1e9, "peak": hbm, "unit": "GB/s",
                                        "frac": fps * (4 * D4 + 4) / 1e9 / hbm, "algorithmic_bytes_per_frame": 4 * D4 + 4}}
    # the same call as the reference's script makes it (scripts/gmm_acc_stats_ali.py:46-56): HOST features and pdf ids
    # (pinned), H2D inside the call, the statistics read back to the host — wall clock, PCIe-bound
    ne = min(n, 8_000_000)
    hf = torch.empty((ne, D4), dtype=torch.float32, pin_memory=True)
    hp = torch.empty(ne, dtype=torch.int32, pin_memory=True)
    hf.copy_(feats4[:ne])
    hp.copy_(pdf4[:ne])
    hfn, hpn = hf.numpy(), hp.numpy()
    st.acc_stats_ali(hfn, hpn, want_total=True)  # sizes the staging buffers
    e_reps = 3
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e_reps):
        st.zero()
        st.acc_stats_ali(hfn, hpn, want_total=True)
        got = st.download()
    dt = (time.perf_counter() - t0) / e_reps
    assert abs(got["tot_frames"] - ne) < 0.5, (got["tot_frames"], ne)
    # and from PAGEABLE host memory (a numpy array): the library stages it through pinned slots with several host threads
    pf, pp = hfn.copy(), hpn.copy()
    st.acc_stats_ali(pf, pp, want_total=True)
    t0 = time.perf_counter()
    for _ in range(e_reps):
        st.zero()
        st.acc_stats_ali(pf, pp, want_total=True)
        got_p = st.download()
    dt_p = (time.perf_counter() - t0) / e_reps
    assert abs(got_p["tot_frames"] - ne) < 0.5
    out["w_aligned_c4"]["e2e_pageable_host_feats"] = {"value": ne / dt_p, "unit": "frames/s", "ms_per_call": dt_p * 1e3,
                                                      "h2d_GBps": ne * (4 * D4 + 4) / dt_p / 1e9}
    del pf, pp
    out["w_aligned_c4"]["e2e_host_feats"] = {"value": ne / dt, "unit": "frames/s", "frames_per_call": ne, "ms_per_call": dt * 1e3,
                                             "h2d_bytes_per_call": int(ne * (4 * D4 + 4)), "d2h_bytes_per_call": int(st.as_torch().numel() * 8),
                                             "h2d_GBps": ne * (4 * D4 + 4) / dt / 1e9}
    del st, hf, hp, hfn, hpn

    # -- the north star's literal 3xTF32 kernel at C4 (dense block only)
    if dm4.dense_kernel() != 2:
        hm4 = host_model(D4, P4, G4)
        dmt = DeviceModel(D4, hm4["offsets"])
        dmt.set_kernel(2)
        dmt.upload(hm4["weights"], hm4["miv"], hm4["iv"])
        nt = min(feats4.shape[0], block4.shape[1])
        ms, ck, reps = timed(torch, lambda: dmt.loglikes_all_pdfs(feats4[:nt], layout=_cabi.KHG_PDF_MAJOR, out=block4), 4, local, 1.0)
        fps = nt / (ms * 1e-3)
        out["k1_tf32_c4"] = {"value": fps, "unit": "frames/s", "ms_per_call": ms, "frames_per_call": nt, "reps": reps, "clocks": ck,
                             "roofline": tensor_roofline(2, D4, G4, fps, peaks, tf32_peak)}
        dmt.sync()
        del dmt

    # -- E-step at C2 and C3 (dense + stats, like the headline)
    for name in ("c2", "c3"):
        D, P, G, T = CONFIGS[name]
        hm = host_model(D, P, G)
        dm = DeviceModel(D, hm["offsets"])
        dm.upload(hm["weights"], hm["miv"], hm["iv"])
        f, p = device_frames(hm, T, 20230615, dev)
        chunk = min(args.chunk, T)
        blk = torch.empty((P, chunk), dtype=torch.float32, device=dev)
        st = DeviceStats(dm)

        def estep():
            st.zero()
            st.estep(f, p, blk, chunk_frames=chunk)

        ms, ck, reps = timed(torch, estep, 5, local, 1.0)
        fps = T / (ms * 1e-3)
        # (timed for 0.7 s like the step itself: a burst of a few launches runs at boost clocks, the step at the power cap)
        dms, _, _ = timed(torch, lambda: dm.loglikes_all_pdfs(f[:chunk], layout=_cabi.KHG_PDF_MAJOR, out=blk), 5, local, 0.7)
        out[f"estep_{name}"] = {"value": fps, "unit": "frames/s", "ms_per_step": ms, "frames_per_step": T, "reps": reps, "clocks": ck,
                                "dense_frames_per_s": chunk / (dms * 1e-3),
                                "roofline": tensor_roofline(dm.dense_kernel(), D, G, chunk / (dms * 1e-3), peaks, tf32_peak)}
        dm.sync()
        del st, blk, f, p, dm

    # -- C5: the dense block and the batched aligner (2000 utterances)
    out.update(align_workload(torch, dev, local, peaks, tf32_peak, 0, 1))
    return out


def align_workload(torch, dev, local, peaks, tf32_peak, rank, world, n_utts=2000):
    import torch.distributed as dist

    from kaldi_hmm_gmm_b200 import DeviceModel, GraphBatch, _cabi, align_batch
    from kaldi_hmm_gmm_b200 import parallel as par

    D, P, G, _ = CONFIGS["c5"]
    hm = host_model(D, P, G)
    dm = DeviceModel(D, hm["offsets"])
    dm.upload(hm["weights"], hm["miv"], hm["iv"])
    graphs, lens, t2p, tid_seq = alignment_workload(hm, n_utts)
    u0, u1 = par.shard_utterances(lens, world)[rank]
    fo = np.concatenate([[0], np.cumsum(lens)])
    tids = tid_seq[fo[u0]:fo[u1]]
    T, T_total = int(tids.size), int(fo[-1])
    f, _ = device_frames(hm, T, 20230615 + rank, dev, pdf_seq=t2p[tids])
    gb = GraphBatch(graphs[u0:u1], lens[u0:u1])
    pdf_out = torch.zeros(T, dtype=torch.int32, device=dev)
    res = {}

    def run(x):
        res["out"] = align_batch(dm, gb, x, t2p, 1.0, 10.0, 40.0, want_paths=False, pdf_ids_out=pdf_out)

    def wall(x, reps=4):
        run(x)
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            run(x)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        t = torch.tensor([min(ts)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local, 500).start()  # (a call is ~10 ms: NVML queries every 100 ms on every rank perturb it)
    t_dev = wall(f, reps=6)
    ck = sampler.stop()
    prep_cached = int(_cabi.lib().khg_align_last_prep_cached())
    os.environ["KHG_ALIGN_PREP_CACHE"] = "0"  # the same with the graph preparation redone by every call (a first alignment)
    t_dev_cold = wall(f, reps=3)
    del os.environ["KHG_ALIGN_PREP_CACHE"]
    hf = torch.empty((T, D), dtype=torch.float32, pin_memory=True)
    hf.copy_(f)
    t_host = wall(hf.numpy())
    exact_utts = int(_cabi.lib().khg_align_last_exact_count())
    tile_frac = float(_cabi.lib().khg_align_last_tile_fraction())
    correct = float((res["out"]["alignment"] == tids).mean())
    status = np.bincount(res["out"]["status"], minlength=3).tolist()
    out = {}
    key = "align_c5" if world == 1 else f"align_c5_sharded_n{world}"
    out[key] = {"value": T_total / t_dev, "unit": "frames/s", "ms_per_call": t_dev * 1e3, "utterances": n_utts, "frames": T_total,
                "graph_preparation_reused": prep_cached, "value_first_alignment": T_total / t_dev_cold, "ms_per_call_first_alignment": t_dev_cold * 1e3,
                "e2e_host_feats": {"value": T_total / t_host, "unit": "frames/s", "h2d_bytes_per_call": T_total * 4 * D,
                                   "d2h_bytes_per_call": T_total * 4 + n_utts * 8},
                "frames_equal_to_generating_path": correct, "exact_host_pass_utterances_rank0": exact_utts,
                "dense_tile_units_computed_rank0": tile_frac, "status_counts_rank0": status, "beam": [10.0, 40.0], "clocks": ck,
                "what": "khg_align_batch: likelihood block (K1, only the 240-Gaussian model tiles that hold a pdf of the graphs of the "
                        "frames' utterances: dense_tile_units_computed) + device Viterbi (+ exact host FasterDecoder re-run of "
                        "flagged utterances); wall clock, max over ranks, utterances sharded with no exchange; value = a REalignment of "
                        "the same graphs (their host preparation is reused, as in every realignment pass of an EM recipe), "
                        "value_first_alignment = with the preparation redone"}
    if world == 1:
        nd = min(T, 148 * 128 * 8)
        blk = torch.empty((P, nd), dtype=torch.float32, device=dev)
        ms, ck2, reps = timed(torch, lambda: dm.loglikes_all_pdfs(f[:nd], layout=_cabi.KHG_PDF_MAJOR, out=blk), 5, local, 1.0)
        fps = nd / (ms * 1e-3)
        out["w_dense_c5"] = {"value": fps, "unit": "frames/s", "ms_per_call": ms, "frames_per_call": nd, "reps": reps, "clocks": ck2,
                             "roofline": tensor_roofline(dm.dense_kernel(), D, G, fps, peaks, tf32_peak)}
    dm.sync()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats, _cabi
    from kaldi_hmm_gmm_b200 import parallel as par

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    _cabi.check(_cabi.lib().khg_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    D, P, G, T_total = CONFIGS[args.config]
    if args.frames:
        T_total = args.frames
    shard_a, shard_b = par.shard_frames(T_total, rank, world)
    T = shard_b - shard_a
    hm = host_model(D, P, G)
    if world > 1:  # model parameters are broadcast once per EM iteration
        hm.update(par.broadcast_model({k: hm[k] for k in ("weights", "miv", "iv")}, device=dev))
    dm = DeviceModel(D, hm["offsets"])
    dm.set_kernel(KERNELS[args.kernel])
    dm.upload(hm["weights"], hm["miv"], hm["iv"])
    st = DeviceStats(dm)
    stats_view = st.as_torch()
    feats, pdf = device_frames(hm, T, 20230615 + rank, dev)
    chunk = min(args.chunk, max(128, T))
    block = torch.empty((P, chunk), dtype=torch.float32, device=dev)
    n_chunks = (T + chunk - 1) // chunk
    dense_events = []

    def step(record):
        st.zero()
        for c in range(n_chunks):
            a, b = c * chunk, min(T, (c + 1) * chunk)
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            dm.loglikes_all_pdfs(feats[a:b], layout=_cabi.KHG_PDF_MAJOR, out=block)
            if record:
                e1.record()
                dense_events.append((e0, e1, b - a))
            st.acc_stats_ali(feats[a:b], pdf[a:b], want_total=False)
        if world > 1:
            par.allreduce_packed(stats_view)  # AccumAmDiagGmm::Add across ranks, NCCL over NVLink

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    dm.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = _cabi.lib().khg_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(True)
    ev1.record()
    barrier()
    dm.sync()  # raises if any log-likelihood was NaN/Inf
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    launches = _cabi.lib().khg_launch_count() - launches0
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = T_total / (ms_per_step * 1e-3)
    dense_ms = sum(e0.elapsed_time(e1) for e0, e1, _ in dense_events)
    dense_frames = sum(n for _, _, n in dense_events)
    got = st.download()
    assert abs(got["tot_frames"] - T_total) < 0.5, (got["tot_frames"], T_total)
    ran = dm.dense_kernel()  # 1 SIMT, 2 tcgen05 3xTF32, 3 tcgen05 3xFP16

    # ---- parity of the timed output (rank 0's shard; outside the timed region) ----
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import khg_oracle as ko  # the checker, never the thing measured

        a = (n_chunks - 1) * chunk
        parity = parity_check(torch, ko, hm, dm, feats, pdf, block, a, T, D)

    # ---- e2e: same metric through the C ABI with HOST buffers, every rank its shard ----
    # (all ranks take part: nothing below sits inside a per-rank try/except, so no rank can skip a collective)
    e2e = None
    if not args.no_e2e:
        hf = torch.empty((T, D), dtype=torch.float32, pin_memory=True)
        hp = torch.empty(T, dtype=torch.int32, pin_memory=True)
        hf.copy_(feats)
        hp.copy_(pdf)
        hfn, hpn = hf.numpy(), hp.numpy()
        st2 = DeviceStats(dm)
        sv2 = st2.as_torch()
        e2e_steps = max(1, min(args.steps, 3))
        n_warm = min(T, (16 << 20) + 2 * chunk)  # sizes the staging buffers (two device halves of one statistics group each)
        st2.estep(hfn[:n_warm], hpn[:n_warm], block, chunk_frames=chunk)
        barrier()
        t0 = time.perf_counter()
        res = None
        for _ in range(e2e_steps):
            st2.zero()
            st2.estep(hfn, hpn, block, chunk_frames=chunk, want_total=True)
            if world > 1:
                par.allreduce_packed(sv2)
            if rank == 0:
                res = st2.download()
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        if rank == 0:
            assert abs(res["tot_frames"] - T_total) < 0.5, (res["tot_frames"], T_total)
        e2e = {"value": T_total / dt, "unit": "frames/s", "h2d_bytes_per_step": int(T_total * (4 * D + 4)),
               "d2h_bytes_per_step": int(stats_view.numel() * 8 + 8), "steps": e2e_steps, "ms_per_step": dt * 1e3,
               "what": "every rank: khg_estep(KHG_HOST) on its shard from pinned host memory (H2D overlapped with compute), "
                       "NCCL all-reduce of the statistics, khg_stats_download on rank 0; wall clock, max over ranks; the "
                       "T x P dense block stays on the device"}
        del hf, hp, hfn, hpn, st2, sv2

    # ---- alignment sharded over the ranks (no exchange), N > 1 ----
    sharded_align = None
    if world > 1 and not args.no_workloads and args.config == "c4":
        sharded_align = align_workload(torch, dev, local, load_peaks(), 0.0, rank, world)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (dense log-likelihoods) ----
    tf32_peak = measure_tf32_peak(torch)
    peaks = load_peaks()
    dense_fps = dense_frames / (dense_ms * 1e-3)
    roofline = tensor_roofline(ran, D, G, dense_fps, peaks, tf32_peak)
    traffic, traffic_src = None, None
    try:
        rt = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = rt["dram_bytes_per_frame"] * dense_frames / max(1, len(dense_events))
        traffic_src = "from profiles/roofline_traffic.json (ncu --set full capture of this kernel), NOT measured in this run: " + rt.get("source", "")
    except Exception:
        pass
    roofline.update({
        "traffic": traffic, "traffic_source": traffic_src,
        "tf32_peak_measured_in_run": tf32_peak, "bf16_peak_measured": peaks.get("bf16_tflops"),
        "bf16_peak_sustained_measured": peaks.get("bf16_tflops_sustained"),
        "algorithmic_flops_per_frame": 2.0 * G * (2 * D + 1),
        "launches_timed": len(dense_events), "avg_launch_ms": dense_ms / max(1, len(dense_events)),
        "share_of_step": dense_ms / (ms_per_step * args.steps), "dense_frames_per_s": dense_fps,
    })

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None,
        "dtype": {1: "f32", 2: "tf32 (3-term hi/lo split, fp32 accumulate)", 3: "f16 (3-term hi/lo split, fp32 accumulate)"}[ran] + " likelihoods, f64 stats",
        "data": "synthetic",
        "config": {"workload": f"{args.config}: D={D} P={P} G={G} T={T_total} E-step = dense all-pdf loglikes + "
                               f"alignment stats (+ NCCL all-reduce of {stats_view.numel() * 8} B stats at N>1)",
                   "frames_per_gpu": T, "dense_block_frames": chunk, "kernel": args.kernel,
                   "l2": "inputs exceed L2 (features %.1f GB per GPU)" % (T * D * 4 / 1e9)},
        "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
    }
    if parity is not None:
        line["parity_check"] = parity
    if e2e is not None:
        line["e2e"] = e2e
    if sharded_align is not None:
        line["workloads"] = sharded_align

    if world == 1 and not args.no_workloads and args.config == "c4":
        try:
            line["workloads"] = run_workloads(torch, args, dev, local, peaks, tf32_peak, dm, feats, pdf, block)
        except Exception as ex:  # report, never fake
            line["workloads"] = {"error": repr(ex)[:300]}

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        try:
            r, Ts, dt = cpu_estep_rate(hm, threads, target_s=10.0)
            line["cpu_baseline"] = {"value": r, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"{Ts} frames of the same workload in {dt:.1f} s, {threads} OpenMP threads; {CPU_KIND_NOTE}"}
            line["cpu_matrix"] = cpu_matrix(threads)
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": threads, "kind": "port", "error": repr(ex)[:200]}
    emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: parity_check FAILED: %s\n" % json.dumps(parity))
        sys.exit(1)


_JSON_FD = None


def emit_line(line):
    """The ONE JSON line goes to the process's original stdout; everything libraries print to fd 1
    during the run (NCCL's "NCCL version ..." banner at N>1) was routed to stderr in __main__."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


if __name__ == "__main__":
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
