#!/usr/bin/env python3
"""bench.py — E-step throughput of the B200-native diag-GMM path (frames/sec).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm

Workload (BASELINE.json `metric`, configs[3], "C4"): LDA+MLLT-scale model, D=40,
4200 pdfs / 40k Gaussians, 100M synthetic frames, sharded over N GPUs (strong
scaling: the total is fixed) with an NCCL all-reduce of the packed fp64 statistics.
One step = one E-step pass over the whole batch = dense all-pdf log-likelihoods
(tcgen05 3xTF32 kernel; the block gmm-align-compiled consumes, kept on the device)
+ alignment-driven statistics (bucketing + posteriors + fp64 stats; what
gmm-acc-stats-ali produces).  Synthetic data per SURVEY.md §8(d).

The JSON line also carries: `roofline` for the dominant kernel (dense log-likelihoods,
tensor-bound; algorithmic flops = 2*G*(2D+1) per frame, counted once — not x3 for the
3xTF32 split), timed live with CUDA events inside the timed region; `cpu_baseline`
(the oracle port on the box's host cores, bounded sample, rank 0 at N=1 only);
`e2e` (same metric through the C ABI with HOST buffers: H2D of the step's inputs and
D2H of the statistics inside the timed region).
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

CONFIGS = {
    # name: (D, P, G, T_total)
    "c2": (39, 130, 1000, 1_000_000),
    "c3": (39, 2000, 10_000, 10_000_000),
    "c4": (40, 4200, 40_000, 100_000_000),
    "c5": (40, 5000, 100_000, 1_000_000),  # dense block feeding the batched aligner (tools/bench_align.py)
}
METRIC = "frames/sec (loglike+E-step stats, 1/2/4/8 B200); % tensor-core peak"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", choices=sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="override total frames (debug)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05", "tcgen05_f16"])
    ap.add_argument("--chunk", type=int, default=148 * 128 * 16, help="frames per dense block")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic data --
def host_model(D, P, G, seed=20230414):
    """SURVEY.md §8(d) model: g_p = G//P (+1 for the first G%P pdfs); mean ~ 3*N(0,1),
    var ~ U(0.5,2), weights = softmax(N(0,1)) within each pdf."""
    rng = np.random.default_rng(seed)
    gp = np.full(P, G // P, np.int32)
    gp[: G % P] += 1
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


def device_frames(hm, T, seed, device):
    """Each frame = a sample from a random Gaussian of a random pdf; alignment = that pdf."""
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
    step = 4_000_000
    for t0 in range(0, T, step):
        n = min(step, T - t0)
        p = torch.randint(0, P, (n,), generator=gen, device=device)
        g = offs[p] + (torch.rand(n, generator=gen, device=device) * gp[p]).long().clamp_(max=int(hm["gp"].max()) - 1).minimum(gp[p] - 1)
        feats[t0:t0 + n] = means[g] + std[g] * torch.randn((n, D), generator=gen, device=device)
        pdf[t0:t0 + n] = p.int()
    return feats, pdf


def host_frames(hm, T, seed):
    rng = np.random.default_rng(seed)
    P = hm["offsets"].size - 1
    D = hm["means"].shape[1]
    p = rng.integers(0, P, T).astype(np.int32)
    g = hm["offsets"][p] + np.minimum((rng.random(T) * hm["gp"][p]).astype(np.int32), hm["gp"][p] - 1)
    x = hm["means"][g] + np.sqrt(hm["vars"][g]) * rng.standard_normal((T, D)).astype(np.float32)
    return x.astype(np.float32), p


# ------------------------------------------------------------------ clocks sampler --
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

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
def cpu_estep_rate(hm, D, P, G, threads, target_s, seed=20230615):
    """The oracle port (oracle/khg_oracle.c, -O3 AVX2 build) on host cores: dense all-pdf
    log-likelihoods + alignment statistics over a bounded sample. Returns (frames/s, sample T, secs)."""
    from oracle import khg_oracle as ko

    ko.build()
    ora = ko.Oracle(fast=True)
    gc = np.concatenate([ora.compute_gconsts(hm["weights"][a:b], hm["miv"][a:b], hm["iv"][a:b])[0]
                         for a, b in zip(hm["offsets"][:-1], hm["offsets"][1:])])
    model = ko.PackedModel(hm["offsets"], hm["weights"], hm["miv"], hm["iv"], gc)

    def run(T):
        x, p = host_frames(hm, T, seed)
        t0 = time.perf_counter()
        ora.loglikes_all_pdfs(model, x, pdf_major=True, threads=threads)
        ora.acc_stats_ali(model, x, p, threads=threads, want_per_frame=False)
        return time.perf_counter() - t0

    probe = max(64, threads * 16)
    run(probe)  # spins up the OpenMP team
    dt = run(probe)
    mid = int(max(probe, min(5_000_000, probe * 1.0 / max(dt, 1e-6))))  # ~1 s
    dt = run(mid)
    T = int(max(mid, min(5_000_000, mid * target_s / max(dt, 1e-6))))
    if T > mid * 1.5:
        dt = run(T)
    else:
        T = mid
    return T / dt, T, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    D, P, G, T_total = CONFIGS[args.config]
    hm = host_model(D, P, G)
    threads = os.cpu_count() or 1
    rates = []
    for i in range(args.warmup + args.steps):
        r, T, dt = cpu_estep_rate(hm, D, P, G, threads, target_s=6.0 if i >= args.warmup else 1.5)
        if i >= args.warmup:
            rates.append((r, T, dt))
    value = float(np.mean([r for r, _, _ in rates]))
    T = rates[-1][1]
    sample = f"{T} frames/step of the {T_total}-frame workload (dense all-pdf log-likes + stats), {threads} OpenMP threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean([dt for _, _, dt in rates])),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (f64 stats)", "data": "synthetic",
        "config": {"workload": f"{args.config}: D={D} P={P} G={G} T={T_total} E-step (dense loglikes + stats)",
                   "sample_frames_per_step": T},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference's Eigen build is unbuildable here (Eigen 3.4.0 is network-fetched); this is the oracle port of its algorithm",
    }
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
    K = 2 * D + 1
    hm = host_model(D, P, G)
    if world > 1:  # model parameters are broadcast once per EM iteration
        hm.update(par.broadcast_model({k: hm[k] for k in ("weights", "miv", "iv")}, device=dev))
    dm = DeviceModel(D, hm["offsets"])
    dm.set_kernel({"auto": 0, "simt": 1, "tcgen05": 2, "tcgen05_f16": 3}[args.kernel])
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

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (dense log-likelihoods) ----
    tf32_peak = measure_tf32_peak(torch)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    ran = dm.dense_kernel()  # 1 SIMT, 2 tcgen05 3xTF32, 3 tcgen05 3xFP16
    flops_per_frame = 2.0 * G * K
    achieved = flops_per_frame * dense_frames / (dense_ms * 1e-3) / 1e12
    # physical tensor work per logical MAC: hi.hi over 2D+2 columns + two cross products over 2D
    # columns, each rounded up to the instruction's K (8 for tf32, 16 for fp16)
    uk = 16 if ran == 3 else 8
    phys = ((2 * D + 2 + uk - 1) // uk * uk + 2 * ((2 * D + uk - 1) // uk * uk)) / K if ran != 1 else 1.0
    if ran == 3:  # kind::f16 operands: the measured bf16 dense figure is the denominator
        peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks
                    else "fallback 1400 TFLOP/s sustained bf16 (B200_PROFILING.md)")
    else:
        peak = tf32_peak
        peak_src = ("TF32 dense measured in this run (torch.matmul fp32 8192^3, allow_tf32, best of 10); "
                    "MEASURED_PEAKS.json has no TF32 entry")
    traffic = None
    try:
        # ncu-measured DRAM bytes per frame (profiles/) x frames of an average bench launch
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["dram_bytes_per_frame"] * dense_frames / max(1, len(dense_events))
    except Exception:
        pass
    roofline = {
        "kernel": {1: "loglikes_simt_kernel (fp32 FMA)", 2: "loglikes_tc_kernel<tf32> (tcgen05 3xTF32 split)",
                   3: "loglikes_tc_kernel<f16> (tcgen05 3xFP16 split, device-gated fallback to 3xTF32)"}[ran],
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
        "tf32_peak_measured_in_run": tf32_peak, "bf16_peak_measured": peaks.get("bf16_tflops"),
        "bf16_peak_sustained_measured": peaks.get("bf16_tflops_sustained"),
        "algorithmic_flops_per_frame": flops_per_frame,
        "physical_per_algorithmic": phys,
        "physical_tensor_tflops": achieved * phys,
        "physical_frac_of_peak": achieved * phys / peak if peak else None,
        "launches_timed": len(dense_events), "avg_launch_ms": dense_ms / max(1, len(dense_events)),
        "share_of_step": dense_ms / (ms_per_step * args.steps),
        "dense_frames_per_s": dense_frames / (dense_ms * 1e-3),
    }

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
        "stats_path_frames_per_s": None,
    }

    # ---- e2e: same metric through the C ABI with HOST buffers ----
    if not args.no_e2e and world == 1:
        try:
            hf = torch.empty((T, D), dtype=torch.float32, pin_memory=False)
            hf.copy_(feats)
            hp = pdf.cpu()
            hfn, hpn = hf.numpy(), hp.numpy()
            st2 = DeviceStats(dm)
            e2e_steps = max(1, min(args.steps, 2))
            st2.estep(hfn[: chunk * 2], hpn[: chunk * 2], block, chunk_frames=chunk)  # warm staging buffers
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                st2.zero()
                st2.estep(hfn, hpn, block, chunk_frames=chunk, want_total=True)
                res = st2.download()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e2e_steps
            assert abs(res["tot_frames"] - T) < 0.5
            line["e2e"] = {"value": T / dt, "unit": "frames/s", "h2d_bytes_per_step": int(T * (4 * D + 4)),
                           "d2h_bytes_per_step": int(stats_view.numel() * 8 + 8), "steps": e2e_steps,
                           "what": "khg_estep(KHG_HOST): pinned double-buffered H2D overlapped with compute, then "
                                   "khg_stats_download; the T x P dense block stays on the device"}
            del hf, hfn
        except Exception as ex:  # report, never fake
            line["e2e"] = {"value": None, "unit": "frames/s", "error": repr(ex)[:200]}
    elif world > 1:
        line["e2e"] = {"value": None, "unit": "frames/s", "note": "measured at N=1 only"}

    # ---- stats path alone (W-aligned; HBM-bound) for DESIGN.md ----
    try:
        n = min(T, 20_000_000)
        st3 = DeviceStats(dm)
        st3.acc_stats_ali(feats[:n], pdf[:n], want_total=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st3.acc_stats_ali(feats[:n], pdf[:n], want_total=False)
        e1.record()
        torch.cuda.synchronize()
        fps = n / (e0.elapsed_time(e1) * 1e-3)
        line["stats_path_frames_per_s"] = fps
        line["stats_path_hbm_frac"] = fps * (4 * D + 4) / 1e9 / peaks.get("hbm_gbs", 6650.0)
    except Exception as ex:
        line["stats_path_error"] = repr(ex)[:200]

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    if not args.no_cpu and world == 1:
        threads = os.cpu_count() or 1
        try:
            r, Ts, dt = cpu_estep_rate(hm, D, P, G, threads, target_s=12.0)
            line["cpu_baseline"] = {"value": r, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"{Ts} frames of the same workload in {dt:.1f} s, oracle/khg_oracle.c "
                                              f"-O3 AVX2, {threads} OpenMP threads"}
        except Exception as ex:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": threads, "kind": "port", "error": repr(ex)[:200]}
    emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
