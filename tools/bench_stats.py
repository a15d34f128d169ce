"""W-aligned alone: khg_acc_stats_ali = K2 bucketing + K3 statistics on device-resident
frames of a BASELINE config; frames/s and the fraction of the HBM roof (4*D + 4 bytes per frame).
usage: tools/bench_stats.py [config c2|c3|c4|c5] [frames]   (KHG_B200_LIB=tools/ab/x.so for A/B)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8_000_000
D, P, G, _ = bench.CONFIGS[cfg]
# KHG_BENCH_SIZES=lo,hi: pdf sizes uniform in [lo, hi] instead of G / P each (a model after mix-up)
sizes = tuple(int(x) for x in os.environ["KHG_BENCH_SIZES"].split(",")) if os.environ.get("KHG_BENCH_SIZES") else None
hm = bench.host_model(D, P, G, size_range=sizes)
if sizes:
    G = int(hm["gp"].sum())
dm = DeviceModel(D, hm["offsets"])
dm.upload(hm["weights"], hm["miv"], hm["iv"])
feats, pdf = bench.device_frames(hm, n, 1, torch.device("cuda"))
st = DeviceStats(dm)
for _ in range(2):
    st.acc_stats_ali(feats, pdf, want_total=False)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st.acc_stats_ali(feats, pdf, want_total=False)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 1e3)
peak = 6537.3
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
gbs = n * (4 * D + 4) / best / 1e9
print(json.dumps({"workload": f"W-aligned {cfg}: D={D} P={P} G={G}, {n} frames resident in HBM", "frames_per_s": n / best,
                  "ms": best * 1e3, "algorithmic_GBps": gbs, "hbm_peak_GBps": peak, "frac_of_hbm_roof": gbs / peak,
                  "lib": os.environ.get("KHG_B200_LIB", "in-tree"), "stats_kernel": int(dm.stats_kernel()),
                  "pdf_sizes": "uniform %d..%d, %.0f %% of the Gaussians in pdfs of more than 32" % (sizes[0], sizes[1], 100.0 * hm["gp"][hm["gp"] > 32].sum() / G) if sizes else "G / P"}))

if os.environ.get("KHG_BENCH_HOST"):
    # the same call with HOST inputs: pageable numpy arrays (what a Python caller hands in) and pinned ones
    import time

    ne = min(n, 8_000_000)
    hf_pin = torch.empty((ne, D), dtype=torch.float32, pin_memory=True)
    hp_pin = torch.empty(ne, dtype=torch.int32, pin_memory=True)
    hf_pin.copy_(feats[:ne])
    hp_pin.copy_(pdf[:ne])
    hf_page, hp_page = hf_pin.numpy().copy(), hp_pin.numpy().copy()
    for name, (f, p) in (("pageable", (hf_page, hp_page)), ("pinned", (hf_pin.numpy(), hp_pin.numpy()))):
        st.acc_stats_ali(f, p, want_total=True)
        t0 = time.perf_counter()
        for _ in range(3):
            st.acc_stats_ali(f, p, want_total=True)
        dt = (time.perf_counter() - t0) / 3
        print(json.dumps({"host_inputs": name, "frames_per_s": ne / dt, "ms": dt * 1e3, "h2d_GBps": ne * (4 * D + 4) / dt / 1e9,
                          "stage_threads": os.environ.get("KHG_STAGE_THREADS", "default")}))
