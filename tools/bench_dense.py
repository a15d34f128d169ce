"""W-dense alone: khg_loglikes_all_pdfs (K1) on device-resident frames of a BASELINE-sized model, optionally with ragged
pdf sizes (KHG_BENCH_SIZES=lo,hi: uniform in [lo, hi], a model after mix-up).  frames/s and Gaussian-frames/s.
usage: tools/bench_dense.py [config c2|c3|c4|c5] [frames]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, _cabi  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 606208
D, P, G, _ = bench.CONFIGS[cfg]
sizes = tuple(int(x) for x in os.environ["KHG_BENCH_SIZES"].split(",")) if os.environ.get("KHG_BENCH_SIZES") else None
hm = bench.host_model(D, P, G, size_range=sizes)
G = int(hm["gp"].sum())
dm = DeviceModel(D, hm["offsets"])
dm.upload(hm["weights"], hm["miv"], hm["iv"])
feats, _ = bench.device_frames(hm, n, 1, torch.device("cuda"))
block = torch.empty((P, n), device="cuda")
for _ in range(3):
    dm.loglikes_all_pdfs(feats, layout=_cabi.KHG_PDF_MAJOR, out=block)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 40
e0.record()
for _ in range(reps):
    dm.loglikes_all_pdfs(feats, layout=_cabi.KHG_PDF_MAJOR, out=block)
e1.record()
torch.cuda.synchronize()
dt = e0.elapsed_time(e1) / 1e3 / reps
print(json.dumps({"workload": f"W-dense {cfg}: D={D} P={P} G={G}, {n} frames per call", "pdf_sizes": "G / P" if not sizes else f"uniform {sizes[0]}..{sizes[1]}",
                  "frames_per_s": n / dt, "gaussian_frames_per_s": n * G / dt, "algorithmic_TFLOPs": n * G * (4 * D + 2) / dt / 1e12,
                  "dense_kernel": int(dm.dense_kernel()), "ms": dt * 1e3}))
