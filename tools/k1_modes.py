#!/usr/bin/env python3
"""K1 pipeline-ceiling experiments: the dense kernel alone on the C4 model, sustained for ~2 s
per setting, with the SM clock sampled during the run.
KHG_TC_DEBUG_MODE: 0 normal, 1 epilogue skips the LSE, 2 no MMAs (TMA only), 3 no TMA (MMA only).
usage: tools/k1_modes.py [kernel(0 auto,2 tf32,3 f16)] [modes, e.g. 0,1,2,3] [extra env K=V ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, _cabi  # noqa: E402

kernel = int(sys.argv[1]) if len(sys.argv) > 1 else 3
modes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3").split(",")]
for kv in sys.argv[3:]:
    k, v = kv.split("=", 1)
    os.environ[k] = v
T = 148 * 128 * 16
if os.environ.get("K1_CONFIG") in ("c5", "c3", "c2"):
    T = 148 * 128 * 8
D, P, G, _ = bench.CONFIGS[os.environ.get("K1_CONFIG", "c4")]
hm = bench.host_model(D, P, G)
dm = DeviceModel(D, hm["offsets"])
dm.set_kernel(kernel)
dm.upload(hm["weights"], hm["miv"], hm["iv"])
feats, pdf = bench.device_frames(hm, T, 1, torch.device("cuda"))
block = torch.empty((P, T), device="cuda")
for mode in modes:
    os.environ["KHG_TC_DEBUG_MODE"] = str(mode)
    for _ in range(3):
        dm.loglikes_all_pdfs(feats, layout=_cabi.KHG_PDF_MAJOR, out=block)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(0)
    sampler.start()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 2.5:
        for _ in range(4):
            dm.loglikes_all_pdfs(feats, layout=_cabi.KHG_PDF_MAJOR, out=block)
        n += 4
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ck = sampler.stop()
    ms = e0.elapsed_time(e1) / n
    print(f"kernel={kernel} mode={mode}: {ms:.3f} ms/launch  {T / ms / 1e3:.2f} M frames/s  sm_mhz={ck.get('sm_mhz')} "
          f"power_max={ck.get('power_w_max')} reasons={ck.get('reasons')}", flush=True)
os.environ["KHG_TC_DEBUG_MODE"] = "0"
