#!/usr/bin/env python3
"""Runs a few launches of the dense log-likelihood kernel and of the stats path on the
C4 model — the command captured by ncu for profiles/ (see profiles/README.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats, _cabi  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 0
D, P, G, _ = bench.CONFIGS["c4"]
hm = bench.host_model(D, P, G)
dm = DeviceModel(D, hm["offsets"])
dm.set_kernel(kernel)
dm.upload(hm["weights"], hm["miv"], hm["iv"])
feats, pdf = bench.device_frames(hm, T, 1, torch.device("cuda"))
block = torch.empty((P, T), device="cuda")
st = DeviceStats(dm)
for _ in range(3):
    dm.loglikes_all_pdfs(feats, layout=_cabi.KHG_PDF_MAJOR, out=block)
    st.acc_stats_ali(feats, pdf, want_total=False)
dm.sync()
print("done", T)
