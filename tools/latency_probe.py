#!/usr/bin/env python3
"""Per-call latency of the utterance-sized entry points (what the reference's scripts call once
per utterance): khg_acc_stats_ali and khg_loglikes_all_pdfs with HOST buffers, T = 500 frames."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import numpy as np  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats  # noqa: E402

for cfg in ("c2", "c4"):
    D, P, G, _ = bench.CONFIGS[cfg]
    hm = bench.host_model(D, P, G)
    dm = DeviceModel(D, hm["offsets"])
    dm.upload(hm["weights"], hm["miv"], hm["iv"])
    st = DeviceStats(dm)
    for T in (500, 5000):
        x, p = bench.host_frames(hm, T, 1)
        for _ in range(5):
            st.acc_stats_ali(x, p)
            dm.loglikes_all_pdfs(x)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            st.acc_stats_ali(x, p)
        t1 = time.perf_counter()
        for _ in range(n):
            dm.loglikes_all_pdfs(x)
        t2 = time.perf_counter()
        print(f"{cfg} T={T}: acc_stats_ali {1e6 * (t1 - t0) / n:8.1f} us/call ({T * n / (t1 - t0) / 1e6:6.2f} M frames/s)   "
              f"loglikes_all_pdfs {1e6 * (t2 - t1) / n:8.1f} us/call ({T * n / (t2 - t1) / 1e6:6.2f} M frames/s)")
