"""Gaussian selection throughput: a 2048-component, 40-dim UBM, top 50 per frame
(khg_gaussian_selection = dense kernel on the one-Gaussian-per-pdf view + gselect_kernel)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

from kaldi_hmm_gmm_b200 import DeviceModel  # noqa: E402
from oracle import khg_oracle as ko  # noqa: E402

ng, D, k, T = 2048, 40, 50, 500_000
model, means, vars_ = ko.make_synthetic_model(D, 1, ng)
rng = np.random.default_rng(1)
feats = (means[rng.integers(0, ng, T)] + np.sqrt(vars_[0]) * rng.standard_normal((T, D))).astype(np.float32)
dm = DeviceModel(D, model.offsets)
dm.upload(model.weights, model.means_invvars, model.inv_vars)
dfe = torch.from_numpy(feats).cuda()
dm.gaussian_selection(0, dfe, k)
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tot, idx, fl = dm.gaussian_selection(0, dfe, k)
    ts.append(time.perf_counter() - t0)
# CPU restatement on a sample (numpy loglikes + per-frame selection)
n = 2000
t0 = time.perf_counter()
ll = ko.np_loglikes_matrix(model.gconsts, model.means_invvars, model.inv_vars, feats[:n])
agree = 0
for t in range(n):
    agree += ko.np_gaussian_selection(ll[t], k)[1] == idx[t].tolist()
cpu = time.perf_counter() - t0
print(json.dumps({"workload": f"UBM {ng} Gaussians, dim {D}, top {k}, {T} frames (indices + per-frame log-like to the host)",
                  "frames_per_s": T / min(ts), "ms": min(ts) * 1e3, "numpy_oracle_frames_per_s": n / cpu,
                  "frames_with_identical_selection_vs_oracle": agree / n}))
