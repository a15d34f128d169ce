#!/usr/bin/env python3
"""Quick sanity run of the Gaussian-stationary dense kernel (khg_loglikes_gs.cu) against the fp32 SIMT
kernel on a few shapes; prints max |diff| per shape.  Exit code 1 on a mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

from kaldi_hmm_gmm_b200 import DeviceModel, _cabi  # noqa: E402
from kaldi_hmm_gmm_b200.synth import device_frames, host_model  # noqa: E402

bad = False
for D, P, G, T in [(40, 37, 350, 3000), (13, 5, 17, 129), (5, 3, 3, 1), (60, 9, 200, 513), (39, 130, 1000, 40000),
                   (40, 420, 4000, 700), (40, 4200, 40000, 148 * 128 * 3 + 77), (40, 5000, 100000, 148 * 128 + 5)]:
    hm = host_model(D, P, G)
    outs = []
    for k in (4, 1):
        dm = DeviceModel(D, hm["offsets"])
        dm.set_kernel(k)
        dm.upload(hm["weights"], hm["miv"], hm["iv"])
        f, _ = device_frames(hm, T, 7, torch.device("cuda"))
        o = dm.loglikes_all_pdfs(f, layout=_cabi.KHG_PDF_MAJOR)
        dm.sync()
        outs.append(o)
    d = (outs[0] - outs[1]).abs().max().item()
    fin = bool(torch.isfinite(outs[0]).all())
    print(f"D={D} P={P} G={G} T={T}: max|gs - simt| = {d:.3e} finite={fin}", flush=True)
    bad = bad or not fin or d > 1e-3
sys.exit(1 if bad else 0)
