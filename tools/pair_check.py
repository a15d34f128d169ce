"""Compares the CTA-pair forms of K1 (KHG_TC_CLUSTER=2 multicast, =3 cta_group::2 MMAs) with the plain launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, _cabi  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 128 * 301 + 77
D, P, G, _ = bench.CONFIGS[os.environ.get("K1_CONFIG", "c4")]
hm = bench.host_model(D, P, G)
feats, pdf = bench.device_frames(hm, T, 1, torch.device("cuda"))
outs = {}
for kernel in (3, 2):
    dm = DeviceModel(D, hm["offsets"])
    dm.set_kernel(kernel)
    dm.upload(hm["weights"], hm["miv"], hm["iv"])
    for mode in ("0", "2", "3"):
        os.environ["KHG_TC_CLUSTER"] = mode
        out = dm.loglikes_all_pdfs(feats, layout=_cabi.KHG_PDF_MAJOR)
        dm.sync()
        outs[mode] = out
        print("kernel", kernel, "mode", mode, "finite", bool(torch.isfinite(out).all()), "mean", float(out.mean()), flush=True)
    for mode in ("2", "3"):
        d = (outs[mode] - outs["0"]).abs().max().item()
        print("kernel", kernel, "mode", mode, "max |diff| vs plain", d, "equal", bool(torch.equal(outs[mode], outs["0"])), flush=True)
