"""K3t (tensor-core statistics kernel, khg_stats_tc.cu) against the fp32 kernel and the CPU oracle on the same
frames: statistics, totals and per-frame log-likes.  usage: tools/stats_tc_check.py [config] [frames]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, DeviceStats  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400_000
D, P, G, _ = bench.CONFIGS[cfg]
hm = bench.host_model(D, P, G)
feats, pdf = bench.device_frames(hm, n, 1, torch.device("cuda"))
w = torch.rand(n, device="cuda") * 2.0


def run(kernel, weights):
    os.environ["KHG_STATS_KERNEL"] = kernel
    dm = DeviceModel(D, hm["offsets"])
    dm.upload(hm["weights"], hm["miv"], hm["iv"])
    st = DeviceStats(dm)
    pf = torch.zeros(n, device="cuda")
    tot = st.acc_stats_ali(feats, pdf, frame_weights=weights, per_frame=pf, want_total=True)
    torch.cuda.synchronize()
    d = st.download()
    return d, tot, pf.cpu().numpy()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    floor = 1e-6 * np.abs(b).max()
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())


for wt in (None, w):
    ds, ts, ps = run("simt", wt)
    dt, tt, pt = run("tc", wt)
    print(f"{cfg} n={n} weights={'yes' if wt is not None else 'no'}: tot simt {ts:.6f} tc {tt:.6f} rel {abs(ts - tt) / abs(ts):.2e}; "
          f"per-frame max abs {np.abs(ps - pt).max():.2e}; "
          + "; ".join(f"{k} max rel {rel(dt[k], ds[k]):.2e}" for k in ("occ", "mean", "var"))
          + f"; frames {ds['tot_frames']:.3f} / {dt['tot_frames']:.3f}")

if os.environ.get("STK_DETAIL"):
    ds, ts, ps = run("simt", None)
    dt, tt, pt = run("tc", None)
    for k in ("mean", "var"):
        a_, b_ = np.asarray(dt[k], np.float64).reshape(-1, D), np.asarray(ds[k], np.float64).reshape(-1, D)
        err = np.abs(a_ - b_) / np.maximum(np.abs(b_), 1e-6 * np.abs(b_).max())
        print(k, "max rel err per dimension:", " ".join(f"{e:.1e}" for e in err.max(0)))
        flat = np.argsort(err.ravel())[::-1][:8]
        occ = np.asarray(ds["occ"], np.float64)
        for f_ in flat:
            g_, d_ = divmod(int(f_), D)
            print(f"  g={g_} (in-pdf {g_ - int(hm['offsets'][np.searchsorted(hm['offsets'], g_, side='right') - 1])}) d={d_} simt={b_[g_, d_]:.6f} tc={a_[g_, d_]:.6f} "
                  f"rel={err[g_, d_]:.2e} occ={occ[g_]:.1f}")
