"""K1 rate on the same frames in three orders: i.i.d. (the bench's synthetic order), sorted by aligned
pdf (neighbouring frames alike, as in real speech), and one tile repeated.  Same work, same model."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")]
import torch  # noqa: E402

import bench  # noqa: E402
from kaldi_hmm_gmm_b200 import DeviceModel, _cabi  # noqa: E402

T = 148 * 128 * 16
D, P, G, _ = bench.CONFIGS["c4"]
hm = bench.host_model(D, P, G)
dm = DeviceModel(D, hm["offsets"])
dm.set_kernel(3)
dm.upload(hm["weights"], hm["miv"], hm["iv"])
feats, pdf = bench.device_frames(hm, T, 1, torch.device("cuda"))
order = torch.argsort(pdf.long(), stable=True)
variants = {"iid": feats, "sorted_by_pdf": feats[order].contiguous(), "one_tile_repeated": feats[:128].repeat(T // 128, 1).contiguous()}
block = torch.empty((P, T), device="cuda")
for name, f in variants.items():
    for _ in range(3):
        dm.loglikes_all_pdfs(f, layout=_cabi.KHG_PDF_MAJOR, out=block)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(0)
    sampler.start()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 2.5:
        for _ in range(4):
            dm.loglikes_all_pdfs(f, layout=_cabi.KHG_PDF_MAJOR, out=block)
        n += 4
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ck = sampler.stop()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: {ms:.3f} ms/launch  {T / ms / 1e3:.2f} M frames/s  sm_mhz={ck.get('sm_mhz')} power_max={ck.get('power_w_max')} "
          f"reasons={ck.get('reasons')}", flush=True)
