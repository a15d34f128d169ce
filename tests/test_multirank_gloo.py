"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: frame sharding,
the packed-statistics layout and the one exchange step (sum over ranks).  Per-shard
statistics come from the oracle here (no GPU in this container); the property checked is
the one the NCCL path relies on: all-reduce(sum) of per-shard packed stats == stats of the
whole batch == AccumAmDiagGmm::Add of the shards (reference csrc/mle-am-diag-gmm.cc:119-128)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, flags, out_dir):
    for p in (ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kaldi_hmm_gmm_b200 import parallel as par
    from oracle import khg_oracle as ko

    ora = ko.Oracle()
    # rank 0 owns the model; everyone else receives it by broadcast
    if rank == 0:
        model, means, vars_ = ko.make_synthetic_model(13, 9, 40, oracle=ora)
        arrays = dict(weights=model.weights, means_invvars=model.means_invvars, inv_vars=model.inv_vars, gconsts=model.gconsts)
    else:
        model, means, vars_ = ko.make_synthetic_model(13, 9, 40, seed=999, oracle=ora)  # wrong on purpose
        arrays = dict(weights=np.zeros_like(model.weights), means_invvars=np.zeros_like(model.means_invvars),
                      inv_vars=np.zeros_like(model.inv_vars), gconsts=np.zeros_like(model.gconsts))
    arrays = par.broadcast_model(arrays)
    ref_model, means, vars_ = ko.make_synthetic_model(13, 9, 40, oracle=ora)
    for k in arrays:
        assert np.array_equal(arrays[k], getattr(ref_model, k)), k
    model = ko.PackedModel(ref_model.offsets, arrays["weights"], arrays["means_invvars"], arrays["inv_vars"], arrays["gconsts"])
    T = 1001
    feats, pdf = ko.make_synthetic_frames(model, means, vars_, T)
    a, b = par.shard_frames(T, rank, world)
    st = ora.acc_stats_ali(model, feats[a:b], pdf[a:b], flags=flags)
    buf = torch.from_numpy(par.pack_stats(st["occ"], st["mean"], st["var"], st["tot_like"], st["tot_frames"], flags))
    par.allreduce_packed(buf)
    got = par.unpack_stats(buf.numpy(), model.num_gauss, model.dim, flags)
    full = ora.acc_stats_ali(model, feats, pdf, flags=flags)
    for k in ("occ", "mean", "var"):
        if full[k] is None:
            assert got[k] is None
        else:
            np.testing.assert_allclose(got[k], full[k], rtol=1e-12, atol=1e-12)
    assert got["tot_frames"] == T and abs(got["tot_like"] - full["tot_like"]) < 1e-9 * abs(full["tot_like"])
    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")


@pytest.mark.parametrize("world,flags", [(2, 15), (3, 4), (2, 1)])
def test_allreduce_of_sharded_stats_equals_full(tmp_path, world, flags):
    mp.spawn(_worker, args=(world, _free_port(), flags, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), f"ok{r}")) for r in range(world))


def test_shard_helpers():
    for p in (ROOT, os.path.join(ROOT, "kaldi-hmm-gmm_b200", "python")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from kaldi_hmm_gmm_b200 import parallel as par

    for T, W in [(10, 3), (0, 2), (7, 8), (100_000_000, 8)]:
        spans = [par.shard_frames(T, r, W) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == T
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    lens = [500] * 20 + [37, 1200, 3]
    parts = par.shard_utterances(lens, 4)
    assert parts[0][0] == 0 and parts[-1][1] == len(lens) and all(parts[i][1] == parts[i + 1][0] for i in range(3))
    lay = par.packed_layout(10, 4, 2)
    assert lay == {"occ": 0, "mean": 10, "var": 50, "tot": 90, "size": 92}
    assert par.packed_layout(10, 4, 4)["size"] == 12 and par.packed_layout(10, 4, 0)["mean"] == -1
