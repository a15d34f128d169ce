"""Multi-GPU plumbing of the E-step: one process per GPU (torch.distributed), frames
sharded across ranks, ONE exchange step — the sum of the packed fp64 statistics buffer
over ranks (NCCL all-reduce over NVLink on GPUs; gloo in the CPU tests).

The reference has no distributed code at all; the CPU-semantics equivalent of this
exchange is AccumAmDiagGmm::Add(1.0, other) (reference csrc/mle-am-diag-gmm.cc:119-128,
Kaldi's gmm-sum-accs).  Model parameters are broadcast once per EM iteration.
"""
from typing import Dict, Optional, Tuple

import numpy as np

kGmmMeans, kGmmVariances, kGmmWeights = 1, 2, 4


def augment_flags(flags: int) -> int:
    """AugmentGmmFlags, reference csrc/model-common.cc:72-84."""
    flags &= 0xF
    if flags & kGmmVariances:
        flags |= kGmmMeans
    if flags & kGmmMeans:
        flags |= kGmmWeights
    if not flags & kGmmWeights:
        flags |= kGmmWeights
    return flags


def shard_frames(num_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames [start, stop) owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(num_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_utterances(utt_lengths, world: int):
    """Greedy contiguous split of utterances so that every rank gets ~equal frames;
    returns a list of (first_utt, last_utt_exclusive) per rank."""
    lens = np.asarray(utt_lengths, np.int64)
    cum = np.concatenate([[0], np.cumsum(lens)])
    total = int(cum[-1])
    cuts = [int(np.searchsorted(cum, total * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(lens)
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def packed_layout(num_gauss: int, dim: int, flags: int) -> Dict[str, int]:
    """Offsets (in doubles) inside the packed statistics buffer of khg_stats
    (include/khg_b200.h): [occ G | mean G*D (if m) | var G*D (if v) | tot_like | tot_frames]."""
    flags = augment_flags(flags)
    n = num_gauss
    lay = {"occ": 0, "mean": -1, "var": -1}
    if flags & kGmmMeans:
        lay["mean"] = n
        n += num_gauss * dim
    if flags & kGmmVariances:
        lay["var"] = n
        n += num_gauss * dim
    lay["tot"] = n
    lay["size"] = n + 2
    return lay


def pack_stats(occ, mean, var, tot_like: float, tot_frames: float, flags: int) -> np.ndarray:
    G = occ.shape[0]
    D = mean.shape[1] if mean is not None else (var.shape[1] if var is not None else 1)
    lay = packed_layout(G, D, flags)
    buf = np.zeros(lay["size"], np.float64)
    buf[:G] = occ
    if lay["mean"] >= 0:
        buf[lay["mean"]:lay["mean"] + G * D] = np.asarray(mean).reshape(-1)
    if lay["var"] >= 0:
        buf[lay["var"]:lay["var"] + G * D] = np.asarray(var).reshape(-1)
    buf[lay["tot"]] = tot_like
    buf[lay["tot"] + 1] = tot_frames
    return buf


def unpack_stats(buf, num_gauss: int, dim: int, flags: int):
    lay = packed_layout(num_gauss, dim, flags)
    buf = np.asarray(buf)
    assert buf.shape[0] == lay["size"], (buf.shape, lay)
    G, D = num_gauss, dim
    return dict(
        occ=buf[:G].copy(),
        mean=buf[lay["mean"]:lay["mean"] + G * D].reshape(G, D).copy() if lay["mean"] >= 0 else None,
        var=buf[lay["var"]:lay["var"] + G * D].reshape(G, D).copy() if lay["var"] >= 0 else None,
        tot_like=float(buf[lay["tot"]]), tot_frames=float(buf[lay["tot"] + 1]))


def allreduce_packed(buf, group=None):
    """Sum of the packed statistics over all ranks, in place (torch tensor on any device)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def allreduce_stats(stats, group=None):
    """All-reduce a DeviceStats in place (its packed device buffer, zero-copy)."""
    return allreduce_packed(stats.as_torch(), group)


def broadcast_model(arrays: Dict[str, np.ndarray], device=None, src: int = 0, group=None) -> Dict[str, np.ndarray]:
    """Broadcast the packed model parameters (weights, means_invvars, inv_vars[, gconsts])
    from rank `src`; returns host arrays on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arrays
    out = {}
    for k in sorted(arrays):
        t = torch.from_numpy(np.ascontiguousarray(arrays[k]))
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src, group=group)
        out[k] = t.cpu().numpy()
    return out
