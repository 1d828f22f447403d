"""Multi-GPU plumbing: frames shard by batch, every GPU runs K1..K5 independently, only the compact per-defect
records and per-frame results are gathered.  No collective sits on the data path."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous frame range [lo, hi) of `rank`; the first n % world ranks get one extra frame."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(dets: torch.Tensor, counts: torch.Tensor, results: torch.Tensor, group=None):
    """All-gather fixed-stride records from every rank (equal per-rank batch).  Works on NCCL (device tensors) and
    gloo (CPU tensors).  Returns (dets[W*B,...], counts[W*B], results[W*B,...]) in global frame order."""
    world = dist.get_world_size(group)
    outs = []
    for t in (dets, counts, results):
        t = t.contiguous()
        buf = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t, group=group)      # concatenated along dim 0 = global frame order
        outs.append(buf)
    return tuple(outs)
