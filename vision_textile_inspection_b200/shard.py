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


def alloc_packed(B: int, max_det: int, device):
    """One contiguous buffer holding a rank's records, per-frame results and counts, plus typed views into it, so
    that the per-step exchange is ONE all-gather: (packed, (dets, counts, results, None))."""
    from ._lib import DET_DTYPE, RESULT_DTYPE
    nd, nr = B * max_det * DET_DTYPE.itemsize, B * RESULT_DTYPE.itemsize
    packed = torch.zeros((nd + nr + 4 * B,), dtype=torch.uint8, device=device)
    dets = packed[:nd].view(B, max_det, DET_DTYPE.itemsize)
    results = packed[nd:nd + nr].view(B, RESULT_DTYPE.itemsize)
    counts = packed[nd + nr:].view(torch.int32)
    return packed, (dets, counts, results, None)


def gather_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of the packed per-rank buffers: (W, bytes).  Row r holds rank r's frames (global frame order = rank
    order); unpack_packed() gives the typed views."""
    world = dist.get_world_size(group)
    out = torch.empty((world * packed.numel(),), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)      # flat: concatenation in rank order
    return out.view(world, packed.numel())


def unpack_packed(gathered: torch.Tensor, B: int, max_det: int):
    """(W, bytes) -> dets (W*B, max_det, 160) u8, counts (W*B,) i32, results (W*B, 56) u8."""
    from ._lib import DET_DTYPE, RESULT_DTYPE
    W = gathered.shape[0]
    nd, nr = B * max_det * DET_DTYPE.itemsize, B * RESULT_DTYPE.itemsize
    dets = gathered[:, :nd].reshape(W * B, max_det, DET_DTYPE.itemsize)
    results = gathered[:, nd:nd + nr].reshape(W * B, RESULT_DTYPE.itemsize)
    counts = gathered[:, nd + nr:].contiguous().view(torch.int32).reshape(W * B)
    return dets, counts, results


class PeerGather:
    """Gather-to-root of the packed per-rank records over NVLink peer memory, with no collective kernel on any SM.

    SURVEY.md 8e asks for the records on rank 0 (the host logic of main.py:225-293 runs once, in frame order), so an
    all-gather moves W times more bytes than needed and its NCCL kernel takes SMs from K1.  Here every rank owns a
    symmetric-memory buffer [slots][W][bytes]; a rank pushes its packed records into ROOT's copy of slot s, row
    `rank`, with one device-to-device copy through the peer mapping (copy engine over NVLink / NVSwitch), then raises
    a signal that root's stream waits on.

    Flow control.  put_signal(n + 1) completes only once root's wait_signal(n) has consumed signal n, and a rank's
    copy of step n + k is ordered (same stream) after its put_signal(n + k - 1).  Root's wait_signal(n) precedes its
    READ of slot n, so with two slots the copy of step n + 2 could land in the row root is still reading.  With THREE
    slots (step n uses slot n % 3) the copy of step n + 3 is ordered after root consumed signal n + 1, which root's
    stream issues after its read of slot n -- provided root reads on the SAME stream that issued the waits, which is
    what push()/consume() do.  tests/test_peer_gather.py reads the slot every step on 2 GPUs.
    """

    def __init__(self, nbytes: int, device, group=None, root: int = 0, slots: int = 3):
        import torch.distributed._symmetric_memory as symm
        if slots < 3:
            raise ValueError("PeerGather needs >= 3 slots (see the flow-control note in the class docstring)")
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world, self.root, self.slots = dist.get_rank(group), dist.get_world_size(group), root, slots
        self.nbytes = nbytes
        self.step = 0
        self.buf = symm.empty((slots, self.world, nbytes), dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        self.root_buf = self.hdl.get_buffer(root, (slots, self.world, nbytes), torch.uint8)
        torch.cuda.synchronize(device)
        self.hdl.barrier()

    def push(self, packed: torch.Tensor, slot: int | None = None) -> int:
        """Enqueue, on the current stream, this rank's copy into root's slot and the signal exchange.  Every rank must
        call push() the same number of times, from one stream; returns the slot used (step % slots)."""
        slot = self.step % self.slots if slot is None else slot
        self.step += 1
        self.root_buf[slot, self.rank].copy_(packed.view(-1))
        if self.rank != self.root:
            self.hdl.put_signal(self.root, channel=self.rank)
        else:
            for r in range(self.world):
                if r != self.root:
                    self.hdl.wait_signal(r, channel=r)
        return slot

    def gathered(self, slot: int) -> torch.Tensor:
        """Root only: (W, bytes) view of a slot; unpack_packed() gives the typed views in global frame order.  Read it
        on the stream push() ran on, before the push() two steps later (or use consume())."""
        return self.buf[slot]

    def consume(self, slot: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """Root only: copy a gathered slot out on the current stream (the one push() ran on), so that the slot may be
        reused; returns the (W, bytes) copy."""
        src = self.buf[slot]
        if out is None:
            out = torch.empty_like(src)
        out.copy_(src)
        return out
