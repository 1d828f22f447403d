"""Worker of tests/test_peer_gather.py: N ranks (one per GPU), every rank pushes a buffer whose bytes encode (rank, step)
through shard.PeerGather for many steps with NO host synchronisation in between; root copies the slot out every step
(consume, on the stream of the waits) and checks at the end that every step's copy holds every rank's bytes of THAT step
-- the slot-reuse race the two-slot version had would show up as a later step's bytes."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision_textile_inspection_b200 import shard  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nbytes, steps = 1 << 20, 64
    peer = shard.PeerGather(nbytes, dev)
    src = [torch.full((nbytes,), (rank * 16 + s) % 251, dtype=torch.uint8, device=dev) for s in range(steps)]
    got = torch.empty((steps, world, nbytes), dtype=torch.uint8, device=dev) if rank == 0 else None
    busy = torch.empty((4096, 4096), device=dev)
    for s in range(steps):
        if rank != 0 and s % 3 == 0:
            busy = busy @ busy.clamp(-1, 1) * 0 + 1.0          # uneven pacing between the ranks
        slot = peer.push(src[s])
        if rank == 0:
            peer.consume(slot, out=got[s])
    torch.cuda.synchronize()
    dist.barrier()
    ok = True
    if rank == 0:
        for s in range(steps):
            for r in range(world):
                v = (r * 16 + s) % 251
                if not bool((got[s, r] == v).all()):
                    ok = False
                    print(f"step {s} row {r}: expected {v}, got {got[s, r].unique().tolist()[:4]}")
        print("PEER_GATHER_OK" if ok else "PEER_GATHER_BAD")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
