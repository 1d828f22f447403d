"""One rank of the world_size-2 gloo test (spawned by test_shard_gloo.py)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_textile_inspection_b200 import shard  # noqa: E402
from vision_textile_inspection_b200._lib import DET_DTYPE, RESULT_DTYPE  # noqa: E402

rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
dist.init_process_group("gloo", rank=rank, world_size=world)
B, max_det = 3, 5
dets = torch.full((B, max_det, DET_DTYPE.itemsize), rank + 1, dtype=torch.uint8)
counts = torch.arange(B, dtype=torch.int32) + 10 * rank
results = torch.full((B, RESULT_DTYPE.itemsize), 7 + rank, dtype=torch.uint8)
d, c, r = shard.gather_records(dets, counts, results)
# the packed single-gather path must give the same thing
packed, (pd, pc, pr, _) = shard.alloc_packed(B, max_det, 'cpu')
pd.copy_(dets); pc.copy_(counts); pr.copy_(results)
d2, c2, r2 = shard.unpack_packed(shard.gather_packed(packed), B, max_det)
assert torch.equal(d, d2) and torch.equal(c, c2) and torch.equal(r, r2)
print(json.dumps(dict(rank=rank, shape=list(d.shape), counts=c.tolist(), first=int(d[0, 0, 0]), last=int(d[-1, 0, 0]),
                      r_mid=int(r[B, 0]))))
dist.destroy_process_group()
