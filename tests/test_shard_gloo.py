"""world_size-2 gloo test of the multi-GPU plumbing (runs on CPU): shard ranges + record gather in frame order."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import torch

from vision_textile_inspection_b200 import shard
from vision_textile_inspection_b200._lib import DET_DTYPE, RESULT_DTYPE


def test_shard_range_partitions():
    for n in (1, 7, 64, 256, 257):
        for w in (1, 2, 3, 4, 8):
            r = [shard.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_gather_records_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    worker = os.path.join(os.path.dirname(__file__), "gloo_worker.py")
    ps = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port)], stdout=subprocess.PIPE, text=True)
          for r in range(2)]
    outs = [json.loads(p.communicate(timeout=180)[0].strip().splitlines()[-1]) for p in ps]
    assert all(p.returncode == 0 for p in ps)
    for o in outs:
        assert o["shape"] == [6, 5, DET_DTYPE.itemsize]
        assert o["counts"] == [0, 1, 2, 10, 11, 12]            # global frame order = rank-major
        assert (o["first"], o["last"], o["r_mid"]) == (1, 2, 8)
