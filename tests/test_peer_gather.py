"""shard.PeerGather on real peers (needs >= 2 GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_peer_gather.py -m gpu`).
Root reads the gathered slot EVERY step without any host synchronisation (the advisor's round-1 finding: with two
slots a rank's copy two steps later could land in the slot root was still reading; three slots + consume() on the
stream of the waits close it).  The single-GPU round trip lives in tests/test_gpu_parity.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_gather_root_reads_every_step_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "peer_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PEER_GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_peer_gather_refuses_two_slots():
    from vision_textile_inspection_b200 import shard
    with pytest.raises(ValueError, match="3 slots"):
        shard.PeerGather(16, "cpu", slots=2)
