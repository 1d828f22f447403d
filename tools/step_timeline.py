#!/usr/bin/env python
"""In-step timeline of the cfg2 hot path: K1 on one stream beside K2 -> K3 -> K4 -> K5 on a high-priority stream, joined
per step like the CUDA-graph form of bench.py.  Prints when K1 and the post chain end relative to the step start and the
in-step durations of K2..K5 (library events), to see which branch is the critical path."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision_textile_inspection_b200 import synth  # noqa: E402
from vision_textile_inspection_b200.engine import EngineConfig, InspectionEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    cfg = synth.CONFIGS[name]
    B = cfg.batch
    dev = torch.device("cuda:0")
    batch = synth.make_batch(cfg, B, seed0=1000 * cfg.cfg_id, n_unique=min(B, 16))
    eng = InspectionEngine(EngineConfig.for_workload(cfg, None, max_batch=B), device=dev)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    frames, lv, coef, proto = d(batch["frames"]), [d(l) for l in batch["levels"]], d(batch["coef"]), d(batch["proto"])
    net_in = torch.empty((B, 3, eng.LH, eng.LW), dtype=torch.float32, device=dev)
    outs = eng.post_measure(lv[0], lv[1], lv[2], coef, proto)
    s_pre, s_post = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
    cur = torch.cuda.current_stream(dev)
    eng.set_profiling(True)
    rows = []
    for it in range(30):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(cur)
        s_pre.wait_event(ev[0]); s_post.wait_event(ev[0])
        with torch.cuda.stream(s_post):
            eng.post_measure(lv[0], lv[1], lv[2], coef, proto, outputs=outs)
            ev[2].record(s_post)
        with torch.cuda.stream(s_pre):
            eng.preprocess(frames, out=net_in)
            ev[1].record(s_pre)
        cur.wait_event(ev[1]); cur.wait_event(ev[2])
        ev[3].record(cur)
        torch.cuda.synchronize()
        if it >= 10:
            rows.append([ev[0].elapsed_time(ev[1]), ev[0].elapsed_time(ev[2]), ev[0].elapsed_time(ev[3])] + eng.stage_ms())
    r = np.median(np.array(rows), axis=0)
    print(json.dumps({"config": name, "B": B, "k1_end_ms": r[0], "post_end_ms": r[1], "step_ms": r[2],
                      "in_step_ms": dict(zip(["K1", "K2", "K3", "K4", "K5"], r[3:].tolist()))}))


if __name__ == "__main__":
    main()
