#!/usr/bin/env python
"""Single-frame latency of the outer drop-in call (BASELINE.json configs[0] / native: what main.py:211 does per frame):
StitchMeasurementApp.process_frame(frame) with the backbone replaced by head tensors that are already on the device,
so the number is pre + post + measure + the host side of the call (H2D of the frame, launches, D2H of the records)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision_textile_inspection_b200 import app as A, synth  # noqa: E402
from vision_textile_inspection_b200.engine import load_reference_calibration  # noqa: E402


def main():
    out = {}
    calib = load_reference_calibration()
    tmp = tempfile.mkdtemp()
    cp, ep = os.path.join(tmp, "c.json"), os.path.join(tmp, "e.json")
    json.dump({"camera_matrix": calib["camera_matrix"], "dist_coeffs": calib["dist_coeffs"]}, open(cp, "w"))
    for name in ("native", "cfg1", "cfg2"):
        cfg = synth.CONFIGS[name]
        json.dump(calib[cfg.extrinsics], open(ep, "w"))
        head = synth.planted_head(cfg, 7)
        dev = torch.device("cuda:0")
        t = [torch.from_numpy(head["levels"][l][None]).to(dev) for l in range(3)] + \
            [torch.from_numpy(head["coef"][None]).to(dev), torch.from_numpy(head["proto"][None]).to(dev)]
        app = A.StitchMeasurementApp(cp, ep, "best_Model.pt", camera_index=None, calib_w=cfg.frame_w, calib_h=cfg.frame_h,
                                     backbone=lambda net_in: tuple(t), roi=cfg.roi(), imgsz=cfg.imgsz, annotate=False,
                                     undistort=cfg.undistort)
        frame = synth.fabric_frame(cfg, 7)
        for use_graph in (False, True):
            app.model.use_graph = use_graph
            for _ in range(20):
                app.process_frame(frame)
            ts = []
            for _ in range(300):
                t0 = time.perf_counter()
                _, m = app.process_frame(frame)
                ts.append((time.perf_counter() - t0) * 1e3)
            ts = np.array(ts)
            out[f"{name}_{'graph' if use_graph else 'streams'}"] = {
                "ms_p50": float(np.median(ts)), "ms_p10": float(np.percentile(ts, 10)), "ms_p90": float(np.percentile(ts, 90)),
                "stitch_count": m["stitch_count"], "frame": list(frame.shape)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
