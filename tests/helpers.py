"""Shared helpers for the parity tests: run the CPU oracle on a seeded synthetic scene."""
import json
import os

import numpy as np

from oracle import cv_fixed, measure_port, ultra_ref
from vision_textile_inspection_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_calib():
    with open(os.path.join(ROOT, "vision_textile_inspection_b200", "data", "reference_calibration.json")) as f:
        return json.load(f)


def measure_config(cfg, calib) -> measure_port.MeasureConfig:
    K = cv_fixed.scale_K(np.array(calib["camera_matrix"]), cfg.frame_w, cfg.frame_h)
    dist = np.zeros(5) if cfg.undistort else np.array(calib["dist_coeffs"])
    ex = calib[cfg.extrinsics]
    return measure_port.MeasureConfig(
        K=K, dist=dist, R=measure_port.rodrigues(ex["rvec"]), t=np.array(ex["tvec"]), variant=cfg.variant,
        roi=cfg.roi(), max_px_distance=250 if cfg.variant == 0 else 150)


def oracle_scene(cfg, seed, calib, with_masks=True, return_soft=False, mask_variant="A"):
    """Oracle post + measure for one seeded frame.  Returns (head dict, RefResults, measure dict)."""
    sc = synth.make_scene(cfg, seed)
    hd = synth.planted_head(cfg, seed, sc)
    res = ultra_ref.postprocess([l[None] for l in hd["levels"]], hd["coef"][None], hd["proto"][None],
                                (cfg.frame_h, cfg.frame_w), cfg.conf, cfg.iou, cfg.max_det, cfg.nc,
                                with_masks=with_masks, return_soft=return_soft, mask_variant=mask_variant)[0]
    m = None
    if with_masks:
        m = measure_port.measure_frame(res.boxes.cls.numpy(), res.boxes.xyxy.numpy(), res.masks.data.numpy(),
                                       cfg.frame_h, cfg.frame_w, measure_config(cfg, calib))
    return hd, res, m
