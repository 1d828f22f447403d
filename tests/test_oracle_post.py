"""Pins oracle/post_spec.py (deterministic float32 spec) against the real torch / torchvision operators."""
import numpy as np
import pytest
import torch
import torchvision

from oracle import post_spec as S
from oracle import ultra_ref as U
from vision_textile_inspection_b200 import synth


def test_exp_spec_accuracy():
    x = np.linspace(-86, 88, 200001).astype(np.float32)
    ref = np.exp(x.astype(np.float64))
    got = S.exp_spec(x).astype(np.float64)
    rel = np.abs(got - ref) / ref
    assert rel.max() < 2.5e-7          # ~2 ulp
    xs = np.linspace(-30, 30, 100001).astype(np.float32)
    s = S.sigmoid_spec(xs)
    t = torch.sigmoid(torch.from_numpy(xs)).numpy()
    assert np.abs(s - t).max() < 2e-7


@pytest.mark.parametrize("name,seed", [("native", 0), ("cfg2", 2000), ("cfg3", 3000), ("cfg4", 4000), ("cfg1", 1000)])
def test_spec_matches_torch_path(name, seed):
    cfg = synth.CONFIGS[name]
    hd = synth.planted_head(cfg, seed)
    sp = S.postprocess_spec(hd["levels"], hd["coef"], cfg.conf, cfg.iou, cfg.max_det, cfg.nc, cfg.LH, cfg.LW,
                            cfg.frame_h, cfg.frame_w)
    res = U.postprocess([l[None] for l in hd["levels"]], hd["coef"][None], hd["proto"][None],
                        (cfg.frame_h, cfg.frame_w), cfg.conf, cfg.iou, cfg.max_det, cfg.nc, with_masks=False)[0]
    assert np.array_equal(res.keep_anchor.numpy(), sp["keep_anchor"])          # NMS keep indices bit-exact
    assert np.array_equal(res.boxes.cls.numpy().astype(int), sp["cls"])
    assert np.abs(res.box_lb.numpy() - sp["box_lb"]).max() < 1e-3               # boxes within 1e-3 px
    assert np.abs(res.boxes.xyxy.numpy() - sp["box_frame"]).max() < 1e-3
    assert np.abs(res.boxes.conf.numpy() - sp["conf"]).max() < 3e-7


def test_nms_spec_equals_torchvision_on_identical_inputs():
    """Given the same float32 boxes/scores the spec's greedy NMS is bit-identical to torchvision.ops.nms,
    including ties (stable order) and the float-vs-double threshold compare."""
    rng = np.random.default_rng(3)
    for thr in (0.25, 0.45, 0.7):
        n = 600
        xy = rng.uniform(0, 300, (n, 2)).astype(np.float32)
        wh = rng.uniform(5, 80, (n, 2)).astype(np.float32)
        b = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        s = np.round(rng.uniform(0.2, 1, n), 2).astype(np.float32)            # many exact score ties
        c = rng.integers(0, 2, n).astype(np.int32)
        ref = torchvision.ops.nms(torch.from_numpy(b) + torch.from_numpy(c).float()[:, None] * 7680,
                                  torch.from_numpy(s), thr).numpy()
        got = S.nms_spec(b, s, c, thr, max_det=10 ** 9)
        assert np.array_equal(ref, got)


def test_scale_boxes_spec_equals_torch():
    rng = np.random.default_rng(4)
    for (h, w, S_) in [(960, 1280, 960), (1080, 1920, 640), (640, 640, 960), (2160, 3840, 960), (720, 1280, 960)]:
        g = synth.letterbox_geometry(h, w, S_)
        b = rng.uniform(-20, 1000, (500, 4)).astype(np.float32)
        ref = U.scale_boxes((g["LH"], g["LW"]), torch.from_numpy(b), (h, w)).numpy()
        assert np.array_equal(ref, S.scale_boxes_spec(b, g["LH"], g["LW"], h, w))
