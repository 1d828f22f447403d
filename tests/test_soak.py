"""Soak parity run (opt-in: VTI_SOAK=<scenes per config>): many more seeded scenes than the regular suite, every one
through the FULL oracle (real torch / torchvision post + the measure-stage port), not only the float32 spec --
keep indices and counts bit-exact, every frame's status / stitch counts equal, mm within 0.1 %.  Writes
gpurun_out/soak_parity.json (committed under profiles/ when run).  Skipped in the regular -m gpu run: the CPU oracle
takes about a second per scene."""
import json
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import post_spec
from vision_textile_inspection_b200 import synth
from vision_textile_inspection_b200.engine import EngineConfig, InspectionEngine

pytestmark = pytest.mark.gpu
N = int(os.environ.get("VTI_SOAK", "0"))
MASKS = int(os.environ.get("VTI_SOAK_MASKS", "0"))       # 1: also export every mask and compare it with torch's, pixel by pixel
VARIANT_B = int(os.environ.get("VTI_SOAK_VARIANT_B", "0"))   # 1: newer-Ultralytics mask semantics (logits, > 0, empty masks dropped)
MM_RTOL = 1e-3
TIE_EPS = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.skipif(N <= 0, reason="opt-in: set VTI_SOAK=<scenes per config>")
@pytest.mark.parametrize("name,first,scale", [("cfg2", 20000, 1.0), ("cfg3", 30000, 0.5), ("native", 40000, 0.25),
                                              ("cfg1", 50000, 0.25), ("cfg4", 60000, 0.125), ("cfg5", 70000, 0.125)])
def test_soak(name, first, scale, calib):
    cfg = synth.CONFIGS[name]
    count = max(1, int(N * scale))
    B = min(count, 16)
    ec = EngineConfig.for_workload(cfg, helpers.load_calib(), max_batch=B)
    ec.mask_variant = 1 if VARIANT_B else 0
    eng = InspectionEngine(ec)
    rep = dict(config=name, scenes=0, detections=0, frames_ok=0, frames_no_fabric=0, frames_no_stitch=0, max_rel_mm=0.0)
    if MASKS:
        rep.update(mask_instances=0, mask_pixels=0, flipped_instances=0, flipped_pixels=0, max_flip_margin=0.0, min_iou_ge_1000px=1.0)
    for s0 in range(first, first + count, B):
        seeds = list(range(s0, min(s0 + B, first + count)))
        heads = [synth.planted_head(cfg, s) for s in seeds]
        lv = [dev(np.stack([h["levels"][l] for h in heads])) for l in range(3)]
        dets, counts, results, masks = eng.post_measure(lv[0], lv[1], lv[2], dev(np.stack([h["coef"] for h in heads])),
                                                        dev(np.stack([h["proto"] for h in heads])), export_masks=bool(MASKS))
        dets, counts, results = eng.dets_to_numpy(dets), counts.cpu().numpy(), eng.results_to_numpy(results)
        for b, (seed, hd) in enumerate(zip(seeds, heads)):
            sp = post_spec.postprocess_spec(hd["levels"], hd["coef"], cfg.conf, cfg.iou, cfg.max_det, cfg.nc, cfg.LH,
                                            cfg.LW, cfg.frame_h, cfg.frame_w)
            n = int(counts[b])
            assert n == len(sp["keep_anchor"]), (seed, n)
            assert np.array_equal(dets[b, :n]["anchor"], sp["keep_anchor"]), seed
            assert np.array_equal(dets[b, :n]["box_lb"].view(np.uint32), sp["box_lb"].view(np.uint32)), seed
            _, res, m = helpers.oracle_scene(cfg, seed, calib, return_soft=bool(MASKS), mask_variant="B" if VARIANT_B else "A")
            if VARIANT_B:       # detections whose mask stayed empty are dropped by the newer Ultralytics: compare the survivors
                keep = (dets[b, :n]["flags"] & 512) == 0
                assert np.array_equal(dets[b, :n]["anchor"][keep], np.asarray(res.keep_anchor)), seed
            if MASKS and n and not VARIANT_B:
                ref = res.masks.data.numpy() > 0
                got = eng.unpack_masks(masks, b, n).cpu().numpy() > 0
                soft = res.soft.numpy()
                for k in range(n):
                    diff = np.logical_xor(got[k], ref[k])
                    rep["mask_instances"] += 1
                    rep["mask_pixels"] += int(ref[k].sum())
                    if diff.any():      # only where torch's own value is within 1e-5 of the threshold
                        margin = float(np.abs(soft[k][diff] - 0.5).max())
                        rep["flipped_instances"] += 1
                        rep["flipped_pixels"] += int(diff.sum())
                        rep["max_flip_margin"] = max(rep["max_flip_margin"], margin)
                        assert margin <= TIE_EPS, (seed, k, int(diff.sum()), margin)
                    if ref[k].sum() >= 1000:
                        iou = float((got[k] & ref[k]).sum()) / float((got[k] | ref[k]).sum())
                        rep["min_iou_ge_1000px"] = min(rep["min_iou_ge_1000px"], iou)
                        assert iou >= 0.999, (seed, k, iou)
            if not VARIANT_B:
                assert np.array_equal(dets[b, :n]["anchor"], res.keep_anchor), seed      # the real torchvision NMS
            r = results[b]
            assert (int(r["status"]) & 0xFF) == {"ok": 0, "no_fabric": 2, "no_stitch": 3}[m["status"]], (seed, m["status"])
            rep["scenes"] += 1
            rep["detections"] += n
            rep["frames_" + m["status"]] += 1
            if m["status"] == "ok":
                assert r["n_dist"] == m["n_dist"] and r["n_width"] == m["n_width"], seed
                for key, ref in (("avg_dist", m["avg_dist"]), ("avg_width", m["avg_width"])):
                    if ref is None:
                        assert np.isnan(r[key]), (seed, key)
                    else:
                        rel = abs(r[key] - ref) / ref
                        rep["max_rel_mm"] = max(rep["max_rel_mm"], float(rel))
                        assert rel <= MM_RTOL, (seed, key, r[key], ref)
    path = os.path.join(helpers.ROOT, "gpurun_out", "soak_masks.json" if MASKS else ("soak_variant_b.json" if VARIANT_B else "soak_parity.json"))
    os.makedirs(os.path.dirname(path), exist_ok=True)
    allrep = json.load(open(path)) if os.path.exists(path) else {}
    allrep[name] = rep
    json.dump(allrep, open(path, "w"), indent=1)
    print("soak", json.dumps(rep))


@pytest.mark.skipif(N <= 0, reason="opt-in: set VTI_SOAK=<scenes per config>")
def test_soak_k1_random_geometries():
    """K1's planner (tile footprints, resize taps, undistort boxes, the generic fallback for widths that are not a
    multiple of 8) over random frame sizes / imgsz / undistort: output bit-exact against real cv2 + torch every time."""
    from oracle import cv_fixed, ultra_ref
    calib = helpers.load_calib()
    rng = np.random.default_rng(12345)
    count = max(4, N // 4)
    rep = dict(geometries=0, fast_path_widths=0, generic_widths=0, undistort_on=0, pixels=0)
    for it in range(count):
        h = int(rng.integers(48, 1400))
        w = int(rng.integers(48, 2100))
        if it % 3:
            w &= ~7                                          # the fast path needs w % 8 == 0; every third width is arbitrary
        imgsz = int(rng.choice([320, 480, 640, 960, 1280]))
        undistort = int(rng.integers(0, 2))
        flip = int(rng.integers(0, 2))
        frames = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
        K = cv_fixed.scale_K(np.array(calib["camera_matrix"]), w, h)
        ec = EngineConfig(frame_h=h, frame_w=w, K=K, dist=np.array(calib["dist_coeffs"]), R=np.eye(3), t=np.array([0, 0, 0.1]),
                          imgsz=imgsz, max_batch=2, undistort=undistort, channel_flip=flip, roi=(0, 0, 0, 0, 0))
        eng = InspectionEngine(ec)
        got = eng.preprocess(dev(frames)).cpu().numpy()
        ref = ultra_ref.preprocess(list(frames), imgsz, undistort=(K, ec.dist) if undistort else None, flip_channels=bool(flip)).numpy()
        assert got.shape == ref.shape, (h, w, imgsz, undistort, flip)
        assert np.array_equal(got, ref), (h, w, imgsz, undistort, flip, int((got != ref).sum()))
        rep["geometries"] += 1
        rep["fast_path_widths" if w % 8 == 0 else "generic_widths"] += 1
        rep["undistort_on"] += undistort
        rep["pixels"] += int(got.size)
        del eng
    path = os.path.join(helpers.ROOT, "gpurun_out", "soak_k1.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(rep, open(path, "w"), indent=1)
    print("soak k1", json.dumps(rep))


@pytest.mark.skipif(N <= 0, reason="opt-in: set VTI_SOAK=<scenes per config>")
def test_soak_random_heads():
    """Decode + filter + NMS fuzz on head tensors that are pure noise (not the planted generator): random class counts,
    thresholds, max_det, logit scales -- hundreds to thousands of heavily overlapping candidates per frame.  Keep indices,
    scores, classes and boxes bit-exact against the float32 spec AND the keep set equal to real torchvision's."""
    from oracle import ultra_ref
    calib = helpers.load_calib()
    rng = np.random.default_rng(777)
    count = max(4, N // 4)
    rep = dict(frames=0, candidates=0, kept=0, truncated_at_max_det=0, overflowed=0, torch_ulp_tie_frames=0, torch_ulp_tie_margin_max=0.0)
    for it in range(count):
        imgsz = int(rng.choice([320, 480, 640]))
        nc = int(rng.choice([1, 2, 3, 5, 17, 80]))
        conf = float(rng.choice([0.02, 0.05, 0.2, 0.4]))
        iou = float(rng.choice([0.1, 0.25, 0.45, 0.7, 0.9]))
        max_det = int(rng.choice([10, 100, 300, 1000]))
        K = cv_scale(calib, imgsz)
        ec = EngineConfig(frame_h=imgsz, frame_w=imgsz, K=K, dist=np.array(calib["dist_coeffs"]), R=np.eye(3),
                          t=np.array([0, 0, 0.1]), imgsz=imgsz, nc=nc, conf=conf, iou=iou, max_det=max_det, max_batch=1)
        eng = InspectionEngine(ec)
        LH = LW = imgsz
        cls_mu, box_sd = float(rng.uniform(-5.0, -1.5)), float(rng.uniform(0.5, 3.0))
        levels = []
        for s in (8, 16, 32):
            x = np.empty((64 + nc, LH // s, LW // s), np.float32)
            x[:64] = rng.normal(0, box_sd, x[:64].shape)
            x[64:] = rng.normal(cls_mu, 1.5, x[64:].shape)
            if it % 5 == 0:                                   # exact score ties across anchors
                x[64:] = np.round(x[64:] * 2) / 2
            levels.append(x)
        A = sum(l.shape[1] * l.shape[2] for l in levels)
        coef = rng.normal(0, 1, (32, A)).astype(np.float32)
        proto = rng.normal(0, 1, (32, LH // 4, LW // 4)).astype(np.float32)
        dets, counts, results, _ = eng.post_measure(*[dev(l[None]) for l in levels], dev(coef[None]), dev(proto[None]))
        torch.cuda.synchronize()
        r = eng.results_to_numpy(results)[0]
        sp = post_spec.postprocess_spec(levels, coef, conf, iou, max_det, nc, LH, LW, imgsz, imgsz)
        rep["frames"] += 1
        rep["candidates"] += sp["n_cand"]
        if int(r["status"]) & 0x100 or sp["n_cand"] > eng.g.max_candidates:     # candidate list overflow: flagged, not compared
            rep["overflowed"] += 1
            del eng
            continue
        n = int(counts[0])
        d = eng.dets_to_numpy(dets)[0, :n]
        key = (imgsz, nc, conf, iou, max_det, it)
        assert n == len(sp["keep_anchor"]), (key, n, len(sp["keep_anchor"]))
        assert np.array_equal(d["anchor"], sp["keep_anchor"]), key
        assert np.array_equal(d["cls"], sp["cls"]), key
        assert np.array_equal(d["conf"].view(np.uint32), sp["conf"].view(np.uint32)), key
        assert np.array_equal(d["box_lb"].view(np.uint32), sp["box_lb"].view(np.uint32)), key
        res = ultra_ref.postprocess([l[None] for l in levels], coef[None], proto[None], (imgsz, imgsz), conf, iou, max_det, nc,
                                    with_masks=False)[0]
        if not np.array_equal(d["anchor"], np.asarray(res.keep_anchor)):
            # The keep set equals the float32 spec but not this torch build's: legitimate only when a decision sits within
            # float noise of its threshold -- torch's vectorised exp decodes boxes / scores that differ from the spec's by
            # an ulp, and an IoU (or score) a few 1e-6 from the threshold then falls on the other side.
            m = _tie_margin(sp, iou, conf)
            assert m < 1e-5, (key, m)
            rep["torch_ulp_tie_frames"] += 1
            rep["torch_ulp_tie_margin_max"] = max(rep["torch_ulp_tie_margin_max"], m)
        rep["kept"] += n
        rep["truncated_at_max_det"] += int(n == max_det)
        del eng
    path = os.path.join(helpers.ROOT, "gpurun_out", "soak_random_heads.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(rep, open(path, "w"), indent=1)
    print("soak heads", json.dumps(rep))


def _tie_margin(sp, iou_thr, conf_thr):
    """Smallest distance of any decision of the frame to its threshold: |IoU - thr| over same-class candidate pairs and
    |score - conf| over the candidates (float32 spec values)."""
    b, c, s_ = sp["cand_xyxy"].astype(np.float64), sp["cand_cls"], sp["cand_conf"].astype(np.float64)
    best = float(np.min(np.abs(s_ - conf_thr))) if len(s_) else 1.0
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    for i0 in range(0, len(b), 512):
        bb = b[i0:i0 + 512]
        w = np.clip(np.minimum(bb[:, None, 2], b[None, :, 2]) - np.maximum(bb[:, None, 0], b[None, :, 0]), 0, None)
        h = np.clip(np.minimum(bb[:, None, 3], b[None, :, 3]) - np.maximum(bb[:, None, 1], b[None, :, 1]), 0, None)
        inter = w * h
        v = inter / (area[i0:i0 + 512, None] + area[None, :] - inter)
        same = c[i0:i0 + 512, None] == c[None, :]
        if same.any():
            best = min(best, float(np.min(np.abs(v[same] - iou_thr))))
    return best


def cv_scale(calib, imgsz):
    from oracle import cv_fixed
    return cv_fixed.scale_K(np.array(calib["camera_matrix"]), imgsz, imgsz)


@pytest.mark.skipif(N <= 0, reason="opt-in: set VTI_SOAK=<scenes per config>")
@pytest.mark.parametrize("name,first", [("cfg2", 21000), ("cfg4", 61000), ("cfg3", 31000)])
def test_soak_host_path_equals_device_path(name, first):
    """vti_process_host on pinned buffers (frames by DMA, head tensors read in place over PCIe, only the union rectangle
    of the crop windows of the prototypes fetched) against the device-resident path on the same scenes: identical
    record bytes, counts and frame results, chunk by chunk."""
    cfg = synth.CONFIGS[name]
    B = 16
    rounds = max(1, N // 64)
    eng = InspectionEngine(EngineConfig.for_workload(cfg, helpers.load_calib(), max_batch=B))
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    rep = dict(config=name, scenes=0, detections=0)
    for r in range(rounds):
        batch = synth.make_batch(cfg, B, seed0=first + r * B)
        keep = [pin(batch["frames"])] + [pin(l) for l in batch["levels"]] + [pin(batch["coef"]), pin(batch["proto"])]
        d1, c1, r1, n1 = eng.process_host(*[t.numpy() for t in keep], want_net_in=True)
        dv = [t.cuda() for t in keep]
        n2 = eng.preprocess(dv[0]).cpu().numpy()
        d2, c2, r2, _ = eng.post_measure(*dv[1:])
        d2, c2, r2 = eng.dets_to_numpy(d2), c2.cpu().numpy(), eng.results_to_numpy(r2)
        assert np.array_equal(n1, n2) and np.array_equal(c1, c2) and r1.tobytes() == r2.tobytes(), (name, r)
        for b in range(B):
            assert d1[b, :c1[b]].tobytes() == d2[b, :c2[b]].tobytes(), (name, r, b)
        rep["scenes"] += B
        rep["detections"] += int(c1.sum())
    print("soak host", json.dumps(rep))
