"""K6: the annotated overlay rasterised on the GPU and the nvJPEG encode, against cv2 itself.

The reference draws with cv2 (measurement.py:230-236, 268-272, 292-296, 358-368, 460-462) and main.py:314 writes the
annotated frame as a JPEG.  Rectangles, discs and the axis-aligned lines must be cv2's exact pixel sets; the thick envelope
polyline covers >= 95 % of cv2's pixels; the JPEG must decode (with cv2) to the annotated frame."""
import cv2
import numpy as np
import pytest
import torch

import helpers
from vision_textile_inspection_b200 import _lib, synth
from vision_textile_inspection_b200.engine import EngineConfig, InspectionEngine

pytestmark = pytest.mark.gpu


def _run(cfg, seed):
    hd = synth.planted_head(cfg, seed)
    frame = synth.fabric_frame(cfg, seed)
    eng = InspectionEngine(EngineConfig.for_workload(cfg, helpers.load_calib(), max_batch=1))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dets, counts, results, _ = eng.post_measure(*[dev(l[None]) for l in hd["levels"]], dev(hd["coef"][None]), dev(hd["proto"][None]))
    d_frame = dev(frame[None])
    ann = eng.annotate(d_frame, dets, counts)
    torch.cuda.synchronize()
    return eng, frame, ann, eng.dets_to_numpy(dets)[0, :int(counts[0])], eng.results_to_numpy(results)[0], d_frame, dets, counts


def _cv_reference(cfg, frame, d):
    """The reference's drawing calls, in its order, with cv2 (no text, no contours)."""
    img = frame.copy()
    h, w = frame.shape[:2]
    en, x0, x1, y0, y1 = cfg.roi()
    layers = {}
    if en:
        x0, x1 = max(0, min(x0, w - 1)), max(0, min(x1, w - 1))
        y0, y1 = max(0, min(y0, h - 1)), max(0, min(y1, h - 1))
        cv2.rectangle(img, (x0, y0), (x1, y1), (144, 238, 144), 2)
    for k in d:
        if (k["flags"] & _lib.F_IN_ROI) and (k["flags"] & _lib.F_FABRIC):
            cv2.rectangle(img, tuple(int(v) for v in k["box_int"][:2]), tuple(int(v) for v in k["box_int"][2:]), (255, 0, 255), 2)
    for k in d:
        if (k["flags"] & _lib.F_IN_ROI) and (k["flags"] & _lib.F_STITCH):
            cv2.rectangle(img, tuple(int(v) for v in k["box_int"][:2]), tuple(int(v) for v in k["box_int"][2:]), (255, 255, 0), 1)
    layers["boxes"] = img.copy()
    return img, layers


def test_boxes_and_markers_are_cv2_pixel_sets():
    """Everything except the (thick, sloped) envelope polyline must be cv2's drawing, pixel for pixel: the reference's calls
    replayed with cv2 in the reference's order, compared wherever our image does not show the envelope colour."""
    cfg = synth.CONFIGS["native"]
    eng, frame, ann, d, r, *_ = _run(cfg, 0)
    got = ann[0].cpu().numpy()
    assert got.shape == frame.shape and (got != frame).any()
    ref, _ = _cv_reference(cfg, frame, d)                       # ROI, fabric boxes, stitch boxes
    for k in d:                                                 # measurement.py:358-363
        if (k["flags"] & _lib.F_STITCH) and (k["flags"] & _lib.F_IN_ROI) and not np.isnan(k["cx"]):
            cy = int(round(k["cy"]))
            cv2.circle(ref, (int(round(k["left_px"])), cy), 3, (200, 200, 0), -1)
            cv2.circle(ref, (int(round(k["right_px"])), cy), 3, (200, 200, 0), -1)
            cv2.line(ref, (int(round(k["left_px"])), cy), (int(round(k["right_px"])), cy), (200, 200, 0), 1)
            cv2.circle(ref, (int(round(k["cx"])), cy), 3, (200, 0, 0), -1)
    n_dist = 0
    for k in d:                                                 # measurement.py:460-462
        if (k["flags"] & _lib.F_HAS_DIST) and (k["flags"] & _lib.F_STITCH):
            cx, cy, ey = int(round(k["cx"])), int(round(k["cy"])), int(round(k["edge_y"]))
            cv2.line(ref, (cx, ey), (cx, cy), (0, 255, 0), 1)
            cv2.circle(ref, (cx, ey), 2, (255, 0, 255), -1)
            n_dist += 1
    assert n_dist >= 3 and (ref != frame).any()
    not_env = ~np.all(got == np.array((255, 128, 0), np.uint8), axis=2)
    bad = np.any(got != ref, axis=2) & not_env
    if bad.any():
        ys, xs = np.where(bad)
        msg = [(int(y), int(x), got[y, x].tolist(), ref[y, x].tolist(), frame[y, x].tolist()) for y, x in list(zip(ys, xs))[:12]]
        raise AssertionError(f"{int(bad.sum())} pixels differ from cv2 (y, x, ours, cv2, frame): {msg}")
    # and the primitives are really there (not everything hidden behind the envelope)
    for col in ((144, 238, 144), (255, 0, 255), (255, 255, 0), (200, 200, 0), (200, 0, 0), (0, 255, 0)):
        assert np.all(got == np.array(col, np.uint8), axis=2).sum() > 0, col


def test_envelope_polyline_covers_cv2s():
    cfg = synth.CONFIGS["native"]
    eng, frame, ann, d, r, *_ = _run(cfg, 1)
    got = np.all(ann[0].cpu().numpy() == np.array((255, 128, 0), np.uint8), axis=2)
    # the envelope K5 built, read back through the records: median rows are not it -- rebuild from the oracle instead
    _, res, _ = helpers.oracle_scene(cfg, 1, helpers.load_calib())
    from oracle import measure_port
    fab = None
    for k in range(res.boxes.cls.shape[0]):
        if int(res.boxes.cls[k]) == 1:
            bm = measure_port.instance_bitmap(res.masks.data[k].numpy(), cfg.frame_h, cfg.frame_w)
            if bm is not None:
                fab = bm if fab is None else np.maximum(fab, bm)
    env = measure_port.fabric_envelope(fab, upper=False)
    pts = [(x, int(env[x])) for x in range(cfg.frame_w) if env[x] >= 0]
    ref = np.zeros((cfg.frame_h, cfg.frame_w, 3), np.uint8)
    cv2.polylines(ref, [np.array(pts, np.int32)], False, (255, 128, 0), 2)
    want = np.all(ref == np.array((255, 128, 0), np.uint8), axis=2)
    painted_over = np.any(ann[0].cpu().numpy() != frame, axis=2) & ~got
    cover = (want & (got | painted_over)).sum() / want.sum()
    assert cover >= 0.95, cover
    assert (got & ~cv2.dilate(want.astype(np.uint8), np.ones((3, 3), np.uint8)).astype(bool)).sum() <= 0.02 * got.sum()


def test_text_and_jpeg_roundtrip():
    cfg = synth.CONFIGS["native"]
    eng, frame, ann, d, r, d_frame, dets, counts = _run(cfg, 2)
    texts = [[(10, 16, "Edge Dist: 7.87mm | Avg Width: 1.19mm (n_d=29)", 2, (0, 0, 255)), (10, cfg.frame_h - 20, "Stitches: 30 | Fabric: 1", 1, (0, 0, 0))]]
    ann = eng.annotate(d_frame, dets, counts, texts=texts)
    torch.cuda.synchronize()
    img = ann[0].cpu().numpy()
    red = np.all(img[10:40, :700] == np.array((0, 0, 255), np.uint8), axis=2)
    assert 800 < red.sum() < 6000                                     # the text line is there
    # the digit "1" of the font: a full-height column (0x7F) -- rendered at scale 1 into a blank image
    blank = torch.zeros_like(d_frame)
    from vision_textile_inspection_b200._lib import check
    check(eng.lib.vti_draw_text(eng._h, blank.data_ptr(), 0, 5, 5, b"1", 1, 255, 255, 255, None), "vti_draw_text")
    torch.cuda.synchronize()
    g = blank[0, 5:12, 5:10, 0].cpu().numpy() > 0
    assert g[:, 2].all() and g.sum() == 7 + 2 + 1                       # 0x00,0x42,0x7F,0x40,0x00
    jpg = eng.encode_jpeg(ann[0], quality=95)
    assert jpg[:2] == b"\xff\xd8" and jpg[-2:] == b"\xff\xd9" and len(jpg) < img.nbytes // 3
    dec = cv2.imdecode(np.frombuffer(jpg, np.uint8), cv2.IMREAD_COLOR)
    assert dec.shape == img.shape
    psnr = lambda x: 10 * np.log10(255.0 ** 2 / np.mean((x.astype(np.float32) - img.astype(np.float32)) ** 2))
    ok, cvjpg = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, 95])
    cvdec = cv2.imdecode(cvjpg, cv2.IMREAD_COLOR)
    assert psnr(dec) >= psnr(cvdec) - 1.5, (psnr(dec), psnr(cvdec))     # as faithful as cv2.imwrite's own JPEG (noisy texture)
    assert 0.5 < len(jpg) / len(cvjpg) < 2.0                            # same ballpark as cv2.imwrite's output


def test_app_gpu_annotation_and_jpeg(tmp_path):
    import json
    import test_app
    from vision_textile_inspection_b200 import app as A
    cfg = synth.CONFIGS["native"]
    cp, ep = test_app.write_calibration(tmp_path)
    bb = test_app.PlantedBackbone(cfg)
    app = A.StitchMeasurementApp(cp, ep, "best_Model.pt", camera_index=None, calib_w=cfg.frame_w, calib_h=cfg.frame_h,
                                 backbone=bb, roi=cfg.roi(), annotate="gpu", jpeg_quality=90)
    bb.queue.append(10)
    frame = synth.fabric_frame(cfg, 10)
    annotated, m = app.process_frame(frame)
    assert annotated.shape == frame.shape and annotated.dtype == np.uint8 and (annotated != frame).any()
    assert np.all(annotated[cfg.roi()[3], 400] == np.array((144, 238, 144), np.uint8))     # the ROI border
    dec = cv2.imdecode(np.frombuffer(app.last_jpeg, np.uint8), cv2.IMREAD_COLOR)
    assert dec.shape == frame.shape


def test_jpeg_ingest_decodes_like_cv2():
    """camera MJPEG frames decoded by nvJPEG straight into the device buffer K1 reads."""
    cfg = synth.CONFIGS["cfg2"]
    eng = InspectionEngine(EngineConfig.for_workload(cfg, helpers.load_calib(), max_batch=2))
    frames = [synth.fabric_frame(cfg, 70 + i) for i in range(2)]
    for flags, tol_max, tol_mean in (([cv2.IMWRITE_JPEG_QUALITY, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444], 5, 0.5),    # IDCT + colour rounding differ by a few LSB
                                     ([cv2.IMWRITE_JPEG_QUALITY, 90], 40, 2.0)):           # 4:2:0: chroma upsampling filters differ
        jp = [cv2.imencode(".jpg", f, flags)[1].tobytes() for f in frames]
        for one_by_one in (True, False):          # the single-image decoder and the batched one (best backend of the box)
            dev_frames = eng.decode_jpeg_batch(jp, one_by_one=one_by_one)
            torch.cuda.synchronize()
            for b in range(2):
                ref = cv2.imdecode(np.frombuffer(jp[b], np.uint8), cv2.IMREAD_COLOR)
                d = np.abs(dev_frames[b].cpu().numpy().astype(np.int32) - ref.astype(np.int32))
                print(f"jpeg ingest one_by_one={one_by_one} backend={eng.jpeg_backend()} max {d.max()} mean {d.mean():.3f}")
                assert d.max() <= tol_max and d.mean() <= tol_mean, (one_by_one, eng.jpeg_backend(), d.max(), d.mean())
        assert eng.jpeg_backend() in ("nvjpeg-hardware", "nvjpeg-gpu-hybrid", "nvjpeg-hybrid")
        net_in = eng.preprocess(dev_frames)                                 # and K1 takes it from there
        assert net_in.shape == (2, 3, eng.LH, eng.LW)
    with pytest.raises(_lib.VtiError, match="the handle was created for"):
        eng.decode_jpeg_batch([cv2.imencode(".jpg", frames[0][:100, :100])[1].tobytes()] * 2)
