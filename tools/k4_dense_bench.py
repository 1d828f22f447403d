"""Crossover measurement for the two forms of K4 (the mask contraction): the default UNIT form (CUDA cores, organised
around the box crop) against the tcgen05 TILE form (vti_params.k4_dense = 2) on a synthetic DENSE-OVERLAP scene family.

Class-offset NMS never lets boxes of different classes suppress each other, so K detections of K different classes,
each covering a fraction `cover` of the letterboxed frame, all survive: every prototype pixel inside is wanted by
~K x cover detections.  The raw head tensors are planted so that K2 / K3 produce exactly those K boxes; VTI_ALL_DETS=1
makes K3 / K4 compute mask statistics for every kept detection (normally only routed stitch / fabric ones get them).

    VTI_ALL_DETS=1 python tools/k4_dense_bench.py [--frames 8]
Prints one JSON line per (K, cover): K4 time of both forms (CUDA events around the kernel, vti_get_stage_ms), coverage
(summed crop-window cells / plane cells) and whether the per-detection statistics of the two forms agree."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VTI_ALL_DETS", "1")


def dense_head(K, cover, LH=640, LW=640, seed=0, solid=False):
    rng = np.random.default_rng(seed)
    nc = max(K, 2)
    lv = []
    for s in (8, 16, 32):
        a = np.zeros((64 + nc, LH // s, LW // s), np.float32)
        a[64:] = -12.0
        lv.append(a)
    g = LH // 32
    cells = sorted(((y, x) for y in range(g) for x in range(g)), key=lambda p: (p[0] - g / 2 + 0.5) ** 2 + (p[1] - g / 2 + 0.5) ** 2)
    half = 0.5 * np.sqrt(cover) * min(LH, LW)
    for k in range(K):
        gy, gx = cells[k]
        ax, ay = (gx + 0.5) * 32, (gy + 0.5) * 32
        jit = rng.uniform(-8, 8, 4)
        x1, y1 = LW / 2 - half + jit[0], LH / 2 - half + jit[1]
        x2, y2 = LW / 2 + half + jit[2], LH / 2 + half + jit[3]
        for side, d in enumerate(((ax - x1) / 32, (ay - y1) / 32, (x2 - ax) / 32, (y2 - ay) / 32)):
            d = float(np.clip(d, 0.0, 14.99))
            lo = int(np.floor(d))
            fr = d - lo
            w = np.full(16, -20.0, np.float32)
            w[lo] = np.log(max(1 - fr, 1e-6))
            w[lo + 1] = np.log(max(fr, 1e-6))
            lv[2][side * 16:(side + 1) * 16, gy, gx] = w
        lv[2][64 + k, gy, gx] = 6.0
    A = sum(l.shape[1] * l.shape[2] for l in lv)
    coef = rng.normal(0, 1, (32, A)).astype(np.float32)
    ph, pw = LH // 4, LW // 4
    base = rng.normal(0, 1, (32, ph + 8, pw + 8)).astype(np.float32)
    c = np.cumsum(np.cumsum(np.pad(base, ((0, 0), (1, 0), (1, 0))), 1), 2)
    proto = ((c[:, 9:, 9:] - c[:, :-9, 9:] - c[:, 9:, :-9] + c[:, :-9, :-9]) / 9.0)[:, :ph, :pw].astype(np.float32)
    if solid:            # solid masks: logit = +4 + small texture everywhere inside the box (edges only at the crop)
        proto *= 0.05
        proto[0] = 1.0
        coef[0] = 4.0
    return lv, coef, np.ascontiguousarray(proto), nc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--solid", action="store_true", help="solid masks (every pixel inside the box set) instead of random ones")
    ap.add_argument("--only", type=int, default=0, help="run only this K (for ncu captures)")
    a = ap.parse_args()
    import torch
    from vision_textile_inspection_b200.engine import EngineConfig, InspectionEngine
    B = a.frames
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for cover in (0.1, 0.5):
        for K in ((a.only,) if a.only else (4, 8, 16, 32, 64, 128, 250)):
            heads = [dense_head(K, cover, seed=100 * K + i, solid=a.solid) for i in range(min(B, 2))]
            nc = heads[0][3]
            lvs = [dev(np.stack([heads[i % len(heads)][0][l] for i in range(B)])) for l in range(3)]
            coef = dev(np.stack([heads[i % len(heads)][1] for i in range(B)]))
            proto = dev(np.stack([heads[i % len(heads)][2] for i in range(B)]))
            res = {}
            for mode in (0, 2):
                ec = EngineConfig(frame_h=640, frame_w=640, K=np.array([[600.0, 0, 320], [0, 600, 320], [0, 0, 1]]),
                                  dist=np.zeros(5), R=np.eye(3), t=np.array([0, 0, 0.3]), imgsz=640, nc=nc, conf=0.25, iou=0.45,
                                  max_det=300, max_batch=B, roi=(0, 0, 0, 0, 0), k4_dense=mode)
                eng = InspectionEngine(ec)
                outs = eng.alloc_outputs(B)
                for _ in range(3):
                    eng.post_measure(lvs[0], lvs[1], lvs[2], coef, proto, outputs=outs)
                eng.set_profiling(True)
                ts = []
                for _ in range(a.reps):
                    eng.post_measure(lvs[0], lvs[1], lvs[2], coef, proto, outputs=outs)
                    ts.append(eng.stage_ms()[3] * 1e3)
                eng.set_profiling(False)
                torch.cuda.synchronize()
                d = eng.dets_to_numpy(outs[0])
                n = outs[1].cpu().numpy()
                res[mode] = (float(np.median(ts)), d, n)
                eng.close()
            (t0, d0, n0), (t2, d2, n2) = res[0], res[2]
            same = bool(np.array_equal(n0, n2)) and all(
                np.array_equal(d0[b, :n0[b]][key], d2[b, :n0[b]][key]) for b in range(B) for key in ("m00", "m10", "m01", "col_min", "col_max"))
            ndiff = int(sum((d0[b, :n0[b]]["m00"] != d2[b, :n0[b]]["m00"]).sum() for b in range(B)))
            maxrel = float(max((np.abs(d0[b, :n0[b]]["m00"] - d2[b, :n0[b]]["m00"]) / np.maximum(d0[b, :n0[b]]["m00"], 1)).max()
                               for b in range(B)))
            bx = d0[0, :n0[0]]["box_lb"] * 0.25
            win = float(((np.ceil(bx[:, 2]) - np.ceil(bx[:, 0]) + 1).clip(0) * (np.ceil(bx[:, 3]) - np.ceil(bx[:, 1]) + 1).clip(0)).sum())
            print(json.dumps({"K_planted": K, "cover": cover, "kept_per_frame": float(n0.mean()), "frames": B,
                              "plane_covers": round(win / (161 * 161), 2), "k4_unit_us": round(t0, 1), "k4_tcgen05_us": round(t2, 1),
                              "speedup_tcgen05": round(t0 / t2, 2), "statistics_identical": same, "masks": "solid" if a.solid else "random",
                              "dets_with_other_m00": ndiff, "max_rel_m00_diff": maxrel}), flush=True)


if __name__ == "__main__":
    main()
