"""CPU simulation of the shared-memory wavefronts of K1's remap loads for a given raw-box pitch / footprint pitch /
cell-to-lane mapping, from the exact undistort map of a config (no GPU needed).  One LDS.32 of a warp costs as many
wavefronts as the largest number of DISTINCT words that fall into one bank.
    python tools/k1_bank_sim.py [cfg2]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from vision_textile_inspection_b200 import _lib, synth          # noqa: E402
from vision_textile_inspection_b200.engine import EngineConfig  # noqa: E402


def plan(cfgname, FTX=64, FTY=32):
    lib = _lib.load()
    cfg = synth.CONFIGS[cfgname]
    ec = EngineConfig.for_workload(cfg)
    fh, fw = cfg.frame_h, cfg.frame_w
    g = _lib.VtiGeometry()
    lib.vti_plan_geometry(fh, fw, cfg.imgsz, 32, cfg.max_det, 0, C.byref(g))
    xi = np.zeros(g.new_w, np.int32); a0 = np.zeros(g.new_w, np.int16); a1 = np.zeros(g.new_w, np.int16)
    i0 = np.zeros(g.new_h, np.int32); i1 = np.zeros(g.new_h, np.int32); b0 = np.zeros(g.new_h, np.int16); b1 = np.zeros(g.new_h, np.int16)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    lib.vti_plan_resize_taps_x(fw, g.new_w, P(xi, C.c_int32), P(a0, C.c_int16), P(a1, C.c_int16))
    lib.vti_plan_resize_taps_y(fh, g.new_h, P(i0, C.c_int32), P(i1, C.c_int32), P(b0, C.c_int16), P(b1, C.c_int16))
    ix = np.zeros((fh, fw), np.int32); iy = np.zeros((fh, fw), np.int32)
    K = (C.c_double * 9)(*np.asarray(ec.K).reshape(9)); D = (C.c_double * 5)(*np.asarray(ec.dist).reshape(-1)[:5])
    lib.vti_plan_undistort_map(K, D, fh, fw, P(ix, C.c_int32), P(iy, C.c_int32))
    tiles = []
    for ty in range((g.LH + FTY - 1) // FTY):
        for tx in range((g.LW + FTX - 1) // FTX):
            X0, Y0 = tx * FTX, ty * FTY
            ry_lo, ry_hi = max(Y0 - g.top, 0), min(Y0 + FTY - 1 - g.top, g.new_h - 1)
            rx_lo, rx_hi = max(X0 - g.left, 0), min(X0 + FTX - 1 - g.left, g.new_w - 1)
            if ry_lo > ry_hi or rx_lo > rx_hi:
                continue
            r_lo = min(i0[ry_lo:ry_hi + 1].min(), i1[ry_lo:ry_hi + 1].min())
            r_hi = max(i0[ry_lo:ry_hi + 1].max(), i1[ry_lo:ry_hi + 1].max())
            c_lo = xi[rx_lo:rx_hi + 1].min(); c_hi = min(xi[rx_lo:rx_hi + 1].max() + 1, fw - 1)
            tiles.append((r_lo, r_hi - r_lo + 1, c_lo, c_hi - c_lo + 1))
    return fh, fw, ix, iy, tiles


def wavefronts(words):
    """words: (nwarps, 32) int array of word addresses (-1 = inactive lane) -> wavefronts per warp."""
    out = np.zeros(len(words), np.int32)
    for k, w in enumerate(words):
        w = np.unique(w[w >= 0])
        if len(w):
            out[k] = np.bincount(w % 32, minlength=32).max()
    return out


def sim(cfgname, raw_pitch, pitch_u, mapping):
    fh, fw, ix, iy, tiles = plan(cfgname)
    PU = max(t[3] for t in tiles) if pitch_u is None else pitch_u
    tot_wf = tot_inst = 0
    for (r_lo, nrows, c_lo, _) in tiles:
        c_end = min(c_lo + PU, fw)
        x = np.clip(ix[r_lo:r_lo + nrows, c_lo:c_end] >> 5, -2, fw)
        y = np.clip(iy[r_lo:r_lo + nrows, c_lo:c_end] >> 5, -2, fh)
        bx0 = (x.min() & ~7) if x.min() >= 0 else -8
        by0 = y.min()
        word = np.full((nrows, PU), -1, np.int64)
        word[:, :c_end - c_lo] = (y - by0) * raw_pitch + (x - bx0)
        if mapping == "flat":
            flat = word.reshape(-1)
            n = (len(flat) + 31) // 32 * 32
            w = np.full(n, -1, np.int64); w[:len(flat)] = flat
            warps = w.reshape(-1, 32)
        elif mapping == "b4x8":
            nb_r, nb_c = (nrows + 3) // 4, (PU + 7) // 8
            pad = np.full((nb_r * 4, nb_c * 8), -1, np.int64); pad[:nrows, :PU] = word
            warps = pad.reshape(nb_r, 4, nb_c, 8).transpose(0, 2, 1, 3).reshape(-1, 32)
        elif mapping == "b2x16":
            nb_r, nb_c = (nrows + 1) // 2, (PU + 15) // 16
            pad = np.full((nb_r * 2, nb_c * 16), -1, np.int64); pad[:nrows, :PU] = word
            warps = pad.reshape(nb_r, 2, nb_c, 16).transpose(0, 2, 1, 3).reshape(-1, 32)
        elif mapping == "row32":
            nb_c = (PU + 31) // 32
            pad = np.full((nrows, nb_c * 32), -1, np.int64); pad[:, :PU] = word
            warps = pad.reshape(-1, 32)
        for d in (0, 1, raw_pitch, raw_pitch + 1):          # t00, t01, t10, t11
            ww = np.where(warps >= 0, warps + d, -1)
            tot_wf += wavefronts(ww).sum()
            tot_inst += (warps >= 0).any(axis=1).sum()
    return tot_wf / tot_inst, tot_inst / len(tiles) / 4


if __name__ == "__main__":
    cfgname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    for mapping, pu, rps in (("flat", None, (104, 112, 116, 118, 120, 124, 128)), ("flat", 88, (112, 120, 128)),
                             ("row32", None, (112, 128)), ("b4x8", None, (104, 112, 120, 136)), ("b2x16", None, (112, 144))):
        for rp in rps:
            wf, slots = sim(cfgname, rp, pu, mapping)
            print(f"{cfgname} {mapping:6s} pitch_u={pu} raw_pitch={rp:4d}: {wf:.3f} wavefronts / LDS, {slots:.1f} warp slots / tile")
