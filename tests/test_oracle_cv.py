"""Pins oracle/cv_fixed.py (the integer kernel spec) against the real OpenCV, bit-exact."""
import cv2
import numpy as np
import pytest

from oracle import cv_fixed as F

SHAPES = [(640, 640, 960), (960, 1280, 960), (720, 1280, 960), (1080, 1920, 640), (1080, 1920, 960),
          (480, 640, 960), (123, 457, 960), (2160, 3840, 960)]


@pytest.mark.parametrize("h,w,S", SHAPES)
def test_resize_linear_bit_exact(h, w, S):
    rng = np.random.default_rng(h * 7 + w)
    g = F.letterbox_geometry(h, w, S)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = cv2.resize(img, (g["new_w"], g["new_h"]), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ref, F.resize_linear_u8(img, g["new_w"], g["new_h"]))


@pytest.mark.parametrize("h,w,S", SHAPES)
def test_nearest_map_bit_exact(h, w, S):
    g = F.letterbox_geometry(h, w, S)
    m = np.random.default_rng(1).random((g["LH"], g["LW"])).astype(np.float32)
    ref = cv2.resize(m, (w, h), interpolation=cv2.INTER_NEAREST)
    assert np.array_equal(ref, m[F.nearest_map(h, g["LH"])][:, F.nearest_map(w, g["LW"])])


def test_letterbox_geometry_matches_survey_table():
    assert F.letterbox_geometry(960, 1280, 960)["LH"] == 736
    assert F.letterbox_geometry(720, 1280, 960) == dict(new_w=960, new_h=540, top=2, bottom=2, left=0, right=0,
                                                        LH=544, LW=960)
    assert F.letterbox_geometry(1080, 1920, 640)["LH"] == 384
    assert F.letterbox_geometry(640, 640, 640)["LH"] == 640


@pytest.mark.parametrize("h,w", [(960, 1280), (720, 1280), (640, 640), (1080, 1920)])
def test_undistort_bit_exact(h, w, calib):
    K = F.scale_K(np.array(calib["camera_matrix"]), w, h)
    dist = np.array(calib["dist_coeffs"])
    img = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(cv2.undistort(img, K, dist), F.undistort_u8(img, K, dist))


def test_yuyv_to_bgr_bit_exact_over_every_triple():
    """K0's spec (camera-native YUYV ingest) against the real cv2.cvtColor: every (Y, U, V) triple, both
    pixel positions of a chroma pair, plus random frames."""
    u, v = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    img = np.empty((65536, 256, 2), np.uint8)                 # row = (U, V), columns = Y (pairs (2k, 2k+1) share the chroma)
    img[:, :, 0] = np.arange(256, dtype=np.uint8)[None, :]
    img[:, 0::2, 1] = u.reshape(-1, 1)
    img[:, 1::2, 1] = v.reshape(-1, 1)
    assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_YUV2BGR_YUY2), F.yuyv_to_bgr(img))
    rng = np.random.default_rng(5)
    for h, w in ((720, 1280), (3, 2), (17, 6)):
        img = rng.integers(0, 256, (h, w, 2), dtype=np.uint8)
        assert np.array_equal(cv2.cvtColor(img, cv2.COLOR_YUV2BGR_YUY2), F.yuyv_to_bgr(img))


def test_camera_stand_in_round_trip_is_close():
    """bgr_to_yuyv (the test / bench camera stand-in) followed by the cv2 decode gives back the frame up to 4:2:2 chroma
    loss -- a sanity check that the synthetic YUYV frames still look like the scene the head tensors were planted for."""
    from vision_textile_inspection_b200 import synth
    frame = synth.fabric_frame(synth.CONFIGS["cfg2"], 7)
    back = F.yuyv_to_bgr(synth.bgr_to_yuyv(frame))
    d = np.abs(back.astype(np.int32) - frame.astype(np.int32))
    assert d.mean() < 3.0

