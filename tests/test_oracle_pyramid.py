"""Pins the oracle's integer paths: cv::pyrDown against the live cv2 4.13 (third-party oracle),
Simd::AbsGradientSaturatedSum and the grid argmax against independent numpy statements, and against the
committed golden fixtures in tests/golden/ (made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("shape", [(376, 1241), (188, 621), (94, 311), (47, 156), (1080, 1920), (480, 640), (5, 7),
                                   (5, 5), (3, 3), (3, 4), (8, 3), (2, 2)])
def test_pyrdown_matches_cv2(orc, shape):
    rng = np.random.default_rng(shape[0] * 4099 + shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(orc.pyrdown(img), cv2.pyrDown(img))


def test_pyramid_stack_matches_cv2(orc, synth):
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (synth.KITTI_H, synth.KITTI_W), dtype=np.uint8)
    ip, gp = orc.build_pyramid(img, 4)
    ipl = orc.unpack_pyramid(ip, synth.KITTI_W, synth.KITTI_H, 4)
    gpl = orc.unpack_pyramid(gp, synth.KITTI_W, synth.KITTI_H, 4)
    assert [a.shape for a in ipl] == [(376, 1241), (188, 621), (94, 311), (47, 156)]
    cur_i, cur_g = img, synth.abs_gradient_np(img)
    for l in range(4):
        assert np.array_equal(ipl[l], cur_i) and np.array_equal(gpl[l], cur_g), l
        cur_i, cur_g = cv2.pyrDown(cur_i), cv2.pyrDown(cur_g)


@pytest.mark.parametrize("shape", [(376, 1241), (3, 3), (2, 9), (1, 5), (17, 4)])
def test_abs_gradient_matches_numpy(orc, synth, shape):
    rng = np.random.default_rng(shape[0] + 31 * shape[1])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    if shape[0] > 2 and shape[1] > 2:
        img[1, 0], img[1, 2], img[0, 1], img[2, 1] = 0, 255, 0, 255  # 510 saturates to 255
        want = synth.abs_gradient_np(img)
        assert want[1, 1] == 255
    else:
        want = np.zeros(shape, np.uint8)
    assert np.array_equal(orc.abs_gradient(img), want)


@pytest.mark.parametrize("cell,thr", [(30, 50), (20, 50), (30, 0), (7, 200), (400, 10)])
def test_grid_select_matches_numpy(orc, synth, cell, thr):
    rng = np.random.default_rng(cell * 7 + thr)
    grad = (rng.integers(0, 6, (376, 1241)) * 51).astype(np.uint8)  # heavy ties
    grad[:35, :35] = 0
    assert np.array_equal(orc.grid_select(grad, cell, thr), synth.grid_argmax_np(grad, cell, thr))


def test_grid_geometry_and_occupancy(orc):
    # src/feature_selection.cpp:19-25: rows = h/cell + 1, cols = w/cell + 1 -> 13 x 42 = 546 cells at 1241 x 376 / 30
    grad = np.full((376, 1241), 200, np.uint8)
    sel = orc.grid_select(grad, 30, 50)
    assert len(sel) == 546
    assert tuple(sel[0]) == (0, 0, 200) and tuple(sel[41]) == (1230, 0, 200) and tuple(sel[-1]) == (1230, 360, 200)
    occ = np.zeros(546, np.uint8)
    occ[5] = occ[100] = 1
    sel2 = orc.grid_select(grad, 30, 50, occupancy=occ)
    assert len(sel2) == 544 and (150, 0, 200) not in [tuple(s) for s in sel2]
    assert len(orc.grid_select(grad, 30, 200)) == 0  # strict >


def test_golden_fixtures(orc):
    g = np.load(os.path.join(GOLD, "pyramid_golden.npz"))
    img = g["img"]
    ip, gp = orc.build_pyramid(img, 4)
    h, w = img.shape
    for l, (a, b) in enumerate(zip(orc.unpack_pyramid(ip, w, h, 4), orc.unpack_pyramid(gp, w, h, 4))):
        assert np.array_equal(a, g["img_l%d" % l]) and np.array_equal(b, g["grad_l%d" % l])
    assert np.array_equal(orc.grid_select(g["grad_l0"], 30, 50), g["select_30_50"])
