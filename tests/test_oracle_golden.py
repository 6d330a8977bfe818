"""CPU tests that pin the oracle (no GPU): the three restatements agree bit for bit, and the
oracle reproduces the reference's only known answers, the two golden flow plots
(HornSchunckOF/img/resimage/0000{40,50}_10.pnghsbresenhamLineFlow.png; parameters from
HornSchunckOF/main.cpp:94-96,104)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

REF_IMG = "/root/reference/HornSchunckOF/img"
SENTINEL = 7  # a colour plotFlow never draws


def test_bgr2gray_matches_cv2(oracle):
    import cv2
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 41, 3), dtype=np.uint8)
    assert np.array_equal(oracle.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("shape,w,iters,alpha", [
    ((37, 53), 3, 20, 1.0), ((40, 31), 5, 15, 1.0), ((16, 16), 4, 9, 0.5), ((9, 20), 2, 7, 10.0),
    ((5, 5), 7, 5, 1.0), ((1, 7), 3, 3, 1.0), ((2, 3), 5, 4, 1.0), ((1, 1), 3, 2, 1.0), ((12, 9), 1, 4, 2.0),
])
def test_three_restatements_bit_identical(oracle, c_oracle, shape, w, iters, alpha):
    rng = np.random.default_rng(hash((shape, w)) % 2**32)
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    b = rng.integers(0, 256, shape, dtype=np.uint8)
    gx, gy, gt, u, v = oracle.cv_flow(a, b, w, iters, alpha)
    nx, ny, nt, nu, nv = oracle.np_flow(a, b, w, iters, alpha)
    cu, cv_ = c_oracle.flow(a, b, w, iters, alpha)
    cgx, cgy, cgt = c_oracle.gradients(a, b)
    for x, y in ((gx, nx), (gy, ny), (gt, nt), (u, nu), (v, nv), (u, cu), (v, cv_), (gx, cgx), (gy, cgy), (gt, cgt)):
        assert np.array_equal(x, y)


def test_alpha_zero_gives_ieee_specials_like_opencv(oracle):
    a = np.full((6, 6), 9, np.uint8)
    b = a.copy(); b[2, 2] = 12
    _, _, _, u, _ = oracle.np_flow(a, b, 3, 1, 0.0)
    _, _, _, cu, _ = oracle.cv_flow(a, b, 3, 1, 0.0)
    assert np.isnan(u).any() and np.array_equal(np.isnan(u), np.isnan(cu))


def _render(oracle, u, v, shape):
    canvas = np.full(shape + (3,), SENTINEL, np.uint8)
    return oracle.plot_bresenham(canvas, u, v, delta=20, scale=20.0, outlier=5)


@pytest.mark.parametrize("pair", ["000040", "000050"])
def test_oracle_reproduces_golden_plot(oracle, kitti, pair):
    prev, nxt = kitti(pair)
    gold = np.load(os.path.join(GOLDEN, f"plot_{pair}.npz"))
    *_, u, v = oracle.cv_flow(prev, nxt, 5, 100, 1.0)          # main.cpp:94-96
    img = _render(oracle, u, v, prev.shape)
    yx = gold["yx"].astype(np.int64)
    # every pixel the reference drew is drawn by the oracle in the same colour
    assert np.array_equal(img[yx[:, 0], yx[:, 1]], gold["bgr"])
    # pixels the oracle drew beyond that can only be ones whose colour equals the raw frame's
    drawn = (img != SENTINEL).any(axis=2)
    extra = drawn.sum() - len(yx)
    assert 0 <= extra <= 3, extra


@pytest.mark.parametrize("pair", ["000050"])
@pytest.mark.parametrize("w,iters,alpha", [(3, 100, 1.0), (5, 10, 1.0), (5, 100, 10.0)])
def test_golden_plot_discriminates_parameters(oracle, kitti, pair, w, iters, alpha):
    prev, nxt = kitti(pair)
    gold = np.load(os.path.join(GOLDEN, f"plot_{pair}.npz"))
    *_, u, v = oracle.cv_flow(prev, nxt, w, iters, alpha)
    img = _render(oracle, u, v, prev.shape)
    yx = gold["yx"].astype(np.int64)
    wrong = (img[yx[:, 0], yx[:, 1]] != gold["bgr"]).any(axis=1).sum()
    assert wrong > 100


@pytest.mark.skipif(not os.path.isdir(REF_IMG), reason="reference checkout not mounted (GPU box)")
@pytest.mark.parametrize("pair", ["000040", "000050"])
def test_oracle_reproduces_golden_png_pixel_exact(oracle, pair):
    import cv2
    raw_prev = cv2.imread(f"{REF_IMG}/leftimage/{pair}_10.png")
    raw_next = cv2.imread(f"{REF_IMG}/leftimage/{pair}_11.png")
    gold = cv2.imread(f"{REF_IMG}/resimage/{pair}_10.pnghsbresenhamLineFlow.png")
    prev, nxt = oracle.bgr2gray(raw_prev), oracle.bgr2gray(raw_next)
    *_, u, v = oracle.cv_flow(prev, nxt, 5, 100, 1.0)
    img = oracle.plot_bresenham(raw_prev, u, v, 20, 20.0, 5)
    assert int((img != gold).any(axis=2).sum()) == 0
    # and the committed gray fixtures are exactly these frames
    assert np.array_equal(prev, cv2.imread(os.path.join(GOLDEN, f"kitti_{pair}_10_gray.png"), cv2.IMREAD_UNCHANGED))
    assert np.array_equal(nxt, cv2.imread(os.path.join(GOLDEN, f"kitti_{pair}_11_gray.png"), cv2.IMREAD_UNCHANGED))


def test_crop_oracle_is_exact_in_the_interior(oracle):
    """a window with margin r*T+1 reproduces the interior of the full solve exactly."""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, (120, 140), dtype=np.uint8)
    b = rng.integers(0, 256, (120, 140), dtype=np.uint8)
    w, iters = 3, 12
    m = oracle.crop_margin(w, iters)
    *_, u, v = oracle.np_flow(a, b, w, iters, 1.0)
    y0, y1, x0, x1 = 50, 70, 60, 80
    *_, cu, cv_ = oracle.np_flow(a[y0 - m:y1 + m, x0 - m:x1 + m], b[y0 - m:y1 + m, x0 - m:x1 + m], w, iters, 1.0)
    assert np.array_equal(cu[m:-m, m:-m], u[y0:y1, x0:x1])
    assert np.array_equal(cv_[m:-m, m:-m], v[y0:y1, x0:x1])
