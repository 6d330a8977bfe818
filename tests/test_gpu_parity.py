"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(libhs_b200.so via ctypes); the oracle is only the checker.

Tolerances (BASELINE.json north_star): gradients bit-exact; flow max|du|,|dv| <= 1e-4 px and mean
endpoint-error difference <= 1e-5 px against the fp64 oracle.  Invariances between kernel variants
are bit-exact because every variant uses the same canonical arithmetic (hs_kernels.cuh)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

TOL_MAX = 1e-4
TOL_EPE = 1e-5


def rand_pair(shape, seed, jitter=20):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-jitter, jitter + 1, shape), 0, 255).astype(np.uint8)
    return a, b


def assert_flow_close(u, v, ou, ov):
    du, dv = np.abs(u - ou).max(), np.abs(v - ov).max()
    epe = np.abs(np.hypot(u, v) - np.hypot(ou, ov)).mean()
    assert du <= TOL_MAX and dv <= TOL_MAX, (du, dv)
    assert epe <= TOL_EPE, epe


# ---------------------------------------------------------------- gradients (getGradients :19-41)
@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (7, 1), (2, 3), (3, 2), (5, 5), (64, 96), (37, 131), (375, 1242)])
def test_gradients_bit_exact_random(pkg, oracle, shape):
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    a = rng.integers(0, 256, shape, dtype=np.uint8)
    b = rng.integers(0, 256, shape, dtype=np.uint8)
    hs = pkg.hornSchunck(3, 1, 1.0)
    gx, gy, gt = hs.getGradients(a, b)
    ogx, ogy, ogt = oracle.np_gradients(a, b)
    assert gx.dtype == np.float64 and gx.shape == shape
    assert np.array_equal(gx, ogx) and np.array_equal(gy, ogy) and np.array_equal(gt, ogt)
    hs.close()


def test_gradients_extreme_values_fit_the_packed_format(pkg, oracle):
    """|Ix|,|Iy| reach 1020 and |It| 255 on black/white steps: the 11/11/10-bit packing must hold them."""
    a = np.zeros((16, 16), np.uint8); a[:, 8:] = 255; a[8:, :] = 255 - a[8:, :]
    b = 255 - a
    hs = pkg.hornSchunck(3, 1, 1.0)
    g = hs.getGradients(a, b)
    o = oracle.np_gradients(a, b)
    assert max(np.abs(o[0]).max(), np.abs(o[1]).max()) == 1020 and np.abs(o[2]).max() == 255
    assert all(np.array_equal(x, y) for x, y in zip(g, o))
    hs.close()


@pytest.mark.parametrize("pair", ["000040", "000050"])
def test_gradients_bit_exact_kitti(pkg, oracle, kitti, pair):
    a, b = kitti(pair)
    hs = pkg.hornSchunck(5, 1, 1.0)
    g = hs.getGradients(a, b)
    o = oracle.cv_gradients(a, b)
    assert all(np.array_equal(x, y) for x, y in zip(g, o))
    hs.close()


# ---------------------------------------------------------------- flow (getFlow :43-75)
@pytest.mark.parametrize("w", [1, 2, 3, 4, 5, 6, 7, 8, 9, 11])
@pytest.mark.parametrize("alpha", [0.5, 1.0, 10.0])
@pytest.mark.parametrize("iters", [1, 2, 7, 100])
def test_flow_matches_oracle(pkg, c_oracle, w, alpha, iters):
    a, b = rand_pair((97, 150), seed=w * 100 + iters)
    ou, ov = c_oracle.flow(a, b, w, iters, alpha)
    hs = pkg.hornSchunck(w, iters, alpha)
    u, v = hs.getFlow(a, b)
    hs.close()
    assert u.dtype == np.float64 and u.shape == a.shape
    assert_flow_close(u, v, ou, ov)


def test_zero_iterations_gives_zeros(pkg):
    a, b = rand_pair((20, 30), 1)
    hs = pkg.hornSchunck(3, 0, 1.0)
    u, v = hs.getFlow(a, b)
    hs.close()
    assert not u.any() and not v.any()


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (7, 1), (2, 3), (3, 2), (5, 5), (47, 129), (49, 121), (96, 240)])
@pytest.mark.parametrize("w", [3, 5])
def test_flow_edge_shapes(pkg, oracle, shape, w):
    a, b = rand_pair(shape, seed=shape[0] * 31 + shape[1])
    *_, ou, ov = oracle.np_flow(a, b, w, 9, 1.0)
    hs = pkg.hornSchunck(w, 9, 1.0)
    u, v = hs.getFlow(a, b)
    hs.close()
    assert_flow_close(u, v, ou, ov)


@pytest.mark.parametrize("pair,w,iters", [("000050", 5, 100), ("000040", 5, 100), ("000050", 3, 300)])
def test_flow_kitti_authors_run(pkg, oracle, kitti, pair, w, iters):
    """The reference author's own configuration (main.cpp:42-43,94-96) on the bundled frames."""
    a, b = kitti(pair)
    *_, ou, ov = oracle.cv_flow(a, b, w, iters, 1.0)
    hs = pkg.hornSchunck(w, iters, 1.0)
    u, v = hs.getFlow(a, b)
    hs.close()
    assert np.abs(ou).max() > 50          # real-image flows are large: the hard case for fp32
    assert_flow_close(u, v, ou, ov)
    if w == 5 and iters == 100:
        # ... and the GPU result draws the reference's golden plot (plotFlow.cpp restated in the oracle)
        gold = np.load(os.path.join(GOLDEN, f"plot_{pair}.npz"))
        canvas = np.full(a.shape + (3,), 7, np.uint8)
        img = oracle.plot_bresenham(canvas, u, v, 20, 20.0, 5)
        yx = gold["yx"].astype(np.int64)
        wrong = int((img[yx[:, 0], yx[:, 1]] != gold["bgr"]).any(axis=1).sum())
        assert wrong <= 20, wrong           # a (int) truncation may flip for a value within 1e-4 of k/20


def test_alpha_zero_nan_pattern_matches_reference(pkg, oracle):
    a = np.full((12, 12), 9, np.uint8)
    b = a.copy(); b[3, 4] = 12; a[8, 8] = 200
    *_, ou, ov = oracle.np_flow(a, b, 3, 1, 0.0)
    hs = pkg.hornSchunck(3, 1, 0.0)
    u, v = hs.getFlow(a, b)
    hs.close()
    assert np.isnan(ou).any()
    assert np.array_equal(np.isnan(u), np.isnan(ou)) and np.array_equal(np.isnan(v), np.isnan(ov))
    m = ~np.isnan(ou)
    assert np.allclose(u[m], ou[m], atol=1e-4) and np.allclose(v[m], ov[m], atol=1e-4)


@pytest.mark.parametrize("shape", [(33, 47), (64, 96), (120, 161), (375, 1242)])
def test_solve_bgr_equals_preprocess_then_solve(pkg, oracle, shape):
    """main.cpp's preprocess() (BGR2GRAY) fused in front of the path, on the device."""
    import cv2
    rng = np.random.default_rng(shape[1])
    a = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-12, 13, a.shape), 0, 255).astype(np.uint8)
    ga, gb = cv2.cvtColor(a, cv2.COLOR_BGR2GRAY), cv2.cvtColor(b, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(ga, oracle.bgr2gray(a))
    with pkg.Solver(shape[1], shape[0], 3, 9, 1.0) as s:
        u, v = s.solve_bgr(a, b)
        gu, gv = s.solve(ga, gb)
        u1, _ = pkg.Solver(shape[1], shape[0], 3, 1, 1.0).solve_bgr(a, b)   # one sweep: u = -Ix*It*inv, gray-exact
        g1, _ = pkg.Solver(shape[1], shape[0], 3, 1, 1.0).solve(ga, gb)
    assert np.array_equal(u, gu) and np.array_equal(v, gv) and np.array_equal(u1, g1)
    *_, ou, ov = oracle.np_flow(ga, gb, 3, 9, 1.0)
    assert_flow_close(u, v, ou, ov)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_video_stream_equals_pairwise_solves(pkg, dtype):
    """frame sequence front-end (main.cpp:53-59); flows arrive one pair late, in order."""
    from cpp_optical_flow_b200 import synth
    frames = [synth.frame_pair(150, 260, seed=3, shift=(0.3 * i, -0.2 * i))[1] for i in range(6)]
    with pkg.Solver(260, 150, 3, 25, 1.0) as ref:
        want = [ref.solve(frames[i], frames[i + 1], dtype) for i in range(5)]
    got = {}
    with pkg.Solver(260, 150, 3, 25, 1.0) as s:
        for rep in range(2):                                  # a second sequence after reset
            s.video_reset()
            got.clear()
            order = []
            for f in frames:
                r = s.video_push(f, dtype)
                if r is not None:
                    got[r[0]] = (r[1], r[2]); order.append(r[0])
            r = s.video_flush(dtype)
            got[r[0]] = (r[1], r[2]); order.append(r[0])
            assert order == [0, 1, 2, 3, 4] and s.video_flush(dtype) is None
            for i in range(5):
                assert np.array_equal(got[i][0], want[i][0]) and np.array_equal(got[i][1], want[i][1])
        u, v = s.solve(frames[0], frames[1], dtype)           # the plain path still works afterwards
        assert np.array_equal(u, want[0][0])


def test_grid_sampler_feeds_the_reference_plot(pkg, oracle, kitti):
    """plotFlow reads only every 20th row/column (plotFlow.cpp:70-75).  The device-side
    sampler returns exactly those values, and they draw the reference's golden plot."""
    a, b = kitti("000050")
    with pkg.Solver(a.shape[1], a.shape[0], 5, 100, 1.0) as s:
        s.upload(a, b); s.solve_device()
        gu, gv = s.sample_grid(20)
        u, v = s.download(np.float64)
    assert gu.shape == (19, 63)
    assert np.array_equal(gu, u[::20, ::20]) and np.array_equal(gv, v[::20, ::20])
    # scatter the samples into otherwise empty fields: the plot only looks at the grid
    su = np.zeros_like(u); sv = np.zeros_like(v)
    su[::20, ::20] = gu; sv[::20, ::20] = gv
    gold = np.load(os.path.join(GOLDEN, "plot_000050.npz"))
    img = oracle.plot_bresenham(np.full(a.shape + (3,), 7, np.uint8), su, sv, 20, 20.0, 5)
    yx = gold["yx"].astype(np.int64)
    assert int((img[yx[:, 0], yx[:, 1]] != gold["bgr"]).any(axis=1).sum()) <= 20


@pytest.mark.parametrize("shape,iters,k", [((64, 96), 1, 0), ((97, 150), 40, 0), ((200, 333), 25, 4), ((375, 1242), 60, 6)])
def test_textbook_mode_matches_its_own_oracle(pkg, oracle, shape, iters, k):
    """SURVEY 8f row 4 (NOT a parity item): cube gradients + 1/6-1/12 weighted average, the scheme
    BASELINE.json's prose describes.  Checked against oracle.np_flow_textbook and fused == generic."""
    from cpp_optical_flow_b200 import hs_ctypes as H, synth
    a, b = synth.frame_pair(shape[0], shape[1], seed=shape[0])
    ogx, ogy, ogt, ou, ov = oracle.np_flow_textbook(a, b, iters, 1.0)
    with pkg.Solver(shape[1], shape[0], 3, iters, 1.0, temporal_k=k, flags=H.FLAG_TEXTBOOK) as s:
        gx, gy, gt = s.gradients(a, b)
        u, v = s.solve(a, b, np.float64)
        assert s.timing().kernel_id == 1
    assert np.array_equal(gx, ogx) and np.array_equal(gy, ogy) and np.array_equal(gt, ogt)   # multiples of 1/4: exact
    assert_flow_close(u, v, ou, ov)
    with pkg.Solver(shape[1], shape[0], 3, iters, 1.0, flags=H.FLAG_TEXTBOOK | H.FLAG_FORCE_GENERIC) as s:
        gu, gv = s.solve(a, b, np.float64)
    assert np.array_equal(u, gu) and np.array_equal(v, gv)
    # a pure translation by (1.0, 0.5) px: the textbook gradients are in pixel units (no factor 8)
    if iters >= 40:
        assert abs(np.median(u) - 1.0) < 0.35 and abs(np.median(v) - 0.5) < 0.35


def test_early_exit_stops_on_the_residual(pkg, oracle):
    """Contract extension: residual-based early exit.  The residual equals the oracle's, the stop
    happens at the first check that meets the tolerance, and the state is that of a fixed-T solve."""
    from cpp_optical_flow_b200 import synth
    a, b = synth.frame_pair(120, 200, seed=4)
    with pkg.Solver(200, 120, 3, 400, 1.0) as s:
        s.upload(a, b); s.prepare()
        done, res = s.iterate_until(400, 2e-4, check_every=20)
        u, v = s.download(np.float32)
        assert 20 <= done < 400 and done % 20 == 0 and res <= 2e-4
        s.prepare(); s.iterate(done); u2, v2 = s.download(np.float32)
        assert np.array_equal(u, u2) and np.array_equal(v, v2)
        s.prepare()
        d2, r2 = s.iterate_until(60, 0.0, check_every=25)       # tolerance never met: runs to the cap
        assert d2 == 60 and r2 > 0
    # the oracle's residual history: first multiple of 20 with residual <= tol is where we stopped
    *_, pu, pv = oracle.np_flow(a, b, 3, done - 1, 1.0)
    *_, cu, cv_ = oracle.np_flow(a, b, 3, done, 1.0)
    assert abs(max(np.abs(cu - pu).max(), np.abs(cv_ - pv).max()) - res) < 2e-5
    *_, pu, pv = oracle.np_flow(a, b, 3, done - 21, 1.0)
    *_, cu, cv_ = oracle.np_flow(a, b, 3, done - 20, 1.0)
    assert max(np.abs(cu - pu).max(), np.abs(cv_ - pv).max()) > 2e-4 - 2e-5


def test_textbook_mode_needs_window_3(pkg):
    from cpp_optical_flow_b200 import hs_ctypes as H
    with pytest.raises(H.HsError):
        pkg.Solver(32, 32, 5, 3, 1.0, flags=H.FLAG_TEXTBOOK)


# ---------------------------------------------------------------- invariances (bit-exact)
@pytest.mark.parametrize("w,k", [(3, 1), (3, 2), (3, 3), (3, 4), (3, 7), (3, 12), (5, 1), (5, 2), (5, 3), (5, 5),
                                 (2, 4), (2, 9), (4, 2), (4, 3), (3, 10), (4, 5), (2, 12),
                                 (6, 1), (6, 2), (6, 3), (7, 1), (7, 2), (7, 3), (8, 1), (8, 2), (8, 3), (9, 1), (9, 2),
                                 # deepest k per window: beyond what the dataflow launch's 3x3 dependency rule
                                 # covers, so these must fall back to chained single-phase launches
                                 (3, 15), (3, 18), (4, 10), (4, 12), (5, 8), (5, 9), (7, 5), (9, 4)])
def test_fused_kernel_equals_generic_sweep(pkg, w, k):
    from cpp_optical_flow_b200 import hs_ctypes as H
    a, b = rand_pair((333, 517), seed=w * 10 + k)
    iters = 2 * k + 1                      # full launches plus a remainder launch
    with pkg.Solver(517, 333, w, iters, 1.0, flags=H.FLAG_FORCE_GENERIC) as s:
        gu, gv = s.solve(a, b, np.float32)
        assert s.timing().kernel_id == 0
    with pkg.Solver(517, 333, w, iters, 1.0, temporal_k=k) as s:
        tu, tv = s.solve(a, b, np.float32)
        assert s.timing().kernel_id == 1 and s.timing().temporal_k == k
    assert np.array_equal(gu, tu) and np.array_equal(gv, tv)


def test_created_context_uses_the_documented_default_k(pkg):
    """hs_create picks exactly what hs_default_temporal_k (the host-only rule DESIGN.md documents) returns for the
    device's SM count."""
    import ctypes
    import torch
    from cpp_optical_flow_b200 import hs_ctypes as H
    lib = pkg.load_library()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for (wid, hgt, w, T) in [(1920, 1080, 3, 1000), (1920, 1080, 5, 1000), (1242, 375, 5, 100), (1242, 375, 3, 1000),
                             (640, 480, 3, 50), (1280, 720, 4, 200), (3840, 2160, 2, 10), (800, 600, 7, 30)]:
        cfg = H.HsConfig(struct_size=ctypes.sizeof(H.HsConfig), width=wid, height=hgt, window_size=w, max_iterations=T, alpha=1.0)
        want = lib.hs_default_temporal_k(ctypes.byref(cfg), sms)
        with pkg.Solver(wid, hgt, w, T, 1.0) as s:
            a, b = rand_pair((hgt, wid), seed=w)
            s.upload(a, b); s.prepare(); s.iterate(1); s.sync()
            assert s.timing().temporal_k == want, (wid, hgt, w, T, s.timing().temporal_k, want)


@pytest.mark.parametrize("shape", [(40, 64), (375, 1242), (240, 320), (480, 640), (720, 1280)])
@pytest.mark.parametrize("w", [3, 5])
def test_default_k_on_small_frames(pkg, shape, w):
    """Small frames pick their own k from a launch-cost model (hs_create); whatever it picks, the
    result is the generic sweep's, bit for bit."""
    from cpp_optical_flow_b200 import hs_ctypes as H
    a, b = rand_pair(shape, seed=shape[0] + w)
    with pkg.Solver(shape[1], shape[0], w, 37, 1.0, flags=H.FLAG_FORCE_GENERIC) as s:
        gu, gv = s.solve(a, b, np.float32)
    with pkg.Solver(shape[1], shape[0], w, 37, 1.0) as s:
        tu, tv = s.solve(a, b, np.float32)
        assert s.timing().kernel_id == 1 and 1 <= s.timing().temporal_k <= 12
    assert np.array_equal(gu, tu) and np.array_equal(gv, tv)


def test_fuzz_fused_equals_generic(pkg):
    """60 random geometries / windows / k / batch sizes / launch modes: fused kernel == generic sweep,
    bit for bit (any halo, border, alignment or scheduling bug shows up here)."""
    from cpp_optical_flow_b200 import hs_ctypes as H
    rng = np.random.default_rng(int(os.environ.get("HS_FUZZ_SEED", 20260118)))
    for case in range(int(os.environ.get("HS_FUZZ_CASES", 60))):      # a longer soak: HS_FUZZ_CASES=600 HS_FUZZ_SEED=...
        w = int(rng.choice([2, 3, 3, 4, 5, 5, 6, 7, 9]))
        hgt = int(rng.integers(1, 260)); wid = int(rng.integers(1, 420))
        if case % 7 == 0:
            hgt, wid = int(rng.integers(300, 700)), int(rng.integers(300, 900))    # several rounds of tiles
        batch = int(rng.choice([1, 1, 1, 2, 3]))
        k = int(rng.integers(1, 9))
        iters = int(rng.integers(1, 4 * k + 3))
        flags = int(rng.choice([0, 0, H.FLAG_SINGLE_PHASE]))
        alpha = float(rng.choice([0.5, 1.0, 3.0]))
        a = rng.integers(0, 256, (batch, hgt, wid), dtype=np.uint8)
        b = np.clip(a.astype(int) + rng.integers(-25, 26, a.shape), 0, 255).astype(np.uint8)
        if batch == 1:
            a, b = a[0], b[0]
        with pkg.Solver(wid, hgt, w, iters, alpha, batch=batch, flags=H.FLAG_FORCE_GENERIC) as s:
            gu, gv = s.solve(a, b, np.float32)
        with pkg.Solver(wid, hgt, w, iters, alpha, batch=batch, temporal_k=k, flags=flags) as s:
            tu, tv = s.solve(a, b, np.float32)
        assert np.array_equal(gu, tu) and np.array_equal(gv, tv), (case, w, hgt, wid, batch, k, iters, flags)


def test_dataflow_launch_is_race_free_under_repetition(pkg):
    """The multi-phase launch synchronises tiles through per-tile counters, fences and TMA reads of
    data written by other SMs.  A memory-ordering bug would be intermittent: 60 whole solves
    (167 phases x 540 tiles each) must all produce the same bits as the generic sweep."""
    from cpp_optical_flow_b200 import hs_ctypes as H, synth
    a, b = synth.frame_pair(1080, 1920, seed=9)
    with pkg.Solver(1920, 1080, 3, 1000, 1.0, flags=H.FLAG_FORCE_GENERIC) as s:
        gu, gv = s.solve(a, b, np.float32)
    with pkg.Solver(1920, 1080, 3, 1000, 1.0) as s:
        s.upload(a, b)
        for rep in range(60):
            s.solve_device()
            u, v = s.download(np.float32)
            assert np.array_equal(u, gu) and np.array_equal(v, gv), rep


def test_iterate_is_a_semigroup_and_deterministic(pkg):
    a, b = rand_pair((211, 390), 5)
    with pkg.Solver(390, 211, 3, 50, 1.0) as s:
        s.upload(a, b); s.prepare(); s.iterate(50); u1, v1 = s.download(np.float32)
        s.prepare(); s.iterate(13); s.iterate(4); s.iterate(33); u2, v2 = s.download(np.float32)
        s.solve_device(); u3, v3 = s.download(np.float32)
    assert np.array_equal(u1, u2) and np.array_equal(v1, v2) and np.array_equal(u1, u3) and np.array_equal(v1, v3)


def test_batch_equals_single_solves(pkg):
    pairs = [rand_pair((120, 200), 40 + i) for i in range(3)]
    prev = np.stack([p[0] for p in pairs]); nxt = np.stack([p[1] for p in pairs])
    with pkg.Solver(200, 120, 3, 21, 1.0, batch=3) as s:
        bu, bv = s.solve(prev, nxt, np.float32)
    for i, (a, b) in enumerate(pairs):
        with pkg.Solver(200, 120, 3, 21, 1.0) as s:
            u, v = s.solve(a, b, np.float32)
        assert np.array_equal(bu[i], u) and np.array_equal(bv[i], v)


def test_async_solves_on_a_shared_stream_equal_synchronous_solves(pkg):
    """hs_solve_async / hs_solve_wait: two contexts on ONE compute stream, two calls in flight, several rounds;
    every flow equals the synchronous hs_solve of the same pair bit for bit, and a second call on a busy context
    is refused."""
    import torch
    from cpp_optical_flow_b200 import hs_ctypes as H
    Hh, Ww = 333, 517                                         # rows of 517 bytes: the flat-upload path too
    pairs = [rand_pair((Hh, Ww), 70 + i) for i in range(5)]
    want = []
    with pkg.Solver(Ww, Hh, 3, 29, 1.0) as s:
        for a, b in pairs:
            want.append(s.solve(a, b, np.float64))
    stream = torch.cuda.Stream()
    hp = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a, _ in pairs]
    hn = [torch.from_numpy(np.ascontiguousarray(b)).pin_memory() for _, b in pairs]
    outs = [(torch.empty((Hh, Ww), dtype=torch.float64).pin_memory(), torch.empty((Hh, Ww), dtype=torch.float64).pin_memory())
            for _ in range(2)]
    ctxs = [pkg.Solver(Ww, Hh, 3, 29, 1.0, stream=stream.cuda_stream) for _ in range(2)]

    def enqueue(i):
        ou, ov = outs[i & 1]
        ctxs[i & 1].solve_async_raw(hp[i].data_ptr(), hn[i].data_ptr(), Ww, 0, ou.data_ptr(), ov.data_ptr(), Ww * 8, 0, H.HS_F64)

    enqueue(0)
    with pytest.raises(H.HsError):
        ctxs[0].solve_async_raw(hp[0].data_ptr(), hn[0].data_ptr(), Ww, 0, outs[0][0].data_ptr(), outs[0][1].data_ptr(), Ww * 8, 0, H.HS_F64)
    with pytest.raises(H.HsError):                            # ... and so is any other use of its buffers
        ctxs[0].solve(pairs[0][0], pairs[0][1], np.float32)
    for i in range(len(pairs)):
        if i + 1 < len(pairs):
            enqueue(i + 1)
        ctxs[i & 1].solve_wait()
        ou, ov = outs[i & 1]
        assert np.array_equal(ou.numpy(), want[i][0]) and np.array_equal(ov.numpy(), want[i][1]), i
    ctxs[0].solve_wait()                                      # nothing pending: a no-op
    for c in ctxs:
        c.close()


def test_strided_inputs(pkg):
    big_a, big_b = rand_pair((90, 300), 9)
    a, b = big_a[:, 10:210], big_b[:, 10:210]            # row stride 300, width 200
    assert not a.flags.c_contiguous
    hs = pkg.hornSchunck(3, 11, 1.0)
    u1, v1 = hs.getFlow(a, b)
    u2, v2 = hs.getFlow(np.ascontiguousarray(a), np.ascontiguousarray(b))
    hs.close()
    assert np.array_equal(u1, u2) and np.array_equal(v1, v2)


def test_f32_and_f64_outputs_agree(pkg):
    a, b = rand_pair((64, 80), 2)
    with pkg.Solver(80, 64, 5, 15, 1.0) as s:
        u64, v64 = s.solve(a, b, np.float64)
        u32, v32 = s.solve(a, b, np.float32)
    assert np.array_equal(u64, u32.astype(np.float64)) and np.array_equal(v64, v32.astype(np.float64))


# ---------------------------------------------------------------- row slabs (one context per slab)
@pytest.mark.parametrize("w,k,nslab,depth", [(3, 4, 2, 1), (3, 3, 3, 1), (5, 2, 2, 1), (4, 2, 3, 1), (3, 1, 4, 1),
                                             (3, 4, 2, 2), (3, 2, 3, 3), (5, 2, 2, 2), (4, 2, 2, 2)])
def test_row_slabs_equal_single_solve(pkg, w, k, nslab, depth):
    """N slab contexts on one GPU with host-side halo copies == one whole-image solve, bit for bit."""
    from cpp_optical_flow_b200 import slab
    a, b = rand_pair((260, 300), seed=w + k)
    iters = 3 * k + 2
    with pkg.Solver(300, 260, w, iters, 1.0, temporal_k=k) as s:
        u, v = s.solve(a, b, np.float32)
    su, sv = slab.solve_slabs_single_process(a, b, w, iters, 1.0, nslab, temporal_k=k, depth=depth)
    assert np.array_equal(u, su) and np.array_equal(v, sv)


@pytest.mark.parametrize("w,k,nslab,shape,iters", [
    (3, 4, 2, (260, 300), 14), (3, 6, 3, (400, 517), 31), (5, 3, 2, (300, 260), 11), (5, 2, 4, (333, 700), 9),
    (4, 2, 3, (270, 200), 7), (2, 4, 2, (200, 333), 13), (7, 2, 2, (280, 300), 5), (3, 4, 4, (1080, 1920), 60),
    (3, 0, 2, (1080, 1920), 100), (5, 0, 3, (900, 1600), 33), (3, 1, 2, (97, 150), 5)])
def test_row_slab_group_in_kernel_exchange_equals_single_solve(pkg, w, k, nslab, shape, iters):
    """ONE context over N row slabs (hs_config.num_devices, HS_DECOMP_ROW_SLAB) with the in-kernel halo
    exchange: seam tiles store straight into the neighbour's halo rows and signal per-tile flags.  Listing
    the same device N times runs the N slabs in one cooperative launch, so the whole seam protocol (peer
    stores, system-scope release/acquire flags, reversed tile order of odd slabs) runs on this one GPU.
    Bit-identical to the plain solve, also for a second solve and for split hs_iterate calls."""
    from cpp_optical_flow_b200 import hs_ctypes as H
    a, b = rand_pair(shape, seed=w * 7 + nslab)
    with pkg.Solver(shape[1], shape[0], w, iters, 1.0, temporal_k=k) as s:
        u, v = s.solve(a, b, np.float32)
        kk = s.timing().temporal_k
    with pkg.Solver(shape[1], shape[0], w, iters, 1.0, temporal_k=k, devices=[0] * nslab,
                    decomposition=H.DECOMP_ROW_SLAB) as g:
        gu, gv = g.solve(a, b, np.float32)
        assert g.timing().kernel_id == 1
        assert np.array_equal(u, gu) and np.array_equal(v, gv), (np.abs(u - gu).max(), np.argwhere(u != gu)[:4])
        gu2, gv2 = g.solve(b, a, np.float32)                   # the context is reusable
        gu3, gv3 = g.solve(a, b, np.float64)
        assert np.array_equal(gu3, u.astype(np.float64)) and np.array_equal(gv3, v.astype(np.float64))
        g.upload(a, b); g.prepare()                           # seam flags are absolute phase counts: split calls
        g.iterate(iters // 3); g.iterate(iters - iters // 3)
        su, sv = g.download(np.float32)
        if g.timing().temporal_k == kk:
            assert np.array_equal(u, su) and np.array_equal(v, sv)
        else:
            assert np.abs(u - su).max() < 1e-4
    with pkg.Solver(shape[1], shape[0], w, iters, 1.0, temporal_k=k) as s:
        u2, v2 = s.solve(b, a, np.float32)
    assert np.array_equal(u2, gu2) and np.array_equal(v2, gv2)


def _build_c(tmp_path, name):
    import subprocess
    from conftest import ROOT
    libdir = os.path.join(ROOT, "cpp-optical-flow_b200")
    exe = str(tmp_path / name)
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-O1", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", name + ".c"), "-o", exe, "-L", libdir, "-l:libhs_b200.so",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.parametrize("n", [2, 3, 4])
def test_plain_c_caller_drives_row_slabs_through_the_abi(pkg, tmp_path, n):
    """tests/c/slab_smoke.c: hs_config.num_devices / device_ids / decomposition from pedantic C99; the n slabs
    share device 0 here (on a multi-GPU box run it with `distinct`, tools/gpu_multi.sh does)."""
    import subprocess
    pkg.load_library()
    r = subprocess.run([_build_c(tmp_path, "slab_smoke"), str(n), "same"], capture_output=True, text=True)
    assert r.returncode == 0 and "bit-identical" in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_slab_rank_contexts_plan_export_and_connect(pkg):
    """One process per GPU API (slab_world / slab_rank + hs_slab_export / hs_slab_connect): plan, handles and
    wiring.  (Running two connected slabs as separate launches on ONE GPU is not allowed - kernels that wait
    on each other must be co-resident - so the sweeps of this path are covered by the group context above
    and, on real multi-GPU boxes, by bench.py's slab_bit_identical.)"""
    from cpp_optical_flow_b200 import hs_ctypes as H
    ranks = [pkg.Solver(300, 1000, 5, 30, 1.0, temporal_k=3, slab=(r, 3)) for r in range(3)]
    try:
        infos = [s.slab_info() for s in ranks]
        assert [(i.own_begin, i.own_end) for i in infos] == [(0, 334), (334, 668), (668, 1000)]
        assert [(i.buf_begin, i.buf_end) for i in infos] == [(0, 340), (328, 674), (662, 1000)]
        assert [(i.frame_begin, i.frame_end) for i in infos] == [(0, 341), (327, 675), (661, 1000)]
        assert all(i.temporal_k == 3 for i in infos) and (infos[1].halo_top, infos[1].halo_bottom) == (6, 6)
        handles = [s.slab_export() for s in ranks]
        assert all(len(h) == 256 for h in handles)
        with pytest.raises(H.HsError):                        # a middle slab needs both neighbours
            ranks[1].slab_connect(handles[0], None)
        with pytest.raises(H.HsError):                        # wrong neighbour
            ranks[1].slab_connect(handles[2], handles[0])
        ranks[0].slab_connect(None, handles[1])
        ranks[1].slab_connect(handles[0], handles[2])
        ranks[2].slab_connect(handles[1], None)
        with pytest.raises(H.HsError):
            ranks[0].slab_connect(None, handles[1])           # already connected
        a, b = rand_pair((1000, 300), 3)
        ranks[1].upload(a[327:675], b[327:675])
        with pytest.raises(ValueError):
            ranks[1].upload(a, b)                             # wants exactly its frame rows
    finally:
        for s in ranks:
            s.close()


def test_multi_device_context_argument_errors(pkg):
    from cpp_optical_flow_b200 import hs_ctypes as H
    with pytest.raises(H.HsError):                            # slabs thinner than their halo
        pkg.Solver(64, 20, 3, 5, 1.0, temporal_k=6, devices=[0, 0, 0, 0], decomposition=H.DECOMP_ROW_SLAB)
    with pytest.raises(H.HsError):                            # the generic kernel has no slab path
        pkg.Solver(64, 200, 11, 5, 1.0, devices=[0, 0], decomposition=H.DECOMP_ROW_SLAB)
    with pytest.raises(H.HsError):                            # batch split needs distinct devices
        pkg.Solver(64, 64, 3, 5, 1.0, batch=2, devices=[0, 0], decomposition=H.DECOMP_BATCH)
    with pkg.Solver(200, 260, 3, 5, 1.0, devices=[0, 0], decomposition=H.DECOMP_ROW_SLAB) as g:
        with pytest.raises(H.HsError) as e:
            g.iterate_rows(1, 0, 10, True)
        assert e.value.status == 4


def test_partial_row_launches_compose_to_a_full_launch(pkg):
    """hs_iterate_rows (the overlap helper of the row-slab path): strips + interior == one launch."""
    a, b = rand_pair((300, 260), 77)
    with pkg.Solver(260, 300, 3, 12, 1.0, temporal_k=4) as s:
        s.upload(a, b); s.prepare(); s.iterate(12); u, v = s.download(np.float32)
        s.prepare()
        for _ in range(3):
            s.iterate_rows(4, 0, 7, False)
            s.iterate_rows(4, 280, 300, False)
            s.iterate_rows(4, 7, 280, True)
        u2, v2 = s.download(np.float32)
    assert np.array_equal(u, u2) and np.array_equal(v, v2)


# ---------------------------------------------------------------- BASELINE configs at full size
@pytest.mark.parametrize("w", [3, 5])
def test_config2_1080p_full_size(pkg, c_oracle, w):
    """configs[1]: 1920x1080, alpha=1, 1000 sweeps.  Oracle at full size for 60 sweeps (seconds on one
    core); at 1000 sweeps size-independent properties: fused == generic bit-exact, known translation."""
    from cpp_optical_flow_b200 import hs_ctypes as H, synth
    a, b = synth.frame_pair(1080, 1920)
    ou, ov = c_oracle.flow(a, b, w, 60, 1.0)
    with pkg.Solver(1920, 1080, w, 60, 1.0) as s:
        u, v = s.solve(a, b, np.float64)
    assert_flow_close(u, v, ou, ov)
    with pkg.Solver(1920, 1080, w, 1000, 1.0) as s:
        u, v = s.solve(a, b, np.float32)
    with pkg.Solver(1920, 1080, w, 1000, 1.0, flags=H.FLAG_FORCE_GENERIC) as s:
        gu, gv = s.solve(a, b, np.float32)
    assert np.array_equal(u, gu) and np.array_equal(v, gv)
    # the pair is a pure translation by (1.0, 0.5) px; flow is in units of px/8 (unnormalised Sobel)
    assert abs(np.median(u) - 1.0 / 8) < 0.01 and abs(np.median(v) - 0.5 / 8) < 0.01


def test_config3_4k_crop_oracle(pkg, oracle):
    """configs[2] size (3840x2160): a crop oracle with margin r*T+1 must match the interior exactly enough."""
    from cpp_optical_flow_b200 import synth
    a, b = synth.frame_pair(2160, 3840)
    w, iters = 3, 40
    with pkg.Solver(3840, 2160, w, iters, 1.0) as s:
        u, v = s.solve(a, b, np.float64)
    m = oracle.crop_margin(w, iters)
    for (y0, x0) in ((0, 0), (1000, 1900), (2160 - 64, 3840 - 64)):
        ya, yb, xa, xb = max(0, y0 - m), min(2160, y0 + 64 + m), max(0, x0 - m), min(3840, x0 + 64 + m)
        *_, cu, cv_ = oracle.cv_flow(a[ya:yb, xa:xb], b[ya:yb, xa:xb], w, iters, 1.0)
        sl = (slice(y0 - ya, y0 - ya + 64), slice(x0 - xa, x0 - xa + 64))
        assert_flow_close(u[y0:y0 + 64, x0:x0 + 64], v[y0:y0 + 64, x0:x0 + 64], cu[sl], cv_[sl])


def test_large_image_8192_square(pkg):
    """Maximum-size style case (67 Mpixel, 1.7 GB of planes): fused == generic bit for bit, so no index
    arithmetic overflows and the dataflow launch holds with 13 000 tiles per phase."""
    from cpp_optical_flow_b200 import hs_ctypes as H
    rng = np.random.default_rng(8192)
    a = rng.integers(0, 256, (8192, 8192), dtype=np.uint8)
    b = np.roll(a, 1, axis=1)
    with pkg.Solver(8192, 8192, 3, 13, 1.0, flags=H.FLAG_FORCE_GENERIC) as s:
        gu, gv = s.solve(a, b, np.float32)
    with pkg.Solver(8192, 8192, 3, 13, 1.0) as s:
        u, v = s.solve(a, b, np.float32)
        assert s.timing().kernel_id == 1
    assert np.array_equal(u, gu) and np.array_equal(v, gv)


# ---------------------------------------------------------------- oracle parity AT THE DEPTH the configs name
def deltas(u, v, ou, ov):
    du, dv = float(np.abs(u - ou).max()), float(np.abs(v - ov).max())
    epe = float(np.abs(np.hypot(u, v) - np.hypot(ou, ov)).mean())
    return du, dv, epe


@pytest.mark.parametrize("w", [3, 5])
def test_config2_1080p_1000_sweeps_vs_fp64_oracle(pkg, c_oracle, record, w):
    """configs[1] exactly: 1920x1080, alpha=1, 1000 sweeps (hornSchunck.cpp:56-74), hs_solve against the
    fp64 C oracle at full size and full depth (oracle: ~2 G pixel-iterations, OpenMP over rows)."""
    from cpp_optical_flow_b200 import synth
    a, b = synth.frame_pair(1080, 1920)
    ou, ov = c_oracle.flow(a, b, w, 1000, 1.0)
    with pkg.Solver(1920, 1080, w, 1000, 1.0) as s:
        u, v = s.solve(a, b, np.float64)
        k = s.timing().temporal_k
    du, dv, epe = deltas(u, v, ou, ov)
    record(case=f"1080p w={w} T=1000", k=k, max_du=du, max_dv=dv, epe=epe, umax=float(np.abs(ou).max()))
    assert du <= TOL_MAX and dv <= TOL_MAX and epe <= TOL_EPE, f"max|du|={du:.3e} max|dv|={dv:.3e} mean EPE diff={epe:.3e}"


@pytest.mark.parametrize("pair", ["000050", "000040"])
def test_kitti_w3_1000_sweeps_vs_fp64_oracle(pkg, c_oracle, record, kitti, pair):
    """The fp32 worst case SURVEY found (real images, |u| > 100 px, w=3, 1000 sweeps: 7.5e-5 with a
    different summation order).  Full fp64 oracle."""
    a, b = kitti(pair)
    ou, ov = c_oracle.flow(a, b, 3, 1000, 1.0)
    with pkg.Solver(a.shape[1], a.shape[0], 3, 1000, 1.0) as s:
        u, v = s.solve(a, b, np.float64)
        k = s.timing().temporal_k
    du, dv, epe = deltas(u, v, ou, ov)
    record(case=f"kitti {pair} w=3 T=1000", k=k, max_du=du, max_dv=dv, epe=epe, umax=float(np.abs(ou).max()))
    assert np.abs(ou).max() > 50
    assert du <= TOL_MAX and dv <= TOL_MAX and epe <= TOL_EPE, f"max|du|={du:.3e} max|dv|={dv:.3e} mean EPE diff={epe:.3e}"


def test_config3_4k_2000_sweeps_vs_fp64_oracle(pkg, c_oracle, record):
    """configs[2] exactly: 3840x2160, 2000 sweeps, against the full fp64 C oracle (16.6 G pixel-iterations;
    about a minute on the GPU box's host cores)."""
    from cpp_optical_flow_b200 import synth
    a, b = synth.frame_pair(2160, 3840)
    ou, ov = c_oracle.flow(a, b, 3, 2000, 1.0)
    with pkg.Solver(3840, 2160, 3, 2000, 1.0) as s:
        u, v = s.solve(a, b, np.float64)
        k = s.timing().temporal_k
    du, dv, epe = deltas(u, v, ou, ov)
    record(case="4k w=3 T=2000", k=k, max_du=du, max_dv=dv, epe=epe, umax=float(np.abs(ou).max()))
    assert du <= TOL_MAX and dv <= TOL_MAX and epe <= TOL_EPE, f"max|du|={du:.3e} max|dv|={dv:.3e} mean EPE diff={epe:.3e}"


# ---------------------------------------------------------------- HS_PREC_F64: the reference's arithmetic, bit for bit
@pytest.mark.parametrize("w,iters,shape", [(3, 100, (97, 150)), (5, 100, (120, 161)), (1, 7, (33, 47)), (2, 9, (64, 96)),
                                           (4, 9, (37, 131)), (7, 30, (64, 80)), (9, 5, (40, 64)), (3, 1, (1, 1)), (5, 3, (1, 7))])
def test_f64_path_is_bit_identical_to_the_fp64_oracle(pkg, c_oracle, oracle, w, iters, shape):
    """precision = HS_PREC_F64 repeats hornSchunck.cpp:19-75 operation by operation in fp64: gradients AND flow
    equal the oracle's bits (C oracle for any w; cv2's own filter2D / Sobel for w <= 7)."""
    from cpp_optical_flow_b200 import hs_ctypes as H
    a, b = rand_pair(shape, seed=w * 13 + iters)
    ou, ov = c_oracle.flow(a, b, w, iters, 0.7)
    with pkg.Solver(shape[1], shape[0], w, iters, 0.7, precision=H.PREC_F64) as s:
        gx, gy, gt = s.gradients(a, b)
        u, v = s.solve(a, b, np.float64)
        u32, v32 = s.solve(a, b, np.float32)
        assert s.timing().kernel_id == 2
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    assert np.array_equal(u32, ou.astype(np.float32)) and np.array_equal(v32, ov.astype(np.float32))
    og = oracle.np_gradients(a, b)
    assert all(np.array_equal(x, y) for x, y in zip((gx, gy, gt), og))
    if w <= 7 and min(shape) >= 2:
        *_, cu, cv_ = oracle.cv_flow(a, b, w, iters, 0.7)
        assert np.array_equal(u, cu) and np.array_equal(v, cv_)


def test_f64_path_separates_fp32_rounding_from_kernel_bugs(pkg, c_oracle, kitti, record):
    """The A/B the fp32 worst case calls for: KITTI 000050, w=3, 1000 sweeps.  fp64 path == oracle
    bit for bit, so the whole 4e-5 of the fp32 path is rounding."""
    from cpp_optical_flow_b200 import hs_ctypes as H
    a, b = kitti("000050")
    ou, ov = c_oracle.flow(a, b, 3, 1000, 1.0)
    with pkg.Solver(a.shape[1], a.shape[0], 3, 1000, 1.0, precision=H.PREC_F64) as s:
        u, v = s.solve(a, b, np.float64)
        ms = s.timing().iterate_ms
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    record(case="kitti 000050 w=3 T=1000 HS_PREC_F64", max_du=0.0, max_dv=0.0, epe=0.0, iterate_ms=ms,
           gpixit_s=a.size * 1000 / ms / 1e6)


@pytest.mark.parametrize("dtype,scale", [(np.float32, 1.0 / 255), (np.float64, 1.0 / 255), (np.uint16, 257), (np.int16, -3),
                                         (np.int32, 1000), (np.int8, None)])
def test_frames_of_any_depth_like_the_reference(pkg, c_oracle, oracle, dtype, scale):
    """hornSchunck.cpp:23-24 converts frames of ANY depth to CV_64FC1.  CV_32F frames in [0,1], 16-bit frames,
    negative values: the mirror class routes them to the fp64 path; result == oracle on the same values."""
    a8, b8 = rand_pair((90, 140), seed=5)
    if scale is None:
        a, b = (a8.astype(np.int16) - 128).astype(np.int8), (b8.astype(np.int16) - 128).astype(np.int8)
    else:
        a, b = (a8.astype(np.float64) * scale).astype(dtype), (b8.astype(np.float64) * scale).astype(dtype)
    assert a.dtype == dtype
    ou, ov = c_oracle.flow_real(a.astype(np.float64), b.astype(np.float64), 3, 40, 0.5)
    hs = pkg.hornSchunck(3, 40, 0.5)
    u, v = hs.getFlow(a, b)
    gx, gy, gt = hs.getGradients(a, b)
    hs.close()
    assert np.array_equal(u, ou) and np.array_equal(v, ov)
    ogx, ogy, ogt = oracle.np_gradients(a.astype(np.float64), b.astype(np.float64))
    assert np.array_equal(gx, ogx) and np.array_equal(gy, ogy) and np.array_equal(gt, ogt)
    *_, cu, cv_ = oracle.cv_flow(a, b, 3, 40, 0.5)        # OpenCV's own Sobel sums floats in another order
    assert np.abs(u - cu).max() <= 1e-9 * max(1.0, np.abs(cu).max()) and np.abs(v - cv_).max() <= 1e-9 * max(1.0, np.abs(cv_).max())


def test_fp32_path_refuses_frames_it_cannot_represent(pkg):
    from cpp_optical_flow_b200 import hs_ctypes as H
    with pytest.raises(H.HsError) as e:
        pkg.Solver(32, 32, 3, 5, 1.0, frame_dtype=np.float32)             # fp32 path + float frames
    assert e.value.status == 4 and "HS_PREC_F64" in str(e.value)
    with pkg.Solver(32, 32, 3, 5, 1.0, precision=H.PREC_F64, frame_dtype=np.float32) as s:
        with pytest.raises(ValueError):
            s.solve(np.zeros((32, 32), np.uint8), np.zeros((32, 32), np.uint8))
        with pytest.raises(H.HsError):
            s.iterate_rows(1, 0, 4, True)


# ---------------------------------------------------------------- error behaviour
def test_argument_errors(pkg):
    from cpp_optical_flow_b200 import hs_ctypes as H
    with pytest.raises(H.HsError) as e:
        pkg.Solver(0, 10, 3, 1, 1.0)
    assert e.value.status == 1
    with pytest.raises(H.HsError):
        pkg.Solver(10, 10, 0, 1, 1.0)
    with pytest.raises(H.HsError):
        pkg.Solver(10, 10, 3, -1, 1.0)
    with pkg.Solver(32, 16, 3, 1, 1.0) as s:
        with pytest.raises(ValueError):                      # main.cpp:70-73: sizes must agree
            s.solve(np.zeros((16, 32), np.uint8), np.zeros((16, 31), np.uint8))
        with pytest.raises(H.HsError) as e:
            s.iterate(1)                                     # nothing uploaded / prepared yet
        assert e.value.status == 5
