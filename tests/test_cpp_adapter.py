"""The C++ host side: cpp-optical-flow_b200/adapter/hornSchunck.cpp (same class surface as the
reference's hornSchunck.cpp) compiled against the minicv stub (no OpenCV C++ SDK in this image) and
driven by the three lines of HornSchunckOF/main.cpp:93-98 (tests/cpp/adapter_driver.cpp)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu


@pytest.fixture(scope="module")
def driver(tmp_path_factory, pkg):
    pkg.load_library()
    exe = str(tmp_path_factory.mktemp("cpp") / "adapter_driver")
    libdir = os.path.join(ROOT, "cpp-optical-flow_b200")
    cmd = ["/usr/bin/g++", "-std=c++14", "-O2", "-Wall", "-Werror",
           "-I", os.path.join(ROOT, "tests", "minicv"), "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(libdir, "adapter"), os.path.join(ROOT, "tests", "cpp", "adapter_driver.cpp"),
           "-o", exe, "-L", libdir, "-l:libhs_b200.so", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def run_driver(exe, tmp_path, a, b, w, iters, alpha, roi=False, mode=None):
    pa, pb, po = str(tmp_path / "a.raw"), str(tmp_path / "b.raw"), str(tmp_path / "o.bin")
    a.tofile(pa); b.tofile(pb)
    args = [exe, pa, pb, str(a.shape[0]), str(a.shape[1]), str(w), str(iters), repr(alpha), po] + \
        (["roi"] if roi else []) + ([mode] if mode else [])
    r = subprocess.run(args, capture_output=True, text=True)
    return r, po


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_adapter_compiles_links_and_throws_without_gpu(driver, tmp_path):
    a = np.zeros((16, 16), np.uint8)
    r, _ = run_driver(driver, tmp_path, a, a, 3, 2, 1.0)
    assert r.returncode == 1 and "cv::Exception" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("w,iters,roi", [(5, 100, False), (3, 33, True)])
def test_adapter_matches_oracle_and_python_mirror(driver, tmp_path, pkg, oracle, w, iters, roi):
    rng = np.random.default_rng(21)
    a = rng.integers(0, 256, (90, 134), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-15, 16, a.shape), 0, 255).astype(np.uint8)
    r, po = run_driver(driver, tmp_path, a, b, w, iters, 1.0, roi)
    assert r.returncode == 0, r.stdout + r.stderr
    if roi:
        a, b = np.ascontiguousarray(a[3:-3, 3:-3]), np.ascontiguousarray(b[3:-3, 3:-3])
    out = np.fromfile(po, np.float64).reshape(5, *a.shape)
    gx, gy, gt, ou, ov = oracle.np_flow(a, b, w, iters, 1.0)
    assert np.array_equal(out[2], gx) and np.array_equal(out[3], gy) and np.array_equal(out[4], gt)
    assert np.abs(out[0] - ou).max() <= 1e-4 and np.abs(out[1] - ov).max() <= 1e-4
    hs = pkg.hornSchunck(w, iters, 1.0)
    u, v = hs.getFlow(a, b)
    hs.close()
    assert np.array_equal(out[0], u) and np.array_equal(out[1], v)      # C++ and Python hosts: same library, same bits


@pytest.mark.gpu
def test_adapter_takes_float_frames_like_the_reference(driver, tmp_path, c_oracle, oracle):
    """CV_32F frames in [0,1]: upstream converts any depth to CV_64FC1 (hornSchunck.cpp:23-24).  The adapter
    routes them to the library's fp64 path; flow and gradients equal the fp64 oracle on the same values, bit
    for bit.  The same path on 8-bit frames (hs.precision = HS_PREC_F64) equals OpenCV's own result."""
    rng = np.random.default_rng(22)
    a = rng.integers(0, 256, (70, 110), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-15, 16, a.shape), 0, 255).astype(np.uint8)
    r, po = run_driver(driver, tmp_path, a, b, 5, 50, 0.05, mode="f32")
    assert r.returncode == 0, r.stdout + r.stderr
    out = np.fromfile(po, np.float64).reshape(5, *a.shape)
    fa, fb = (a * (1.0 / 255)).astype(np.float32), (b * (1.0 / 255)).astype(np.float32)
    ou, ov = c_oracle.flow_real(fa.astype(np.float64), fb.astype(np.float64), 5, 50, 0.05)
    assert np.abs(ou).max() > 1e-3
    assert np.array_equal(out[0], ou) and np.array_equal(out[1], ov)
    gx, gy, gt = oracle.np_gradients(fa.astype(np.float64), fb.astype(np.float64))
    assert np.array_equal(out[2], gx) and np.array_equal(out[3], gy) and np.array_equal(out[4], gt)
    r, po = run_driver(driver, tmp_path, a, b, 5, 50, 1.0, mode="f64prec")
    assert r.returncode == 0, r.stdout + r.stderr
    out = np.fromfile(po, np.float64).reshape(5, *a.shape)
    *_, cu, cv_ = oracle.cv_flow(a, b, 5, 50, 1.0)
    assert np.array_equal(out[0], cu) and np.array_equal(out[1], cv_)
