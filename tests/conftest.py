import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    import hs_oracle
    return hs_oracle


@pytest.fixture(scope="session")
def c_oracle():
    """The plain-C oracle (oracle/hs_oracle.c), built on demand with the committed Makefile."""
    lib_path = os.path.join(ROOT, "oracle", "_build", "libhs_oracle.so")
    src = os.path.join(ROOT, "oracle", "hs_oracle.c")
    if not os.path.exists(lib_path) or os.path.getmtime(lib_path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    lib = ctypes.CDLL(lib_path)

    class COracle:
        def flow(self, prev, nxt, w, iters, alpha, dtype=np.float64, threads=0):
            prev = np.ascontiguousarray(prev, np.uint8)
            nxt = np.ascontiguousarray(nxt, np.uint8)
            h, wd = prev.shape
            u = np.zeros((h, wd), dtype)
            v = np.zeros((h, wd), dtype)
            # OpenMP over rows (bit-identical for any thread count: every pixel is summed by one thread)
            lib.hs_oracle_set_threads(int(threads) if threads else min(os.cpu_count() or 1, 32))
            fn = lib.hs_oracle_flow_f64 if np.dtype(dtype) == np.float64 else lib.hs_oracle_flow_f32
            vp = ctypes.c_void_p
            rc = fn(vp(prev.ctypes.data), vp(nxt.ctypes.data), h, wd, int(w), int(iters),
                    ctypes.c_double(alpha), vp(u.ctypes.data), vp(v.ctypes.data))
            assert rc == 0
            return u, v

        def flow_real(self, prev, nxt, w, iters, alpha, threads=0):
            """Frames of any depth, converted to float64 first as hornSchunck.cpp:23-24 does."""
            prev = np.ascontiguousarray(prev, np.float64)
            nxt = np.ascontiguousarray(nxt, np.float64)
            h, wd = prev.shape
            u = np.zeros((h, wd)); v = np.zeros((h, wd))
            lib.hs_oracle_set_threads(int(threads) if threads else min(os.cpu_count() or 1, 32))
            vp = ctypes.c_void_p
            rc = lib.hs_oracle_flow_real_f64(vp(prev.ctypes.data), vp(nxt.ctypes.data), h, wd, int(w), int(iters),
                                             ctypes.c_double(alpha), vp(u.ctypes.data), vp(v.ctypes.data))
            assert rc == 0
            return u, v

        def gradients(self, prev, nxt, dtype=np.float64):
            prev = np.ascontiguousarray(prev, np.uint8)
            nxt = np.ascontiguousarray(nxt, np.uint8)
            h, wd = prev.shape
            g = [np.zeros((h, wd), dtype) for _ in range(3)]
            fn = lib.hs_oracle_gradients_f64 if np.dtype(dtype) == np.float64 else lib.hs_oracle_gradients_f32
            vp = ctypes.c_void_p
            fn(vp(prev.ctypes.data), vp(nxt.ctypes.data), h, wd, *[vp(x.ctypes.data) for x in g])
            return tuple(g)

    return COracle()


@pytest.fixture(scope="session")
def kitti():
    """The two bundled frame pairs the reference author ran, as gray uint8 (tests/golden)."""
    import cv2

    def load(pair):
        a = cv2.imread(os.path.join(GOLDEN, f"kitti_{pair}_10_gray.png"), cv2.IMREAD_UNCHANGED)
        b = cv2.imread(os.path.join(GOLDEN, f"kitti_{pair}_11_gray.png"), cv2.IMREAD_UNCHANGED)
        assert a is not None and b is not None and a.ndim == 2
        return a, b

    return load


@pytest.fixture(scope="session")
def pkg():
    import cpp_optical_flow_b200
    return cpp_optical_flow_b200


@pytest.fixture(scope="session")
def record():
    """Append measured parity deltas to gpurun_out/parity_deltas.jsonl (read back into DESIGN.md)."""
    import json
    path = os.path.join(ROOT, "gpurun_out", "parity_deltas.jsonl")

    def rec(**kw):
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "a") as f:
                f.write(json.dumps(kw) + "\n")
        except OSError:
            pass
    return rec
