"""Output formats after the hot path: .flo round trip, and the OpenCV FileStorage
text the reference writes (HornSchunckOF/main.cpp:99-102) is readable by OpenCV itself."""
import numpy as np


def test_flo_round_trip(pkg, tmp_path):
    rng = np.random.default_rng(0)
    u = rng.normal(size=(37, 53)).astype(np.float32)
    v = rng.normal(size=(37, 53)).astype(np.float32)
    p = str(tmp_path / "f.flo")
    pkg.flowio.write_flo(p, u, v)
    ru, rv = pkg.flowio.read_flo(p)
    assert np.array_equal(ru, u) and np.array_equal(rv, v)


def test_opencv_yaml_is_read_back_by_opencv(pkg, tmp_path):
    import cv2
    rng = np.random.default_rng(1)
    m = rng.normal(scale=50, size=(19, 23))
    m[3, 4] = 0.0
    p = str(tmp_path / "uMatrixHS.txt")                       # the reference uses a .txt name, too
    pkg.flowio.write_opencv_yaml(p, "u matrix", m)
    fs = cv2.FileStorage(p, cv2.FILE_STORAGE_READ | cv2.FILE_STORAGE_FORMAT_YAML)
    back = fs.getNode("u matrix").mat()
    fs.release()
    assert back is not None and back.dtype == np.float64 and back.shape == m.shape
    assert np.array_equal(back, m)                             # %.16e round-trips float64
