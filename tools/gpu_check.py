"""First-contact GPU diagnostics (run under gpurun): parity of every kernel variant against the
oracle and against each other, with error locations, plus a quick timing sweep.  Writes
gpurun_out/gpu_check.json.  Not part of the product."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hs_oracle as O                      # noqa: E402
import cpp_optical_flow_b200 as P          # noqa: E402
from cpp_optical_flow_b200 import hs_ctypes as H   # noqa: E402

out = {"cases": [], "timing": []}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def where(d):
    i = np.unravel_index(np.argmax(d), d.shape)
    return [int(x) for x in i]


def check(shape, w, T, alpha, k, seed=0, kitti=None):
    rng = np.random.default_rng(seed)
    if kitti is None:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        # smooth-ish second frame so flows stay moderate
        b = np.clip(a.astype(int) + rng.integers(-20, 21, shape), 0, 255).astype(np.uint8)
    else:
        a, b = kitti
    Hh, Ww = a.shape
    rec = {"shape": [Hh, Ww], "w": w, "T": T, "alpha": alpha, "k": k}
    try:
        gx, gy, gt, ou, ov = O.np_flow(a, b, w, T, alpha) if Hh * Ww * T < 3e7 else O.cv_flow(a, b, w, T, alpha)
        with P.Solver(Ww, Hh, w, T, alpha, flags=H.FLAG_FORCE_GENERIC) as s:
            g = s.gradients(a, b)
            rec["grad_exact"] = bool(all(np.array_equal(x, y) for x, y in zip(g, (gx, gy, gt))))
            gu, gv = s.solve(a, b, np.float64)
        rec["generic_max_du"] = float(np.abs(gu - ou).max()); rec["generic_max_dv"] = float(np.abs(gv - ov).max())
        with P.Solver(Ww, Hh, w, T, alpha, temporal_k=k) as s:
            tu, tv = s.solve(a, b, np.float64)
            tm = s.timing()
            rec["kernel_id"] = tm.kernel_id; rec["k_used"] = tm.temporal_k
        du = np.abs(tu - gu); dv = np.abs(tv - gv)
        rec["tile_vs_generic_max"] = float(max(du.max(), dv.max()))
        rec["tile_vs_generic_where"] = where(np.maximum(du, dv))
        rec["tile_vs_generic_nbad"] = int(((du > 0) | (dv > 0)).sum())
        rec["tile_max_du"] = float(np.abs(tu - ou).max()); rec["tile_max_dv"] = float(np.abs(tv - ov).max())
        rec["umax"] = float(np.abs(ou).max())
    except Exception as e:  # noqa: BLE001
        rec["error"] = repr(e)
    out["cases"].append(rec)
    print(json.dumps(rec), flush=True)


cases = [
    ((64, 96), 3, 1, 1.0, 1), ((64, 96), 3, 2, 1.0, 2), ((64, 96), 3, 8, 1.0, 4), ((64, 96), 3, 9, 1.0, 4),
    ((200, 300), 3, 12, 1.0, 4), ((200, 300), 3, 12, 1.0, 3), ((200, 300), 3, 16, 1.0, 8),
    ((200, 300), 5, 12, 1.0, 2), ((200, 300), 5, 12, 1.0, 3), ((200, 300), 5, 9, 0.5, 4),
    ((130, 250), 2, 10, 1.0, 4), ((130, 250), 4, 10, 1.0, 2), ((77, 131), 7, 6, 1.0, 0), ((50, 61), 1, 5, 1.0, 0),
    ((1, 1), 3, 3, 1.0, 0), ((1, 7), 3, 3, 1.0, 0), ((2, 3), 5, 4, 1.0, 0), ((3, 2), 3, 4, 1.0, 0), ((5, 5), 3, 5, 1.0, 2),
    ((375, 1242), 3, 20, 1.0, 4), ((375, 1242), 5, 20, 10.0, 2),
]
for c in cases:
    check(*c)

# the author's own run: pair 000050, w=5, T=100, alpha=1 (main.cpp:94-96)
try:
    import cv2
    g = os.path.join(ROOT, "tests", "golden")
    kp = (cv2.imread(f"{g}/kitti_000050_10_gray.png", cv2.IMREAD_UNCHANGED), cv2.imread(f"{g}/kitti_000050_11_gray.png", cv2.IMREAD_UNCHANGED))
    check(None, 5, 100, 1.0, 2, kitti=kp)
    check(None, 3, 200, 1.0, 4, kitti=kp)
except Exception as e:  # noqa: BLE001
    print("kitti failed", e)

# timing sweep on 1080p / 4K (device-resident, T sweeps, CUDA events inside the library are per
# hs_solve; here wall clock around solve_device+sync with frames already uploaded)
from cpp_optical_flow_b200 import synth  # noqa: E402
for (Hh, Ww, T) in ((1080, 1920, 1000), (2160, 3840, 400)):
    a, b = synth.frame_pair(Hh, Ww)
    for w in (3, 5):
        for k, flags in ((1, H.FLAG_FORCE_GENERIC), (1, 0), (2, 0), (3, 0), (4, 0), (6, 0), (8, 0), (12, 0)):
            try:
                with P.Solver(Ww, Hh, w, T, 1.0, temporal_k=k, flags=flags) as s:
                    s.upload(a, b)
                    s.solve_device(); s.sync()
                    best = 1e9
                    for _ in range(3):
                        t0 = time.perf_counter(); s.solve_device(); s.sync(); best = min(best, time.perf_counter() - t0)
                    tm = s.timing()
                    rec = {"H": Hh, "W": Ww, "w": w, "T": T, "k_req": k, "k": tm.temporal_k, "kernel": tm.kernel_id,
                           "ms": best * 1e3, "mpixit_s": Hh * Ww * T / best / 1e6}
            except Exception as e:  # noqa: BLE001
                rec = {"H": Hh, "W": Ww, "w": w, "k_req": k, "error": repr(e)}
            out["timing"].append(rec)
            print(json.dumps(rec), flush=True)

json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)
