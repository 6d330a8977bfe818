"""Timing sweep of the fused kernel over k (device-resident, wall clock around solve_device+sync,
best of 3) - a tuning aid, not the benchmark.  Usage: python tools/gpu_sweep.py [1080p|4k|kitti] ..."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_optical_flow_b200 as P
from cpp_optical_flow_b200 import hs_ctypes as H, synth
SIZES = {"1080p": (1080, 1920, 1000), "4k": (2160, 3840, 400), "kitti": (375, 1242, 1000), "8k": (4320, 7680, 100), "720p": (720, 1280, 1000),
         "900p": (900, 1600, 1000), "640x480": (480, 640, 1000), "1440p": (1440, 2560, 600)}
names = sys.argv[1:] or ["1080p", "4k"]
ks = [int(x) for x in os.environ.get("KS", "1,2,3,4,5,6,8,10,12").split(",")]
ws = [int(x) for x in os.environ.get("WS", "3,5").split(",")]
for name in names:
    Hh, Ww, T = SIZES[name]
    a, b = synth.frame_pair(Hh, Ww)
    for w in ws:
        for k in ks:
            try:
                with P.Solver(Ww, Hh, w, T, 1.0, temporal_k=k) as s:
                    s.upload(a, b); s.solve_device(); s.sync()
                    best = 1e9
                    for _ in range(3):
                        s.solve_device(); s.sync(); best = min(best, s.timing().iterate_ms)
                    tm = s.timing()
                    print(json.dumps({"size": name, "H": Hh, "W": Ww, "T": T, "w": w, "k": tm.temporal_k, "iterate_ms": round(best, 3),
                                      "gpixit_s": round(Hh * Ww * T / best / 1e6, 1)}), flush=True)
            except Exception as e:
                print(json.dumps({"size": name, "w": w, "k": k, "error": repr(e)[:200]}), flush=True)
