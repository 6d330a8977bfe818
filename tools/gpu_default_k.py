"""Which k does the library pick by default, and how fast is it?  (tuning aid)  python tools/gpu_default_k.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cpp_optical_flow_b200 as P
from cpp_optical_flow_b200 import synth
for (Hh, Ww) in [(240, 320), (375, 1242), (480, 640), (720, 1280), (768, 1366), (900, 1600), (1080, 1920), (1200, 1600), (1440, 2560)]:
    a, b = synth.frame_pair(Hh, Ww)
    for w in (3, 5):
        T = 1000
        for kk in (0, 4 if w == 3 else 3):        # library default, then the large-frame default
          with P.Solver(Ww, Hh, w, T, 1.0, temporal_k=kk) as s:
            s.upload(a, b); s.solve_device(); s.sync()
            best = 1e9
            for _ in range(3):
                s.solve_device(); s.sync(); best = min(best, s.timing().iterate_ms)
            print(json.dumps({"size": f"{Ww}x{Hh}", "w": w, "k_arg": kk, "k": s.timing().temporal_k,
                              "iterate_ms": round(best, 3), "gpixit_s": round(Hh * Ww * T / best / 1e6, 1)}), flush=True)
