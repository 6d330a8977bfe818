"""DRAM traffic of the fused Jacobi kernel in the HBM-resident regime.

Runs ONE fused launch per configuration (so that an ncu capture of `-k regex:k_jacobi_tile` maps
1:1 onto the list printed here) and prints one JSON line per configuration with the CUDA-event
time of that launch.  Run it plain first (timings), then under
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
and merge both with tools/traffic_merge.py.

    python tools/traffic_probe.py [out.jsonl]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cpp_optical_flow_b200 as P
from cpp_optical_flow_b200 import synth

out = open(sys.argv[1], "w") if len(sys.argv) > 1 else None
SWEEPS = 48                                     # a multiple of every k below: whole phases only
CASES = []
for k in (1, 2, 3, 4, 6, 8, 12, 16):            # BASELINE configs[2]: 4K, temporal-blocking depth sweep
    CASES.append(("4k", 2160, 3840, 1, 3, k, SWEEPS))
for k in (1, 2, 3, 4, 6):
    CASES.append(("4k", 2160, 3840, 1, 5, k, SWEEPS if k != 6 else 48))
CASES.append(("1080p", 1080, 1920, 1, 3, 4, 1000))        # the headline launch itself (L2-resident)
CASES.append(("1080p", 1080, 1920, 1, 5, 3, 999))
CASES.append(("1080p", 1080, 1920, 1, 3, 6, 996))         # the default k since the r02w sweep
CASES.append(("kitti", 375, 1242, 1, 5, 7, 98))
CASES.append(("batch256", 1080, 1920, 4, 3, 6, SWEEPS))   # 4 pairs per launch = 200 MB working set
CASES.append(("batch256", 1080, 1920, 4, 3, 4, SWEEPS))
CASES.append(("kitti", 375, 1242, 1, 5, 4, 100))
CASES.append(("slab16k", 16384, 16384, 1, 3, 6, 12))      # one 16K^2 single-slab launch (6.4 GB of planes)
CASES.append(("slab16k", 16384, 16384, 1, 3, 4, 12))

cache = {}
for name, Hh, Ww, B, w, k, T in CASES:
    key = (Hh, Ww)
    if key not in cache:
        cache.clear()
        cache[key] = synth.frame_pair(Hh, Ww)
    a, b = cache[key]
    if B > 1:
        a = np.stack([a] * B); b = np.stack([b] * B)
    with P.Solver(Ww, Hh, w, T, 1.0, batch=B, temporal_k=k) as s:
        s.upload(a, b)
        s.solve_device(); s.sync()
        tm = s.timing()
        rec = {"workload": name, "H": Hh, "W": Ww, "batch": B, "window": w, "k": tm.temporal_k, "sweeps": T,
               "tile_launches": tm.launches - 1, "iterate_ms": tm.iterate_ms,
               "gpixit_s": Hh * Ww * B * T / tm.iterate_ms / 1e6}
    line = json.dumps(rec)
    print(line, flush=True)
    if out:
        out.write(line + "\n"); out.flush()
