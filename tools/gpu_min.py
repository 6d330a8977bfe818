"""Smallest possible fused-kernel launch (a debugging aid: one tile, a few sweeps)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import hs_oracle as O
import cpp_optical_flow_b200 as P
w = int(os.environ.get("W_", 3)); T = int(os.environ.get("T_", 2)); k = int(os.environ.get("K_", 2))
Hh = int(os.environ.get("H_", 64)); Ww = int(os.environ.get("WID_", 96))
rng = np.random.default_rng(0)
a = rng.integers(0, 256, (Hh, Ww), dtype=np.uint8)
b = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
*_, ou, ov = O.np_flow(a, b, w, T, 1.0)
with P.Solver(Ww, Hh, w, T, 1.0, temporal_k=k) as s:
    u, v = s.solve(a, b)
    print("kernel", s.timing().kernel_id, "k", s.timing().temporal_k)
print("max err", np.abs(u - ou).max(), np.abs(v - ov).max())
