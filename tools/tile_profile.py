"""Where does a tile's time go?  Builds a -DHS_TILE_PROFILE copy of the library (clock64 stamps in
the compute warps), runs one workload and prints the per-tile cycle breakdown.  Debug aid only."""
import ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cpp_optical_flow_b200 import _build, hs_ctypes, synth
import cpp_optical_flow_b200 as P
lib_path = os.path.join(ROOT, "build", "libhs_b200_prof.so")
if "--build" in sys.argv or not os.path.exists(lib_path):
    os.makedirs(os.path.dirname(lib_path), exist_ok=True)
    cmd = [_build.nvcc_path(), "-ccbin", "/usr/bin/g++"] + _build.NVCC_FLAGS + ["-DHS_TILE_PROFILE", "-o", lib_path] + _build.SOURCES
    subprocess.run(cmd, check=True)
    if "--build" in sys.argv:
        sys.exit(0)
hs_ctypes._lib = None
lib = hs_ctypes.load_library(lib_path)
hs_ctypes._lib = lib
lib.hs_debug_tile_profile.argtypes = [ctypes.c_void_p, ctypes.c_int]
for (Hh, Ww, T, w, k) in [(1080, 1920, 1000, 3, 4), (1080, 1920, 1000, 3, 6), (2160, 3840, 400, 3, 6), (1080, 1920, 1000, 5, 3)]:
    a, b = synth.frame_pair(Hh, Ww)
    with P.Solver(Ww, Hh, w, T, 1.0, temporal_k=k) as s:
        s.upload(a, b); s.solve_device(); s.sync()
        buf = np.zeros((256, 8), np.int64)
        lib.hs_debug_tile_profile(buf.ctypes.data, 1)
        s.solve_device(); s.sync()
        ms = s.timing().iterate_ms
        lib.hs_debug_tile_profile(buf.ctypes.data, 1)
    act = buf[buf[:, 4] > 0]
    tiles = act[:, 4].sum()
    names = ["wait_tma", "load_unpack_sync", "sweeps", "stores"]
    per = {n: act[:, i].sum() / tiles for i, n in enumerate(names)}
    tot = sum(per.values())
    print(f"{Ww}x{Hh} w={w} k={k}: {ms:.3f} ms, {Hh*Ww*T/ms/1e6:.0f} Gpix-it/s, tiles/CTA {tiles/len(act):.1f}, cycles/tile {tot:.0f}: "
          + ", ".join(f"{n} {v:.0f} ({100*v/tot:.0f}%)" for n, v in per.items())
          + f" | barrier->issue of next tile avg {act[:, 7].sum() / tiles:.0f} cyc (early decode+deps+fences took {act[:, 5].sum() / tiles:.0f})", flush=True)
