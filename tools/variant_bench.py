"""A/B timing of library variants (tools/variant_build.sh): python tools/variant_bench.py default pre1 pre2
Each variant runs in its own process (HS_B200_LIB selects the library); prints Gpix-it/s per case."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("1080p", 3, 4), ("1080p", 3, 6), ("4k", 3, 6), ("4k", 3, 4), ("1080p", 5, 2), ("1080p", 5, 3), ("1080p", 4, 4), ("720p", 3, 8), ("kitti", 5, 4)]
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import cpp_optical_flow_b200 as P
    from cpp_optical_flow_b200 import synth
    SIZES = {"1080p": (1080, 1920, 1000), "4k": (2160, 3840, 400), "720p": (720, 1280, 1000), "kitti": (375, 1242, 1000)}
    out = {}
    for name, w, k in CASES:
        Hh, Ww, T = SIZES[name]
        a, b = synth.frame_pair(Hh, Ww)
        with P.Solver(Ww, Hh, w, T, 1.0, temporal_k=k) as s:
            s.upload(a, b); s.solve_device(); s.sync()
            best = 1e9
            for _ in range(5):
                s.solve_device(); s.sync(); best = min(best, s.timing().iterate_ms)
        out[f"{name} w={w} k={k}"] = round(Hh * Ww * T / best / 1e6, 1)
    print(json.dumps(out))
    sys.exit(0)
res = {}
for v in sys.argv[1:] or ["default"]:
    env = dict(os.environ)
    if v != "default":
        env["HS_B200_LIB"] = os.path.join(ROOT, "build", f"libhs_{v}.so")
    r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
    try:
        res[v] = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        res[v] = {"error": (r.stdout + r.stderr)[-300:]}
keys = list(next(iter(res.values())).keys())
print("case | " + " | ".join(res.keys()))
for k in keys:
    print(k + " | " + " | ".join(str(res[v].get(k)) for v in res))
