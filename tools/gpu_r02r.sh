#!/bin/bash
# tests + benches after a kernel change
set -u
TAG=${1:-r02r}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== bench"
python bench.py --no-slab > gpurun_out/${TAG}_bench_1080p_w3.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 5 --warmup 3 --window 5 --no-cpu --no-slab > gpurun_out/${TAG}_bench_1080p_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --workload kitti --no-cpu --no-slab > gpurun_out/${TAG}_bench_kitti_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 3 --warmup 3 --workload 4k --no-cpu --no-slab > gpurun_out/${TAG}_bench_4k_w3.json 2>> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
for f in 1080p_w3 1080p_w5 kitti_w5 4k_w3; do python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$f.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("$f", round(d["value"]/1e3,1), "k", d["config"]["temporal_k"], "e2e", round(e["value"]/1e3,1), "f32out", round(e["value_f32_outputs"]/1e3,1), "pageable", round(e["value_pageable_buffers"]/1e3,1), "stream", round(e["stream_value"]/1e3,1), {k: round(v,3) for k,v in e["last_step_ms"].items()}, d["roofline"].get("fp32_pipe",{}).get("frac"))
PY
done
