#!/bin/bash
# final single-GPU record of the round: the driver's three steps (tests, smoke, default bench) + reference arm
set -u
TAG=${1:-r02z}
mkdir -p gpurun_out
rm -f gpurun_out/parity_deltas.jsonl
[ "${SKIP_TESTS:-0}" = "1" ] || timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.txt
T0=$(date +%s); python bench.py > gpurun_out/${TAG}_bench_default_n1.json 2> gpurun_out/${TAG}_bench.err; echo "default bench wall seconds: $(( $(date +%s) - T0 ))"
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_default_n1.json").read().strip().splitlines()[-1])
e=d["e2e"]
print(round(d["value"]/1e3,1), "k", d["config"]["temporal_k"], {k:(round(v/1e3,1) if isinstance(v,float) else v) for k,v in e.items() if "value" in k}, "slab", round(d["slab16k"]["value"]/1e3,1), d["clocks"], d["gpu_launches"])
r=json.loads(open("gpurun_out/${TAG}_bench_reference.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"])
PY
