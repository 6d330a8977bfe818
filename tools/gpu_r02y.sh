#!/bin/bash
# full session with the retuned k: tests, smoke, default bench (with slab record), other workloads, launch list,
# full ncu captures (w=3, w=5), DRAM-traffic probe
set -u
TAG=${1:-r02y}
mkdir -p gpurun_out
echo "== pytest -m gpu"
rm -f gpurun_out/parity_deltas.jsonl
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"
python bench.py > gpurun_out/${TAG}_bench_default_n1.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 5 --warmup 3 --window 5 --no-cpu --no-slab > gpurun_out/${TAG}_bench_1080p_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --workload kitti --no-cpu --no-slab > gpurun_out/${TAG}_bench_kitti_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 3 --warmup 3 --workload 4k --no-cpu --no-slab > gpurun_out/${TAG}_bench_4k_w3.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 2 --warmup 3 --workload batch256 --no-cpu --no-slab > gpurun_out/${TAG}_bench_batch256.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
for f in default_n1 1080p_w5 kitti_w5 4k_w3 batch256; do python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$f.json").read().strip().splitlines()[-1])
e=d.get("e2e") or {}
print("$f", round(d["value"]/1e3,1), "k", d["config"].get("temporal_k"), "e2e", round(e.get("value",0)/1e3,1), "slab", round(d.get("slab16k",{}).get("value",0)/1e3,1), d["roofline"].get("fp32_pipe",{}).get("frac"), d["roofline"].get("traffic_source","")[:40])
PY
done
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-cpu --no-slab > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-slab > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
bash tools/gpu_ncu.sh ${TAG} 3
bash tools/gpu_ncu.sh ${TAG}w5 5
echo "== traffic probe"
python tools/traffic_probe.py gpurun_out/${TAG}_traffic_plain.jsonl > gpurun_out/${TAG}_traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none \
    -k regex:k_jacobi_tile --csv --log-file gpurun_out/${TAG}_traffic_ncu.csv python tools/traffic_probe.py > gpurun_out/${TAG}_traffic_ncu.log 2>&1
echo "traffic ncu rc=$?"; tail -2 gpurun_out/${TAG}_traffic_plain.log
