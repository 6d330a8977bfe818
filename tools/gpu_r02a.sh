#!/bin/bash
# Round-2 session A: full GPU test-suite (incl. the depth-of-config parity tests), headline bench, DRAM-traffic probe.
set -u
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt
nproc >> gpurun_out/${TAG}_gpu.txt
echo "== pytest -m gpu"
rm -f gpurun_out/parity_deltas.jsonl
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest.txt
echo "== bench"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1080p_w3.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench_1080p_w3.json
echo "== traffic probe"
python tools/traffic_probe.py gpurun_out/${TAG}_traffic_plain.jsonl > gpurun_out/${TAG}_traffic_plain.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none \
    -k regex:k_jacobi_tile --csv --log-file gpurun_out/${TAG}_traffic_ncu.csv python tools/traffic_probe.py > gpurun_out/${TAG}_traffic_ncu.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/${TAG}_traffic_plain.log
ls -la gpurun_out | tail -8
