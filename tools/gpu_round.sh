#!/bin/bash
# One GPU session: tests, smoke, bench, launch list, one full ncu capture of the dominant kernel.
# Usage (under gpurun): bash tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt
echo "== pytest -m gpu"
timeout -s KILL 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench (1080p w=3, then w=5, 4k)"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1080p_w3.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench_1080p_w3.json
python bench.py --steps 5 --warmup 3 --window 5 --no-cpu > gpurun_out/${TAG}_bench_1080p_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 3 --warmup 3 --workload 4k --no-cpu > gpurun_out/${TAG}_bench_4k_w3.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --workload kitti --no-cpu > gpurun_out/${TAG}_bench_kitti_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
echo "== ncu full capture of the fused kernel"
python bench.py --steps 1 --warmup 1 --no-cpu --iters 120 > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tile -s 2 -c 2 -f -o gpurun_out/${TAG}_prof_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --iters 120 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
