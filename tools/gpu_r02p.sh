#!/bin/bash
# re-entry session: full GPU tests, default bench, launch list, full ncu capture of the fused kernel (w=3, w=5)
set -u
TAG=${1:-r02p}
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"
python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench default"
python bench.py > gpurun_out/${TAG}_bench_default_n1.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_default_n1.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --steps 5 --warmup 3 --window 5 --no-cpu --no-slab > gpurun_out/${TAG}_bench_1080p_w5.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --workload kitti --no-cpu --no-slab > gpurun_out/${TAG}_bench_kitti_w5.json 2>> gpurun_out/${TAG}_bench.err
echo "== ncu launch list"
python bench.py --steps 2 --warmup 1 --no-cpu --no-slab > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-slab > gpurun_out/${TAG}_ncu_list.log 2>&1
echo "ncu list rc=$?"
bash tools/gpu_ncu.sh ${TAG} 3
bash tools/gpu_ncu.sh ${TAG}w5 5
ls -la gpurun_out | tail -20
