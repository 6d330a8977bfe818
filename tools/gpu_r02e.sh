#!/bin/bash
set -u
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "slab" 2>&1 | tail -8
echo "== default bench N=1 (with slab sub-record)"
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench.err; tail -c 2200 gpurun_out/${TAG}_bench_default.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== slab16k workload line, small"
HS_SLAB_SIZE=4096 python bench.py --workload slab16k --iters 300 --steps 2 --warmup 1 2>&1 | tail -c 1500
