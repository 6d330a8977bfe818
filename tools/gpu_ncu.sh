#!/bin/bash
# one full ncu capture of the fused kernel (1080p, 120 sweeps per launch)
TAG=${1:-r02h}; W=${2:-3}
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu --no-slab --iters 120 --window $W > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_jacobi_tile -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_tile \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-slab --iters 120 --window $W > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${TAG}_ncu_full.log
