#!/bin/bash
# Multi-GPU session (gpurun --gpus N): slab correctness, weak-scaling batch bench, slab bench.
N=${1:-2}; TAG=${2:-r01}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
echo "== slab check"; $TR tools/slab_check.py 4096 4096 51 3 6 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/${TAG}_slab_check_n$N.json
$TR tools/slab_check.py 2048 3000 23 5 3 2>&1 | grep -E "^\{|Error|error" | tee -a gpurun_out/${TAG}_slab_check_n$N.json
echo "== batch bench"; $TR bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/${TAG}_bench_1080p_n$N.json
echo "== slab bench"; $TR bench.py --gpus $N --workload slab16k --steps 2 --warmup 1 --iters ${SLAB_ITERS:-600} 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/${TAG}_bench_slab16k_n$N.json
