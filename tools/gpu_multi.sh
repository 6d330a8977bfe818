#!/bin/bash
# Multi-GPU session (gpurun --gpus N): slab correctness (in-kernel exchange), default bench (1080p replicas + the
# 16K^2 row-slab sub-record), slab bench with both exchange variants.
N=${1:-2}; TAG=${2:-r02}
TR="timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n$N.txt 2>&1
echo "== single process, N devices behind ONE hs_ctx (plain C caller): in-kernel exchange, then NCCL"
gcc -std=c99 -O1 -I include tests/c/slab_smoke.c -o /tmp/slab_smoke -L cpp-optical-flow_b200 -l:libhs_b200.so -Wl,-rpath,$PWD/cpp-optical-flow_b200
( timeout -s KILL 120 /tmp/slab_smoke $N distinct peer; echo "rc=$?"; LD_LIBRARY_PATH=$(python -c "import nvidia.nccl, os; print(os.path.join(list(nvidia.nccl.__path__)[0], 'lib'))" 2>/dev/null):${LD_LIBRARY_PATH:-} timeout -s KILL 120 /tmp/slab_smoke $N distinct nccl; echo "rc=$?" ) 2>&1 | tee gpurun_out/${TAG}_slab_smoke_c_n$N.txt
echo "== slab check (in-kernel exchange)"
$TR tools/slab_check.py 4096 4096 51 3 6  2048 3000 23 5 3  3000 2048 40 4 2  4096 4096 400 3 0 2>&1 | grep -E "^\{|Error|error|Traceback" | tee gpurun_out/${TAG}_slab_check_n$N.json
echo "== default bench (with slab16k sub-record)"
$TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_default_n$N.log 2>&1; grep -E "^\{" gpurun_out/${TAG}_bench_default_n$N.log > gpurun_out/${TAG}_bench_default_n$N.json; tail -c 2500 gpurun_out/${TAG}_bench_default_n$N.json; grep -E "Error|error|Traceback" gpurun_out/${TAG}_bench_default_n$N.log | head -5
if [ "${SLAB_AB:-1}" = "1" ]; then
echo "== slab bench: NCCL A/B"
$TR bench.py --gpus $N --workload slab16k --steps 2 --warmup 1 --slab-exchange nccl --no-slab-check 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/${TAG}_bench_slab16k_nccl_n$N.json | cut -c1-600
fi
