#!/bin/bash
set -u
TAG=${1:-r02n}
mkdir -p gpurun_out
echo "== pytest -m gpu (full)"
rm -f gpurun_out/parity_deltas.jsonl
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
echo "== k sweep"
KS=1,2,3,4,5,6,8,10,12 WS=3,5 python tools/gpu_sweep.py kitti 640x480 720p 900p 1080p 1440p 4k > gpurun_out/${TAG}_k_sweep.jsonl 2> gpurun_out/${TAG}_k_sweep.err; wc -l gpurun_out/${TAG}_k_sweep.jsonl; tail -2 gpurun_out/${TAG}_k_sweep.err
echo "== default k per size"
python tools/gpu_default_k.py > gpurun_out/${TAG}_default_k.jsonl 2>&1; tail -3 gpurun_out/${TAG}_default_k.jsonl
