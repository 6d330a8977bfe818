#!/bin/bash
# k sweep with the current kernel: all sizes for w=3,5 (the cost-model fit), large frames for the other windows
set -u
TAG=${1:-r02w}
mkdir -p gpurun_out
KS=1,2,3,4,5,6,7,8,10,12 WS=3,5 python tools/gpu_sweep.py kitti 640x480 720p 900p 1080p 1440p 4k > gpurun_out/${TAG}_k_sweep.jsonl 2> gpurun_out/${TAG}_k_sweep.err
KS=1,2,3,4,5,6,8,10,12 WS=2,4 python tools/gpu_sweep.py 1080p 4k > gpurun_out/${TAG}_k_sweep_w24.jsonl 2>> gpurun_out/${TAG}_k_sweep.err
KS=1,2,3,4 WS=6,7,8,9 python tools/gpu_sweep.py 1080p 4k > gpurun_out/${TAG}_k_sweep_w6789.jsonl 2>> gpurun_out/${TAG}_k_sweep.err
wc -l gpurun_out/${TAG}_k_sweep*.jsonl; tail -2 gpurun_out/${TAG}_k_sweep.err
