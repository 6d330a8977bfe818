#!/bin/bash
set -u
TAG=${1:-r02c}
mkdir -p gpurun_out
echo "== quick tests"
timeout -s KILL 900 python -m pytest tests -m gpu -x -q -k "row_slab_group or fused_kernel_equals or fuzz or race_free or default_k" 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_quick.txt
echo "== bench"
for w in 3 5; do
python bench.py --steps 10 --warmup 3 --no-cpu --window $w > gpurun_out/${TAG}_bench_1080p_w$w.json 2>> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_1080p_w$w.json')); print('1080p w$w', d['value'], d['e2e']['value'], d['iterate_ms'])"
done
python bench.py --steps 5 --warmup 3 --no-cpu --workload kitti > gpurun_out/${TAG}_bench_kitti.json 2>> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_kitti.json')); print('kitti', d['value'], d['e2e']['value'])"
python bench.py --steps 3 --warmup 3 --no-cpu --workload 4k > gpurun_out/${TAG}_bench_4k.json 2>> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_4k.json')); print('4k', d['value'], d['e2e']['value'])"
echo "== tile profile"
python tools/tile_profile.py 2>&1 | tail -6 | tee gpurun_out/${TAG}_tile_profile.txt
tail -5 gpurun_out/${TAG}_bench.err
