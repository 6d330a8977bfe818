#!/bin/bash
# Round-2 session B: kernel refactor (LaunchDesc / seams) - tests + headline bench
set -u
TAG=${1:-r02b}
mkdir -p gpurun_out
echo "== group tests first"
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "row_slab_group or multi_device" 2>&1 | tail -30 | tee gpurun_out/${TAG}_pytest_group.txt
echo "== full pytest -m gpu"
timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
echo "== bench"
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench_1080p_w3.json 2> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_1080p_w3.json')); print('1080p w3', d['value'], d['e2e']['value'])"
python bench.py --steps 5 --warmup 3 --no-cpu --window 5 > gpurun_out/${TAG}_bench_1080p_w5.json 2>> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_1080p_w5.json')); print('1080p w5', d['value'], d['e2e']['value'])"
python bench.py --steps 5 --warmup 3 --no-cpu --window 7 > gpurun_out/${TAG}_bench_1080p_w7.json 2>> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_1080p_w7.json')); print('1080p w7', d['value'], d['e2e']['value'], d['config']['temporal_k'])"
python bench.py --steps 5 --warmup 3 --no-cpu --workload kitti > gpurun_out/${TAG}_bench_kitti.json 2>> gpurun_out/${TAG}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_kitti.json')); print('kitti', d['value'], d['e2e']['value'])"
tail -5 gpurun_out/${TAG}_bench.err
