#!/bin/bash
# quick perf check after a kernel change: bit-identity subset, benches, tile phase profile
set -u
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q -k "fused or fuzz or dataflow or textbook or slab or kitti" 2>&1 | tail -3
for args in "--window 3" "--window 5" "--workload 4k" "--workload kitti"; do
python bench.py --steps 5 --warmup 3 $args --no-cpu --no-slab 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$args', round(d['value']/1e3,1), 'k', d['config']['temporal_k'], 'e2e', round(d['e2e']['value']/1e3,1))
"
done
python tools/tile_profile.py --build && python tools/tile_profile.py 2>&1 | tail -5 | tee gpurun_out/${TAG}_tile_profile.txt
