#!/bin/bash
# memcpy probe + fused-kernel throughput for every window size that has a fused kernel (1080p, T=1000)
set -u
mkdir -p gpurun_out
./tools/memcpy_probe | tee gpurun_out/r02q_memcpy_probe.txt
for w in 1 2 3 4 5 6 7 8 9 11; do
  python bench.py --steps 3 --warmup 3 --window $w --no-cpu --no-slab 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'window': d['config']['window'], 'k': d['config'].get('temporal_k'), 'gpixit_s': round(d['value']/1e3,1), 'e2e_gpixit_s': round(d['e2e']['value']/1e3,1), 'kernel': d['roofline'].get('kernel')}))
" | tee -a gpurun_out/r02q_windows_1080p.jsonl
done
