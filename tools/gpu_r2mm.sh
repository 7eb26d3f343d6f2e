#!/bin/bash
# A/B on one box: tower convolutions on cuDNN (SCAN_B200_TOWERS=cudnn) vs the tcgen05 kernels, same bench.py, stationary workload
mkdir -p gpurun_out
for impl in cudnn scan cudnn scan; do
  SCAN_B200_TOWERS=$impl timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --sustained 2 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$impl', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'sustained', round(d['sustained']['value'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done | tee gpurun_out/mm_ab.txt
