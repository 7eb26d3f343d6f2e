#!/bin/bash
# BASELINE configs[4] per-GPU batches: 16 + 16 images (4 GPUs) and 32 + 32 images (2 GPUs) on one GPU
mkdir -p gpurun_out
for n in 16 32; do
  timeout 900 python bench.py --images $n --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/nn_bench_$n.log 2> gpurun_out/nn_bench_$n.err; echo "images=$n rc=$?"
  tail -2 gpurun_out/nn_bench_$n.err | cut -c1-300
  python - <<PY
import json
l=[x for x in open("gpurun_out/nn_bench_$n.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print($n, round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["source_nodes"], d["target_nodes"], d["dbscan_points_per_level"])
PY
done
