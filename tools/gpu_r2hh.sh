#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=600 TAILN=12 run hh_kern python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -k "alias_chain or head_out_levels or tower_convolution or condconv"
TMO=1200 TAILN=12 run hh_module python -m pytest tests/test_gpu_module.py tests/test_gpu_dist.py -q -m gpu --tb=short
TMO=900 TAILN=1 run hh_bench_n1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/hh_bench_n1.log") if x.startswith("{")][-1])
print(round(d["value"],1), round(d["ms_per_step"],2), round(d["e2e"]["value"],1), d["clocks"], d.get("sustained") and round(d["sustained"]["value"],1), d["gpu_launches"])
PY
