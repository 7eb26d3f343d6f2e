#!/bin/bash
# round 2, GPU call A: bench-size parity tests + baseline bench of the tree
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
nproc; free -g | head -2
TMO=900 TAILN=40 run r2a_fullsize python -m pytest tests/test_gpu_fullsize.py -q --tb=short --durations=8
TMO=1200 TAILN=50 run r2a_module python -m pytest tests/test_gpu_module.py -q --tb=short --durations=8
TMO=600 TAILN=15 run r2a_kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=600 TAILN=3 run r2a_bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
