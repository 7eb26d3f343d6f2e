#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=600 TAILN=25 run kernels python -m pytest tests/test_gpu_kernels.py -q --tb=short
TMO=600 TAILN=60 run microbench python tools/bench_kernels.py 8
TMO=600 TAILN=3 run bench python bench.py --steps 5 --warmup 3 --no-cpu-baseline
