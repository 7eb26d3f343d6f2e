#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=900 TAILN=4 run tests python -m pytest tests -q -m gpu --tb=short
TMO=900 TAILN=2 run bench python bench.py --steps 10 --warmup 3
TMO=900 TAILN=2 run bench_ref python bench.py --impl reference --steps 2 --warmup 1
TMO=600 TAILN=40 run microbench python tools/bench_kernels.py 8
