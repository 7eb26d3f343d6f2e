#!/bin/bash
# round 2, GPU call I: reordered source pass / overlapped head_out, full GPU suite, bench
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=1800 TAILN=12 run r2i_all python -m pytest tests -m gpu -q --tb=short
TMO=900 TAILN=3 run r2i_bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
TMO=900 TAILN=3 run r2i_bench_b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
TMO=600 TAILN=3 run r2i_bench_sim10k python bench.py --steps 10 --warmup 3 --config sim10k --no-cpu-baseline --no-eager-baseline
