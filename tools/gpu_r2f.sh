#!/bin/bash
# round 2, GPU call F: full GPU test suite + e2e diagnosis
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=1800 TAILN=25 run r2f_all python -m pytest tests -m gpu -q --tb=short --durations=6
TMO=600 TAILN=3 run r2f_bench5 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline
TMO=600 TAILN=3 run r2f_bench10 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline
TMO=300 TAILN=5 run r2f_smoke python -c "import __graft_entry__ as g; g.smoke()"
