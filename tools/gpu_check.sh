#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-12}" gpurun_out/$name.log; }
TMO=900 TAILN=6 run gpu_tests python -m pytest tests -m gpu -q --tb=short
TMO=300 TAILN=2 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
TMO=900 TAILN=1 run bench python bench.py --steps 10 --warmup 3 --no-cpu-baseline
