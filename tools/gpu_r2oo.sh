#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-200; }
TMO=1800 TAILN=3 run oo_gpu_tests python -m pytest tests -x -q -m gpu
TMO=300 TAILN=2 run oo_smoke python -c "import __graft_entry__ as g; g.smoke()"
