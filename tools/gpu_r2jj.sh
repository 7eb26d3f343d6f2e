#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; }
TMO=600 TAILN=8 run jj_full python -m pytest tests/test_gpu_fullsize.py -q -m gpu --tb=short -k "cka or tower"
TMO=600 TAILN=8 run jj_next python tools/bench_next_rows.py
