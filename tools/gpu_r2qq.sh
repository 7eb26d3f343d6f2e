#!/bin/bash
# round 2, call QQ: whole suite on the tree with the GroupNorm changes; if it is not green, the module tests with the separate statistics pass
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "rc=$rc"; tail -n "${TAILN:-6}" gpurun_out/$name.log | cut -c1-300; return $rc; }
if ! TMO=900 TAILN=14 run qq_gpu_tests python -m pytest tests -q -m gpu --tb=short; then
  SCAN_B200_GN_STATS=0 TMO=600 TAILN=14 run qq_module_sep python -m pytest tests/test_gpu_module.py -q -m gpu --tb=short
fi
TMO=300 TAILN=2 run qq_smoke python -c "import __graft_entry__ as g; g.smoke()"
