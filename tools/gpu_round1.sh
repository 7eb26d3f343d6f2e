#!/bin/bash
# One gpurun call: kernel tests, module tests (FFMA cond-conv, then tcgen05), each in its own process + timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout -k 10 "$TMO" "$@" > gpurun_out/$name.log 2>&1; echo "rc=$?"; tail -n "${TAILN:-15}" gpurun_out/$name.log; }
TMO=600 run kernels_other python -m pytest tests/test_gpu_kernels.py -q -k "not condconv" --tb=short
TMO=300 TAILN=25 run kernels_condconv python -m pytest tests/test_gpu_kernels.py -q -k "condconv" --tb=short
SCAN_B200_CONDCONV_IMPL=1 TMO=900 run module_simt python -m pytest tests/test_gpu_module.py -q --tb=short
TMO=900 run module_tc python -m pytest tests/test_gpu_module.py -q --tb=short
