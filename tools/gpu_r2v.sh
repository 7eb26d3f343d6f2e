#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -s -k "cka" 2>&1 | tail -40 > gpurun_out/v_cka.log
tail -30 gpurun_out/v_cka.log | cut -c1-300
