#!/bin/bash
# call L: tower convolution kernel — correctness (1-CTA, then CTA pairs), then timing next to cuDNN
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "conv3x3 and -1]" 2>&1 | tail -15 > gpurun_out/l_cg1.log; echo "cg1 rc=$?" >> gpurun_out/l_cg1.log
tail -8 gpurun_out/l_cg1.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "conv3x3 and -2]" 2>&1 | tail -15 > gpurun_out/l_cg2.log; echo "cg2 rc=$?" >> gpurun_out/l_cg2.log
tail -8 gpurun_out/l_cg2.log
timeout 300 python tools/bench_conv.py 16 > gpurun_out/l_bench_conv.log 2>&1
cat gpurun_out/l_bench_conv.log | tail -12
