#!/bin/bash
# call M: tower convolution forward re-check + weight-gradient kernel, timing next to cuDNN
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv3x3_rows" 2>&1 | tail -15 > gpurun_out/m_fwd.log
tail -4 gpurun_out/m_fwd.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "conv3x3_wgrad" 2>&1 | tail -25 > gpurun_out/m_wgrad.log
tail -12 gpurun_out/m_wgrad.log | cut -c1-250
timeout 300 python tools/bench_conv.py 16 > gpurun_out/m_bench_conv.log 2>&1
tail -12 gpurun_out/m_bench_conv.log
