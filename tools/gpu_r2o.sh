#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_wgrad|conv3x3_kernel" -c 6 -o gpurun_out/o_conv python tools/prof_conv.py 16 > gpurun_out/o_ncu.log 2>&1
ncu -i gpurun_out/o_conv.ncu-rep --page raw --csv > gpurun_out/o_conv_raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/o_conv_raw.csv | tail -80
