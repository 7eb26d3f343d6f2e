#!/bin/bash
mkdir -p gpurun_out
SCAN_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ii_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline --sustained 0 > gpurun_out/ii_ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/ii_launches.csv 70 > gpurun_out/ii_launches_step_n8.txt; head -14 gpurun_out/ii_launches_step_n8.txt; grep -n "at::\|void at" gpurun_out/ii_launches_step_n8.txt | head; tail -1 gpurun_out/ii_launches_step_n8.txt
for i in 1 2 3; do timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --sustained 0 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"; done
