#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_ -s 14 -c 7 -o gpurun_out/r2b_filter -f \
  python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --graph off > gpurun_out/r2b_ncu_filter.log 2>&1
ncu -i gpurun_out/r2b_filter.ncu-rep --page raw --csv > gpurun_out/r2b_filter_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_filter.ncu-rep --page source --csv > gpurun_out/r2b_filter_source.csv 2>/dev/null
tail -3 gpurun_out/r2b_ncu_filter.log | cut -c1-300
