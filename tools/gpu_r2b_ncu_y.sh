#!/bin/bash
mkdir -p gpurun_out
for k in YInv YFwd; do
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 2 -c 1 -o gpurun_out/r2b_$k -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_$k.log 2>&1
ncu -i gpurun_out/r2b_$k.ncu-rep --page raw --csv > gpurun_out/r2b_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_$k.ncu-rep --page source --csv > gpurun_out/r2b_${k}_source.csv 2>/dev/null
done
timeout 200 python bench.py --workload c3 > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
python tools/show_bench.py gpurun_out/r2b_bench_c3.json | tail -30
ls -la gpurun_out | grep -E "r2b_Y"
