#!/bin/bash
mkdir -p gpurun_out
{
echo "== zquad, two barriers per unit: 512^3"; timeout 120 python tools/poisson_only.py 512 512 512 5
for g in 32 128; do echo "== L2 fetch granularity $g"; SOPHT_L2_FETCH=$g timeout 120 python tools/poisson_only.py 512 512 512 5; done
} 2>&1 | tee gpurun_out/r2b_batch3_timings.txt
timeout 120 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "pow2_path_vs_oracle" 2>&1 | tail -2
for k in YInv; do
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 2 -c 2 -o gpurun_out/r2b_$k -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_$k.log 2>&1
ncu -i gpurun_out/r2b_$k.ncu-rep --page raw --csv > gpurun_out/r2b_${k}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_$k.ncu-rep --page source --csv > gpurun_out/r2b_${k}_source.csv 2>/dev/null
done
