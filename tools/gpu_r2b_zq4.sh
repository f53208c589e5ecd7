#!/bin/bash
mkdir -p gpurun_out
SOPHT_P2_ZQUAD=2 timeout 200 python -m pytest tests/test_cuda_parity.py tests/test_simulator_gpu.py -q -m gpu -x -k "pow2_path_vs_oracle or poisson_full_size" 2>&1 | tail -2
{
echo "== zquad (3 quartets + producers): 512^3"; SOPHT_P2_ZQUAD=1 timeout 120 python tools/poisson_only.py 512 512 512 5
echo "== zquad4 (4 quartets): 512^3"; SOPHT_P2_ZQUAD=2 timeout 120 python tools/poisson_only.py 512 512 512 5
} 2>&1 | tee gpurun_out/r2b_zq4_timings.txt
timeout 300 env SOPHT_P2_ZQUAD=2 ncu --set full --clock-control none --import-source on -k regex:zquad4_kernel -c 1 -o gpurun_out/r2b_zquad4 -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_zquad4.log 2>&1
ncu -i gpurun_out/r2b_zquad4.ncu-rep --page raw --csv > gpurun_out/r2b_zquad4_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_zquad4.ncu-rep --page source --csv > gpurun_out/r2b_zquad4_source.csv 2>/dev/null
