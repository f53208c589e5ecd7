#!/bin/bash
# warp-quartet z pass: parity, timings against the row kernel, ncu capture
mkdir -p gpurun_out
( time timeout 500 python -m pytest tests/test_cuda_parity.py tests/test_simulator_gpu.py tests/test_slab_gpu.py -q -m gpu -x -k "poisson or slab or step" ) > gpurun_out/r2b_pytest_zquad.log 2>&1
tail -3 gpurun_out/r2b_pytest_zquad.log
{
echo "== zrow (SOPHT_P2_ZQUAD=0): 512^3"; SOPHT_P2_ZQUAD=0 timeout 120 python tools/poisson_only.py 512 512 512 5
echo "== zquad packed: 512^3"; timeout 120 python tools/poisson_only.py 512 512 512 5

} 2>&1 | tee gpurun_out/r2b_zquad_timings.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:zquad_kernel -c 1 -o gpurun_out/r2b_zquad -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_zquad.log 2>&1
ncu -i gpurun_out/r2b_zquad.ncu-rep --page raw --csv > gpurun_out/r2b_zquad_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_zquad.ncu-rep --page source --csv > gpurun_out/r2b_zquad_source.csv 2>/dev/null
