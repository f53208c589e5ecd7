#!/bin/bash
# z-pass experiment pass (run under gpurun): parity of the Poisson paths, solve timings with the row-mode z kernel on / off,
# a sweep of the group stagger, one ncu capture of the new kernel.
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_cuda_parity.py tests/test_slab_gpu.py tests/test_periodic_poisson.py -q -m gpu -x -k "poisson or slab or periodic" ) > gpurun_out/pytest_zrow.log 2>&1
tail -3 gpurun_out/pytest_zrow.log
{
for s in 0 3000; do
  echo "== zrow stagger $s: 512^3"; SOPHT_P2_ZROW_STAGGER=$s timeout 120 python tools/poisson_only.py 512 512 512 5
done
echo "== zrow off: 512^3"; SOPHT_P2_ZROW=0 timeout 120 python tools/poisson_only.py 512 512 512 5
for g in "256 256 256" "128 128 256"; do
  echo "== zrow on: $g";  timeout 120 python tools/poisson_only.py $g 5
  echo "== zrow off: $g"; SOPHT_P2_ZROW=0 timeout 120 python tools/poisson_only.py $g 5
done
} 2>&1 | grep -v "^poisson" | tee gpurun_out/zrow_timings.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:zrow_kernel -c 1 -o gpurun_out/zrow_512 -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/ncu_zrow.log 2>&1
ncu -i gpurun_out/zrow_512.ncu-rep --page raw --csv > gpurun_out/zrow_512_raw.csv 2>/dev/null
ncu -i gpurun_out/zrow_512.ncu-rep --page source --csv > gpurun_out/zrow_512_source.csv 2>/dev/null
