#!/bin/bash
mkdir -p gpurun_out
{
for zb in 3 1 2 4 5 6; do
  echo "== zb_shift $zb"; SOPHT_P2_ZB_SHIFT=$zb timeout 120 python tools/poisson_only.py 512 512 512 5 | head -1
done
echo "== stage=1 minb=1"; SOPHT_P2_STAGE=1 SOPHT_P2_MINB=1 timeout 120 python tools/poisson_only.py 512 512 512 5 | head -1
echo "== stage=0 minb=1"; SOPHT_P2_STAGE=0 SOPHT_P2_MINB=1 timeout 120 python tools/poisson_only.py 512 512 512 5 | head -1
} 2>&1 | tee gpurun_out/r2b_yinv_sweep.txt
nvidia-smi --query-gpu=name,serial,uuid,pci.bus_id --format=csv | tee -a gpurun_out/r2b_yinv_sweep.txt
