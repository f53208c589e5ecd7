#!/bin/bash
mkdir -p gpurun_out


for wl in c2 c3 u512; do
for pdl in 0 1; do
  SOPHT_PDL=$pdl timeout 300 python bench.py --workload $wl --no-cpu-baseline --no-parity > gpurun_out/r2b_pdl${pdl}_$wl.json 2> gpurun_out/r2b_pdl${pdl}_$wl.err
  echo "PDL=$pdl $wl: $(python tools/show_bench.py gpurun_out/r2b_pdl${pdl}_$wl.json 2>/dev/null | head -1 | cut -c1-120)"
done
done 2>&1 | tee gpurun_out/r2b_pdl_timings.txt
