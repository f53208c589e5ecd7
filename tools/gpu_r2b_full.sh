#!/bin/bash
# Round-2 single-GPU pass: full GPU suite, default bench line, ncu launch list of the bench, full captures of the z pass
# (warp-quartet kernel) and the y inverse pass.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=8 ) > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -14 gpurun_out/r2b_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2b_bench_u512.json 2> gpurun_out/r2b_bench_u512.err; tail -c 400 gpurun_out/r2b_bench_u512.err
python tools/show_bench.py gpurun_out/r2b_bench_u512.json 2>/dev/null | tail -30
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 400 --csv --log-file gpurun_out/r2b_launches_u512.csv python bench.py --steps 2 --warmup 3 \
    --no-cpu-baseline --no-parity > gpurun_out/r2b_ncu_bench_u512.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:zquad_kernel -c 1 -o gpurun_out/r2b_zquad -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_zquad.log 2>&1
ncu -i gpurun_out/r2b_zquad.ncu-rep --page raw --csv > gpurun_out/r2b_zquad_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_zquad.ncu-rep --page source --csv > gpurun_out/r2b_zquad_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:YInv -c 1 -o gpurun_out/r2b_yinv -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_yinv.log 2>&1
ncu -i gpurun_out/r2b_yinv.ncu-rep --page raw --csv > gpurun_out/r2b_yinv_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_yinv.ncu-rep --page source --csv > gpurun_out/r2b_yinv_source.csv 2>/dev/null
rm -f gpurun_out/r2b_zrow_m0.ncu-rep gpurun_out/r2b_zrow_m3.ncu-rep
