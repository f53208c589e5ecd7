#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_simulator_gpu.py tests/test_cuda_ib.py tests/test_slab_gpu.py tests/test_cuda_parity.py -q -m gpu -x ) > gpurun_out/pytest_check3.log 2>&1
tail -4 gpurun_out/pytest_check3.log
timeout 120 python tools/ib_rate.py 2>&1 | tee gpurun_out/ib_rate.txt
for wl in u512 c2; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 300 gpurun_out/bench_$wl.err
done
python tools/show_bench.py gpurun_out/bench_u512.json gpurun_out/bench_c2.json
