#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_cuda_ib.py tests/test_slab_gpu.py -q -m gpu -x ) > gpurun_out/pytest_small.log 2>&1
tail -4 gpurun_out/pytest_small.log
{ timeout 120 python tools/ib_rate.py; SOPHT_IB_SPREAD=atomics timeout 120 python tools/ib_rate.py; } 2>&1 | tee gpurun_out/ib_rate.txt
for wl in c1 c2 c3; do
  timeout 300 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 300 gpurun_out/bench_$wl.err
done
timeout 300 python bench.py --workload c2 --graph off --no-cpu-baseline --no-parity > gpurun_out/bench_c2_eager.json 2> gpurun_out/bench_c2_eager.err
python tools/show_bench.py gpurun_out/bench_c1.json gpurun_out/bench_c2.json gpurun_out/bench_c2_eager.json gpurun_out/bench_c3.json 2>&1 | grep -E "Gcell|ib\."
