#!/bin/bash
# small-grid pass: IB tests (privatised spread), simulator tests, bench lines c1 / c2 / c3 with and without CUDA graphs
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_cuda_ib.py tests/test_simulator_gpu.py tests/test_forcing_grids.py tests/test_rod_forcing_grids.py tests/test_drag_gpu.py tests/test_slab_gpu.py -q -m gpu -x ) > gpurun_out/pytest_small.log 2>&1
tail -6 gpurun_out/pytest_small.log
for wl in c1 c2 c3; do
  timeout 300 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -c 300 gpurun_out/bench_$wl.err
  timeout 300 python bench.py --workload $wl --graph off --no-cpu-baseline --no-parity > gpurun_out/bench_${wl}_eager.json 2> gpurun_out/bench_${wl}_eager.err
done
SOPHT_IB_SPREAD=atomics timeout 300 python bench.py --workload c2 --no-cpu-baseline --no-parity > gpurun_out/bench_c2_atomics.json 2> gpurun_out/bench_c2_atomics.err
python tools/show_bench.py gpurun_out/bench_c1.json gpurun_out/bench_c1_eager.json gpurun_out/bench_c2.json gpurun_out/bench_c2_eager.json gpurun_out/bench_c2_atomics.json gpurun_out/bench_c3.json gpurun_out/bench_c3_eager.json 2>&1 | grep -v "cpu_baseline"
