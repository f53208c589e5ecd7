#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_filter_fused.py tests/test_cuda_parity.py -q -m gpu -x -k "filter" 2>&1 | tail -2
timeout 200 python bench.py --workload c3 --no-cpu-baseline --no-parity > gpurun_out/r2b_bench_c3_f.json 2> gpurun_out/r2b_bench_c3_f.err
python tools/show_bench.py gpurun_out/r2b_bench_c3_f.json | grep -E "Gcell|filter"
