#!/bin/bash
# Full GPU pass (run under gpurun): the GPU test suite, the default bench line (u512) and its reference arm.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
tail -14 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_u512.json 2> gpurun_out/bench_u512.err; tail -c 400 gpurun_out/bench_u512.err
python tools/show_bench.py gpurun_out/bench_u512.json 2>/dev/null | tail -40
