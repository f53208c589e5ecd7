#!/bin/bash
# periodic (config 4) pass: tests, solver timings (pow2 pipeline vs cuFFT path), bench lines tg256 / tg512
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_periodic_poisson.py -q -m gpu -x ) > gpurun_out/pytest_periodic.log 2>&1
tail -5 gpurun_out/pytest_periodic.log
timeout 300 python bench.py --workload tg512 --steps 10 > gpurun_out/bench_tg512.json 2> gpurun_out/bench_tg512.err; tail -c 300 gpurun_out/bench_tg512.err
timeout 300 python bench.py --workload tg256 --steps 20 --no-cpu-baseline > gpurun_out/bench_tg256.json 2> gpurun_out/bench_tg256.err
SOPHT_PERIODIC_FORCE_CUFFT=1 timeout 300 python bench.py --workload tg512 --steps 10 --no-cpu-baseline --no-parity > gpurun_out/bench_tg512_cufft.json 2> gpurun_out/bench_tg512_cufft.err
python tools/show_bench.py gpurun_out/bench_tg512.json gpurun_out/bench_tg256.json gpurun_out/bench_tg512_cufft.json | grep -v cpu_baseline
