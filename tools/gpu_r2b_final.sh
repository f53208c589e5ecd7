#!/bin/bash
# Round-2 closing pass on one GPU: the full GPU suite and one bench line per workload.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -q -m gpu -x --durations=5 ) > gpurun_out/r2b_pytest_gpu.log 2>&1
tail -9 gpurun_out/r2b_pytest_gpu.log
for wl in u512 c2; do
  timeout 400 python bench.py --workload $wl > gpurun_out/r2b_bench_$wl.json 2> gpurun_out/r2b_bench_$wl.err
  python tools/show_bench.py gpurun_out/r2b_bench_$wl.json 2>/dev/null | head -1 | cut -c1-110
done
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err
cut -c1-400 gpurun_out/r2b_bench_reference.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
