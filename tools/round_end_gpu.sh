#!/bin/bash
# Round-end GPU pass (run under gpurun): full GPU test suite, the two single-GPU bench lines, an ncu launch list of
# the rod case and a memcheck pass over the kernels added last. Everything lands in gpurun_out/.
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -q -m gpu -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 300 gpurun_out/bench_c2.err
timeout 200 python bench.py --workload c3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -c 600 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 \
    --no-cpu-baseline > gpurun_out/ncu_bench_c3.log 2>&1
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu -x \
    "tests/test_filter_fused.py::test_fused_convolution_filter_vector_and_views" \
    "tests/test_rod_forcing_grids.py::test_cuda_rod_grids_match_restatement_on_random_rods" \
    "tests/test_fastdiag.py::test_cuda_fastdiag_matches_oracle" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/memcheck.log
python tools/show_bench.py gpurun_out/bench_c2.json gpurun_out/bench_c3.json 2>/dev/null | tail -60
