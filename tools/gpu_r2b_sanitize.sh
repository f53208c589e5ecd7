#!/bin/bash
# memcheck + racecheck of the warp-quartet z pass on a small 2 nz = 1024 grid
mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/poisson_only.py 512 16 16 1 > gpurun_out/r2b_memcheck_zquad.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2b_memcheck_zquad.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/poisson_only.py 512 16 16 1 > gpurun_out/r2b_racecheck_zquad.log 2>&1
echo "racecheck rc=$?"; tail -6 gpurun_out/r2b_racecheck_zquad.log
