#!/bin/bash
# Packed-FP32 FFT butterflies + z-pass modes (run under gpurun): parity per mode, 512^3 solve timings per mode and for the
# scalar build, ncu captures of two modes.
mkdir -p gpurun_out
for m in 0 1 2 3; do
  echo "== parity mode $m"
  SOPHT_P2_ZROW_MODE=$m timeout 300 python -m pytest tests/test_cuda_parity.py -q -m gpu -x -k "pow2_path_vs_oracle" 2>&1 | tail -2
done > gpurun_out/r2b_parity_modes.log 2>&1
cat gpurun_out/r2b_parity_modes.log
( time timeout 500 python -m pytest tests/test_cuda_parity.py tests/test_periodic_poisson.py tests/test_simulator_gpu.py -q -m gpu -x -k "poisson or fft" ) > gpurun_out/r2b_pytest_fft.log 2>&1
tail -3 gpurun_out/r2b_pytest_fft.log
{
echo "== scalar build: 512^3"; SOPHT_B200_LIB=$PWD/sopht_b200/lib/libsopht_b200_scalar.so timeout 120 python tools/poisson_only.py 512 512 512 5
for m in 0 1 2 3; do
  echo "== packed, zrow mode $m: 512^3"; SOPHT_P2_ZROW_MODE=$m timeout 120 python tools/poisson_only.py 512 512 512 5
done
for g in "256 256 256" "128 128 256"; do
  echo "== scalar: $g"; SOPHT_B200_LIB=$PWD/sopht_b200/lib/libsopht_b200_scalar.so timeout 120 python tools/poisson_only.py $g 5
  echo "== packed: $g";  timeout 120 python tools/poisson_only.py $g 5
  echo "== packed zrow512: $g";  SOPHT_P2_ZROW_512=1 timeout 120 python tools/poisson_only.py $g 5
done
echo "== periodic 512 scalar"; SOPHT_B200_LIB=$PWD/sopht_b200/lib/libsopht_b200_scalar.so timeout 120 python tools/time_solvers.py 2>&1 | tail -8
echo "== periodic 512 packed"; timeout 120 python tools/time_solvers.py 2>&1 | tail -8
} 2>&1 | tee gpurun_out/r2b_zrow_timings.txt
for m in 0 3; do
timeout 300 env SOPHT_P2_ZROW_MODE=$m ncu --set full --clock-control none --import-source on -k regex:zrow_kernel -c 1 -o gpurun_out/r2b_zrow_m$m -f \
  python tools/poisson_only.py 512 512 512 1 > gpurun_out/r2b_ncu_zrow_m$m.log 2>&1
ncu -i gpurun_out/r2b_zrow_m$m.ncu-rep --page raw --csv > gpurun_out/r2b_zrow_m${m}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2b_zrow_m$m.ncu-rep --page source --csv > gpurun_out/r2b_zrow_m${m}_source.csv 2>/dev/null
done
ls -la gpurun_out | grep r2b
