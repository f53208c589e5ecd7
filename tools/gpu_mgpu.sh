#!/bin/bash
# multi-GPU pass (run under gpurun --gpus N): slab parity checks (unbounded + periodic) and the weak-scaled bench lines
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 600 python -m pytest tests/test_slab_gpu.py -q -m gpu -x -rs ) > gpurun_out/pytest_slab_n$N.log 2>&1
tail -12 gpurun_out/pytest_slab_n$N.log
for args in "32 16 64 3" "128 64 128 2"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/mgpu_slab_check.py $args 2>&1 | grep -E "slab check|SLAB CHECK" | tee -a gpurun_out/slab_check_n$N.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tests/mgpu_periodic_check.py 128 64 256 3 2>&1 | grep -E "slab check|SLAB CHECK" | tee -a gpurun_out/slab_check_n$N.log
for wl in u512 tg512; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  tail -c 600 gpurun_out/bench_${wl}_n$N.err
done
python tools/show_bench.py gpurun_out/bench_u512_n$N.json gpurun_out/bench_tg512_n$N.json
