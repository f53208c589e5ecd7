#!/bin/bash
# multi-GPU pass 2 (gpurun --gpus N): slab parity checks with the pipelined solve, weak-scaled bench lines, fused-transpose
# variant for comparison
N=${1:-2}
mkdir -p gpurun_out
for args in "32 16 64 3" "128 64 128 2"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/mgpu_slab_check.py $args 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee -a gpurun_out/slab_check_pipe_n$N.log
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tests/mgpu_periodic_check.py 128 64 256 3 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee -a gpurun_out/slab_check_pipe_n$N.log
for wl in u512 tg512; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  tail -c 400 gpurun_out/bench_${wl}_n$N.err
done
SOPHT_SLAB_PIPELINE=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --workload u512 --steps 10 --warmup 3 --no-parity > gpurun_out/bench_u512_fused_n$N.json 2> gpurun_out/bench_u512_fused_n$N.err
python tools/show_bench.py gpurun_out/bench_u512_n$N.json gpurun_out/bench_u512_fused_n$N.json gpurun_out/bench_tg512_n$N.json
