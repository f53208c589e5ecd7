#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): slab parity checks, the weak-scaled bench lines (1024^3), fused vs pipelined transposes
N=${1:-8}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/mgpu_slab_check.py 128 64 128 2 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee -a gpurun_out/slab_check_n$N.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tests/mgpu_periodic_check.py 128 64 256 3 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee -a gpurun_out/slab_check_n$N.log
SOPHT_SLAB_PIPELINE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 \
    tests/mgpu_slab_check.py 128 64 128 2 2>&1 | grep -E "SLAB CHECK|rror" | sed 's/^/[pipelined] /' | tee -a gpurun_out/slab_check_n$N.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --workload u512 --steps 10 --warmup 3 > gpurun_out/bench_u512_n$N.json 2> gpurun_out/bench_u512_n$N.err
tail -c 300 gpurun_out/bench_u512_n$N.err
SOPHT_SLAB_PIPELINE=1 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --gpus $N --workload u512 --steps 10 --warmup 3 --no-parity > gpurun_out/bench_u512_pipe_n$N.json 2> gpurun_out/bench_u512_pipe_n$N.err
tail -c 300 gpurun_out/bench_u512_pipe_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 \
    bench.py --gpus $N --workload tg512 --steps 10 --warmup 3 > gpurun_out/bench_tg512_n$N.json 2> gpurun_out/bench_tg512_n$N.err
tail -c 300 gpurun_out/bench_tg512_n$N.err
python tools/show_bench.py gpurun_out/bench_u512_n$N.json gpurun_out/bench_u512_pipe_n$N.json gpurun_out/bench_tg512_n$N.json
