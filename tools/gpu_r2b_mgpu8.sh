#!/bin/bash
# 8-GPU closing pass: slab parity (unbounded + periodic), the 1024^3 bench lines
N=8
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/mgpu_slab_check.py 128 64 128 2 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee gpurun_out/r2b_slab_check_n8.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 \
    tests/mgpu_periodic_check.py 128 64 256 3 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee -a gpurun_out/r2b_slab_check_n8.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --workload u512 --steps 10 --warmup 3 > gpurun_out/r2b_bench_u512_n8.json 2> gpurun_out/r2b_bench_u512_n8.err
tail -c 300 gpurun_out/r2b_bench_u512_n8.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29546 \
    bench.py --gpus $N --workload tg512 --steps 10 --warmup 3 > gpurun_out/r2b_bench_tg512_n8.json 2> gpurun_out/r2b_bench_tg512_n8.err
tail -c 300 gpurun_out/r2b_bench_tg512_n8.err
python tools/show_bench.py gpurun_out/r2b_bench_u512_n8.json 2>/dev/null | head -18
python tools/show_bench.py gpurun_out/r2b_bench_tg512_n8.json 2>/dev/null | head -1
