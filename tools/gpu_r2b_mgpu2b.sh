#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out

for args in "512 32 64 2" "128 64 128 2"; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tests/mgpu_slab_check.py $args 2>&1 | grep -E "slab check|SLAB CHECK|rror" | tee -a gpurun_out/r2b_slab_check_fused_n$N.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 \
    bench.py --gpus $N --workload u512 --steps 10 --warmup 3 > gpurun_out/r2b_bench_u512_n$N.json 2> gpurun_out/r2b_bench_u512_n$N.err
tail -c 300 gpurun_out/r2b_bench_u512_n$N.err
python tools/show_bench.py gpurun_out/r2b_bench_u512_n$N.json | head -8
