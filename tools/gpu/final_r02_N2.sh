#!/bin/bash
# Two B200 of one box: the distributed check under torchrun (incl. the forced-split fused plan), then the bench line
# (METIS-sharded, fused fine-mesh exchange).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    tests/run_distributed_check.py > gpurun_out/dist_check_N2.txt 2>&1
echo "dist check rc $?"
tail -7 gpurun_out/dist_check_N2.txt
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_N2.json 2> gpurun_out/bench_r02_N2.err
echo "bench N2 rc $?"
tail -c 300 gpurun_out/bench_r02_N2.err
