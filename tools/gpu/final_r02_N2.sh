#!/bin/bash
# Two B200 of one box: bench line (METIS-sharded, fused fine-mesh exchange) and the distributed check under torchrun.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_N2.json 2> gpurun_out/bench_r02_N2.err
echo "bench N2 rc $?"
tail -c 300 gpurun_out/bench_r02_N2.err
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    tests/run_distributed_check.py > gpurun_out/dist_check_N2.txt 2>&1
echo "dist check rc $?"
tail -5 gpurun_out/dist_check_N2.txt
