#!/bin/bash
# Last GPU pass of round 2 (one B200): GPU test suite, ncu --set full of the pipelined fine-mesh kernel, bench line.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/gputest_r02.txt 2>&1
echo "pytest rc $?" >> gpurun_out/gputest_r02.txt
tail -3 gpurun_out/gputest_r02.txt
timeout 150 bash tools/gpu/ncu_stream.sh 3 6 > gpurun_out/ncu_stream_sh.log 2>&1
echo "ncu rc $?"
rm -f gpurun_out/r02_fine_stream.ncu-rep   # keep the csv + summary (the report is large)
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
echo "bench rc $?"
head -c 600 gpurun_out/bench_r02.json
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
echo "ref rc $?"
