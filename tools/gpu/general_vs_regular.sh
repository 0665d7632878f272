#!/bin/bash
# k_fine_stream: the general instantiation forced on regular tiles against the regular one (one GPU, config F2)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 40 python tools/run_config.py F2 --steps 10 > gpurun_out/f2_regular.json 2>&1
PD_FINE_GENERAL=1 timeout 40 python tools/run_config.py F2 --steps 10 > gpurun_out/f2_general.json 2>&1
grep -h -o '"mf_vmult_ms": [0-9.]*' gpurun_out/f2_regular.json gpurun_out/f2_general.json
