#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU): launch list of a short bench run, then `ncu --set full`
# captures of the tensor-path kernels on config C and D/8; the reports land in gpurun_out/ and are summarised into
# profiles/ with profiles/ncu_summary.py.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/r02_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-configs --no-mf-vmult > gpurun_out/r02_ncu_launches.log 2>&1
python tools/time_assembly.py C > gpurun_out/r02_plain_C.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"k_cart|k_brick" -s 8 -c 5 -o gpurun_out/r02_cart_C -f \
    python tools/time_assembly.py C > gpurun_out/r02_ncu_C.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cart|k_brick" -s 8 -c 5 -o gpurun_out/r02_cart_D8 -f \
    python tools/time_assembly.py D8 > gpurun_out/r02_ncu_D8.log 2>&1
