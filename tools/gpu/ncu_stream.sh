#!/bin/bash
# ncu --set full capture of the pipelined fine-mesh kernel on config F2 (64^3, DGQ2); summary into gpurun_out/
cd "$(dirname "$0")/../.."
G=${1:-3}; S=${2:-6}
PD_FINE_GROUPS=$G PD_FINE_STAGES=$S python tools/run_config.py F2 --steps 5 > gpurun_out/ncu_stream_plain.log 2>&1 || exit 1
PD_FINE_GROUPS=$G PD_FINE_STAGES=$S ncu --set full --clock-control none --import-source on -k regex:"k_fine_stream" -s 3 -c 2 -o gpurun_out/r02_fine_stream -f \
    python tools/run_config.py F2 --steps 5 > gpurun_out/ncu_stream.log 2>&1
ncu -i gpurun_out/r02_fine_stream.ncu-rep --page raw --csv > gpurun_out/r02_fine_stream_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/r02_fine_stream_raw.csv > gpurun_out/r02_fine_stream_summary.txt
cat gpurun_out/r02_fine_stream_summary.txt | head -40
