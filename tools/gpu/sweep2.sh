#!/bin/bash
cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fine_mesh" > gpurun_out/sw9_tests.log 2>&1
echo "rc $?" >> gpurun_out/sw9_tests.log
tail -5 gpurun_out/sw9_tests.log
rm -f gpurun_out/sw9_times.log
for cfg in F2 F2L; do
 for v in "stream 2 3" "stream 3 6" "stream 4 6" "stream 5 6"; do
  set -- $v
  echo "== $cfg kernel=$1 groups=$2 stages=$3" >> gpurun_out/sw9_times.log
  PD_FINE_KERNEL=$1 PD_FINE_GROUPS=$2 PD_FINE_STAGES=$3 timeout 300 python tools/run_config.py $cfg --steps 10 >> gpurun_out/sw9_times.log 2>&1
 done
done
grep -E "==|mf_vmult_ms" gpurun_out/sw9_times.log | sed 's/.*"fine_kernel_last": \([0-9]\).*"mf_vmult_ms": \([0-9.]*\), "mf_vmult_gdofs": \([0-9.]*\).*/k=\1 ms=\2 gdofs=\3/'
