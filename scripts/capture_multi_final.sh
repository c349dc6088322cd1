#!/bin/bash
# On the GPU box: ncu captures of the MultiSnake step kernels (three state modes), extracted there (the reports are ~11 MB
# each and gpurun brings back at most 64 MiB).  Results under gpurun_out/.  LISTS=1 adds the launch lists.
WORKLOADS="C4 C5" bash scripts/capture_traffic.sh 2>&1 | tail -6
python scripts/update_traffic.py 2>&1 | grep -E "^C[45]" | cut -c1-160
cp profiles/traffic.json gpurun_out/traffic.json
cp profiles/r02_ncu_C4_*.txt profiles/r02_ncu_C5_*.txt gpurun_out/
rm -f gpurun_out/r02_ncu_C4_*.ncu-rep gpurun_out/r02_ncu_C5_dense_scan.ncu-rep
if [ -n "$LISTS" ]; then
  WORKLOADS="C4 C5" bash scripts/launch_list.sh 2>&1 | grep -E "multi_" | head
  rm -f gpurun_out/r02_launches_*.csv
fi
du -sh gpurun_out
