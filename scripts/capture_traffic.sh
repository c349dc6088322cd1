#!/bin/bash
# ncu --set full capture of the step kernel of every BASELINE workload (dense and compact state), steady state (12th launch).
# Run on the GPU box; then scripts/update_traffic.py turns the reports into profiles/r02_ncu_*.txt and profiles/traffic.json.
mkdir -p gpurun_out
for w in ${WORKLOADS:-C2 C3 C4 C5 G1}; do
  for st in dense compact dense_scan; do
    if [ $w = G1 ] && [ $st != dense ]; then continue; fi
    case $w in C4|C5) ;; *) if [ $st = dense_scan ]; then continue; fi ;; esac     # (only MultiSnake has a shadowed dense state)
    case $w in C2|C3|C1) k='regex:single_tile_kernel|single_body_kernel|single_compact_kernel' ;; G1) k='regex:grid_tile_kernel|grid_small_kernel|grid_env_kernel' ;; *) k='regex:multi_env_kernel' ;; esac
    ncu --set full --clock-control none --import-source on -k "$k" -s 11 -c 1 -f -o gpurun_out/r02_ncu_${w}_${st} \
        python scripts/profile_step.py $w $st 14 2>&1 | tail -1
  done
done
