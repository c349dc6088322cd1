#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over the small-size parity tests; logs to gpurun_out/sanitizer_<tool>.txt
K="golden or basic_movement or edge_collision or self_collision or other_snake or eat_food or create_envs or test_reset or agent_observations or boost or respawn or overlapping or nearly_full or known or consistency or a2c or fused_step_reset_equals or stale or compact or packed or second_food or second_head or shadow"
K="($K) and not config4 and not invariants and not 100000 and not stepper"
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest ${FILES:-tests/test_single_gpu.py tests/test_multi_gpu.py \
    tests/test_gridworld_gpu.py tests/test_check_gpu.py tests/test_a2c_gpu.py} -m gpu -x -q -k "$K" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "passed|failed|SUMMARY" gpurun_out/sanitizer_$tool.txt | tail -3
done
