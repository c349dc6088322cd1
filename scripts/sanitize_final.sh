K="(shadow or fused_step_reset_equals or golden or respawn or nearly_full or compact) and not stepper"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_multi_gpu.py tests/test_check_gpu.py -m gpu -x -q -k "$K" > gpurun_out/sanitizer_memcheck_final.txt 2>&1; echo "memcheck rc=$?"; grep -E "passed|failed|SUMMARY" gpurun_out/sanitizer_memcheck_final.txt | tail -3
K2="(shadow or fused_step_reset_equals) and not stepper and not 16-64 and not tensors"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -k "$K2" > gpurun_out/sanitizer_racecheck_final.txt 2>&1; echo "racecheck rc=$?"; grep -E "passed|failed|SUMMARY|hazard" gpurun_out/sanitizer_racecheck_final.txt | tail -5
timeout 300 python scripts/soak.py 120 > gpurun_out/r02_soak_final.txt 2>&1; tail -1 gpurun_out/r02_soak_final.txt
