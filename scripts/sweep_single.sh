#!/bin/bash
# Launch-configuration sweep of the SingleSnake step kernel on C2 (scratch; results go to gpurun_out/).
for cfg in "4 32 44" "2 32 44" "2 64 44" "8 32 44" "4 64 44"; do
  set -- $cfg
  echo -n "G=$1 THREADS=$2 SMEM_KB=$3: "
  WURM_SINGLE_G=$1 WURM_SINGLE_THREADS=$2 WURM_SINGLE_SMEM_KB=$3 python bench.py --workload ${W:-C2} --steps 200 --warmup 10 --no-cpu-baseline 2>&1 | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=l['roofline']
print(f\"value {l['value']:.3e} kernel {r['kernel_ms']:.3f} ms frac {r['frac']:.3f} e2e {l['e2e']['value']:.3e} fused {l['fused_step_reset']['value']:.3e}\")"
done
