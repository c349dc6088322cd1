#!/bin/bash
# Every workload once on one GPU; JSON lines to gpurun_out/final_<W>.json and a one-line summary each.
mkdir -p gpurun_out
for w in C2 C1 C3 C4 C5 C5F G1; do
  python bench.py --workload $w --steps ${STEPS:-500} --warmup 20 > gpurun_out/final_$w.json 2> gpurun_out/final_$w.err
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
try:
    l = json.loads(open(f'gpurun_out/final_{w}.json').read().strip().splitlines()[-1]); r = l['roofline']
    print(w, 'value %.4g  ms/step %.4f  kernel %.4f ms  frac %.3f  e2e %.4g  cpu %.4g (%d cores)  clocks %s' % (
        l['value'], l['ms_per_step'], r['kernel_ms'], r['frac'], l['e2e']['value'], l['cpu_baseline']['value'],
        l['cpu_baseline']['cores'], l['clocks']))
except Exception as ex:
    print(w, 'FAILED', ex, open(f'gpurun_out/final_{w}.err').read()[-500:])
PY
done
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_reference_arm.json; cut -c1-200 gpurun_out/final_reference_arm.json
