"""One-line digest of a bench.py JSON line read from stdin:  python bench.py ... | python scripts/bench_line.py [label]"""
import json, sys
l = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = l.get('roofline') or {}
f = l.get('fused_step_reset') or {}
g = l.get('graphed_step_reset') or {}
print(' '.join(sys.argv[1:]), 'value %.4g  ms/step %.4f  kernel %.4f ms  frac %.3f  e2e %.4g  fused %.4g' % (
    l['value'], l['ms_per_step'], r.get('kernel_ms', float('nan')), r.get('frac', float('nan')), l['e2e']['value'],
    f.get('value', float('nan'))) + ('  graphed %.4g' % g['value'] if g else ''))
