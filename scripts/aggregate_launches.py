"""Per-kernel aggregate of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  python scripts/aggregate_launches.py file.csv [title]"""
import collections, csv, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
for r in rd:
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(unit, 1.0)
    name = r['Kernel Name']
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values()) or 1.0
print('#', ' '.join(sys.argv[2:]))
print(f'# {sum(a[0] for a in agg.values())} launches, {tot:.1f} us of kernel time')
print(f'{"launches":>8} {"total_us":>12} {"mean_us":>10} {"share":>7}  kernel')
for name, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{n:8d} {us:12.1f} {us / n:10.2f} {100 * us / tot:6.1f}%  {name[:120]}')
