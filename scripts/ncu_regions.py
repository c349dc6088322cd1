"""Dynamic warp instructions and stall samples of an ncu report grouped by source-line ranges of one file.

    python scripts/ncu_regions.py report.ncu-rep file.cu  a-b:name  a-b:name ...   [--per N]  (N = divide counts by, e.g. the grid size)
"""
import collections, csv, subprocess, sys
args = sys.argv[1:]
per = 1.0
if '--per' in args:
    i = args.index('--per'); per = float(args[i + 1]); del args[i:i + 2]
rep, fname = args[0], args[1]
regions = []
for a in args[2:]:
    rng, name = a.split(':', 1); lo, hi = rng.split('-'); regions.append((int(lo), int(hi), name))
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
cur = ''; hdr = None; agg = collections.Counter(); samp = collections.Counter()
for r in csv.reader(txt.splitlines()):
    if not r: continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try: n, sm = int(d['Instructions Executed']), int(d['# Samples'])
        except (KeyError, ValueError): continue
        key = cur
        if cur == fname:
            key = fname + ':other'
            for lo, hi, name in regions:
                if lo <= int(r[0]) <= hi: key = name; break
        agg[key] += n; samp[key] += sm
tot, tots = sum(agg.values()), max(sum(samp.values()), 1)
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    print(f'{v / per:10.1f}  {100 * v / tot:5.1f}%   stall samples {100 * samp[k] / tots:5.1f}%   {k}')
print(f'{tot / per:10.1f}  total')
