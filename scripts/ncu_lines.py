"""Per-CUDA-source-line instruction and stall-sample totals from an ncu report.

    python scripts/ncu_lines.py report.ncu-rep [top_n]
"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = ''; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0] not in ('', ) and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            out.append((int(d['Instructions Executed']), int(d['# Samples']), cur_file, int(r[0]), r[1].strip()[:100]))
        except (ValueError, KeyError):
            pass
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print(f'total warp instructions {tot}, stall samples {tots}')
for o in sorted(out, reverse=True)[:top]:
    print(f'{o[0]:>10} {100*o[0]/tot:5.1f}%  samp {100*o[1]/max(tots,1):5.1f}%  {o[2]}:{o[3]:<4} {o[4]}')
