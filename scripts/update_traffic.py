"""Turns gpurun_out/r02_ncu_<W>_<state>.ncu-rep into profiles/r02_ncu_<W>_<state>.txt (key metrics) and records the dram bytes
per launch in profiles/traffic.json together with the hash of the kernel sources they were measured on (bench.py refuses
the figure once wurm_b200/csrc changes).  Run in the container right after scripts/capture_traffic.sh, BEFORE editing csrc."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
path = os.path.join(ROOT, 'profiles', 'traffic.json')
out = json.load(open(path)) if os.path.exists(path) else {}
out = {k: v for k, v in out.items() if isinstance(v, dict) or k.startswith('_')}
out['_comment'] = ('dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel (env.step of the bench loop, 12th '
                   'launch) from `ncu --set full` captures; csrc_hash = bench.csrc_hash(family) of the source files that kernel is compiled from (`sources`) at '
                   'capture time; read by bench.py for roofline.traffic')
for w in ('C1', 'C2', 'C3', 'C4', 'C5', 'G1'):
    for st in ('dense', 'compact', 'dense_scan'):
        rep = os.path.join(ROOT, 'gpurun_out', f'{tag}_ncu_{w}_{st}.ncu-rep')
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
        def mbytes(name):
            v = float(d[name]); unit = u[name].lower()
            return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}[unit]
        total = mbytes('dram__bytes_read.sum') + mbytes('dram__bytes_write.sum')
        key = w if st == 'dense' else w + ':' + st
        out[key] = {'dram_bytes_per_launch': int(total), 'kernel': d['Kernel Name'], 'csrc_hash': bench.csrc_hash(bench.kernel_family(w)), 'sources': list(bench.KERNEL_SOURCES[bench.kernel_family(w)]),
                    'source': f'profiles/{tag}_ncu_{w}_{st}.txt',
                    'gpu_time_us': float(d['gpu__time_duration.sum']) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}[u['gpu__time_duration.sum']]}
        summary = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_summary.py'), rep,
                                  f'{w} {st} state, step kernel of the bench loop (12th launch), kernel sources {bench.csrc_hash(bench.kernel_family(w))}'],
                                 capture_output=True, text=True).stdout
        open(os.path.join(ROOT, 'profiles', f'{tag}_ncu_{w}_{st}.txt'), 'w').write(summary)
        print(key, out[key])
json.dump(out, open(path, 'w'), indent=1)
