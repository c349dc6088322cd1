"""Per-kernel SASS evidence of the built library: instruction count, registers, static shared memory and the counts of
the mnemonics that show which hardware paths a kernel uses (UBLKCP = TMA bulk copies, SYNCS = mbarrier, LDG.E.128 /
LD.E.128 = 128-bit global loads, STG.E.128, SHFL, VOTE = ballot, ATOMS = shared atomics, HMMA/UTCMMA = tensor cores: none).
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
lib = os.path.join(ROOT, 'wurm_b200', '_C', 'libwurm_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
res = subprocess.run(['cuobjdump', '-res-usage', lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r'Function (\S+):', line)
    if m:
        cur = m.group(1); continue
    m = re.search(r'REG:(\d+).*?SHARED:(\d+)', line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2))); cur = None
keys = ['UBLKCP', 'SYNCS', 'LDG.E.128', 'LD.E.128', 'STG.E.128', 'LDS', 'STS', 'ATOMS', 'SHFL', 'VOTE', 'BAR.SYNC', 'HMMA', 'UTCMMA']
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', line)
    if m and cur:
        ins = m.group(1)
        counts[cur]['_total'] += 1
        for k in keys:
            if re.search(r'(^|\s|\.)' + re.escape(k) + r'(\.|\s|$)', ins) or ins.split()[0].startswith(k) or (len(ins.split()) > 1 and ins.split()[0].startswith('@') and ins.split()[1].startswith(k)):
                counts[cur][k] += 1
print(f'# SASS summary of wurm_b200/_C/libwurm_b200.so (sm_100a), kernel sources csrc {bench.csrc_hash()}')
print('# columns: instructions, registers/thread, static smem bytes, then mnemonic counts (zero counts omitted)')
for name in sorted(counts, key=lambda n: demangle(n)):
    c = counts[name]
    reg, smem = usage.get(name, (None, None))
    extra = '  '.join(f'{k}={c[k]}' for k in keys if c[k])
    print(f'{demangle(name)[:110]:<110}  instr={c["_total"]:<6} regs={reg} smem={smem}  {extra}')
