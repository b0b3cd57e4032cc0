"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library (cuobjdump -sass; needs no GPU):
UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops,
HMMA = legacy mma.sync, LDGSTS = cp.async.   usage: python scripts/sass_summary.py [path.so] > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'timbre_trap_b200', 'libtimbretrap_b200.so')
text = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', text)), capture_output=True, text=True).stdout.split('\n')
OPS = ('UTCHMMA', 'UTMALDG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'HMMA', 'LDGSTS', 'FFMA2', 'FADD2', 'REDG', 'ATOMG')
rows, cur, i = [], None, -1
for line in text.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        i += 1
        cur = [names[i], collections.Counter(), 0]
        rows.append(cur)
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur is not None:
        cur[2] += 1
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + '.'):
                cur[1][o] += 1
print(f'# {os.path.relpath(so, ROOT)}: {len(rows)} kernels; columns = instruction counts in the SASS of each kernel')
print(f'{"kernel":100s} {"instr":>6s} ' + ' '.join(f'{o:>7s}' for o in OPS))
tot = collections.Counter()
for name, c, n in sorted(rows, key=lambda r: r[0]):
    short = re.sub(r'\(.*', '', name)[:100]
    print(f'{short:100s} {n:6d} ' + ' '.join(f'{c[o]:7d}' for o in OPS))
    tot.update(c)
print(f'{"TOTAL":100s} {sum(r[2] for r in rows):6d} ' + ' '.join(f'{tot[o]:7d}' for o in OPS))
