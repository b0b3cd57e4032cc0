"""SASS-level hot spots (stall samples / executed instructions) of the first kernel in an .ncu-rep captured with --import-source on."""
import csv, collections, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
H = rows[hi]; ia = H.index('Source'); isamp = H.index('# Samples'); iex = H.index('Instructions Executed')
data = []
for r in rows[hi + 1:]:
    if len(r) > iex and r[0].startswith('0x'): data.append((r[ia].strip(), int(r[isamp] or 0), int(r[iex] or 0)))
    elif r and r[0] == 'Kernel Name': break
ts, te = sum(d[1] for d in data), sum(d[2] for d in data)
print('samples', ts, 'warp-instructions', te, 'sass lines', len(data))
op, ops = collections.Counter(), collections.Counter()
for s, sa, ex in data:
    t = s.split(); o = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; op[o] += ex; ops[o] += sa
for o, c in op.most_common(14): print(f'  {o:10s} exec {c/te*100:5.1f}%  samples {ops[o]/ts*100:5.1f}%')
print('top by samples')
for i, (s, sa, ex) in sorted(enumerate(data), key=lambda t: -t[1][1])[:topn]: print(f'  #{i:4d} {sa/ts*100:5.1f}% ex={ex:9d} {s[:90]}')
