"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]; ki = H.index('Kernel Name'); vi = H.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        agg.setdefault(r[ki][:64], []).append(float(r[vi].replace(',', '')))
tot = sum(sum(v) for k, v in agg.items() if 'tt::' in k)
print(f'{"kernel":64s} {"n":>5s} {"mean us":>10s} {"total ms":>9s} {"share":>6s}')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if 'tt::' in k:
        print(f'{k:64s} {len(v):5d} {sum(v)/len(v)/1e3:10.1f} {sum(v)/1e6:9.2f} {sum(v)/tot*100:5.1f}%')
print('total (tt kernels) ms', round(tot / 1e6, 2))
