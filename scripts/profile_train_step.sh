ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_train_launches6.csv python scripts/train_step_bench.py 8 1 > gpurun_out/r02_train_ncu6.log 2>&1; tail -1 gpurun_out/r02_train_ncu6.log; python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/r02_train_launches6.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi: agg.setdefault(r[ki][:100], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("total ms", tot/1e6, "launches", sum(len(v) for v in agg.values()))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:30]: print(f"{k:100s} {len(v):5d} {sum(v)/len(v)/1e3:9.1f} {sum(v)/1e6:8.2f} {sum(v)/tot*100:5.1f}%")
PY
