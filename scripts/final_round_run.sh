set -u
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-eager --no-train --no-hour > gpurun_out/r02_bench_under_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/r02_bench_launches.csv > gpurun_out/r02_bench_launch_shares.txt 2>&1
head -8 gpurun_out/r02_bench_launch_shares.txt
ncu --set full --clock-control none --import-source on -k regex:res_rs_kernel -s 2 -c 1 -o gpurun_out/r02_res_rs_c4_fold \
    python scripts/time_res.py 4 540 1 256 fold 3 > gpurun_out/r02_res_ncu.log 2>&1
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench11.json 2> gpurun_out/r02_bench11.err
tail -c 300 gpurun_out/r02_bench11.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench11.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['family_frac'], d['gpu_launches'], d['clocks'], d['cqt']['forward']['frac'], d['cqt']['inverse']['frac'], d['train_step']['ms_per_step'])"
