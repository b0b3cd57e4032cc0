#!/bin/bash
# End-of-round run on one GPU: the whole GPU test suite, the standard bench line, and the launch list of the bench command.
set -u
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['family_frac'], d['gpu_launches'], d['clocks'], d['cqt']['forward']['frac'], d['cqt']['inverse']['frac'], d['train_step']['ms_per_step'], d['train_step']['tflops'])"
