#!/bin/bash
# Round profile captures on the GPU box (one GPU): launch list of the bench command, --set full captures of the dominant inference kernel,
# the fused convout / cross-fade kernel and the tensor-core weight-gradient kernel.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -u
R=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${R}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-eager --no-train --no-hour > gpurun_out/${R}_bench_under_ncu.log 2>&1
python scripts/launch_shares.py gpurun_out/${R}_bench_launches.csv > gpurun_out/${R}_bench_launch_shares.txt 2>&1
head -45 gpurun_out/${R}_bench_launch_shares.txt
ncu --set full --clock-control none --import-source on -k regex:res_rs_kernel -s 2 -c 1 -o gpurun_out/${R}_res_rs_c4_fold \
    python scripts/time_res.py 4 540 1 256 fold 3 > gpurun_out/${R}_res_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_out_xfade -s 1 -c 1 -o gpurun_out/${R}_convout_xfade \
    python scripts/one_hour.py 300 > gpurun_out/${R}_xfade_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 6 -c 4 -o gpurun_out/${R}_wgrad \
    python scripts/train_step_bench.py 8 1 > gpurun_out/${R}_wgrad_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bins_inv|rows_inv|cols_inv" -s 9 -c 6 -o gpurun_out/${R}_cqt_inv \
    python scripts/time_cqt.py 256 > gpurun_out/${R}_cqt_inv_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
