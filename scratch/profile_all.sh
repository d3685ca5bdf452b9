#!/bin/bash
# One GPU call: the ncu launch list of the bench and one `--set full` capture per kernel family (round 2).
# Reports land in gpurun_out/; profiles/summarize_ncu.py turns them into the tracked summaries afterwards.
set -u
mkdir -p gpurun_out
export PYTHONPATH=.
NCU="ncu --set full --clock-control none --import-source on --launch-count 1 -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --passes 8 --no-cpu-baseline --no-secondary > gpurun_out/r02_launches_bench.log 2>&1
$NCU --kernel-name regex:iqbb_fold_f32_win --launch-skip 4 -o gpurun_out/r02_fold_win python bench.py --steps 1 --warmup 3 --passes 8 --no-cpu-baseline --no-secondary --no-c5 > /dev/null 2>&1
$NCU --kernel-name regex:conv8k --launch-skip 2 -o gpurun_out/r02_conv8k python scratch/c3_only.py > /dev/null 2>&1
$NCU --kernel-name regex:fft8k --launch-skip 1 -o gpurun_out/r02_fft8k python scratch/c3_only.py > /dev/null 2>&1
$NCU --kernel-name regex:iqbb_accum_int_warp --launch-skip 2 -o gpurun_out/r02_int16_warp python scratch/c1_only.py > /dev/null 2>&1
$NCU --kernel-name regex:iqbb_accum_int_warp --launch-skip 25 -o gpurun_out/r02_int16_warp_realsym python scratch/c1_only.py > /dev/null 2>&1
$NCU --kernel-name regex:bank_accum --launch-skip 2 -o gpurun_out/r02_bank python scratch/bank_probe.py > /dev/null 2>&1
$NCU --kernel-name regex:perwin16 --launch-skip 2 -o gpurun_out/r02_perwin16_ss64 python scratch/float_sweep.py 64,32 > /dev/null 2>&1
$NCU --kernel-name regex:perwin1_ --launch-skip 2 -o gpurun_out/r02_perwin1_ss16 python scratch/float_sweep.py 16,15 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
