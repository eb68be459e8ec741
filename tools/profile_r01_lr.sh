#!/bin/bash
# Run under gpurun (1 GPU).  ncu evidence for the low-rank radial path: per-launch duration list of one forward (depth-1
# slice of cfg2) and full-set captures of the dominant kernels.  Numbers printed by runs under ncu are NOT bench values.
set -u
mkdir -p gpurun_out
W=${1:-cfg2_depth1}
B="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --profile-range"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_lr_$W.csv \
    $B > gpurun_out/ncu_launch_run.log 2>&1
# edge-aligned path: a (+m,-m) x (a,b) launch (P = 2, F = 2) of the attention block (the 11th of that instantiation)
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pairwise_lr_kernel<.int.2' -s 10 -c 1 -o gpurun_out/prof_pairwise_lr_$W \
    $B > gpurun_out/ncu_pairwise_run.log 2>&1
# an m = 0 launch (P = 1, F = 1) of the attention block that accumulates (13th of that instantiation)
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pairwise_lr_kernel<.int.1' -s 12 -c 1 -o gpurun_out/prof_pairwise_lr_p1_$W \
    $B > gpurun_out/ncu_pairwise_p1_run.log 2>&1
ls -la gpurun_out/ | tail -8
