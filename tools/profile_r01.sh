#!/bin/bash
# Run under gpurun (1 GPU).  Produces the ncu evidence for profiles/: a per-launch duration list of one bench step and
# full-set captures of the dominant kernels.  Numbers printed by runs under ncu are NOT bench values.
set -u
mkdir -p gpurun_out
W=${1:-cfg2_depth1}
B="python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --profile-range"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$W.csv \
    $B > gpurun_out/ncu_launch_run.log 2>&1
# (3,3) pair of to_k / to_v: 35th and 36th pairwise launch of the forward (4 conv_in pairs, then 16 pairs x {k,v})
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pairwise_tc_kernel -s 34 -c 1 -o gpurun_out/prof_pairwise_$W \
    $B > gpurun_out/ncu_pairwise_run.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tbuild -s 19 -c 1 -o gpurun_out/prof_tbuild_$W \
    $B > gpurun_out/ncu_tbuild_run.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_kernel -s 3 -c 1 -o gpurun_out/prof_attn_$W \
    $B > gpurun_out/ncu_attn_run.log 2>&1
ls -la gpurun_out/
