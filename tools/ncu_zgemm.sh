#!/bin/bash
# ncu evidence for the production kernels on a depth-1 slice of cfg2 (same widths, same launches as the headline, 1/6 of the layers):
#   (1) launch list with device times, (2) full-set capture of zgemm launches (mode 1 and mode 3), exported to CSV for profiles/.
tag=${1:-r02}
export SE3B200_BENCH_QUIET=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_cfg2_depth1.csv \
    python bench.py --workload cfg2_depth1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:zgemm -s 60 -c 4 -f -o gpurun_out/${tag}_prof_zgemm \
    python bench.py --workload cfg2_depth1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full_bench.log 2>&1
ncu -i gpurun_out/${tag}_prof_zgemm.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_zgemm_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_prof_zgemm* gpurun_out/${tag}_launches_cfg2_depth1.csv
