#!/bin/bash
# ncu evidence for the production kernels on a depth-1 slice of cfg2 (same widths, same launches as the headline, 1/6 of the layers):
#   (1) launch list with device times of ONE timed forward (cudaProfilerStart/Stop around the resident steps: bench.py --profile-range),
#   (2) full-set capture of zgemm launches (mode 1 and mode 3) with source correlation, raw metrics exported to CSV for profiles/.
tag=${1:-r02}
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches_cfg2_depth1.csv \
    python bench.py --workload cfg2_depth1 --steps 1 --warmup 3 --no-cpu-baseline --profile-range > gpurun_out/${tag}_ncu_launch_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:zgemm -s 8 -c 4 -f -o gpurun_out/${tag}_prof_zgemm \
    python bench.py --workload cfg2_depth1 --steps 1 --warmup 3 --no-cpu-baseline --profile-range > gpurun_out/${tag}_ncu_full_bench.log 2>&1
ncu -i gpurun_out/${tag}_prof_zgemm.ncu-rep --page raw --csv > gpurun_out/${tag}_prof_zgemm_raw.csv 2>/dev/null
ls -la gpurun_out/${tag}_prof_zgemm* gpurun_out/${tag}_launches_cfg2_depth1.csv
