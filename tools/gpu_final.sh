#!/bin/bash
# Final validation round trip: weight-multicast width experiment (depth-1 slice of cfg2), the whole GPU suite with durations,
# the headline bench line, the ncu launch list of one timed forward.
tag=${1:-final}
for c in 2 4 1; do
  SE3B200_Z_CLUSTER=$c python bench.py --workload cfg2_depth1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_csz${c}.json 2> gpurun_out/r02_csz${c}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_csz$c.json").read().strip().splitlines()[-1])
print("cluster $c", d["ms_per_step"], d["kernel_ms_per_step"]["zgemm"], d["clocks"]["sm_mhz"])
PY
done
SE3B200_Z_CLUSTER=4 timeout 600 python -m pytest tests/test_gpu_zgemm.py -x -q 2>&1 | tail -3 | tee gpurun_out/r02_pytest_csz4.log
timeout 1800 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu_$tag.log
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_$tag.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["kernel_ms_per_step"], d["untimed_share_of_step"], d["clocks"])
print(d["parity"]["rel_err"]); r=d["roofline"]; print(r["kernel"], r["achieved"], r["frac"])
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches_cfg2_depth1.csv \
    python bench.py --workload cfg2_depth1 --steps 1 --warmup 3 --no-cpu-baseline --profile-range > gpurun_out/${tag}_ncu_launch_bench.log 2>&1
ls -la gpurun_out/${tag}_launches_cfg2_depth1.csv
