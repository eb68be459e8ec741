#!/bin/bash
# One GPU round trip of the development loop: kernel + headline + model parity tests, then the cfg2 bench line.
# usage (through gpurun): bash tools/gpu_cycle.sh <tag> [pytest files...]
tag=${1:-x}; shift
files=${@:-tests/test_gpu_zgemm.py tests/test_gpu_headline.py tests/test_gpu_model.py}
timeout 1500 python -m pytest $files -x -q 2>&1 | tail -12 | tee gpurun_out/r02_pytest_$tag.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_$tag.json 2> gpurun_out/r02_bench_$tag.err
tail -c 600 gpurun_out/r02_bench_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_$tag.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["kernel_ms_per_step"], d["untimed_share_of_step"], d["lowrank_K_histogram"], d["clocks"])
print(d["parity"])
r=d["roofline"]; print(r["kernel"], r["achieved"], r["frac"], r["avg_launch_ms"])
PY
