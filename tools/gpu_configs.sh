#!/bin/bash
# Single-GPU bench lines of every BASELINE.json config that fits one GPU (cfg2 is the headline; cfg3 / cfg4 / cfg5-per-GPU shapes),
# the weights-independent direct path, and the weights-sensitivity experiment.  Outputs: gpurun_out/r02_cfg_*.json
run() { tag=$1; shift; echo "== $tag: $@"; "$@" > gpurun_out/r02_cfg_$tag.json 2> gpurun_out/r02_cfg_$tag.err || tail -5 gpurun_out/r02_cfg_$tag.err;
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_cfg_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],4), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],4), "K", d.get("lowrank_K_histogram"), "parity", (d.get("parity") or {}).get("rel_err"), "top", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"])
    print("   kernels", {k: round(v,2) for k,v in list(d["kernel_ms_per_step"].items())[:6]})
except Exception as e:
    print("$tag FAILED", e)
PY
}
run cfg5_1gpu python bench.py --workload cfg5 --steps 3 --warmup 3
run cfg3 python bench.py --workload cfg3 --steps 20 --warmup 5
run cfg3_graph python bench.py --workload cfg3 --steps 20 --warmup 5 --cuda-graph --no-cpu-baseline
run cfg4 python bench.py --workload cfg4 --steps 5 --warmup 3
run cfg1_graph python bench.py --workload cfg1 --steps 50 --warmup 5 --cuda-graph --no-cpu-baseline
run cfg2_depth1_rs8 python bench.py --workload cfg2_depth1 --steps 3 --warmup 3 --radial-scale 8 --no-cpu-baseline
run cfg2_depth1_rs1 python bench.py --workload cfg2_depth1 --steps 3 --warmup 3 --no-cpu-baseline
SE3B200_NO_LOWRANK=1 run cfg2_direct python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline
