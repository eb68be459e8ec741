#!/bin/bash
# Multi-GPU lines (one process per GPU, NCCL): usage: bash tools/gpu_multi.sh <ngpus> <tag> [bench args...]
n=$1; tag=$2; shift; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" \
    > gpurun_out/r02_multi_$tag.json 2> gpurun_out/r02_multi_$tag.err
tail -c 400 gpurun_out/r02_multi_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_multi_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "n_gpus", d["n_gpus"], "value", round(d["value"],4), "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],4), d["scaling"], d["config"]["global_batch"])
except Exception as e:
    print("$tag FAILED", e)
PY
