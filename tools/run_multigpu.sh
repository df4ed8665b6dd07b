#!/bin/bash
# Multi-GPU evidence for the BASELINE configurations that name 8 GPUs (run under `gpurun --gpus N`):
#   config 5  2M Gaussians / 1600^2 / 8 views forward+backward, views sharded rank::N (ours and the reference)
#   config 4  the eval flow, 8 objects sharded over the ranks (ours: drop-in loop + fused; reference)
# usage: tools/run_multigpu.sh N [tag]      -> gpurun_out/<tag>_c{4,5}_<N>gpu_{ours,ref}.json (last stdout line = the JSON)
N=${1:-8}
tag=${2:-r2}
run() {  # name, bench args...
    name=$1; shift
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
        bench.py --gpus $N "$@" > gpurun_out/${name}.log 2> gpurun_out/${name}.err
    tail -n 1 gpurun_out/${name}.log > gpurun_out/${name}.json
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${name}.json"))
    print("${name}", "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "ms/step", d.get("ms_per_step"),
          {k: v.get("views_per_s") for k, v in (d.get("arms") or {}).items()})
except Exception as e:
    print("${name}", "FAILED", e)
PY
}
mkdir -p gpurun_out
run ${tag}_c5_${N}gpu_ours --config 5 --steps 5 --warmup 3
run ${tag}_c5_${N}gpu_ref --config 5 --steps 5 --warmup 3 --impl reference
run ${tag}_c4_${N}gpu_ours --config 4
run ${tag}_c4_${N}gpu_ref --config 4 --impl reference
