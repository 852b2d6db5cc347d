#!/bin/bash
# round 2, session r (8 GPUs): weak and strong iteration sharding, c5 with the cells sharded
mkdir -p gpurun_out
nproc; python -c "import os; print('cores', len(os.sched_getaffinity(0)))"
run() {  # tag, args...
    tag=$1; shift
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 "$@" > gpurun_out/r2r_$tag.json 2> gpurun_out/r2r_$tag.err
    grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/r2r_$tag.err | tail -3
}
run weak --steps 3 --warmup 3 --no-extra --no-cpu-baseline
run strong --steps 3 --warmup 3 --scaling strong --no-extra --no-cpu-baseline
run cells_c5 --steps 1 --warmup 1 --shard cells --workload c5 --iters 8 --no-extra
python - <<'PY'
import json
for tag in ("weak", "strong", "cells_c5"):
    try:
        txt = [l for l in open(f"gpurun_out/r2r_{tag}.json").read().splitlines() if l.startswith("{")][-1]
        l = json.loads(txt)
        print(tag, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 1), "e2e", l.get("e2e", {}).get("value"), l.get("scaling"), l["config"].get("workload"), l.get("e2e", {}).get("host_ms_last_fit"))
    except Exception as e:
        print(tag, "failed", e)
PY
