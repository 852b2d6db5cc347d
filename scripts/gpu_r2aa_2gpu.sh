#!/bin/bash
# round 2, session aa (2 GPUs): cluster-ordered kNN under cell-block sharding
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s 2>&1 | grep -E "passed|failed|SHARDED|Error|error|assert|differs" | tail -12 | tee gpurun_out/r2aa_tests.log
run() {
    tag=$1; shift
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 "$@" > gpurun_out/r2aa_$tag.json 2> gpurun_out/r2aa_$tag.err
    grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/r2aa_$tag.err | tail -3
}
run cells_c3 --steps 2 --warmup 2 --shard cells --no-extra
DD_KNN_DENSE=1 run cells_c3_dense --steps 2 --warmup 2 --shard cells --no-extra
run cells_c5 --steps 1 --warmup 1 --shard cells --workload c5 --iters 5 --no-extra
python - <<'PY'
import json
for tag in ("cells_c3", "cells_c3_dense", "cells_c5"):
    try:
        txt = [l for l in open(f"gpurun_out/r2aa_{tag}.json").read().splitlines() if l.startswith("{")][-1]
        l = json.loads(txt)
        print(tag, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 1), {k: round(v, 1) for k, v in list(l["kernel_ms_total"].items())[:8]})
    except Exception as e:
        print(tag, "failed", e)
PY
