#!/bin/bash
# round 2, session o (2 GPUs): cell-sharded parity test, weak / strong iteration sharding, c3 and c5 with the cells sharded
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s 2>&1 | tail -6 | tee gpurun_out/r2o_tests.log
run() {  # tag, args...
    tag=$1; shift
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/r2o_$tag.json 2> gpurun_out/r2o_$tag.err
    tail -c 600 gpurun_out/r2o_$tag.err | tail -3
}
run weak --steps 3 --warmup 3 --no-extra
run strong --steps 3 --warmup 3 --scaling strong --no-extra
run cells_c3 --steps 2 --warmup 2 --shard cells --no-extra
run cells_c5 --steps 1 --warmup 1 --shard cells --workload c5 --iters 5 --no-extra
python - <<'PY'
import json
for tag in ("weak", "strong", "cells_c3", "cells_c5"):
    try:
        txt = [l for l in open(f"gpurun_out/r2o_{tag}.json").read().splitlines() if l.startswith("{")][-1]
        l = json.loads(txt)
        print(tag, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 1), "e2e", l.get("e2e", {}).get("value"), l.get("scaling"), l["config"].get("workload"), l.get("e2e", {}).get("host_ms_last_fit"))
    except Exception as e:
        print(tag, "failed", e)
PY
