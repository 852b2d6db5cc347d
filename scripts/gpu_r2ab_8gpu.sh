#!/bin/bash
# round 2, session ab (8 GPUs): config 5 with the cluster-ordered kNN under cell-block sharding
mkdir -p gpurun_out
run() {
    tag=$1; shift
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 "$@" > gpurun_out/r2ab_$tag.json 2> gpurun_out/r2ab_$tag.err
    grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/r2ab_$tag.err | tail -3
}
run cells_c5 --steps 1 --warmup 1 --shard cells --workload c5 --iters 8 --no-extra
python - <<'PY'
import json
for tag in ("cells_c5",):
    try:
        txt = [l for l in open(f"gpurun_out/r2ab_{tag}.json").read().splitlines() if l.startswith("{")][-1]
        l = json.loads(txt)
        print(tag, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 1), {k: round(v, 1) for k, v in list(l["kernel_ms_total"].items())[:10]}, l.get("stage_ms_per_step"))
    except Exception as e:
        print(tag, "failed", e)
PY
