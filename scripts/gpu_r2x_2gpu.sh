#!/bin/bash
# round 2, session x (2 GPUs): the sharded parity test and one weak-scaling line with the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -s 2>&1 | grep -E "passed|failed|SHARDED|Error|error" | tail -6 | tee gpurun_out/r2x_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2x_weak.json 2> gpurun_out/r2x_weak.err
python - <<'PY'
import json
txt = [l for l in open("gpurun_out/r2x_weak.json").read().splitlines() if l.startswith("{")][-1]
l = json.loads(txt)
print("weak 2 GPUs: value", round(l["value"]), "ms/step", round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["e2e"]["host_ms_last_fit"])
PY
