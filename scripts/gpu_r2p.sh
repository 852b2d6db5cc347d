#!/bin/bash
# round 2, session p: fallback test, blocking-sync workers, sanitizer pass over the round-2 kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn_clustered.py -m gpu -q -s -k "auto_mode or headline" 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert" | cut -c1-300 | tee gpurun_out/r2p_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2p_bench_louvain.json 2> gpurun_out/r2p_bench_louvain.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2p_bench_louvain.json"))
print(round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["stage_ms_per_step"])
PY
bash scripts/gpu_sanitize_r2.sh
