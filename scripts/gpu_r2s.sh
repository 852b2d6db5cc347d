#!/bin/bash
# round 2, session s: stream priority A/B (PCA above kNN / build), launch-A list order, hook test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_knn_clustered.py -m gpu -q -s -k "hooks or headline or equals_dense" 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert" | cut -c1-300 | tee gpurun_out/r2s_tests.log
timeout 300 python scripts/knn_clustered_bench.py c3 2>&1 | grep -E "stage|knn_tc_listed|identical" | tee gpurun_out/r2s_knn_bench.log
for cfg in "X=1" "DD_PRIO_MAIN=1" "DD_PRIO_MAIN=1 DD_PRIO_BUILD=1" "DD_PRIO_MAIN=2 DD_PRIO_BUILD=1" "DD_PRIO_KNN=1" "DD_PIPELINES=3"; do
    tag=$(echo "$cfg" | tr ' =' '__')
    env $cfg python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2s_bench_$tag.json 2> gpurun_out/r2s_bench_$tag.err
    python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    l = json.load(open(f"gpurun_out/r2s_bench_{tag}.json"))
    print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["stage_ms_per_step"])
except Exception as e:
    print(tag, "failed", e, open(f"gpurun_out/r2s_bench_{tag}.err").read()[-400:])
PY
done 2>&1 | tee gpurun_out/r2s_prio.log
