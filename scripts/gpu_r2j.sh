#!/bin/bash
# round 2, session j: cluster-ordered kNN (device k-means ordering, two listed launches) as the fit loop's default
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn_clustered.py -m gpu -q -s -x 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert" | cut -c1-300 | tee gpurun_out/r2j_tests.log
timeout 300 python scripts/knn_clustered_bench.py c3 2>&1 | tee gpurun_out/r2j_knn_bench.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "phenograph or pipeline or end_to_end or knn" 2>&1 | tail -3 | tee -a gpurun_out/r2j_tests.log
for tag in clustered dense; do
    env="X=1"
    [ $tag = dense ] && env="DD_KNN_DENSE=1"
    env $env python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2j_bench_$tag.json 2> gpurun_out/r2j_bench_$tag.err
done
python - <<'PY'
import json
for tag in ("clustered", "dense"):
    try:
        l = json.load(open(f"gpurun_out/r2j_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in l["kernel_ms_total"].items()}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2j_bench_{tag}.err").read()[-600:])
PY
