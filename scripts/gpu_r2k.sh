#!/bin/bash
# round 2, session k: cluster-ordered kNN tuned (persistent assign, coalesced list kernel), k = 31 variant, full test suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn_clustered.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert" | cut -c1-300 | tee gpurun_out/r2k_tests.log
timeout 300 python scripts/knn_clustered_bench.py c3 2>&1 | tee gpurun_out/r2k_knn_bench.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee -a gpurun_out/r2k_tests.log
for tag in louvain pheno; do
    extra=""
    [ $tag = pheno ] && extra="--clustering phenograph"
    python bench.py --steps 3 --warmup 3 $extra --no-cpu-baseline --no-extra > gpurun_out/r2k_bench_$tag.json 2> gpurun_out/r2k_bench_$tag.err
done
python - <<'PY'
import json
for tag in ("louvain", "pheno"):
    try:
        l = json.load(open(f"gpurun_out/r2k_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in l["kernel_ms_total"].items()}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2k_bench_{tag}.err").read()[-600:])
PY
