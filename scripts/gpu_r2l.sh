#!/bin/bash
# round 2, session l: kNN filter-width check (k = 31), two pipelines on one GPU, fused Jacobi
mkdir -p gpurun_out
timeout 900 python scripts/knn_recall_check.py c3 3000 2>&1 | tee gpurun_out/r2l_knn_recall.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_exact_pca.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2l_tests.log
timeout 600 python scripts/two_pipelines_probe.py c3 3 2>&1 | tee gpurun_out/r2l_two_pipelines.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2l_bench_louvain.json 2> gpurun_out/r2l_bench_louvain.err
python - <<'PY'
import json
for tag in ("louvain",):
    try:
        l = json.load(open(f"gpurun_out/r2l_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in l["kernel_ms_total"].items()}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2l_bench_{tag}.err").read()[-600:])
PY
