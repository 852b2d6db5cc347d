#!/bin/bash
# round 2, session i: 8-lane weighted Louvain propose; HVG on the device; pseudocount == 1 (exact PCA); benches
mkdir -p gpurun_out
python -m pytest tests/test_gpu_hvg.py tests/test_gpu_exact_pca.py tests/test_gpu_pheno_level0.py tests/test_gpu_zz_leiden.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error" | cut -c1-400 | tee gpurun_out/r2i_tests.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee -a gpurun_out/r2i_tests.log
for tag in pheno louvain; do
    extra="--clustering phenograph"
    [ $tag = louvain ] && extra=""
    env="X=1"
    env $env python bench.py --steps 2 --warmup 3 $extra --no-cpu-baseline --no-extra > gpurun_out/r2i_bench_$tag.json 2> gpurun_out/r2i_bench_$tag.err
done
python - <<'PY'
import json
for tag in ("pheno", "louvain"):
    try:
        l = json.load(open(f"gpurun_out/r2i_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in l["kernel_ms_total"].items()}, l["stage_ms_per_step"], l.get("host_ms_last_fit"))
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2i_bench_{tag}.err").read()[-600:])
PY
