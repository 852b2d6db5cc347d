#!/bin/bash
# round 2, session n: lists of 40 for k > 13, two pipelines per GPU (default), ncu --set full of the kNN / dense / k-means kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_knn_clustered.py tests/test_gpu_parity.py tests/test_gpu_pheno_level0.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert" | cut -c1-300 | tee gpurun_out/r2n_tests.log
for tag in louvain louvain_p1 pheno pheno_p1; do
    extra=""; env="X=1"
    case $tag in pheno*) extra="--clustering phenograph";; esac
    case $tag in *_p1) env="DD_PIPELINES=1";; esac
    env $env python bench.py --steps 3 --warmup 3 $extra --no-cpu-baseline --no-extra > gpurun_out/r2n_bench_$tag.json 2> gpurun_out/r2n_bench_$tag.err
done
python - <<'PY'
import json
for tag in ("louvain", "louvain_p1", "pheno", "pheno_p1"):
    try:
        l = json.load(open(f"gpurun_out/r2n_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in list(l["kernel_ms_total"].items())[:14]}, l["stage_ms_per_step"], l["roofline_kernel"], round(l["roofline"]["frac"], 4))
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2n_bench_{tag}.err").read()[-600:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_dense_rows|k_km_assign|k_jacobi" -s 5 -c 10 -o gpurun_out/r2n_full python scripts/stage_bench.py c3 2 > gpurun_out/r2n_ncu_full.log 2>&1
ls -la gpurun_out | tail -4
