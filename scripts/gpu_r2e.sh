#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "louvain_level0 or pipeline_matches or end_to_end" 2>&1 | tail -3 | tee gpurun_out/r2e_tests.log
DD_LOUVAIN_WHILE=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "louvain_level0 or pipeline_matches or end_to_end" 2>&1 | tail -3 | tee -a gpurun_out/r2e_tests.log
for v in pdl while nopdl_while; do
    echo "=== $v"
    if [ $v = pdl ]; then python scripts/lv_probe.py c3; elif [ $v = while ]; then DD_LOUVAIN_WHILE=1 python scripts/lv_probe.py c3; else DD_LOUVAIN_WHILE=1 DD_LOUVAIN_NO_PDL=1 python scripts/lv_probe.py c3; fi
done 2>&1 | tee gpurun_out/r2e_lv_probe.log
for tag in pdl while while_inline; do
    env=""
    [ $tag = while ] && env="DD_LOUVAIN_WHILE=1"
    [ $tag = while_inline ] && env="DD_LOUVAIN_WHILE=1 DD_KNN_INLINE=1"
    env $env python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2e_bench_$tag.json 2> gpurun_out/r2e_bench_$tag.err
done
python - <<'PY'
import json
for tag in ("pdl", "while", "while_inline"):
    try:
        l = json.load(open(f"gpurun_out/r2e_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: v for k, v in l["kernel_ms_total"].items() if k.startswith("lv") or k in ("knn_tc", "tc_gemm_dq", "tc_gemm_dty")}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e)
PY
