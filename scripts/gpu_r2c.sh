#!/bin/bash
# round 2, session c: 8-lane Louvain propose + parallel graph scan; kNN stream A/B; parity
mkdir -p gpurun_out
python -m pytest tests/test_gpu_e2e_parity.py tests/test_gpu_parity.py tests/test_gpu_zz_leiden.py -m gpu -q -s > gpurun_out/r2c_parity_full.log 2>&1
grep -E "^\[|passed|failed|FAILED|Error" gpurun_out/r2c_parity_full.log | cut -c1-700 > gpurun_out/r2c_parity.log
for tag in default warp inline; do
    env=""
    [ $tag = warp ] && env="DD_LOUVAIN_WARP=1"
    [ $tag = inline ] && env="DD_KNN_INLINE=1"
    env $env python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2c_bench_$tag.json 2> gpurun_out/r2c_bench_$tag.err
done
cat gpurun_out/r2c_parity.log
python - <<'PY'
import json
for tag in ("default", "warp", "inline"):
    try:
        l = json.load(open(f"gpurun_out/r2c_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: v for k, v in l["kernel_ms_total"].items() if k.startswith("lv") or k in ("knn_tc", "tc_gemm_dq", "tc_gemm_dty")}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e)
PY
