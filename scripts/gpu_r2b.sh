#!/bin/bash
# round 2, session b: the asserted end-to-end parity tests + the thread-per-node Louvain propose kernel (A/B against the
# warp-per-node one, DD_LOUVAIN_WARP=1) + the kNN on its own stream (A/B against DD_KNN_INLINE=1)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_e2e_parity.py "tests/test_gpu_parity.py::test_classifier_end_to_end_vs_golden" \
    tests/test_gpu_parity.py::test_gpu_louvain_level0_matches_host_twin tests/test_gpu_parity.py::test_pipeline_matches_stagewise_calls \
    tests/test_gpu_zz_leiden.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -70 > gpurun_out/r2b_parity.log
for tag in default warp inline inline_warp; do
    env=""
    [ $tag = warp ] && env="DD_LOUVAIN_WARP=1"
    [ $tag = inline ] && env="DD_KNN_INLINE=1"
    [ $tag = inline_warp ] && env="DD_KNN_INLINE=1 DD_LOUVAIN_WARP=1"
    env $env python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2b_bench_$tag.json 2> gpurun_out/r2b_bench_$tag.err
done
cat gpurun_out/r2b_parity.log
python - <<'PY'
import json
for tag in ("default", "warp", "inline", "inline_warp"):
    try:
        l = json.load(open(f"gpurun_out/r2b_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: v for k, v in l["kernel_ms_total"].items() if k.startswith("lv") or k in ("knn_tc", "tc_gemm_dq", "tc_gemm_dty")}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e)
PY
