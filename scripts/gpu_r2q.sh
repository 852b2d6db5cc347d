#!/bin/bash
# round 2, session q: filter certificate + exact fix-up, fallback test, c5 on ONE GPU (cluster-ordered kNN at 1.25 M rows)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_knn_clustered.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|error|assert" | cut -c1-300 | tee gpurun_out/r2q_tests.log
timeout 300 python scripts/knn_clustered_bench.py c3 2>&1 | tee gpurun_out/r2q_knn_bench.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2q_bench_louvain.json 2> gpurun_out/r2q_bench_louvain.err
python bench.py --steps 2 --warmup 2 --clustering phenograph --no-cpu-baseline --no-extra > gpurun_out/r2q_bench_pheno.json 2> gpurun_out/r2q_bench_pheno.err
python - <<'PY'
import json
for tag in ("louvain", "pheno"):
    try:
        l = json.load(open(f"gpurun_out/r2q_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in list(l["kernel_ms_total"].items())[:12]}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2q_bench_{tag}.err").read()[-600:])
PY
timeout 900 python bench.py --shard cells --workload c5 --iters 4 --steps 1 --warmup 1 --no-extra > gpurun_out/r2q_c5_1gpu.json 2> gpurun_out/r2q_c5_1gpu.err
tail -c 1500 gpurun_out/r2q_c5_1gpu.json; tail -3 gpurun_out/r2q_c5_1gpu.err
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_e2e_parity.py -m gpu -q -x 2>&1 | tail -3 | tee -a gpurun_out/r2q_tests.log
