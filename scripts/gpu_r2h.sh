#!/bin/bash
# round 2, session h: PhenoGraph's first level on the device by default (graph replay), new goldens; baseline benches
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pheno_level0.py tests/test_gpu_zz_leiden.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2h_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2h_bench_louvain.json 2> gpurun_out/r2h_bench_louvain.err
for tag in pheno pheno_nograph pheno_hostlevel; do
    env="X=1"
    [ $tag = pheno_nograph ] && env="DD_LVW_NO_GRAPH=1"
    [ $tag = pheno_hostlevel ] && env="DD_PHENO_LEVEL0=0"
    env $env python bench.py --steps 2 --warmup 3 --clustering phenograph --no-cpu-baseline --no-extra > gpurun_out/r2h_bench_$tag.json 2> gpurun_out/r2h_bench_$tag.err
done
python - <<'PY'
import json
for tag in ("louvain", "pheno", "pheno_nograph", "pheno_hostlevel"):
    try:
        l = json.load(open(f"gpurun_out/r2h_bench_{tag}.json"))
        print(tag, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), {k: round(v, 1) for k, v in l["kernel_ms_total"].items()}, l["stage_ms_per_step"])
    except Exception as e:
        print(tag, "failed", e, open(f"gpurun_out/r2h_bench_{tag}.err").read()[-600:])
PY
# per-launch durations of the list-driven kNN experiment (is launch A's 0.4 ms real?)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knn_tc --csv --log-file gpurun_out/r2h_knn_listed_launches.csv python scripts/knn_listed_experiment.py c3 128 > gpurun_out/r2h_knn_listed_ncu.log 2>&1
tail -5 gpurun_out/r2h_knn_listed_ncu.log
