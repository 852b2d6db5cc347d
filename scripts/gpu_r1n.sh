#!/bin/bash
# round 1n GPU session: full parity suite, bench line, PhenoGraph timing, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r1n_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r1n_bench_c3.json 2> gpurun_out/r1n_bench_c3.err
python - > gpurun_out/r1n_phenograph.log 2>&1 <<PY
import time, warnings, numpy as np, bench
from doubletdetection_b200 import BoostClassifier
for wl in ("c2", "c3"):
    counts = bench.make_counts(bench.WORKLOADS[wl])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_iters=10, n_jobs=16)
        clf.fit(counts)
        t = time.perf_counter(); clf.fit(counts); dt = time.perf_counter() - t
    n_aug = counts.shape[0] * 1.25
    print(wl, "BoostClassifier() defaults (phenograph, n_iters=10): %.3f s/fit = %.2f M aug-cells/s" % (dt, 10 * n_aug / dt / 1e6),
          {k: round(v, 1) for k, v in clf.stage_ms_.items()}, "NaN frac %.4f" % np.isnan(clf.all_scores_).mean(), flush=True)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1n_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r1n_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc|k_tc_gemm|k_dense_rows|k_jaccard" -s 10 -c 6 -o gpurun_out/r1n_full python scripts/stage_bench.py c3 2 > gpurun_out/r1n_ncu_full.log 2>&1
cat gpurun_out/r1n_tests.log gpurun_out/r1n_phenograph.log; tail -c 1200 gpurun_out/r1n_bench_c3.json; ls -la gpurun_out | tail -8
