#!/bin/bash
# compute-sanitizer over the kernels that were written without hardware access (dense-build variant 4, the list-driven kNN
# kernel, the Leiden branch of the fit loop) at a small size -- memcheck finds out-of-bounds / misaligned accesses, racecheck
# shared-memory hazards, synccheck barrier misuse.  ~20-50x slower than a plain run: keep the sizes small.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
log=gpurun_out/sanitize.log
: > $log
run() {  # tool, env..., -- command
    tool=$1; shift
    echo "=== compute-sanitizer --tool $tool $*" | tee -a $log
    timeout 900 env "${@:1:$(($#-3))}" compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 "${@: -3}" 2>&1 | tail -25 | tee -a $log
}
# dense build: default, variant 3, variant 4 (c2: 12.5k rows x 3k genes)
for v in 1 3 4; do
    for tool in memcheck racecheck; do
        run $tool DD_DENSE_V=$v python scripts/dense_bench.py c2
    done
done
# the whole pipeline once per clustering algorithm at the smoke size, memcheck + synccheck
cat > gpurun_out/_sanitize_fit.py <<'PY'
import sys, warnings, numpy as np
sys.path.insert(0, ".")
from doubletdetection_b200 import BoostClassifier
counts = np.random.default_rng(0).poisson(1.0, (700, 150))
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for algo in ("louvain", "phenograph", "leiden"):
        clf = BoostClassifier(n_iters=2, clustering_algorithm=algo, n_jobs=2).fit(counts)
        print(algo, "ok", np.nanmean(clf.doublet_score()))
PY
for tool in memcheck synccheck; do
    run $tool DD_X=0 python gpurun_out/_sanitize_fit.py -
done
# list-driven kNN kernel
run memcheck DD_X=0 python scripts/knn_listed_experiment.py c2
# weighted first Louvain level (three smallest cases)
for tool in memcheck racecheck; do
    run $tool DD_CHECK_CASES=3 python tests/gpu_weighted_level_check.py -
done
grep -c "ERROR SUMMARY: 0 errors" $log; grep "ERROR SUMMARY" $log
