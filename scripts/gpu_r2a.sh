#!/bin/bash
# First GPU session of round 2, one call (~20-25 min of box time):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_r2a.sh'
# 1. the whole GPU suite (new since the last hardware run: tests/test_gpu_zz_leiden.py, per-device kernel attributes, the
#    upload thread in fit(), the dense first aggregation of the host Louvain);
# 2. the bench line (e2e should move: predict() 42 -> 8 ms, upload underneath the parent draws);
# 3. dense-build variants incl. the never-run variant 4;  4. the list-driven kNN experiment.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2a_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
bash scripts/gpu_dense_variants.sh > gpurun_out/r2a_dense_variants.out 2>&1
timeout 900 python scripts/knn_listed_experiment.py c3 64 > gpurun_out/r2a_knn_listed.log 2>&1
timeout 600 python scripts/knn_listed_experiment.py c3 128 >> gpurun_out/r2a_knn_listed.log 2>&1
timeout 600 python tests/gpu_weighted_level_check.py > gpurun_out/r2a_weighted_level.log 2>&1
cat gpurun_out/r2a_tests.log gpurun_out/dense_variants.log gpurun_out/r2a_knn_listed.log gpurun_out/r2a_weighted_level.log
tail -c 1500 gpurun_out/r2a_bench_c3.json
