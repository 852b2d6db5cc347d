#!/bin/bash
# round 2, session al: SURVEY Q7 (at most 50 genes: neighbours on X instead of X_pca) on the exact-PCA route
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_exact_pca.py -q -m gpu -s -k "at_most_50" 2>&1 | tail -12 | tee gpurun_out/r2al_tests.log
