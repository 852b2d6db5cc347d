#!/bin/bash
mkdir -p gpurun_out
python scripts/hbm_write_bw.py 1.5 > gpurun_out/r1k_hbm.log 2>&1
DD_KNN_EPI=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn or end_to_end or pipeline or config2" 2>&1 | tail -5 > gpurun_out/r1k_tests_epi1.log
{
for epi in 0 1; do echo "== DD_KNN_EPI=$epi"; DD_KNN_EPI=$epi python scripts/stage_bench.py c3 3 2>&1 | grep -E "knn_tc|rep 2"; done
} > gpurun_out/r1k_stage.log 2>&1
cat gpurun_out/r1k_hbm.log gpurun_out/r1k_tests_epi1.log gpurun_out/r1k_stage.log
