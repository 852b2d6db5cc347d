#!/bin/bash
# round 1j GPU session: parity, kNN epilogue A/B, dense-build experiments, ncu captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r1j_tests.log
DD_KNN_EPI=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "knn or end_to_end or pipeline or config2" 2>&1 | tail -5 > gpurun_out/r1j_tests_epi1.log
{
for epi in 0 1; do echo "== DD_KNN_EPI=$epi"; DD_KNN_EPI=$epi python scripts/stage_bench.py c3 3 2>&1 | grep -E "knn_tc|rep 2"; done
for dbg in 0 1 2; do echo "== DD_DENSE_DBG=$dbg"; DD_DENSE_DBG=$dbg python scripts/stage_bench.py c3 3 2>&1 | grep -E "dense_rows"; done
} > gpurun_out/r1j_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_dense_rows|k_knn_tc" -c 3 -o gpurun_out/r1j_full_a python scripts/stage_bench.py c3 1 > gpurun_out/r1j_ncu_a.log 2>&1
DD_KNN_EPI=1 DD_DENSE_V=0 ncu --set full --clock-control none --import-source on -k regex:"k_dense_rows|k_knn_tc" -c 3 -o gpurun_out/r1j_full_b python scripts/stage_bench.py c3 1 > gpurun_out/r1j_ncu_b.log 2>&1
ls -la gpurun_out
cat gpurun_out/r1j_tests.log gpurun_out/r1j_tests_epi1.log gpurun_out/r1j_stage.log
