#!/bin/bash
# round 2, session af: ncu launch list of the pipelined loop with the final code (one loop, 2 iterations x 3 calls: the
# two-pipeline bench command stops under ncu after a few iterations); graph-hook tests after the knn_last_k check
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2af_launches.csv \
    python scripts/one_iter.py c3 2 > gpurun_out/r2af_ncu_launches.log 2>&1
grep "^==" gpurun_out/r2af_launches.csv | head -8
tail -3 gpurun_out/r2af_ncu_launches.log
wc -l gpurun_out/r2af_launches.csv
timeout 600 python -m pytest tests/test_gpu_zz_leiden.py tests/test_gpu_pheno_level0.py tests/test_gpu_parity.py -q -m gpu -k "umap or jaccard or pheno or leiden" 2>&1 | tail -3 | tee gpurun_out/r2af_tests.log
