#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_leiden.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2f_tests.log
for lanes in 1 2 3 4; do
  for v in pdl nopdl; do
    for knn in own inline; do
      env="DD_LV_LANES=$lanes"
      [ $v = nopdl ] && env="$env DD_LOUVAIN_NO_PDL=1"
      [ $knn = inline ] && env="$env DD_KNN_INLINE=1"
      echo "=== lanes=$lanes $v knn=$knn"
      env $env python scripts/lv_probe.py c3 2>&1 | tail -1
    done
  done
done 2>&1 | tee gpurun_out/r2f_lanes.log
