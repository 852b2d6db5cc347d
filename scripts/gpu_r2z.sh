#!/bin/bash
# round 2, session z: HVG on the device vs the host lines at a realistic size; ncu --set full of the PhenoGraph level / HVG kernels
mkdir -p gpurun_out
timeout 600 python scripts/hvg_bench.py 50000 20000 0.05 3000 2>&1 | tee gpurun_out/r2z_hvg_bench.log
cat > gpurun_out/_pheno_once.py <<'PY'
import sys, warnings, numpy as np
sys.path.insert(0, ".")
import bench
from doubletdetection_b200 import BoostClassifier
counts = bench.make_counts(bench.WORKLOADS["c2"])
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    BoostClassifier(n_iters=2, n_jobs=4).fit(counts)
print("ok")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_lvw_propose_g|k_jaccard_weights|k_lv_propose_g|k_hvg_moments|k_hvg_scatter" -c 10 -o gpurun_out/r2z_full python gpurun_out/_pheno_once.py > gpurun_out/r2z_ncu.log 2>&1
tail -2 gpurun_out/r2z_ncu.log
