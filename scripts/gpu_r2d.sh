#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_exact_pca.py -m gpu -q -s 2>&1 | grep -E "^\[|passed|failed|FAILED|Error|assert" | cut -c1-400 > gpurun_out/r2d_exact.log
cat gpurun_out/r2d_exact.log
for cv in none 100 50 0; do
    echo "=== DD_LV_CARVEOUT=$cv"
    if [ $cv = none ]; then python scripts/lv_probe.py c3; else DD_LV_CARVEOUT=$cv python scripts/lv_probe.py c3; fi
done 2>&1 | tee gpurun_out/r2d_lv_probe.log
