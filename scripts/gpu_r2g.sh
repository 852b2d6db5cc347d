#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_leiden.py tests/test_gpu_e2e_parity.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2g_tests.log
for cfg in "default" "DD_KNN_WIDE=1" "DD_KNN_INLINE=1" "DD_LV_LANES=1" "DD_LV_LANES=1 DD_LOUVAIN_NO_PDL=1" "DD_LOUVAIN_NO_PDL=1"; do
    echo "=== $cfg"
    if [ "$cfg" = default ]; then python scripts/lv_probe.py c3 2>&1 | tail -1; else env $cfg python scripts/lv_probe.py c3 2>&1 | tail -1; fi
done 2>&1 | tee gpurun_out/r2g_probe.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2g_bench.json"))
print(round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["kernel_ms_total"], l["stage_ms_per_step"], l["roofline_kernel"])
PY
