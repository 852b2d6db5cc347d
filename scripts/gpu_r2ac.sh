#!/bin/bash
# round 2, session ac: umap's connectivities on the device (Leiden branch) -- tests, then c3 with clustering="leiden",
# device graph vs the host graph (DD_UMAP_HOST=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_leiden.py -x -q -m gpu -s 2>&1 | tail -40 > gpurun_out/r2ac_leiden_tests.log
tail -15 gpurun_out/r2ac_leiden_tests.log
for mode in 0 1; do
    DD_UMAP_HOST=$mode timeout 900 python bench.py --steps 1 --warmup 1 --clustering leiden --no-cpu-baseline --no-extra \
        > gpurun_out/r2ac_bench_c3_leiden_umaphost$mode.json 2> gpurun_out/r2ac_bench_c3_leiden_umaphost$mode.err
    python - "$mode" <<'PY'
import json, sys
mode = sys.argv[1]
try:
    txt = [l for l in open(f"gpurun_out/r2ac_bench_c3_leiden_umaphost{mode}.json").read().splitlines() if l.startswith("{")][-1]
    l = json.loads(txt)
    print("DD_UMAP_HOST", mode, "value", round(l["value"]), "ms/step", round(l["ms_per_step"], 1), l.get("stage_ms_per_step"),
          {k: round(v, 1) for k, v in list(l["kernel_ms_total"].items())[:8]}, "e2e", l.get("e2e", {}).get("value"))
except Exception as e:
    print("mode", mode, "failed", e)
    print(open(f"gpurun_out/r2ac_bench_c3_leiden_umaphost{mode}.err").read()[-2000:])
PY
done
