#!/bin/bash
# round 2, session v: dense-build variant 5 as the default; 18-warp flavour
mkdir -p gpurun_out
: > gpurun_out/r2v_dense.log
for cfg in "5 18" "5 16"; do
    set -- $cfg; v=$1; w=$2
    echo "=== DD_DENSE_V=$v warps=$w: parity" | tee -a gpurun_out/r2v_dense.log
    env DD_DENSE_V=$v DD_DENSE_WARPS=$w timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
        -k "dense or doublets or end_to_end or wide_matrix or pipeline_matches" 2>&1 | tail -2 | tee -a gpurun_out/r2v_dense.log
    echo "=== DD_DENSE_V=$v warps=$w: c3 timing" | tee -a gpurun_out/r2v_dense.log
    env DD_DENSE_V=$v DD_DENSE_WARPS=$w timeout 300 python scripts/dense_bench.py c3 6 2>&1 | tail -1 | tee -a gpurun_out/r2v_dense.log
done
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2v_tests.log
for w in 16 18; do
    DD_DENSE_WARPS=$w python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2v_bench_w$w.json 2> gpurun_out/r2v_bench_w$w.err
    python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
l = json.load(open(f"gpurun_out/r2v_bench_w{w}.json"))
print("warps", w, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), "dense_rows", l["rooflines"].get("dense_rows"))
PY
done 2>&1 | tee -a gpurun_out/r2v_dense.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_dense_rows" -s 1 -c 2 -o gpurun_out/r2v_dense_full python scripts/dense_bench.py c3 3 > gpurun_out/r2v_ncu.log 2>&1
