#!/bin/bash
# round 2, session u: dense-build variant 5 (variant 4 + next row's gathers prefetched into registers)
mkdir -p gpurun_out
: > gpurun_out/r2u_dense.log
for cfg in "5 16" "5 12" "5 8" "4 16" "1 0"; do
    set -- $cfg; v=$1; w=$2
    wenv="X=1"; [ "$w" != 0 ] && wenv="DD_DENSE_WARPS=$w"
    echo "=== DD_DENSE_V=$v warps=$w: parity" | tee -a gpurun_out/r2u_dense.log
    env DD_DENSE_V=$v $wenv timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
        -k "dense or doublets or end_to_end or wide_matrix or pipeline_matches" 2>&1 | tail -2 | tee -a gpurun_out/r2u_dense.log
    echo "=== DD_DENSE_V=$v warps=$w: c3 timing" | tee -a gpurun_out/r2u_dense.log
    env DD_DENSE_V=$v $wenv timeout 300 python scripts/dense_bench.py c3 6 2>&1 | tail -1 | tee -a gpurun_out/r2u_dense.log
done
for v in 5 1; do
    env DD_DENSE_V=$v python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2u_bench_v$v.json 2> gpurun_out/r2u_bench_v$v.err
    python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
l = json.load(open(f"gpurun_out/r2u_bench_v{v}.json"))
print("DD_DENSE_V=" + v, round(l["value"]), round(l["ms_per_step"], 1), "dense_rows ms/launch", l["rooflines"].get("dense_rows", {}).get("ms_per_launch"), l["rooflines"].get("dense_rows", {}).get("frac"))
PY
done 2>&1 | tee -a gpurun_out/r2u_dense.log
