#!/bin/bash
# round 2, session w: variant 5 with the Newton-corrected division; rooflines_alone in the bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense or doublets or end_to_end or wide_matrix or pipeline_matches or lib_size" 2>&1 | tail -2 | tee gpurun_out/r2w_tests.log
timeout 300 python scripts/dense_bench.py c3 6 2>&1 | tail -1 | tee gpurun_out/r2w_dense.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2w_bench.json"))
print(round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]))
print("in the loop", {k: (round(v["ms_per_launch"], 3), round(v["frac"], 3)) for k, v in l["rooflines"].items()})
print("alone      ", {k: (round(v["ms_per_launch"], 3), round(v["frac"], 3)) for k, v in l["rooflines_alone"].items()})
PY
tail -3 gpurun_out/r2w_bench.err
