#!/bin/bash
# round 2, session aj: the driver's default bench command once more with the final bench.py
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2aj_bench_c3.json 2> gpurun_out/r2aj_bench_c3.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2aj_bench_c3.json"))
print("ours", round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["roofline_kernel"], round(l["roofline"]["frac"], 4),
      "critical", l["roofline_critical_path"]["kernel"], round(l["roofline_critical_path"]["frac"], 3), l["roofline_critical_path"]["frac_alone"], l["clocks"])
PY
tail -3 gpurun_out/r2aj_bench_c3.err
