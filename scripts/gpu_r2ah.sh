#!/bin/bash
# round 2, session ah: what the driver runs at round end -- full GPU suite, smoke(), the default bench line (+ reference arm)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2ah_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2ah_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2ah_bench_reference.json 2> gpurun_out/r2ah_bench_reference.err
timeout 900 python bench.py > gpurun_out/r2ah_bench_c3.json 2> gpurun_out/r2ah_bench_c3.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/r2ah_bench_c3.json"))
r = json.load(open("gpurun_out/r2ah_bench_reference.json"))
print("ours", round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["e2e"]["host_ms_last_fit"], "cpu", round(l["cpu_baseline"]["value"]), l["parity_c2"], l["roofline_kernel"], l["roofline"]["frac"], l["clocks"])
print({k: round(v["frac"], 3) for k, v in l["rooflines"].items()})
print("reference arm", round(r["value"]), r["config"]["workload"] == l["config"]["workload"])
PY
