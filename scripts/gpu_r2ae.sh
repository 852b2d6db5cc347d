#!/bin/bash
# round 2, session ae: ncu launch list of the bench command itself with the final code; PhenoGraph-default bench line
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80000 --csv --log-file gpurun_out/r2ae_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2ae_ncu_bench.log 2>&1
wc -l gpurun_out/r2ae_launches.csv
tail -c 600 gpurun_out/r2ae_ncu_bench.log
timeout 600 python bench.py --clustering phenograph --no-cpu-baseline --no-extra > gpurun_out/r2ae_bench_c3_phenograph.json 2> gpurun_out/r2ae_bench_c3_phenograph.err
python - <<'PY'
import json
l = json.loads([x for x in open("gpurun_out/r2ae_bench_c3_phenograph.json") if x.startswith("{")][-1])
print("phenograph", round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["roofline_kernel"], {k: round(v, 1) for k, v in list(l["kernel_ms_total"].items())[:6]})
PY
# keep the pulled file small: kernel name + duration only
python - <<'PY'
import csv, sys
rows = []
with open("gpurun_out/r2ae_launches.csv", newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
out = open("gpurun_out/r2ae_launches_compact.csv", "w")
out.write("id,kernel,stream,duration_ns\n")
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    val = r["Metric Value"].replace(",", "")
    unit = r.get("Metric Unit", "ns")
    ns = float(val) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    out.write(f'{r["ID"]},{name},{r.get("Stream","")},{ns:.0f}\n')
out.close()
PY
rm -f gpurun_out/r2ae_launches.csv
wc -l gpurun_out/r2ae_launches_compact.csv
