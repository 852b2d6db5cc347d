#!/bin/bash
# round 2, session ai: Leiden and PhenoGraph lines after the host-side work (prefetched sweeps, bit-set aggregation)
mkdir -p gpurun_out
for algo in leiden phenograph; do
    steps="--steps 1 --warmup 1"; [ $algo = phenograph ] && steps="--steps 3 --warmup 2"
    timeout 600 python bench.py $steps --clustering $algo --no-cpu-baseline --no-extra > gpurun_out/r2ai_bench_c3_$algo.json 2> gpurun_out/r2ai_bench_c3_$algo.err
    python - $algo <<'PY'
import json, sys
a = sys.argv[1]
try:
    l = json.loads([x for x in open(f"gpurun_out/r2ai_bench_c3_{a}.json") if x.startswith("{")][-1])
    print(a, round(l["value"]), round(l["ms_per_step"], 1), "e2e", round(l["e2e"]["value"]), l["stage_ms_per_step"])
except Exception as e:
    print(a, "failed", e); print(open(f"gpurun_out/r2ai_bench_c3_{a}.err").read()[-1500:])
PY
done
