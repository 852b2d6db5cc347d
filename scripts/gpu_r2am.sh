#!/bin/bash
# round 2, session am: the bench's reference_defaults leg on its own (BoostClassifier() defaults at c3 through the public API)
mkdir -p gpurun_out
timeout 60 python - <<'PY' 2>&1 | tail -4 | tee gpurun_out/r2am_reference_defaults.log
import json, os
import bench
counts = bench.make_counts(bench.WORKLOADS["c3"])
print(json.dumps(bench.reference_defaults_leg(counts, 0, max(1, os.cpu_count() or 1))))
PY
