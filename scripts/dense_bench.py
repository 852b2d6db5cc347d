"""Time the dense build alone (no PCA behind it) at a workload size.  usage: dense_bench.py [c2|c3] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
counts = bench.make_counts(wl)
n = counts.shape[0]
h = _capi.Handle(0)
h.upload_counts(counts)
rng = np.random.default_rng(0)
ts = []
for r in range(reps):
    h.create_doublets(rng.choice(n, size=(n // 4, 2), replace=False))
    h.normalise_log(h.median_lib_size(), 0.1)
    ts.append(h.last_stage_ms("normalise"))
print("dense build ms:", " ".join(f"{t:.3f}" for t in ts), flush=True)
