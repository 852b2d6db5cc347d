"""Run N pipelined iterations at a workload size (for ncu launch lists).  usage: one_iter.py [c2|c3] [n_iters]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
counts = bench.make_counts(wl)
n, g = counts.shape
h = _capi.Handle(0)
h.upload_counts(counts)
rng = np.random.default_rng(0)
omega, npi = _pca_plan(n + n // 4, g, 30, 0)
kw = dict(pseudocount=0.1, standard_scaling=False, n_comp=30, n_power_iter=npi, n_host_threads=8)
for rep in range(3):
    par = bench.draw_parents(rng, n, n_iters)
    t0 = time.perf_counter()
    out = h.fit_iterations(par, omega, **kw)
    print("wall %.1f ms" % (1e3 * (time.perf_counter() - t0)), {k: round(v, 2) for k, v in out["stage_ms"].items()}, flush=True)
