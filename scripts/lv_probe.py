"""GPU probe: the first Louvain level's CUDA-graph replay alone (a one-iteration fit: nothing overlaps it) and inside
the pipelined 25-iteration fit, with the number of rounds it ran.   python scripts/lv_probe.py [c3]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
counts = bench.make_counts(wl)
n, g = counts.shape
omega, npi = _pca_plan(n + n // 4, g, 30, 0)
h = _capi.Handle(0)
h.upload_counts(counts)
rng = np.random.default_rng(0)
kw = dict(pseudocount=0.1, standard_scaling=False, n_comp=30, n_power_iter=npi, n_host_threads=8)
h.fit_iterations(bench.draw_parents(rng, n, 2), omega, **kw)
h.set_kernel_timing(True)
for n_it in (1, 1, 25):
    before = h.kernel_timing_report()
    out = h.fit_iterations(bench.draw_parents(rng, n, n_it), omega, **kw)
    after = h.kernel_timing_report()
    d = {k: (after[k][0] - before.get(k, (0, 0))[0]) / n_it for k in after if k.startswith("lv") or k in ("knn_tc",)}
    print(f"n_iters={n_it}: per iteration ms {dict((k, round(v, 3)) for k, v in d.items())}; rounds {h.last_stage_ms('lv_rounds')}; "
          f"wall per iteration {out['stage_ms']['wall'] / n_it:.2f} ms", flush=True)
h.close()
