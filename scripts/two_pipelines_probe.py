"""GPU probe: does the fit loop get faster when TWO pipelined loops (two handles, two host threads) share one GPU, each
running half of the iterations?  The small latency-bound kernels of one loop's PCA would run underneath the other loop's
HBM-bound products.    python scripts/two_pipelines_probe.py [c3] [n_pipelines]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n_pipes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
counts = bench.make_counts(wl)
n, g = counts.shape
omega, npi = _pca_plan(n + n // 4, g, 30, 0)
kw = dict(pseudocount=0.1, standard_scaling=False, n_comp=30, n_power_iter=npi, n_host_threads=8)
rng = np.random.default_rng(0)
handles = [_capi.Handle(0) for _ in range(n_pipes)]
for h in handles:
    h.upload_counts(counts)
    h.fit_iterations(bench.draw_parents(rng, n, 3), omega, **kw)


def run(n_it_total, pipes):
    parents = bench.draw_parents(rng, n, n_it_total)
    bounds = np.linspace(0, n_it_total, pipes + 1).astype(int)
    out = [None] * pipes

    def work(i):
        out[i] = handles[i].fit_iterations(parents, omega, iter_begin=int(bounds[i]), iter_end=int(bounds[i + 1]), **kw)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(pipes)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0


for pipes in range(1, n_pipes + 1):
    ts = [run(25, pipes) for _ in range(4)]
    print(f"{pipes} pipeline(s): 25 iterations in {1e3 * min(ts):.1f} ms (median {1e3 * np.median(ts):.1f}) = "
          f"{25 * (n + n // 4) / min(ts) / 1e6:.2f} M augmented-cells/s", flush=True)
for h in handles:
    h.close()
