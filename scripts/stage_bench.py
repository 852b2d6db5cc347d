"""Time one stage of the hot path at a workload size through the C ABI (CUDA-event stage timers).
usage: python scripts/stage_bench.py [c2|c3] [reps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
counts = bench.make_counts(wl)
n, g = counts.shape
h = _capi.Handle(0)
h.upload_counts(counts)
rng = np.random.default_rng(0)
omega, npi = _pca_plan(n + n // 4, g, 30, 0)
h.set_kernel_timing(True)
for r in range(reps):
    par = rng.choice(n, size=(n // 4, 2), replace=False)
    h.create_doublets(par)
    med = h.median_lib_size()
    h.normalise_log(med, 0.1)
    t_norm = h.last_stage_ms("normalise")
    h.pca(30, omega, npi)
    t_pca = h.last_stage_ms("pca")
    h.knn(10)
    t_knn = h.last_stage_ms("knn")
    print(f"rep {r}: doublets(csr) {h.last_stage_ms('doublets'):.3f} normalise {t_norm:.3f} pca {t_pca:.3f} knn {t_knn:.3f} ms", flush=True)
rep = h.kernel_timing_report()
for k, (ms, cnt) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:18s} {ms / cnt * 1e3:10.1f} us/launch  x{cnt}")
