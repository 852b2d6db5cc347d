"""GPU timing of the cluster-ordered kNN against the all-tiles kernel on the c3 embedding, kernel by kernel.
    python scripts/knn_clustered_bench.py [c3]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
counts = bench.make_counts(wl)
n_cells, n_genes = counts.shape
h = _capi.Handle(0)
h.upload_counts(counts)
h.create_doublets(np.random.default_rng(0).choice(n_cells, size=(n_cells // 4, 2), replace=False))
h.normalise_log(h.median_lib_size(), 0.1)
omega, n_power = _pca_plan(n_cells + n_cells // 4, n_genes, 30, 0)
h.pca(30, omega, n_power)
h.set_kernel_timing(True)
for mode, tag in ((1, "all tiles"), (2, "cluster-ordered")):
    h.set_knn_mode(mode)
    h.knn(10)
    before = h.kernel_timing_report()
    stage = []
    for _ in range(5):
        idx, _ = h.knn(10)
        stage.append(h.last_stage_ms("knn"))
    after = h.kernel_timing_report()
    print(f"{tag}: stage {np.median(stage):.3f} ms (min {min(stage):.3f})")
    for name in sorted(after):
        a, b = after[name], before.get(name, (0.0, 0))
        if a[1] > b[1]:
            print(f"    {name:18s} {(a[0] - b[0]) / 5:.3f} ms in {(a[1] - b[1]) // 5} launch(es)")
    if mode == 1:
        truth = idx
    else:
        print("    identical to the all-tiles result:", bool((idx == truth).all()), h.knn_clustered_stats())
h.close()
