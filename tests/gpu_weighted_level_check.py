"""GPU check of the EXPERIMENTAL weighted first Louvain level (louvain_gpu_w.cu, dd_louvain_level0_weighted) against its
specification (oracle/louvain_ref.py:level0_parallel with weights) -- label for label -- on umap-weighted and PhenoGraph
(Jaccard) graphs of growing size, with its time next to the host sweep it is meant to replace.  Never run yet (written after
round 1's GPU budget was spent).  A script (not collected by pytest), not a test, until it has passed once; it lives under tests/ because it uses the oracle as
its checker.

    gpurun --timeout 900 -- 'python tests/gpu_weighted_level_check.py > gpurun_out/weighted_level.log 2>&1'
"""
import os
import sys
import time
import warnings

os.environ["DD_PHENO_LEVEL0"] = "1"  # read once by dd_fit_iterations: PhenoGraph's first level on the device (last section)

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from doubletdetection_b200 import _capi  # noqa: E402
from oracle import louvain_ref, upstream  # noqa: E402  (checker only)

h = _capi.Handle(0)
ok_all = True
CASES = [(300, 6, "umap", 1.0, 3), (2000, 10, "umap", 4.0, 0), (3000, 31, "jaccard", 1.0, 1), (20000, 31, "jaccard", 1.0, 2),
         (60000, 31, "jaccard", 1.0, 0)]
for n, k, kind, gamma, seed in CASES[: int(os.environ.get("DD_CHECK_CASES", len(CASES)))]:
    rs = np.random.default_rng(n)
    pts = (rs.normal(size=(n, 8)) + rs.integers(0, 6, size=(n, 1)) * 2.5).astype(np.float32)
    h.upload_embedding(pts)
    idx, dist = h.knn(k)  # exact kNN from the device (the oracle's brute force is too slow at 60k)
    if kind == "jaccard":
        G = h.jaccard_graph(k, prune=True)  # device-built, zeros dropped, sorted rows
    else:
        G = _capi.umap_connectivities(idx, dist).astype(np.float64)
    w = np.asarray(G.data, dtype=np.float64)
    t0 = time.perf_counter()
    want = louvain_ref.level0_parallel(G.indptr, G.indices, gamma, seed, w)
    t_spec = time.perf_counter() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        got, rounds = h.louvain_level0_weighted(G.indptr, G.indices, w, gamma, seed)
        t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    _capi.louvain_csr(G.indptr, G.indices, w, gamma, seed)
    t_seq = time.perf_counter() - t0
    same = bool(np.array_equal(got, want))
    ok_all &= same
    print(f"n={n} k={k} {kind} gamma={gamma}: nnz {G.nnz}, rounds {rounds}, communities {len(np.unique(got))} / {len(np.unique(want))}, "
          f"device == specification: {same}; device call {1e3 * t_dev:.1f} ms (incl. uploads, {rounds} x 17 launches), numpy "
          f"specification {t_spec:.2f} s, sequential host Louvain (all levels) {1e3 * t_seq:.1f} ms", flush=True)
    if not same:
        bad = np.nonzero(got != want)[0]
        print("  first differences:", [(int(i), int(got[i]), int(want[i])) for i in bad[:8]], "of", bad.size)
print("ALL EQUAL" if ok_all else "MISMATCH")
h.close()

# ---- the same level inside the fit loop (DD_PHENO_LEVEL0=1): BoostClassifier(phenograph) against the oracle's PhenoGraph
# restatement with the parallel first level
from doubletdetection_b200 import BoostClassifier  # noqa: E402
from oracle import datasets, reference_path  # noqa: E402

counts = datasets.structured_counts(1500, 300, seed=1234)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t0 = time.perf_counter()
    clf = BoostClassifier(n_iters=3, random_state=0, n_jobs=2).fit(counts)
    t_fit = time.perf_counter() - t0
    ora = reference_path.OracleClassifier(n_iters=3, random_state=0, clustering_algorithm="phenograph",
                                          clustering_kwargs={"level0": "parallel"}).fit(counts)
same = (clf.communities_ == ora.communities_).all(axis=1)
print(f"fit loop with the device level: iterations with communities identical to the oracle (parallel first level): "
      f"{int(same.sum())}/{same.size}; fit {1e3 * t_fit:.0f} ms; stage ms {clf.stage_ms_}", flush=True)
