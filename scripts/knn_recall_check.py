"""GPU check: is the approximate 3xBF16 filter of the tcgen05 kNN wide enough?  Exact float64 brute force on
the host for a sample of query rows of the c3 embedding against the device's lists, k = 10 (lists of 16) and k = 31 (lists of
32), all-tiles and cluster-ordered kernels.    python scripts/knn_recall_check.py [c3] [n_sample]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n_sample = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
counts = bench.make_counts(wl)
n_cells, n_genes = counts.shape
h = _capi.Handle(0)
h.upload_counts(counts)
h.create_doublets(np.random.default_rng(0).choice(n_cells, size=(n_cells // 4, 2), replace=False))
h.normalise_log(h.median_lib_size(), 0.1)
omega, n_power = _pca_plan(n_cells + n_cells // 4, n_genes, 30, 0)
emb, _ = h.pca(30, omega, n_power)
emb = np.ascontiguousarray(emb, dtype=np.float32)
n = emb.shape[0]
e64 = emb.astype(np.float64)
rows = np.sort(np.random.default_rng(1).choice(n, size=min(n_sample, n), replace=False))
sq = (e64 ** 2).sum(1)
for k in (10, 31):
    # exact: float64 distances of the sampled rows to every point, ties by index
    truth = np.empty((rows.size, k), dtype=np.int64)
    for c0 in range(0, rows.size, 250):
        r = rows[c0:c0 + 250]
        d2 = sq[r][:, None] - 2.0 * (e64[r] @ e64.T) + sq[None, :]
        d2[np.arange(r.size), r] = -1.0  # self first
        part = np.argpartition(d2, k + 8, axis=1)[:, :k + 8]
        # exact re-evaluation of the shortlisted candidates (difference form), then order by (distance, index)
        for i in range(r.size):
            cand = part[i]
            dd = ((e64[r[i]][None, :] - e64[cand]) ** 2).sum(1)
            dd[cand == r[i]] = -1.0
            order = np.lexsort((cand, dd))
            truth[c0 + i] = cand[order][:k]
    h.upload_embedding(emb)
    for mode, tag in ((1, "all tiles"), (2, "cluster-ordered")):
        h.set_knn_mode(mode)
        idx, _ = h.knn(k)
        got = idx[rows].astype(np.int64)
        bad_rows = np.nonzero((got != truth).any(axis=1))[0]
        # a row is WRONG if its neighbour SET differs (order differences can only come from exact ties)
        wrong = [i for i in bad_rows if set(got[i]) != set(truth[i])]
        print(f"k = {k:2d} {tag:16s}: {rows.size} sampled rows, {len(bad_rows)} differ in some position, {len(wrong)} have a different "
              f"neighbour SET", flush=True)
        for i in wrong[:3]:
            miss = sorted(set(truth[i]) - set(got[i]))
            extra = sorted(set(got[i]) - set(truth[i]))
            dm = [float(np.sqrt(((e64[rows[i]] - e64[j]) ** 2).sum())) for j in miss]
            de = [float(np.sqrt(((e64[rows[i]] - e64[j]) ** 2).sum())) for j in extra]
            print(f"    row {rows[i]}: missing {miss} at {dm}, reported instead {extra} at {de}; |q| = {np.sqrt(sq[rows[i]]):.1f}")
    h.set_knn_mode(0)
h.close()
