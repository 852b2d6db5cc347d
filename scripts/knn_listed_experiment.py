"""GPU experiment for DESIGN.md section 5 ("next lever"): does the tcgen05 kNN get proportionally faster when a
256-row query block only visits the candidate tiles it can need?  (Written after round 1's GPU budget was spent: the
kernel variant k_knn_tc<LIST, true> and dd_knn_listed have never run.)

    gpurun --timeout 900 -- 'python scripts/knn_listed_experiment.py c3 > gpurun_out/knn_listed.log 2>&1'

Everything except the kernel is done on the HOST here, on purpose -- the question is the kernel's speed and the
exactness of the scheme, not yet the device pre-pass:
  1. one iteration's embedding from the pipeline (upload, doublets, normalise, PCA) and the dense exact kNN as the truth;
  2. points ordered by cluster (k-means on the embedding), every cluster padded to whole 256-row blocks (pad rows sit
     at 1e12 and can never be selected) and sorted along its principal direction; permuted embedding uploaded;
  3. launch A: every block against the tiles of its own cluster -> an upper bound on each query's 10th distance;
  4. launch B: every block against the tiles whose box-to-box distance^2 (all dimensions) is within the block's
     largest bound;
  5. the result, mapped back, must equal the dense kNN index for index; kernel times are printed next to each other;
  6. the same scheme once more with steps 2-4 on the device as well (dd_knn_pruned: gather into the padded order, per-tile
     boxes, thresholds from launch A, list construction, translation back), given only the permutation.
"""
import os
import sys

import numpy as np
from sklearn.cluster import KMeans

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from doubletdetection_b200 import _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
n_clusters = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K, T, QT = 10, 128, 256
counts = bench.make_counts(wl)
n_cells, n_genes = counts.shape
n_synth = n_cells // 4
h = _capi.Handle(0)
h.upload_counts(counts)
h.create_doublets(np.random.default_rng(0).choice(n_cells, size=(n_synth, 2), replace=False))
h.normalise_log(h.median_lib_size(), 0.1)
omega, n_power = _pca_plan(n_cells + n_synth, n_genes, 30, 0)
emb, _ = h.pca(30, omega, n_power)
emb = np.ascontiguousarray(emb, dtype=np.float32)
n = emb.shape[0]
h.set_kernel_timing(True)


def kernel_ms(name):
    rep = h.kernel_timing_report()
    return rep.get(name, (0.0, 0))


for _ in range(3):
    truth_idx, truth_dist = h.knn(K)
t_dense = kernel_ms("knn_tc")
print(f"dense kNN: {n} points, k_knn_tc {t_dense[0] / max(t_dense[1], 1):.3f} ms per launch, stage {h.last_stage_ms('knn'):.3f} ms",
      flush=True)

# ---- host pre-pass: cluster order, padding, permutation
lab = KMeans(n_clusters, n_init=1, max_iter=5, random_state=0).fit(emb).labels_
order = []
for c in range(n_clusters):
    ids = np.nonzero(lab == c)[0]
    if ids.size == 0:
        continue
    x = emb[ids] - emb[ids].mean(0)
    ids = ids[np.argsort(x @ np.linalg.svd(x, full_matrices=False)[2][0])]
    pad = (-ids.size) % QT
    order.append(np.concatenate([ids, np.full(pad, -1, dtype=ids.dtype)]))
blocks_of_cluster = np.concatenate([np.full(len(o) // QT, c) for c, o in enumerate(order)])
perm = np.concatenate(order)  # padded position -> original index or -1
n_pad = perm.size
real = perm >= 0
emb_p = np.zeros((n_pad, emb.shape[1]), dtype=np.float32)
emb_p[real] = emb[perm[real]]
emb_p[~real, 0] = 1e12
n_blocks, n_tiles = n_pad // QT, n_pad // T
print(f"{n_clusters} clusters -> {n_blocks} query blocks, {n_tiles} candidate tiles ({n_pad / n:.3f} x rows)", flush=True)
h.upload_embedding(emb_p)

big = np.float32(3e38)
lo = np.array([np.where(real[t * T:(t + 1) * T, None], emb_p[t * T:(t + 1) * T], big).min(0) for t in range(n_tiles)])
hi = np.array([np.where(real[t * T:(t + 1) * T, None], emb_p[t * T:(t + 1) * T], -big).max(0) for t in range(n_tiles)])
empty_tile = ~np.array([real[t * T:(t + 1) * T].any() for t in range(n_tiles)])
qlo = np.minimum(lo[0::2], lo[1::2])
qhi = np.maximum(hi[0::2], hi[1::2])


def run(lists, tag):
    off = np.zeros(n_blocks + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(x) for x in lists])
    tiles = np.concatenate(lists).astype(np.int32) if off[-1] else np.zeros(0, dtype=np.int32)
    before = kernel_ms("knn_tc_listed")
    for _ in range(3):
        idx, dist = h.knn_listed(K, off, tiles)
    after = kernel_ms("knn_tc_listed")
    ms = (after[0] - before[0]) / max(after[1] - before[1], 1)
    print(f"{tag}: {off[-1]} block-tile pairs = {off[-1] / (n_blocks * n_tiles):.3f} of all, k_knn_tc<listed> {ms:.3f} ms per launch, "
          f"stage {h.last_stage_ms('knn'):.3f} ms", flush=True)
    return idx, dist


# ---- launch A: own cluster
tile_cluster = np.repeat(blocks_of_cluster, QT // T)
lists_a = [np.nonzero((tile_cluster == blocks_of_cluster[b]) & ~empty_tile)[0] for b in range(n_blocks)]
idx_a, dist_a = run(lists_a, "launch A (own cluster)")
tau = np.where(idx_a[:, K - 1] >= 0, dist_a[:, K - 1].astype(np.float64) ** 2, np.inf)
tau[~real] = 0.0
thr = tau.reshape(n_blocks, QT).max(1)

# ---- launch B: every tile the bound cannot exclude
lists_b = []
for b in range(n_blocks):
    gap = np.maximum(0, np.maximum(lo - qhi[b], qlo[b] - hi)).astype(np.float64)
    lb = (gap ** 2).sum(1)
    need = (lb <= thr[b] * (1 + 1e-5)) & ~empty_tile
    order_b = np.nonzero(need)[0]
    lists_b.append(order_b[np.argsort(lb[order_b], kind="stable")])
idx_b, dist_b = run(lists_b, "launch B (bounded)")

# ---- back to the original numbering and compare
got = np.full((n, K), -1, dtype=np.int64)
rows = perm[real]
mapped = np.where(idx_b[real] >= 0, perm[np.maximum(idx_b[real], 0)], -1)
got[rows] = mapped
same = (got == truth_idx).all(1)
print(f"rows identical to the dense kNN: {int(same.sum())} / {n}", flush=True)
if not same.all():
    bad = np.nonzero(~same)[0][:5]
    for r in bad:
        print("  row", r, "got", got[r], "dense", truth_idx[r], "dist", truth_dist[r])
    # equal distances can legitimately come back in another order only if their indices tie-break differently after the
    # permutation: the refine step ranks by (distance, PERMUTED index)
    print("  max |distance difference| over all rows:", float(np.abs(np.sort(dist_b[real], 1)[np.argsort(rows)] - np.sort(truth_dist, 1)).max()))

# ---- the same scheme with the pre-pass on the device too (dd_knn_pruned: gather, boxes, thresholds, lists, translation)
h.upload_embedding(emb)
h.knn(K)  # sizes the output buffers for the original embedding
before = h.kernel_timing_report()
for _ in range(3):
    idx_p, dist_p, stats = h.knn_pruned(K, perm.astype(np.int32), blocks_of_cluster.astype(np.int32), n)
after = h.kernel_timing_report()
print(f"dd_knn_pruned: launch A {stats['pairs_a']} + launch B {stats['pairs_b']} block-tile pairs of {stats['blocks'] * stats['tiles']} "
      f"({(stats['pairs_a'] + stats['pairs_b']) / (stats['blocks'] * stats['tiles']):.3f}); stage {h.last_stage_ms('knn'):.3f} ms; per call:",
      flush=True)
for name in ("prune_gather", "knn_prep", "prune_boxes", "knn_tc_listed", "knn_refine", "prune_threshold", "prune_lists",
             "prune_offsets", "prune_translate"):
    a, b = after.get(name, (0.0, 0)), before.get(name, (0.0, 0))
    if a[1] > b[1]:
        print(f"    {name:16s} {(a[0] - b[0]) / 3:.3f} ms in {(a[1] - b[1]) // 3} launch(es)")
same_p = (idx_p.astype(np.int64) == truth_idx).all(1)
print(f"dd_knn_pruned rows identical to the dense kNN: {int(same_p.sum())} / {n}; lists of launch B equal to the host's: "
      f"{stats['pairs_b'] == int(sum(len(x) for x in lists_b))}", flush=True)
h.close()
