"""CPU study behind DESIGN.md section 5 ("cluster-aware tile pruning"): how many (query tile, candidate tile) pairs of
the exact kNN could be skipped with valid lower bounds on a BASELINE-shaped embedding?  No GPU, no oracle: the augmented
matrix of one iteration is rebuilt with scipy / sklearn from bench.make_counts.

    python scripts/knn_prune_study.py [c2|c3] [--sim]

Points are ordered by k-means clusters of the 8 signal components (tiles never straddle clusters); a tile pair is needed
iff the box-to-box distance^2 in the leading m components is <= the largest exact 10th-neighbour distance^2 of the query
tile.  Measured (tile slots kept, padding included, 256-query x 128-candidate tiles as in k_knn_tc):

    c2 (12.5k points)   8 clusters 0.90   36 clusters 0.96   64 clusters 0.86   (clusters of ~200 points: padding eats it)
    c3 (125k points)    8 clusters 0.50   36 clusters 0.27   64 clusters 0.21   128 clusters 0.21

The synthetic doublets sit BETWEEN the cell types and are what keeps pairs alive (36 % of all pairs involve one); the
cells of different types never need each other.  With points in their natural order (what round 1 tried) every pair is
kept.  At c3 a cluster-ordered kNN would therefore score ~1/5 of the tile pairs -- the largest remaining lever on the kNN
kernel ("next").

--sim replays what the kernel would do, with thresholds it can actually know: every cluster is padded to whole 256-row
query tiles and sorted along its own principal direction, a query tile visits the candidate tiles by increasing box
bound (all 30 dimensions) and stops when the bound exceeds the largest CURRENT 10th-best distance of its 256 queries.
c3, k-means with 5 Lloyd iterations on all dimensions: 64 clusters -> 0.31 of the dense tile pairs visited (0.41 without
the sort inside the clusters, 0.46 without aligning tiles to clusters), 128 clusters 0.30, 256 clusters 0.33; the
result equals the brute-force kNN."""
import os
import sys

import numpy as np
from sklearn.cluster import KMeans
from sklearn.decomposition import PCA
from sklearn.neighbors import NearestNeighbors

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
counts = bench.make_counts(bench.WORKLOADS[args[0] if args else "c2"]).tocsr()
n_cells = counts.shape[0]
par = np.random.default_rng(0).choice(n_cells, size=(n_cells // 4, 2), replace=False)
aug = np.vstack([counts.toarray(), (counts[par[:, 0]] + counts[par[:, 1]]).toarray()]).astype(np.float32)
lib = aug.sum(1, keepdims=True)
aug = np.log(aug / lib * np.median(lib) + 0.1)
emb = PCA(30, random_state=0).fit_transform(aug).astype(np.float64)
n = emb.shape[0]
print("variance per component:", np.round(emb.var(0), 1))
dist, _ = NearestNeighbors(n_neighbors=10, algorithm="brute").fit(emb).kneighbors(emb)
tau = dist[:, -1] ** 2
print("10th-neighbour distance^2: median %.1f (cells %.1f, doublets %.1f), max %.1f"
      % (np.median(tau), np.median(tau[:n_cells]), np.median(tau[n_cells:]), tau.max()))


def tiles(labels, size):
    out = []
    for c in np.unique(labels):
        ids = np.nonzero(labels == c)[0]
        ids = ids[np.argsort(emb[ids, 7])]
        out += [ids[s:s + size] for s in range(0, len(ids), size)]
    return out


def kept(cand, query, m=8):
    lo = np.array([emb[t, :m].min(0) for t in cand])
    hi = np.array([emb[t, :m].max(0) for t in cand])
    qlo = np.array([emb[t, :m].min(0) for t in query])
    qhi = np.array([emb[t, :m].max(0) for t in query])
    qt = np.array([tau[t].max() for t in query])
    gap = np.maximum(0, np.maximum(lo[None] - qhi[:, None], qlo[:, None] - hi[None]))
    need = (gap ** 2).sum(-1) <= qt[:, None]
    pairs = (need * np.array([len(t) for t in query])[:, None] * np.array([len(t) for t in cand])[None]).sum() / (n * n)
    slots = need.sum() * len(query[0]) * len(cand[0]) / (n * n)
    return pairs, slots


for kc in (8, 16, 36, 64, 128):
    lab = KMeans(kc, n_init=2, random_state=0).fit(emb[:, :8]).labels_
    for qt in (128, 256):
        p, s = kept(tiles(lab, 128), tiles(lab, qt))
        print(f"k-means {kc:3d} clusters, query tile {qt}: point pairs kept {p:.2f}, tile slots incl. padding {s:.2f} of n^2")


def simulate(n_clusters, lloyd_iters=5, tile=128, qtile=256):
    lab = KMeans(n_clusters, n_init=1, max_iter=lloyd_iters, random_state=0).fit(emb).labels_
    qtiles = []
    for c in range(n_clusters):
        ids = np.nonzero(lab == c)[0]
        if len(ids) == 0:
            continue
        x = emb[ids] - emb[ids].mean(0)
        ids = ids[np.argsort(x @ np.linalg.svd(x, full_matrices=False)[2][0])]
        qtiles += [ids[s:s + qtile] for s in range(0, len(ids), qtile)]
    ctiles = [t[h:h + tile] for t in qtiles for h in range(0, len(t), tile)]
    lo = np.array([emb[t].min(0) for t in ctiles])
    hi = np.array([emb[t].max(0) for t in ctiles])
    nrm = (emb ** 2).sum(1)
    visited, exact = 0, True
    for qi, qt in enumerate(qtiles):
        q = emb[qt]
        gap = np.maximum(0, np.maximum(lo - q.max(0), q.min(0) - hi))
        lb = (gap ** 2).sum(-1)
        best = np.full((len(qt), 10), np.inf)
        for ti in np.argsort(lb, kind="stable"):
            if lb[ti] > best[:, -1].max() * (1 + 1e-6):
                break
            ct = ctiles[ti]
            d2 = nrm[qt][:, None] - 2 * q @ emb[ct].T + nrm[ct][None]
            best = np.sort(np.concatenate([best, d2], 1), 1)[:, :10]
            visited += 1
        if qi % 16 == 0:
            exact &= bool(np.allclose(np.sqrt(np.maximum(best, 0)), dist[qt], rtol=1e-5, atol=1e-4))
    dense = -(-n // qtile) * -(-n // tile)
    print(f"sim: {n_clusters} clusters -> {len(qtiles)} query tiles, {len(ctiles)} candidate tiles, visited "
          f"{visited / dense:.3f} of the dense tile pairs, equals brute force: {exact}", flush=True)


if "--sim" in sys.argv:
    for kc in (64, 128):
        simulate(kc)
