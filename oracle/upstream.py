"""Restatements of the third-party calls the reference makes into packages that are absent from
this image (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Each function names the reference call site (``/root/reference/doubletdetection/doubletdetection.py``)
and the upstream semantics it restates (upstream sources are not available
here, so these are semantic restatements -- the parts that bottom out in the *installed* sklearn
(PCA, brute-force kNN) call that code directly and are therefore exact).
"""

import numpy as np
import scipy.sparse as sp_sparse

from . import leiden_ref, louvain_ref


class AnnDataLite:
    """Minimal stand-in for ``anndata.AnnData`` as used at doubletdetection.py:300-301,343."""

    def __init__(self, X):
        self.X = X
        self.obs = {}
        self.obsm = {}
        self.obsp = {}
        self.uns = {}

    @property
    def shape(self):
        return self.X.shape

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]


def pp_scale(X, max_value=None):
    """``sc.pp.scale(adata, max_value=15)`` -- doubletdetection.py:302-303; SURVEY Appendix B4.

    Per-gene mean and variance accumulated in float64 (var = E[x^2]-E[x]^2, times n/(n-1)),
    std==0 -> 1, in-place centre and divide on the float32 array, clip to [-max, max]
    (scanpy >= 1.10 with zero_center=True).
    """
    X = np.asarray(X)
    n = X.shape[0]
    mean = np.mean(X, axis=0, dtype=np.float64)
    mean_sq = np.multiply(X, X).mean(axis=0, dtype=np.float64)  # squares in X's own dtype
    var = (mean_sq - mean**2) * (n / (n - 1))
    std = np.sqrt(var)
    std[std == 0] = 1
    X = X.copy()
    # in-place ops with float64 operands: computed in float64, rounded once into X's dtype
    np.subtract(X, mean, out=X, casting="same_kind")
    np.divide(X, std, out=X, casting="same_kind")
    if max_value is not None:
        np.clip(X, -max_value, max_value, out=X)
    return X, mean, std


def tl_pca(X, n_comps, random_state=0, svd_solver="auto"):
    """``sc.tl.pca(adata, n_comps, random_state, svd_solver)`` dense path --
    doubletdetection.py:308-314; SURVEY Appendix B5.  scanpy calls
    ``sklearn.decomposition.PCA(n_components, svd_solver, random_state).fit_transform(X)`` and
    casts to float32; sklearn is installed, so this is the real code.
    """
    from sklearn.decomposition import PCA

    if sp_sparse.issparse(X):
        pca = PCA(n_components=n_comps, svd_solver="arpack", random_state=random_state)
    else:
        pca = PCA(n_components=n_comps, svd_solver=svd_solver, random_state=random_state)
    emb = pca.fit_transform(X)
    return np.ascontiguousarray(emb, dtype=np.float32), pca


def knn_brute(rep, n_neighbors=10):
    """Exact kNN as scanpy does it below 8192 observations:
    ``KNeighborsTransformer(algorithm="brute", metric="euclidean")``; the result has the point
    itself in column 0 and ``n_neighbors - 1`` others.  Returns (indices int64, distances float32).
    """
    from sklearn.neighbors import KNeighborsTransformer

    n = rep.shape[0]
    k = min(n_neighbors, n)
    tr = KNeighborsTransformer(algorithm="brute", metric="euclidean", n_neighbors=k - 1, mode="distance")
    g = tr.fit_transform(rep)  # k explicit entries per row incl. self (distance 0)
    g.sort_indices()
    idx = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    indptr, indices, data = g.indptr, g.indices, g.data
    for i in range(n):
        cols = indices[indptr[i] : indptr[i + 1]]
        d = data[indptr[i] : indptr[i + 1]]
        # self first, then ascending distance (ties by index, as a stable sort gives)
        is_self = cols == i
        key = np.lexsort((cols, d, ~is_self))
        idx[i] = cols[key][:k]
        dist[i] = d[key][:k]
    return idx, dist.astype(np.float32)


def smooth_knn_dist(distances, k, n_iter=64):
    """umap ``smooth_knn_dist(distances, k, local_connectivity=1, bandwidth=1)`` restated;
    ``distances`` float32 with self in column 0.  The arithmetic is pinned the way numba types the upstream code,
    so that the product's C++ (``doubletdetection_b200/csrc/leiden.cpp:umap_weights``) reproduces it bit for bit:
    ``rho`` and the stored ``sigma`` are float32; ``d = dist - rho`` is a float32 subtraction; the bisection
    (``mid``, ``psum``, ``exp``) runs in float64 with libm's ``exp``; sums are sequential.  ``mean(distances)``
    (only used as the sigma floor of rows whose neighbours all coincide with the cell) is the float64 mean of
    the sequential float64 row sums."""
    import math

    SMOOTH_K_TOLERANCE = 1e-5
    MIN_K_DIST_SCALE = 1e-3
    distances = np.asarray(distances, dtype=np.float32)
    n, width = distances.shape
    target = math.log2(float(k))
    rho = np.zeros(n, dtype=np.float32)
    sigma = np.zeros(n, dtype=np.float32)
    row_sums = []
    for i in range(n):
        s = 0.0
        for x in distances[i].tolist():
            s += x
        row_sums.append(s)
    total = 0.0
    for s in row_sums:
        total += s
    mean_distances = total / float(n * width) if n * width else 0.0
    for i in range(n):
        ith = distances[i]
        for x in ith:
            if x > 0.0:
                rho[i] = x  # local_connectivity = 1: the nearest neighbour at a positive distance
                break
        d = [float(x) for x in (ith[1:] - rho[i])]  # float32 subtraction, then exact in float64
        lo, hi, mid = 0.0, math.inf, 1.0
        for _ in range(n_iter):
            psum = 0.0
            for x in d:
                psum += math.exp(-(x / mid)) if x > 0.0 else 1.0
            if abs(psum - target) < SMOOTH_K_TOLERANCE:
                break
            if psum > target:
                hi = mid
                mid = (lo + hi) / 2.0
            else:
                lo = mid
                if hi == math.inf:
                    mid *= 2.0
                else:
                    mid = (lo + hi) / 2.0
        sigma[i] = mid
        floor = MIN_K_DIST_SCALE * (row_sums[i] / float(width) if rho[i] > 0.0 else mean_distances)
        if float(sigma[i]) < floor:
            sigma[i] = floor
    return sigma, rho


def membership_strengths(knn_idx, knn_dist, sigma, rho):
    """umap ``compute_membership_strengths`` restated: float32 directed weights (n x k), 0 for the cell itself, 1 at or
    below rho, else float32(exp(-(float64(d - rho) / float64(sigma)))) with ``d - rho`` a float32 subtraction."""
    import math

    n, k = knn_idx.shape
    d = np.asarray(knn_dist, dtype=np.float32)
    vals = np.zeros((n, k), dtype=np.float32)
    for i in range(n):
        diff = d[i] - rho[i]
        s = float(sigma[i])
        for j in range(k):
            if knn_idx[i, j] == i:
                v = 0.0
            elif diff[j] <= 0.0 or s == 0.0:
                v = 1.0
            else:
                v = math.exp(-(float(diff[j]) / s))
            vals[i, j] = v
    return vals


def fuzzy_connectivities(knn_idx, knn_dist):
    """umap ``fuzzy_simplicial_set(set_op_mix_ratio=1, local_connectivity=1)`` restated:
    membership strengths, then W + W^T - W o W^T in float32, zeros eliminated."""
    knn_idx = np.asarray(knn_idx)
    n, k = knn_idx.shape
    d = np.asarray(knn_dist, dtype=np.float32)
    sigma, rho = smooth_knn_dist(d, float(k))
    vals = membership_strengths(knn_idx, d, sigma, rho).ravel()
    rows = np.repeat(np.arange(n), k)
    cols = knn_idx.ravel()
    W = sp_sparse.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    W.eliminate_zeros()
    T = W.T.tocsr()
    P = W.multiply(T)
    C = (W + T - P).tocsr()
    C.eliminate_zeros()
    C.sort_indices()
    return C


def knn_pattern_graph(knn_idx):
    """Symmetric 0/1 pattern of the connectivities graph: i~j iff j in kNN(i)\\{i} or i in
    kNN(j)\\{j}.  This is all ``sc.tl.louvain`` sees because it ignores weights by default
    (``use_weights=False``)."""
    n, k = knn_idx.shape
    rows = np.repeat(np.arange(n, dtype=np.int64), k)
    cols = np.asarray(knn_idx, dtype=np.int64).ravel()
    keep = rows != cols
    rows, cols = rows[keep], cols[keep]
    data = np.ones(rows.size, dtype=np.float32)
    W = sp_sparse.coo_matrix((data, (rows, cols)), shape=(n, n)).tocsr()
    S = (W + W.T).tocsr()
    S.data[:] = 1.0
    S.sort_indices()
    return S


def pp_neighbors(adata, random_state=0, method="umap", n_neighbors=10, with_weights=False):
    """``sc.pp.neighbors(adata, random_state, method="umap", n_neighbors=10)`` --
    doubletdetection.py:331-336.  Representation: ``X_pca`` when n_vars > 50, else X.
    Exact brute-force kNN is used at every size (upstream switches to approximate NN-descent at
    >= 8192 observations; the oracle defines the reference there through the exact kNN).
    """
    if adata.n_vars > 50 and "X_pca" in adata.obsm:
        rep = adata.obsm["X_pca"]
    else:
        rep = adata.X
    rep = np.asarray(rep)
    idx, dist = knn_brute(rep, n_neighbors)
    adata.uns["knn_indices"] = idx
    adata.uns["knn_distances"] = dist
    if with_weights:
        adata.obsp["connectivities"] = fuzzy_connectivities(idx, dist)
    else:
        adata.obsp["connectivities"] = knn_pattern_graph(idx)
    return adata


def tl_louvain(adata, key_added="clusters", random_state=0, directed=False, resolution=4, louvain_fn=None, **_):
    """``sc.tl.louvain(adata, key_added="clusters", random_state, directed=False, resolution=4)``
    -- doubletdetection.py:337-342; SURVEY Appendix B2.  Unweighted graph from
    ``connectivities.nonzero()``; partition by the in-repo deterministic Louvain
    (``louvain_ref``; PARITY UNPINNED, the louvain package is absent); labels by decreasing size,
    stored as strings like scanpy's categorical."""
    C = adata.obsp["connectivities"].tocsr()
    C.sort_indices()
    fn = louvain_fn or louvain_ref.louvain
    # the kNN-pipeline flavour of the specification: synchronous coloured first level (GPU-friendly), then
    # sequential levels (oracle/louvain_ref.py)
    labels = fn(C.indptr, C.indices, None, resolution=float(resolution), seed=int(random_state), level0="parallel")
    adata.obs[key_added] = np.asarray([str(int(x)) for x in labels])
    return adata


def tl_leiden(adata, key_added="clusters", random_state=0, directed=False, resolution=4, leiden_fn=None, **_):
    """``sc.tl.leiden(adata, key_added="clusters", random_state, directed=False, resolution=4)`` --
    doubletdetection.py:340-342; SURVEY Appendix B2.  Unlike ``tl.louvain`` it uses the connectivities' weights
    (``use_weights=True``) and runs ``leidenalg`` until no iteration improves (``n_iterations=-1``).  The partition
    comes from the in-repo deterministic Leiden (``leiden_ref``; PARITY UNPINNED, leidenalg is absent)."""
    C = adata.obsp["connectivities"].tocsr()
    C.sort_indices()
    fn = leiden_fn or leiden_ref.leiden
    labels = fn(C.indptr, C.indices, C.data.astype(np.float64), resolution=float(resolution), seed=int(random_state))
    adata.obs[key_added] = np.asarray([str(int(x)) for x in labels])
    return adata


def jaccard_graph(nbr_idx, prune=True):
    """PhenoGraph's graph (phenograph/core.py: jaccard_kernel + neighbor_graph, cluster(): prune) from the k nearest
    neighbours WITHOUT self (n x k): directed weight w_ij = s / (2k - s) with s = |N(i) & N(j)| for j in N(i);
    prune=True keeps mutual edges with the product of the two directed weights (``graph.multiply(graph.T)``),
    prune=False averages (``(graph + graph.T) / 2``).  Zero weights are dropped.  float64 CSR, sorted rows.
    **[upstream, absent -- restated from SURVEY Appendix B3]**"""
    nbr_idx = np.asarray(nbr_idx, dtype=np.int64)
    n, k = nbr_idx.shape
    sets = [set(row.tolist()) for row in nbr_idx]
    rows = np.repeat(np.arange(n, dtype=np.int64), k)
    cols = nbr_idx.ravel()
    w = np.empty(n * k, dtype=np.float64)
    for i in range(n):
        si = sets[i]
        for c in range(k):
            s = len(si & sets[nbr_idx[i, c]])
            w[i * k + c] = s / (2.0 * k - s)
    G = sp_sparse.coo_matrix((w, (rows, cols)), shape=(n, n)).tocsr()
    S = G.multiply(G.T) if prune else (G + G.T) / 2.0
    S = S.tocsr()
    S.eliminate_zeros()
    S.sort_indices()
    return S


def phenograph_cluster(X_pca, k=30, prune=True, min_cluster_size=10, seed=0, louvain_fn=None, level0="parallel"):
    """``phenograph.cluster(X_pca, n_jobs=..., prune=...)[0]`` -- doubletdetection.py:320 (defaults k=30,
    jaccard=True, min_cluster_size=10): exact k+1 nearest neighbours, self dropped; Jaccard graph; Louvain on the
    weighted graph (standard modularity = resolution 1); communities numbered by decreasing size, those NOT larger
    than ``min_cluster_size`` relabelled -1 (``sort_by_size``: ``sizes[c] > min_size`` keeps the label).  Upstream runs its bundled Louvain binaries repeatedly with
    time-based seeds and keeps the best modularity, so its labels are not reproducible even against itself;
    the oracle fixes ONE seeded run of the in-repo Louvain specification.  PARITY UNPINNED for this stage."""
    idx, _ = knn_brute(np.asarray(X_pca), k + 1)
    G = jaccard_graph(idx[:, 1:], prune=prune)
    fn = louvain_fn or louvain_ref.louvain
    # level0="parallel" (the specification the product and the goldens follow): the first level by synchronous coloured rounds
    # on fixed-point weights, which the fit loop runs on the device (louvain_gpu_w.cu); "sequential": the classic sweep
    kw = {} if level0 == "sequential" else {"level0": level0}
    labels = np.asarray(fn(G.indptr, G.indices, G.data, resolution=1.0, seed=int(seed), **kw), dtype=np.int64).copy()
    sizes = np.bincount(labels, minlength=labels.max() + 1 if labels.size else 0)
    # phenograph.core.sort_by_size keeps a community only if its size is > min_size: exactly min_size cells -> -1
    labels[sizes[labels] <= min_cluster_size] = -1
    return labels, G
