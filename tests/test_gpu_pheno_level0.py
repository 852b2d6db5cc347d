"""The weighted first Louvain level on the device (louvain_gpu_w.cu: synchronous coloured rounds on fixed-point weights,
PhenoGraph's default path since round 2) against its specification ``oracle/louvain_ref.py:level0_parallel(..., weights)``
-- label for label -- on umap-weighted and PhenoGraph (Jaccard) graphs, and the PhenoGraph branch of the fit loop as a
CHAIN: the oracle's kNN + Jaccard graph + Louvain + scoring run on the GPU's own embedding must reproduce the classifier's
communities and scores of every iteration exactly (doubletdetection.py:318-327, 344-383).  Needs a B200 (`-m gpu`)."""

import warnings

import numpy as np
import pytest

from oracle import datasets, louvain_c, louvain_ref, pca_f64, reference_path, upstream

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k,kind,gamma,seed", [(300, 6, "umap", 1.0, 3), (2000, 10, "umap", 4.0, 0), (3000, 31, "jaccard", 1.0, 1),
                                                 (20000, 31, "jaccard", 1.0, 2), (60000, 31, "jaccard", 1.0, 0)])
def test_weighted_level0_device_equals_specification(handle, n, k, kind, gamma, seed):
    from doubletdetection_b200 import _capi

    rs = np.random.default_rng(n)
    pts = (rs.normal(size=(n, 8)) + rs.integers(0, 6, size=(n, 1)) * 2.5).astype(np.float32)
    handle.upload_embedding(pts)
    idx, dist = handle.knn(k)  # exact kNN from the device (the oracle's brute force is slow at 60k)
    if kind == "jaccard":
        G = handle.jaccard_graph(k, prune=True)  # device-built (bit-exact vs the oracle: test_jaccard_graph_on_device_matches_oracle)
    else:
        G = _capi.umap_connectivities(idx, dist).astype(np.float64)
    w = np.asarray(G.data, dtype=np.float64)
    want = louvain_ref.level0_parallel(G.indptr, G.indices, gamma, seed, w)
    got, rounds = handle.louvain_level0_weighted(G.indptr, G.indices, w, gamma, seed)
    assert 1 <= rounds <= 32
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("shape,n_iters,kwargs", [((1500, 300), 3, {}), ((10000, 3000), 3, {}),
                                                   ((4000, 1200), 2, {"clustering_kwargs": {"prune": False}})])
def test_phenograph_fit_loop_chain_vs_oracle(handle, shape, n_iters, kwargs):
    from doubletdetection_b200 import BoostClassifier

    counts = datasets.structured_counts(*shape, seed=1234)
    n, g_ = counts.shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_iters=n_iters, random_state=0, n_jobs=4, **kwargs).fit(counts)
    assert clf.clustering_algorithm == "phenograph"
    parents = np.asarray(clf._parents_array)
    omega = pca_f64.omega(g_, 30, 0).astype(np.float32)
    n_iter = pca_f64.auto_n_iter(n + parents.shape[1], g_, 30)
    handle.upload_counts(counts)
    ckw = dict(kwargs.get("clustering_kwargs") or {})
    for i in range(n_iters):
        handle.create_doublets(parents[i])
        handle.normalise_log(handle.median_lib_size(), 0.1)
        emb, _ = handle.pca(30, omega, n_iter)
        labels, _ = upstream.phenograph_cluster(emb, seed=0, louvain_fn=louvain_c.louvain, **ckw)
        np.testing.assert_array_equal(clf.communities_[i], labels[:n], err_msg=f"communities of iteration {i}")
        np.testing.assert_array_equal(clf.synth_communities_[i], labels[n:])
        s, lp, _, _ = reference_path.score_communities(labels, n)
        np.testing.assert_array_equal(clf.all_scores_[i], s)
        np.testing.assert_allclose(clf.all_log_p_values_[i], lp, rtol=1e-9, atol=1e-12, equal_nan=True)
