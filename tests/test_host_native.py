"""CPU-side checks of libdd_b200.so: it loads, exports every symbol the header declares, and its host
entry points (Louvain, scoring, hypergeom) agree with the oracle.  No CUDA compute is called."""

import os
import re
import warnings

import numpy as np
import pytest
from scipy.stats import hypergeom

from conftest import ROOT, golden_case, load_golden
from oracle import louvain_c, louvain_ref, reference_path, upstream


def test_library_exports_every_header_symbol(native):
    header = open(os.path.join(ROOT, "include", "dd_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = native.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dd_b200.h but not exported"
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert lib.dd_abi_version() == native.ABI_VERSION


def test_no_cpu_fallback_without_device(native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        native.Handle(0)
    from doubletdetection_b200 import BoostClassifier

    clf = BoostClassifier(n_iters=2, clustering_algorithm="louvain")
    with pytest.raises(RuntimeError):
        clf.fit(np.random.default_rng(0).poisson(1.0, (600, 120)))


@pytest.mark.parametrize("args", [(10, 625, 125, 40), (0, 625, 125, 3), (5, 625, 125, 5), (3, 625, 125, 0),
                                  (0, 12500, 2500, 1), (40, 12500, 2500, 41), (700, 125000, 25000, 1500),
                                  (2, 10, 9, 8), (7, 10, 9, 8)])
def test_hypergeom_logsf_matches_scipy(native, args):
    want = hypergeom.logsf(*args)
    got = native.hypergeom_logsf(*args)
    if np.isinf(want):
        assert got == want
    else:
        assert got == pytest.approx(want, rel=1e-9, abs=1e-12)


def test_hypergeom_logsf_grid(native):
    rs = np.random.default_rng(1)
    for _ in range(300):
        M = int(rs.integers(2, 3000))
        n = int(rs.integers(0, M + 1))
        N = int(rs.integers(0, M + 1))
        k = int(rs.integers(0, N + 1))
        want = hypergeom.logsf(k, M, n, N)
        got = native.hypergeom_logsf(k, M, n, N)
        if np.isinf(want) or np.isnan(want):
            assert (np.isnan(got) and np.isnan(want)) or got == want, (k, M, n, N, got, want)
        else:
            assert got == pytest.approx(want, rel=1e-8, abs=1e-10), (k, M, n, N)


def test_score_matches_oracle(native):
    rs = np.random.default_rng(2)
    n_cells, n_synth = 400, 100
    labels = rs.integers(0, 12, n_cells + n_synth)
    labels[n_cells:][labels[n_cells:] == 3] = 4  # a community without synthetics
    labels[:n_cells][labels[:n_cells] == 7] = 8  # a community without original cells
    s, lp = native.score(labels, n_cells)
    os_, olp, _, _ = reference_path.score_communities(labels, n_cells)
    np.testing.assert_allclose(s, os_, rtol=0, atol=0)
    np.testing.assert_allclose(lp, olp, rtol=1e-9, atol=1e-12)
    # -1 labels become NaN (phenograph's small clusters, doubletdetection.py:379-381)
    labels[:10] = -1
    s, lp = native.score(labels, n_cells)
    os_, olp, _, _ = reference_path.score_communities(labels, n_cells)
    assert np.isnan(s[:10]).all() and np.isnan(lp[:10]).all()
    np.testing.assert_allclose(s, os_, equal_nan=True)
    np.testing.assert_allclose(lp, olp, rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.parametrize("n,k,gamma,seed", [(50, 4, 1.0, 0), (300, 6, 4.0, 3), (1000, 10, 4.0, 0), (1000, 10, 0.5, 9)])
def test_louvain_knn_matches_python_spec(native, n, k, gamma, seed):
    rs = np.random.default_rng(n + k)
    pts = (rs.normal(size=(n, 6)) + rs.integers(0, 5, size=(n, 1)) * 2.5).astype(np.float32)
    idx, _ = upstream.knn_brute(pts, k)
    S = upstream.knn_pattern_graph(idx)
    # kNN entry: parallel (GPU-style) first level + sequential upper levels
    want = louvain_ref.louvain(S.indptr, S.indices, None, resolution=gamma, seed=seed, level0="parallel")
    got = native.louvain_knn(idx.astype(np.int32), resolution=gamma, seed=seed)
    np.testing.assert_array_equal(got, want)
    # explicit-graph entry: fully sequential
    want_seq = louvain_ref.louvain(S.indptr, S.indices, None, resolution=gamma, seed=seed)
    got_csr = native.louvain_csr(S.indptr, S.indices, None, resolution=gamma, seed=seed)
    np.testing.assert_array_equal(got_csr, want_seq)


def test_louvain_weighted_matches_spec(native):
    rs = np.random.default_rng(11)
    pts = (rs.normal(size=(400, 4)) + rs.integers(0, 3, size=(400, 1)) * 3).astype(np.float32)
    idx, dist = upstream.knn_brute(pts, 8)
    C = upstream.fuzzy_connectivities(idx, dist)
    want = louvain_ref.louvain(C.indptr, C.indices, C.data, resolution=1.5, seed=4)
    got = native.louvain_csr(C.indptr, C.indices, C.data.astype(np.float64), resolution=1.5, seed=4)
    np.testing.assert_array_equal(got, want)


def test_louvain_large_matches_c_oracle(native):
    rs = np.random.default_rng(21)
    n = 20000
    pts = (rs.normal(size=(n, 8)) + rs.integers(0, 8, size=(n, 1)) * 2.0).astype(np.float32)
    idx, _ = upstream.knn_brute(pts, 10)
    S = upstream.knn_pattern_graph(idx)
    want = louvain_c.louvain(S.indptr, S.indices, None, resolution=4.0, seed=0, level0="parallel")
    got = native.louvain_knn(idx.astype(np.int32), resolution=4.0, seed=0)
    np.testing.assert_array_equal(got, want)
    want = louvain_c.louvain(S.indptr, S.indices, None, resolution=4.0, seed=0)
    got = native.louvain_csr(S.indptr, S.indices, None, resolution=4.0, seed=0)
    np.testing.assert_array_equal(got, want)


def test_louvain_degenerate_graphs(native):
    # isolated nodes only: every node its own community, labelled by index order
    idx = np.arange(7, dtype=np.int32)[:, None]
    np.testing.assert_array_equal(native.louvain_knn(idx, 4.0, 0), np.arange(7))
    with pytest.raises(ValueError):
        native.louvain_knn(np.array([[0, 9]], dtype=np.int32), 4.0, 0)


def test_louvain_and_score_on_golden_knn(native):
    """Feed the reference run's own kNN graph: labels, scores and log p must equal the golden fit."""
    g = load_golden("structured_1500x300")
    n_cells = g["communities"].shape[1]
    labels = native.louvain_knn(g["knn_indices0"], resolution=4.0, seed=0)
    np.testing.assert_array_equal(labels[:n_cells], g["communities"][0])
    np.testing.assert_array_equal(labels[n_cells:], g["synth_communities"][0])
    s, lp = native.score(labels, n_cells)
    np.testing.assert_array_equal(s, g["all_scores"][0])
    np.testing.assert_allclose(lp, g["all_log_p_values"][0], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("prune", [True, False])
def test_phenograph_host_twin_matches_oracle(native, prune):
    """dd_phenograph_knn (the host twin of the device graph + the weighted Louvain the fit loop runs) against the
    oracle's restatement of phenograph.cluster (oracle/upstream.py) on the same exact kNN lists."""
    from oracle import louvain_c, upstream

    rs = np.random.default_rng(3)
    for case, (sizes, dim, spread) in enumerate([((150, 150, 150, 150), 12, 4.0), ((300, 40, 9, 200), 8, 6.0)]):
        x = np.vstack([rs.normal(spread * c, 1.0, (m, dim)) for c, m in enumerate(sizes)]).astype(np.float32)
        idx, _ = upstream.knn_brute(x, 31)
        want, graph = upstream.phenograph_cluster(x, k=30, prune=prune, min_cluster_size=10, seed=case,
                                                  louvain_fn=louvain_c.louvain)
        got = native.phenograph_knn(idx, prune=prune, min_cluster_size=10, seed=case)
        np.testing.assert_array_equal(got, want)
        assert graph.nnz > 0 and (abs(graph - graph.T)).nnz == 0  # the Jaccard graph is symmetric
    # a larger min_cluster_size sends more cells to -1 (NaN scores downstream, doubletdetection.py:379-381)
    many = native.phenograph_knn(idx, prune=prune, min_cluster_size=100, seed=1)
    assert (many == -1).sum() >= (got == -1).sum()
