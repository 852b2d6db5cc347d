"""CPU-side checks of libdd_b200.so: it loads, exports every symbol the header declares, and its host
entry points (Louvain, scoring, hypergeom) agree with the oracle.  No CUDA compute is called."""

import os
import re
import warnings

import numpy as np
import pytest
from scipy.stats import hypergeom

from conftest import ROOT, golden_case, load_golden
from oracle import leiden_ref, louvain_c, louvain_ref, reference_path, upstream


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


def test_phenograph_min_cluster_size_boundary(native):
    """phenograph.core.sort_by_size keeps a community only if its size is > min_cluster_size: a community of EXACTLY
    min_cluster_size cells is relabelled -1 (NaN score / log p downstream, doubletdetection.py:379-381), one of
    min_cluster_size + 1 keeps its label.  Both the product's host code and the oracle restatement."""
    from oracle import louvain_c, upstream

    rs = np.random.default_rng(5)
    sizes = (60, 11, 10)
    x = np.vstack([rs.normal(40.0 * c, 1.0, (m, 6)) for c, m in enumerate(sizes)]).astype(np.float32)
    idx, _ = upstream.knn_brute(x, 10)
    for prune in (True, False):
        want, _ = upstream.phenograph_cluster(x, k=9, prune=prune, min_cluster_size=10, seed=0, louvain_fn=louvain_c.louvain)
        got = native.phenograph_knn(idx, prune=prune, min_cluster_size=10, seed=0)
        np.testing.assert_array_equal(got, want)
        assert (got[71:] == -1).all(), "a community of exactly min_cluster_size cells must be labelled -1"
        assert len(set(got[60:71].tolist())) == 1 and got[60] >= 0, "a community of min_cluster_size + 1 cells keeps its label"
        # one more cell allowed -> the 10-cell community is kept
        loose = native.phenograph_knn(idx, prune=prune, min_cluster_size=9, seed=0)
        assert (loose[71:] >= 0).all()


# ---- clustering_algorithm="leiden": umap connectivities + Leiden on the host workers (leiden.cpp)
def _blobs(n, dim, seed, spread=2.5, n_types=5):
    rs = np.random.default_rng(seed)
    return (rs.normal(size=(n, dim)) + rs.integers(0, n_types, size=(n, 1)) * spread).astype(np.float32)


@pytest.mark.parametrize("n,k", [(60, 4), (400, 10), (1200, 10), (700, 15)])
def test_umap_connectivities_match_oracle_bit_for_bit(native, n, k):
    idx, dist = upstream.knn_brute(_blobs(n, 6, n + k), k)
    want = upstream.fuzzy_connectivities(idx, dist)
    got = native.umap_connectivities(idx, dist)
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)  # float32, same bits
    assert got.data.dtype == np.float32 and (got.data > 0).all() and (got.data <= 1).all()
    assert (abs(got - got.T)).nnz == 0  # the fuzzy union is symmetric
    # every directed kNN edge survives the union (its membership strength is > 0 unless it underflows)
    pattern = upstream.knn_pattern_graph(idx)
    assert got.nnz <= pattern.nnz and got.nnz >= 0.99 * pattern.nnz


def test_umap_connectivities_duplicates_and_far_neighbours(native):
    """rho = 0 rows (all neighbours coincide with the cell: sigma floor from the global mean), exact duplicates
    (distance 0 -> strength 1) and a neighbour so far away that its strength underflows to 0 and is dropped."""
    x = _blobs(300, 4, 5)
    x[10:16] = x[10]  # six coincident cells: with k = 4 all their neighbours are at distance 0
    x[200] += 1.0e4  # an outlier: its neighbours are all ~1e4 away, and nobody lists it
    idx, dist = upstream.knn_brute(x, 4)
    want = upstream.fuzzy_connectivities(idx, dist)
    got = native.umap_connectivities(idx, dist)
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)
    labels = native.leiden_knn(idx, dist, resolution=1.0, seed=0)
    assert len(set(labels[10:16].tolist())) == 1  # coincident cells stay together


@pytest.mark.parametrize("n,k,gamma,seed", [(50, 4, 1.0, 0), (300, 6, 4.0, 3), (1000, 10, 4.0, 0), (1000, 10, 0.5, 9),
                                            (1500, 10, 4.0, 21)])
def test_leiden_matches_python_spec(native, n, k, gamma, seed):
    idx, dist = upstream.knn_brute(_blobs(n, 6, n + k), k)
    C = upstream.fuzzy_connectivities(idx, dist)
    w = C.data.astype(np.float64)
    want = leiden_ref.leiden(C.indptr, C.indices, w, resolution=gamma, seed=seed)
    np.testing.assert_array_equal(native.leiden_csr(C.indptr, C.indices, w, resolution=gamma, seed=seed), want)
    np.testing.assert_array_equal(native.leiden_knn(idx, dist, resolution=gamma, seed=seed), want)
    # unweighted flavour of the same entry
    S = upstream.knn_pattern_graph(idx)
    want_u = leiden_ref.leiden(S.indptr, S.indices, None, resolution=gamma, seed=seed)
    np.testing.assert_array_equal(native.leiden_csr(S.indptr, S.indices, None, resolution=gamma, seed=seed), want_u)
    # Leiden's guarantees: not worse than the Louvain specification on the same graph, communities connected
    lou = louvain_ref.louvain(C.indptr, C.indices, w, resolution=gamma, seed=seed)
    q_le = leiden_ref.quality(C.indptr, C.indices, w, want, gamma)
    q_lo = leiden_ref.quality(C.indptr, C.indices, w, lou, gamma)
    assert q_le >= q_lo - 1e-3
    from scipy.sparse.csgraph import connected_components

    for c in range(int(want.max()) + 1):
        members = np.nonzero(want == c)[0]
        ncomp, _ = connected_components(C[members][:, members], directed=False)
        assert ncomp == 1, f"community {c} is disconnected"


@pytest.mark.parametrize("n,k,seed", [(400, 10, 0), (1500, 10, 21)])
def test_leiden_from_device_layout_graph(native, n, k, seed):
    """dd_fit_iterations hands its Leiden workers the umap graph as the DEVICE builds it: the symmetric pattern with int32
    offsets, rows in arbitrary order, float64 weights and 0 for "no edge".  The host half (dd_leiden_device_graph) must put
    that into the canonical form and partition it exactly like dd_leiden_knn."""
    idx, dist = upstream.knn_brute(_blobs(n, 6, n + k), k)
    C = upstream.fuzzy_connectivities(idx, dist)
    rs = np.random.default_rng(seed)
    off, adj, w = [0], [], []
    for i in range(n):
        cols = C.indices[C.indptr[i]:C.indptr[i + 1]].tolist()
        vals = C.data[C.indptr[i]:C.indptr[i + 1]].astype(np.float64).tolist()
        for _ in range(int(rs.integers(0, 3))):  # pattern entries whose union underflowed: weight 0
            j = int(rs.integers(0, n))
            if j != i and j not in cols:
                cols.append(j)
                vals.append(0.0)
        order = rs.permutation(len(cols))
        adj += [cols[t] for t in order]
        w += [vals[t] for t in order]
        off.append(len(adj))
    want = native.leiden_knn(idx, dist, resolution=4.0, seed=seed)
    got = native.leiden_device_graph(off, adj, w, resolution=4.0, seed=seed)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        native.leiden_device_graph([0, 1, 2], [1, 5], [1.0, 1.0])  # neighbour out of range
    with pytest.raises(ValueError):
        native.leiden_device_graph([0, 2, 1], [1, 0], [1.0, 1.0])  # offsets not monotone


def test_leiden_degenerate_inputs(native):
    # no edges: every node its own community
    np.testing.assert_array_equal(native.leiden_csr(np.zeros(6, dtype=np.int64), np.zeros(0, dtype=np.int64)), np.arange(5))
    with pytest.raises(ValueError):
        native.leiden_knn(np.array([[0, 7], [1, 0]], dtype=np.int32), np.zeros((2, 2), dtype=np.float32))
    with pytest.raises(ValueError):
        native.leiden_knn(np.zeros((3, 2), dtype=np.int32), np.zeros((3, 3), dtype=np.float32))


def test_oracle_classifier_leiden_path_runs_and_matches_native_twin(native):
    """The oracle's leiden path (reference lines + sklearn PCA / kNN + restated umap weights + Leiden spec) against
    the native host twin fed the oracle's own kNN lists: communities, scores and log p identical."""
    counts, _, _ = golden_case("c1_louvain")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ora = reference_path.OracleClassifier(n_iters=2, random_state=0, clustering_algorithm="leiden", keep_stages=True)
        ora.fit(counts)
    n_cells = counts.shape[0]
    for i, st in enumerate(ora.stages):
        labels = native.leiden_knn(st["knn_indices"], st["knn_distances"], resolution=4.0, seed=0)
        np.testing.assert_array_equal(labels, st["fullcommunities"])
        s, lp = native.score(labels, n_cells)
        np.testing.assert_array_equal(s, ora.all_scores_[i])
        np.testing.assert_allclose(lp, ora.all_log_p_values_[i], rtol=1e-9, atol=1e-12)


def test_louvain_quality_against_an_independent_implementation(native):
    """The louvain / leidenalg / igraph packages are absent, so the clustering stage has no bit-level pin.  networkx's
    Louvain (an independent third-party implementation of the same RB-configuration objective) is the closest check
    available: at resolution 1 on well separated data both find the same partition; at the reference's resolution 4,
    where every cell type is split arbitrarily, the objective values must agree (ours may only be negligibly worse)."""
    nx = pytest.importorskip("networkx")
    from sklearn.metrics import adjusted_rand_score

    rs = np.random.default_rng(5)
    pts = (rs.normal(size=(3000, 8)) + rs.integers(0, 6, size=(3000, 1)) * 2.5).astype(np.float32)
    idx, dist = upstream.knn_brute(pts, 10)
    S = upstream.knn_pattern_graph(idx)
    G = nx.from_scipy_sparse_array(S)
    for gamma in (1.0, 4.0):
        theirs = np.empty(3000, dtype=np.int64)
        for c, members in enumerate(nx.community.louvain_communities(G, resolution=gamma, seed=0)):
            theirs[list(members)] = c
        ours = native.louvain_knn(idx.astype(np.int32), gamma, 0)
        q_theirs = leiden_ref.quality(S.indptr, S.indices, None, theirs, gamma)
        q_ours = leiden_ref.quality(S.indptr, S.indices, None, ours, gamma)
        assert q_ours >= q_theirs - 0.01, (gamma, q_ours, q_theirs)
        if gamma == 1.0:
            assert adjusted_rand_score(theirs, ours) > 0.98
    # the weighted objective (PhenoGraph's Jaccard graph at resolution 1, Leiden's umap graph at resolution 4)
    C = upstream.fuzzy_connectivities(idx, dist)
    Gw = nx.from_scipy_sparse_array(C.astype(np.float64), edge_attribute="weight")
    w = C.data.astype(np.float64)
    for gamma, fn in ((1.0, native.louvain_csr), (4.0, native.leiden_csr)):
        theirs = np.empty(3000, dtype=np.int64)
        for c, members in enumerate(nx.community.louvain_communities(Gw, weight="weight", resolution=gamma, seed=0)):
            theirs[list(members)] = c
        ours = fn(C.indptr, C.indices, w, resolution=gamma, seed=0)
        q_theirs = leiden_ref.quality(C.indptr, C.indices, w, theirs, gamma)
        q_ours = leiden_ref.quality(C.indptr, C.indices, w, ours, gamma)
        assert q_ours >= q_theirs - 0.01, (gamma, q_ours, q_theirs)


@pytest.mark.parametrize("n,k,gamma,seed", [(300, 6, 1.0, 3), (1000, 10, 1.0, 0), (2000, 12, 4.0, 7), (2500, 31, 1.0, 1)])
def test_weighted_parallel_first_level_matches_python_spec(native, n, k, gamma, seed):
    """dd_louvain_csr_level0: synchronous coloured first level on WEIGHTED graphs with fixed-point (2^-32) weights -- the
    specification of the device level PhenoGraph / Leiden will get.  Host twin vs the numpy
    specification, label for label; the unweighted flavour of the same entry equals the kNN pipeline's partition."""
    idx, dist = upstream.knn_brute(_blobs(n, 6, n), k)
    G = upstream.jaccard_graph(idx[:, 1:], prune=True) if k == 31 else upstream.fuzzy_connectivities(idx, dist)
    w = G.data.astype(np.float64)
    want = louvain_ref.louvain(G.indptr, G.indices, w, resolution=gamma, seed=seed, level0="parallel")
    got = native.louvain_csr(G.indptr, G.indices, w, gamma, seed, level0="parallel")
    np.testing.assert_array_equal(got, want)
    S = upstream.knn_pattern_graph(idx)
    np.testing.assert_array_equal(native.louvain_csr(S.indptr, S.indices, None, gamma, seed, level0="parallel"),
                                  native.louvain_knn(idx.astype(np.int32), gamma, seed))
    # same objective value as the fully sequential sweep it is meant to replace (to within the usual Louvain scatter)
    seq = native.louvain_csr(G.indptr, G.indices, w, gamma, seed)
    q_par = leiden_ref.quality(G.indptr, G.indices, w, got, gamma)
    q_seq = leiden_ref.quality(G.indptr, G.indices, w, seq, gamma)
    assert q_par >= q_seq - 0.01


def test_fixed_point_level_is_order_independent(native):
    """The point of the fixed-point weights: relabelling the nodes (which permutes every summation order in the level)
    changes nothing but the labels' names when the colour classes are permuted along."""
    idx, dist = upstream.knn_brute(_blobs(800, 5, 17), 10)
    G = upstream.fuzzy_connectivities(idx, dist).astype(np.float64)
    comm = louvain_ref.level0_parallel(G.indptr, G.indices, 1.0, 0, G.data)
    # shuffle the ORDER OF THE ENTRIES inside every row: same graph, different summation order for w(i, c), k_i, tot
    rs = np.random.default_rng(0)
    indptr, indices, data = G.indptr, G.indices.copy(), G.data.copy()
    for i in range(G.shape[0]):
        p = rs.permutation(indptr[i + 1] - indptr[i]) + indptr[i]
        indices[indptr[i]:indptr[i + 1]], data[indptr[i]:indptr[i + 1]] = indices[p], data[p]
    np.testing.assert_array_equal(louvain_ref.level0_parallel(indptr, indices, 1.0, 0, data), comm)


def test_leiden_medium_graph_matches_python_spec(native):
    """A graph large enough for several Leiden iterations with long tails (6000 cells, 7 types, resolution 4)."""
    rs = np.random.default_rng(1)
    pts = (rs.normal(size=(6000, 8)) + rs.integers(0, 7, size=(6000, 1)) * 2.2).astype(np.float32)
    idx, dist = upstream.knn_brute(pts, 10)
    C = native.umap_connectivities(idx, dist)  # bit-identical to the oracle's graph (tested above), and fast
    w = C.data.astype(np.float64)
    want = leiden_ref.leiden(C.indptr, C.indices, w, resolution=4.0, seed=2)
    np.testing.assert_array_equal(native.leiden_csr(C.indptr, C.indices, w, 4.0, 2), want)
    np.testing.assert_array_equal(native.leiden_knn(idx, dist, 4.0, 2), want)


def test_phenograph_with_parallel_first_level_is_equivalent(native):
    """PhenoGraph's clustering with the first Louvain level by synchronous rounds (the plan for the device, DESIGN.md
    section 10) against today's fully sequential sweep on the same Jaccard graph: oracle == host twin label for label, and
    the two flavours agree on everything but the arbitrary small clusters."""
    from sklearn.metrics import adjusted_rand_score

    rs = np.random.default_rng(9)
    x = np.vstack([rs.normal(4.0 * c, 1.0, (m, 10)) for c, m in enumerate((700, 500, 300, 40, 8))]).astype(np.float32)
    want, G = upstream.phenograph_cluster(x, k=30, prune=True, min_cluster_size=10, seed=0, level0="parallel")
    w = G.data.astype(np.float64)
    got = native.louvain_csr(G.indptr, G.indices, w, 1.0, 0, level0="parallel").astype(np.int64)
    sizes = np.bincount(got)
    got[sizes[got] < 10] = -1
    np.testing.assert_array_equal(got, want)
    seq, _ = upstream.phenograph_cluster(x, k=30, prune=True, min_cluster_size=10, seed=0, louvain_fn=louvain_c.louvain)
    both = (seq >= 0) & (want >= 0)
    assert both.mean() > 0.95 and adjusted_rand_score(seq[both], want[both]) > 0.95
    assert abs(int((seq < 0).sum()) - int((want < 0).sum())) <= 0.02 * x.shape[0]
