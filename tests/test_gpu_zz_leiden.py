"""clustering_algorithm="leiden" on the B200 path (doubletdetection.py:331-342): umap's connectivities of every
iteration's exact kNN lists + distances are built on the GPU (smooth_knn_dist bisection, membership strengths, fuzzy
union) and the native host workers run the in-repo Leiden on them.  Every test needs a B200 (`-m gpu`); nothing here
reads /root/reference.  The host side is covered bit for bit by tests/test_host_native.py.
"""

import warnings

import numpy as np
import pytest

from oracle import datasets, pca_f64, reference_path, upstream

pytestmark = pytest.mark.gpu


def _blobs(n, dim, seed, spread=2.5, n_types=5):
    rs = np.random.default_rng(seed)
    return (rs.normal(size=(n, dim)) + rs.integers(0, n_types, size=(n, 1)) * spread).astype(np.float32)


@pytest.mark.parametrize("n,k", [(400, 10), (3000, 10), (20000, 10), (5000, 15), (60000, 10)])
def test_device_umap_graph_matches_host_twin(handle, native, n, k):
    """dd_umap_graph (device: one thread per cell for smooth_knn_dist + strengths, one warp per cell for the fuzzy union
    on the symmetric pattern) against the host twin and the oracle's restatement on the device's own lists and
    distances: same pattern, same float32 weights.  The device's float64 exp may differ from libm's in the last bit; that
    survives the rounding to float32 with probability 2^-29 per edge and flips a bisection branch only if the membership
    sum lands within 1e-15 of its target, so bit equality is asserted."""
    emb = _blobs(n, 30, n + k, spread=1.5)
    handle.upload_embedding(emb)
    idx, dist = handle.knn(k)
    got = handle.umap_graph(k)
    want = native.umap_connectivities(idx, dist)
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    differ = int(np.count_nonzero(got.data != want.data))
    print(f"\n[umap graph n={n} k={k}] nnz {got.nnz}, weights that differ from the host twin: {differ}, "
          f"max relative difference {float(np.max(np.abs(got.data - want.data) / want.data)):.2e}")
    np.testing.assert_array_equal(got.data, want.data)
    if n <= 5000:
        ora = upstream.fuzzy_connectivities(idx, dist)
        np.testing.assert_array_equal(got.indices, ora.indices)
        np.testing.assert_array_equal(got.data, ora.data)


def test_device_umap_graph_duplicates_and_far_neighbours(handle, native):
    """Coincident cells (rho = 0 rows, distance 0 -> strength 1), and an outlier whose memberships underflow to 0 (entries
    that stay in the device pattern with weight 0 and are dropped by the canonical form)."""
    emb = _blobs(300, 4, 5)
    emb[10:16] = emb[10]
    emb[200] += 1.0e4
    handle.upload_embedding(emb)
    idx, dist = handle.knn(4)
    got = handle.umap_graph(4)
    want = native.umap_connectivities(idx, dist)
    np.testing.assert_array_equal(got.indptr, want.indptr)
    np.testing.assert_array_equal(got.indices, want.indices)
    np.testing.assert_array_equal(got.data, want.data)


def test_pipeline_leiden_matches_stagewise_calls(handle, native):
    """dd_fit_iterations(clustering=leiden) == stage-by-stage entry points + Leiden on the device-built umap graph of the
    device's own lists and distances (checks the pinned-slot layout and the ordering of the graph kernels and copies
    against the next iteration's kNN), and that graph == the host twin's == the oracle's, so == dd_leiden_knn."""
    raw = datasets.structured_counts(3000, 400, seed=7)
    n_cells, n_iters, n_synth = 3000, 4, 750
    rng = np.random.default_rng(5)
    parents = np.stack([rng.choice(n_cells, size=(n_synth, 2), replace=False) for _ in range(n_iters)])
    C = 30
    omega = pca_f64.omega(400, C, 0).astype(np.float32)
    n_power = pca_f64.auto_n_iter(n_cells + n_synth, 400, C)
    handle.upload_counts(raw)
    out = handle.fit_iterations(parents, omega, pseudocount=0.1, standard_scaling=False, n_comp=C, n_power_iter=n_power,
                                n_host_threads=3, clustering="leiden", resolution=4.0, seed=0)
    for i in range(n_iters):
        handle.create_doublets(parents[i])
        handle.normalise_log(handle.median_lib_size(), 0.1)
        handle.pca(C, omega, n_power)
        idx, dist = handle.knn(10)
        dev = handle.umap_graph(10)
        labels = native.leiden_csr(dev.indptr, dev.indices, dev.data.astype(np.float64), resolution=4.0, seed=0)
        np.testing.assert_array_equal(out["communities"][i], labels[:n_cells])
        np.testing.assert_array_equal(out["synth_communities"][i], labels[n_cells:])
        s, lp, _, _ = reference_path.score_communities(labels, n_cells)
        np.testing.assert_array_equal(out["scores"][i], s)
        np.testing.assert_allclose(out["log_p"][i], lp, rtol=1e-9, atol=1e-12)
        # the device graph against the oracle's restatement on the device's own lists: same bits
        want = upstream.fuzzy_connectivities(idx, dist)
        for got in (native.umap_connectivities(idx, dist), dev):
            np.testing.assert_array_equal(got.indices, want.indices)
            np.testing.assert_array_equal(got.data, want.data)
        np.testing.assert_array_equal(native.leiden_knn(idx, dist, resolution=4.0, seed=0), labels)


def test_classifier_leiden_vs_oracle():
    """BoostClassifier(clustering_algorithm="leiden") against the oracle's leiden path.  Parents are bit-exact.  The
    umap weights depend on the float32 kNN DISTANCES, and those inherit the 1e-4 relative difference between the
    GPU embedding and sklearn's own float32 embedding; at resolution 4 Leiden splits every cell type
    into many sub-communities whose borders react to that (measured with the host twin on perturbed oracle distances:
    identical partitions up to 1e-6 relative noise, adjusted Rand index 0.61-0.67 at 1e-4, doublet-score correlation
    0.75, 98.6 % equal calls).  So this comparison is statistical; the exact one is the test above."""
    from sklearn.metrics import adjusted_rand_score

    from doubletdetection_b200 import BoostClassifier

    counts = datasets.structured_counts(1500, 300, seed=1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_iters=3, random_state=0, n_jobs=2, clustering_algorithm="leiden").fit(counts)
        ora = reference_path.OracleClassifier(n_iters=3, random_state=0, clustering_algorithm="leiden").fit(counts)
        labels = clf.predict(p_thresh=1e-3, voter_thresh=0.5)
        want = ora.predict(p_thresh=1e-3, voter_thresh=0.5)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    assert clf.communities_.shape == ora.communities_.shape and clf.all_scores_.shape == ora.all_scores_.shape
    same, aris = 0, []
    for i in range(3):
        full_n = np.concatenate([clf.communities_[i], clf.synth_communities_[i]])
        full_o = np.concatenate([ora.communities_[i], ora.synth_communities_[i]])
        aris.append(adjusted_rand_score(full_o, full_n))
        same += int(np.array_equal(full_n, full_o))
        n_n, n_o = len(np.unique(full_n)), len(np.unique(full_o))
        assert 0.7 * n_o <= n_n <= 1.4 * n_o, (n_n, n_o)
        if np.array_equal(full_n, full_o):
            np.testing.assert_array_equal(clf.all_scores_[i], ora.all_scores_[i])
            np.testing.assert_allclose(clf.all_log_p_values_[i], ora.all_log_p_values_[i], rtol=1e-4, atol=1e-12)
    a = np.ma.filled(np.ma.asarray(clf.doublet_score(), dtype=np.float64), np.nan)
    b = np.ma.filled(np.ma.asarray(ora.doublet_score(), dtype=np.float64), np.nan)
    ok = np.isfinite(a) & np.isfinite(b)
    corr = np.corrcoef(a[ok], b[ok])[0, 1]
    agree = float(np.mean(np.asarray(labels) == np.asarray(want)))
    print(f"\n[leiden] identical communities in {same}/3 iterations; adjusted Rand {np.round(aris, 3)}; "
          f"doublet-score correlation {corr:.3f}; equal calls {agree:.4f}")
    assert min(aris) >= 0.4
    assert ok.mean() > 0.9 and corr > 0.5
    assert agree >= 0.95
    assert (clf.communities_ >= 0).all()  # no -1 labels on this path


def test_reference_package_test_mirrored():
    """The reference's own test (tests/test_package.py:6-48) on this package: 500 x 100 Poisson counts, two iterations
    with standard scaling for each of the three clustering algorithms, predict + doublet_score, and its one value
    assertion -- two leiden fits with random_state=123 give equal scores.  (The plotting calls of :40-42 are out of
    scope; the ValueError check of :45-48 lives in tests/test_api_surface.py.)"""
    from doubletdetection_b200 import BoostClassifier

    counts = np.random.default_rng(8).poisson(1.0, size=(500, 100))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for algo in ("louvain", "phenograph"):
            clf = BoostClassifier(n_iters=2, clustering_algorithm=algo, standard_scaling=True)
            labels = clf.fit(counts).predict(p_thresh=1e-16, voter_thresh=0.5)
            assert labels.shape == (500,) and np.asarray(clf.doublet_score()).shape == (500,)
        scores = []
        for _ in range(2):
            clf = BoostClassifier(n_iters=2, clustering_algorithm="leiden", standard_scaling=True, random_state=123)
            clf.fit(counts).predict(p_thresh=1e-16, voter_thresh=0.5)
            scores.append(clf.doublet_score())
    np.testing.assert_equal(np.ma.filled(scores[0], np.nan), np.ma.filled(scores[1], np.nan))
    assert np.isfinite(np.ma.filled(scores[0], np.nan)).any()


@pytest.mark.parametrize("name", ["c1_phenograph_scaled", "structured_900x200_phenograph"])
def test_classifier_phenograph_vs_reference_golden(name):
    """The PhenoGraph branch against goldens produced by the reference's REAL control flow (doubletdetection.py:317-327,
    tests/golden/make_golden.py) over the restated phenograph.cluster: parents bit-exact; wherever the communities of an
    iteration are identical scores are identical and log p-values agree to 1e-4; EVERY iteration must be identical."""
    from conftest import golden_case, load_golden

    from doubletdetection_b200 import BoostClassifier

    g = load_golden(name)
    counts, kw, pkw = golden_case(name)
    kw = dict(kw, clustering_kwargs=dict(kw.get("clustering_kwargs") or {}))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_jobs=2, **kw).fit(counts)
        labels = np.asarray(clf.predict(**pkw), dtype=np.float64)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), g["parents"])
    same = (clf.communities_ == g["communities"]).all(axis=1)
    print(f"\n[{name}] iterations with identical communities: {int(same.sum())}/{same.size}; "
          f"cells labelled -1 in iteration 0: {int((clf.communities_[0] < 0).sum())} (golden {int((g['communities'][0] < 0).sum())})")
    assert same.all(), f"communities differ from the golden in iterations {np.nonzero(~same)[0]}"
    np.testing.assert_array_equal(clf.synth_communities_, g["synth_communities"])
    np.testing.assert_array_equal(clf.all_scores_, g["all_scores"])
    np.testing.assert_allclose(clf.all_log_p_values_, g["all_log_p_values"], rtol=1e-4, atol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(labels, g["labels"])


def test_classifier_leiden_on_the_exact_pca_route(handle, native):
    """``clustering_algorithm="leiden"`` with a shape for which sklearn's PCA is exact (<= 1000 genes, >= 10x as many
    augmented cells -> covariance_eigh): the iteration-by-iteration route of the shim, which also takes umap's graph from
    the device.  The host twin (``dd_leiden_knn``: graph AND partition on the host) run on the device's own lists and
    distances must reproduce the classifier's communities and scores exactly."""
    from doubletdetection_b200 import BoostClassifier
    from doubletdetection_b200.classifier import _exact_pca, _pca_solver

    counts = datasets.structured_counts(2400, 200, seed=33)
    n = counts.shape[0]
    assert _pca_solver(n + n // 4, 200, 30) == "covariance_eigh"
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_iters=2, clustering_algorithm="leiden", random_state=0, n_jobs=2).fit(counts)
    handle.upload_counts(reference_path.prologue(counts, 10000)["raw"])
    for i in range(2):
        handle.create_doublets(np.asarray(clf._parents_array[i]))
        handle.normalise_log(handle.median_lib_size(), 0.1)
        _exact_pca(handle, 30)
        idx, dist = handle.knn(10)
        labels = native.leiden_knn(idx, dist, resolution=4.0, seed=0)
        np.testing.assert_array_equal(clf.communities_[i], labels[:n])
        np.testing.assert_array_equal(clf.synth_communities_[i], labels[n:])
        s, lp, _, _ = reference_path.score_communities(labels, n)
        np.testing.assert_array_equal(clf.all_scores_[i], s)
        np.testing.assert_allclose(clf.all_log_p_values_[i], lp, rtol=1e-9, atol=1e-12, equal_nan=True)
    assert clf.communities_.dtype == np.float64  # converted on access (the reference keeps float arrays, :188)
