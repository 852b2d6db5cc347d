"""sklearn's EXACT PCA branches on the B200 path (doubletdetection.py:309-314 -> PCA(svd_solver="auto") ->
"covariance_eigh" for <= 1000 genes and >= 10x as many augmented cells, "full" for tiny matrices): device Gram matrix +
projection around a host eigendecomposition.  Needs a B200 (`-m gpu`)."""

import warnings

import numpy as np
import pytest

from oracle import datasets, louvain_c, pca_f64, reference_path, upstream

pytestmark = pytest.mark.gpu


def _blobs(shape, seed):
    rs = np.random.default_rng(seed)
    centres = rs.normal(size=(5, shape[1])) * 2.0
    return (centres[rs.integers(0, 5, shape[0])] + rs.normal(size=shape)).astype(np.float32)


@pytest.mark.parametrize("shape", [(12500, 1000), (2600, 120), (4097, 333), (400, 90)])
def test_centered_gram_and_projection_vs_numpy(handle, shape):
    X = _blobs(shape, shape[0])
    handle.upload_dense(X)
    Xc = X.astype(np.float64) - X.astype(np.float64).mean(axis=0)
    want = Xc.T @ Xc
    got = handle.centered_gram(False)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-9 * np.abs(want).max())
    assert np.array_equal(got, got.T)
    v = np.linalg.qr(np.random.default_rng(1).normal(size=(shape[1], 30)))[0]
    emb = handle.project(v)
    np.testing.assert_allclose(emb, (Xc @ v).astype(np.float32), rtol=2e-6, atol=2e-6)
    idx, _ = handle.knn(10)  # the projection left a valid embedding on the device
    assert (idx[:, 0] == np.arange(shape[0])).all()


@pytest.mark.parametrize("shape", [(60, 300), (36, 64), (300, 450)])
def test_centered_gram_on_the_cell_side(handle, shape):
    X = _blobs(shape, shape[1])
    handle.upload_dense(X)
    Xc = X.astype(np.float64) - X.astype(np.float64).mean(axis=0)
    want = Xc @ Xc.T
    np.testing.assert_allclose(handle.centered_gram(True), want, rtol=1e-11, atol=1e-9 * np.abs(want).max())


@pytest.mark.parametrize("shape", [(12500, 1000), (2600, 120), (400, 90), (60, 300), (36, 64)])
def test_exact_pca_vs_float64_truth_and_sklearn(handle, shape):
    """north_star tolerance 1e-4 against the float64 truth; sklearn's own float32 run is reported beside it."""
    from sklearn.decomposition import PCA

    from doubletdetection_b200.classifier import _exact_pca, _pca_solver

    X = _blobs(shape, 7 + shape[0])
    assert _pca_solver(shape[0], shape[1], 30) in ("covariance_eigh", "full")
    handle.upload_dense(X)
    emb = _exact_pca(handle, 30)
    truth, _ = pca_f64.exact_pca_f64(X, 30)
    err = np.abs(emb - truth).max() / np.abs(truth).max()
    sk = PCA(n_components=30, svd_solver="auto", random_state=0).fit_transform(X)
    err_sk = np.abs(sk - truth).max() / np.abs(truth).max()
    print(f"\n[exact PCA {shape}] GPU vs f64 truth {err:.2e}; sklearn-f32 vs f64 truth {err_sk:.2e}")
    assert err < 1e-4


@pytest.mark.parametrize("algo", ["louvain", "phenograph"])
def test_classifier_with_1000_top_var_genes(handle, algo):
    """``n_top_var_genes=1000`` on a 10k-cell dataset (a common setting): A = 12 500 >= 10 G -> covariance_eigh.  Parents
    bit-exact; per iteration the oracle's downstream stages run on the GPU embedding reproduce the classifier's communities
    and scores exactly; the embedding is within 1e-4 of the float64 truth."""
    from doubletdetection_b200 import BoostClassifier
    from doubletdetection_b200.classifier import _exact_pca

    counts = datasets.structured_counts(10000, 1500, seed=21)
    kw = dict(n_iters=2, n_top_var_genes=1000, clustering_algorithm=algo, random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_jobs=4, **kw).fit(counts)
        ora = reference_path.OracleClassifier(louvain_fn=louvain_c.louvain, **kw).fit(counts)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    np.testing.assert_array_equal(clf.top_var_genes_, ora.top_var_genes_)
    pro = reference_path.prologue(counts, 1000)
    n = counts.shape[0]
    handle.upload_counts(pro["raw"])
    for i in range(2):
        handle.create_doublets(np.asarray(clf._parents_array[i]))
        handle.normalise_log(handle.median_lib_size(), 0.1)
        emb = _exact_pca(handle, 30)
        if i == 0:
            truth, _ = pca_f64.exact_pca_f64(handle.download_dense(), 30)
            err = np.abs(emb - truth).max() / np.abs(truth).max()
            print(f"\n[{algo}, n_top_var_genes=1000] embedding vs f64 truth {err:.2e}")
            assert err < 1e-4
        if algo == "phenograph":
            labels, _ = upstream.phenograph_cluster(emb, seed=0, louvain_fn=louvain_c.louvain, prune=True)
        else:
            idx, _ = upstream.knn_brute(emb, 10)
            g = upstream.knn_pattern_graph(idx)
            labels = louvain_c.louvain(g.indptr, g.indices, None, resolution=4.0, seed=0, level0="parallel")
        np.testing.assert_array_equal(clf.communities_[i], labels[:n])
        s, lp, _, _ = reference_path.score_communities(labels, n)
        np.testing.assert_array_equal(clf.all_scores_[i], s)
        np.testing.assert_allclose(clf.all_log_p_values_[i], lp, rtol=1e-9, atol=1e-12, equal_nan=True)
    # drift against the oracle's own float32 covariance_eigh run
    from sklearn.metrics import adjusted_rand_score

    ari = [adjusted_rand_score(clf.communities_[i], ora.communities_[i]) for i in range(2)]
    print(f"[{algo}] adjusted Rand vs the oracle's float32 run: {np.round(ari, 4)}")
    # sklearn's covariance_eigh forms X^T X and subtracts n * mean mean^T in FLOAT32 (catastrophic cancellation on log counts
    # whose mean is far from 0): its embedding is 1e-3 .. 1e-2 away from the float64 truth, so the oracle's own partition is
    # the noisy side here (measured on B200: adjusted Rand 0.68 for louvain at resolution 4)
    assert min(ari) > 0.5


# ---------------------------------------------------------------- pseudocount == 1: the reference's sparse / arpack branch
def test_pseudocount1_stages_vs_reference_golden(handle):
    """doubletdetection.py:296-297, 308 through the reference's real control flow (golden ``structured_1200x260_pc1``): the
    dense log1p matrix equals the reference's sparse one, and the exact PCA is within 1e-4 of the float64 truth (sklearn's own
    float32 ARPACK run, the golden's X_pca, is 7e-5 away from that truth)."""
    from conftest import golden_case, load_golden

    from doubletdetection_b200.classifier import _exact_pca

    g = load_golden("structured_1200x260_pc1")
    counts, kw, _ = golden_case("structured_1200x260_pc1")
    pro = reference_path.prologue(counts, 10000)
    handle.upload_counts(pro["raw"])
    handle.create_doublets(g["parents"][0])
    handle.normalise_log(handle.median_lib_size(), 1.0)
    dense = handle.download_dense()
    np.testing.assert_allclose(dense, g["pca_input0"], rtol=3e-6, atol=1e-7)
    assert (dense[g["pca_input0"] == 0] == 0).all()  # log1p(0): the sparse matrix's gaps are exact zeros
    emb = _exact_pca(handle, 30)
    truth, _ = pca_f64.exact_pca_f64(dense, 30)
    err = np.abs(emb - truth).max() / np.abs(truth).max()
    err_ref = np.abs(g["X_pca0"] - truth).max() / np.abs(truth).max()
    print(f"\n[pseudocount=1] GPU exact PCA vs f64 truth {err:.2e}; the reference's ARPACK (float32) vs f64 truth {err_ref:.2e}")
    assert err < 1e-4
    np.testing.assert_allclose(emb, g["X_pca0"], atol=5e-4 * np.abs(truth).max())


@pytest.mark.parametrize("shape,n_top", [((1200, 260), 10000), ((6000, 3000), 10000)])
def test_classifier_pseudocount1_chain_vs_oracle(handle, shape, n_top):
    """BoostClassifier(pseudocount=1): parents bit-exact; the oracle's downstream stages on the GPU's exact embedding
    reproduce communities and scores of every iteration; end to end against the reference golden / the oracle's ARPACK run
    the communities agree up to the drift of its float32 solver."""
    from sklearn.metrics import adjusted_rand_score

    from doubletdetection_b200 import BoostClassifier
    from doubletdetection_b200.classifier import _exact_pca

    counts = datasets.structured_counts(*shape, seed=5)
    kw = dict(n_iters=2, clustering_algorithm="louvain", pseudocount=1, random_state=0, n_top_var_genes=n_top)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(**kw).fit(counts)
        ora = reference_path.OracleClassifier(louvain_fn=louvain_c.louvain, **kw).fit(counts)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    n = counts.shape[0]
    handle.upload_counts(reference_path.prologue(counts, n_top)["raw"])
    for i in range(2):
        handle.create_doublets(np.asarray(clf._parents_array[i]))
        handle.normalise_log(handle.median_lib_size(), 1.0)
        emb = _exact_pca(handle, 30)
        idx, _ = upstream.knn_brute(emb, 10)
        gph = upstream.knn_pattern_graph(idx)
        labels = louvain_c.louvain(gph.indptr, gph.indices, None, resolution=4.0, seed=0, level0="parallel")
        np.testing.assert_array_equal(clf.communities_[i], labels[:n])
        s, lp, _, _ = reference_path.score_communities(labels, n)
        np.testing.assert_array_equal(clf.all_scores_[i], s)
        np.testing.assert_allclose(clf.all_log_p_values_[i], lp, rtol=1e-9, atol=1e-12, equal_nan=True)
    ari = [adjusted_rand_score(clf.communities_[i], ora.communities_[i]) for i in range(2)]
    print(f"\n[pseudocount=1 {shape}] adjusted Rand vs the oracle's float32 ARPACK run: {np.round(ari, 4)}")
    assert min(ari) > 0.8


def test_pseudocount1_with_scaling_takes_the_dense_solver(handle):
    """standard_scaling densifies the matrix (sc.pp.scale zero-centres), so :308 picks "auto" again: the pipelined loop with
    the log1p build.  Chain check against the oracle (scaled log1p matrix, randomized PCA on the GPU embedding's stages) and
    determinism."""
    from doubletdetection_b200 import BoostClassifier

    counts = datasets.structured_counts(1500, 600, seed=9)
    kw = dict(n_iters=2, clustering_algorithm="louvain", pseudocount=1, standard_scaling=True, random_state=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = BoostClassifier(**kw).fit(counts)
        b = BoostClassifier(**kw).fit(counts)
        ora = reference_path.OracleClassifier(louvain_fn=louvain_c.louvain, keep_stages=True, **kw).fit(counts)
    np.testing.assert_array_equal(a.communities_, b.communities_)
    assert np.isfinite(a.all_scores_).all()
    np.testing.assert_array_equal(np.asarray(a.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    n, g_ = counts.shape
    omega = pca_f64.omega(g_, 30, 0).astype(np.float32)
    n_iter = pca_f64.auto_n_iter(n + n // 4, g_, 30)
    handle.upload_counts(reference_path.prologue(counts, 10000)["raw"])
    for i in range(2):
        handle.create_doublets(np.asarray(a._parents_array[i]))
        handle.normalise_log(handle.median_lib_size(), 1.0)
        handle.standard_scale(15.0)
        if i == 0:  # the scaled log1p matrix against the oracle's
            np.testing.assert_allclose(handle.download_dense(), ora.stages[0]["aug"], rtol=2e-5, atol=2e-5)
        emb, _ = handle.pca(30, omega, n_iter)
        idx, _ = upstream.knn_brute(emb, 10)
        gph = upstream.knn_pattern_graph(idx)
        labels = louvain_c.louvain(gph.indptr, gph.indices, None, resolution=4.0, seed=0, level0="parallel")
        np.testing.assert_array_equal(a.communities_[i], labels[:n])


@pytest.mark.parametrize("name", ["structured_1200x260_pc1", "structured_1200x260_pc1_scaled"])
def test_classifier_pseudocount1_vs_reference_golden(name):
    """The sparse branch end to end against goldens produced by the reference's REAL control flow (doubletdetection.py:296-297,
    302-303, 308; tests/golden/make_golden.py): parents bit-exact, final labels identical, communities up to the drift of the
    reference's float32 PCA (ARPACK resp. randomized) against the GPU's."""
    from conftest import golden_case, load_golden
    from sklearn.metrics import adjusted_rand_score

    from doubletdetection_b200 import BoostClassifier

    g = load_golden(name)
    counts, kw, pkw = golden_case(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(**kw).fit(counts)
        labels = np.asarray(clf.predict(**pkw), dtype=np.float64)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), g["parents"])
    same = (clf.communities_ == g["communities"]).all(axis=1)
    ari = [adjusted_rand_score(clf.communities_[i], g["communities"][i]) for i in range(same.size)]
    print(f"\n[{name}] iterations with identical communities: {int(same.sum())}/{same.size}; adjusted Rand {np.round(ari, 4)}")
    assert min(ari) >= 0.9
    agree = np.mean((labels == g["labels"]) | (np.isnan(labels) & np.isnan(g["labels"])))
    assert agree >= 0.99, agree
    for i in np.nonzero(same)[0]:
        np.testing.assert_array_equal(clf.all_scores_[i], g["all_scores"][i])
        np.testing.assert_allclose(clf.all_log_p_values_[i], g["all_log_p_values"][i], rtol=1e-4, atol=1e-12)


@pytest.mark.parametrize("algo,n_genes", [("louvain", 40), ("leiden", 24)])
def test_at_most_50_genes_neighbours_on_x(handle, algo, n_genes):
    """with at most 50 genes ``sc.pp.neighbors`` (doubletdetection.py:331-336) takes its neighbours from adata.X
    itself instead of X_pca (scanpy's N_PCS).  The classifier must do the same: the oracle's downstream stages run on the
    DEVICE's normalised matrix reproduce its communities and scores exactly, and no PCA is involved any more, so the
    oracle's own end-to-end run agrees as well (up to near-ties among 2-ulp different log values)."""
    from sklearn.metrics import adjusted_rand_score

    from doubletdetection_b200 import BoostClassifier, _capi
    from oracle import leiden_ref

    rs = np.random.default_rng(77)
    n = 1500
    prof = 25.0 * np.exp(rs.normal(0.0, 0.7, size=(6, n_genes)))  # deep counts: no two cells get identical rows
    counts = rs.poisson(prof[rs.integers(0, 6, n)] * rs.lognormal(0.0, 0.25, size=(n, 1))).astype(np.float32)
    kw = dict(n_iters=2, clustering_algorithm=algo, random_state=0, n_components=min(30, n_genes // 2))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_jobs=2, **kw).fit(counts)
        ora = reference_path.OracleClassifier(louvain_fn=louvain_c.louvain, **kw).fit(counts)
        with pytest.raises(ValueError):  # sklearn's check inside sc.tl.pca: more components than genes
            BoostClassifier(n_iters=2, clustering_algorithm=algo, n_components=n_genes + 1).fit(counts)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    handle.upload_counts(reference_path.prologue(counts, 10000)["raw"])
    for i in range(2):
        handle.create_doublets(np.asarray(clf._parents_array[i]))
        handle.normalise_log(handle.median_lib_size(), 0.1)
        idx, dist = upstream.knn_brute(handle.download_dense(), 10)
        if algo == "leiden":
            C = upstream.fuzzy_connectivities(idx, dist)
            labels = leiden_ref.leiden(C.indptr, C.indices, C.data.astype(np.float64), resolution=4.0, seed=0)
        else:
            g = upstream.knn_pattern_graph(idx)
            labels = louvain_c.louvain(g.indptr, g.indices, None, resolution=4.0, seed=0, level0="parallel")
        labels = np.asarray(labels)
        if algo == "louvain":  # the pattern graph depends on the neighbour SETS only
            np.testing.assert_array_equal(clf.communities_[i], labels[:n])
            s, lp, _, _ = reference_path.score_communities(labels, n)
            np.testing.assert_array_equal(clf.all_scores_[i], s)
        else:  # umap's weights see the float32 distances: sklearn's and the device's differ in the last bits
            assert adjusted_rand_score(clf.communities_[i], labels[:n]) > 0.5
    ari = [adjusted_rand_score(clf.communities_[i], ora.communities_[i]) for i in range(2)]
    print(f"\n[{algo}, {n_genes} genes] adjusted Rand vs the oracle's end-to-end run: {np.round(ari, 4)}")
    assert min(ari) > (0.9 if algo == "louvain" else 0.5)
