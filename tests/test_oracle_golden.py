"""The oracle (oracle/reference_path.py + oracle/upstream.py) against the committed golden vectors,
which were produced by the reference's own code (tests/golden/make_golden.py).  CPU only."""

import warnings

import numpy as np
import pytest

from conftest import CLUSTER_GOLDEN_NAMES, GOLDEN_NAMES, SPARSE_GOLDEN_NAMES, golden_case, load_golden
from oracle import louvain_c, louvain_ref, pca_f64, reference_path, refshim, upstream


def _oracle_fit(name):
    counts, kw, pkw = golden_case(name)
    kw = dict(kw)
    kw["clustering_kwargs"] = dict(kw.get("clustering_kwargs") or {})  # the classifier adds its defaults in place
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = reference_path.OracleClassifier(louvain_fn=louvain_c.louvain, keep_stages=True, **kw)
        clf.fit(counts)
        labels = np.asarray(clf.predict(**pkw), dtype=np.float64)
        score = clf.doublet_score()
    return clf, labels, score


@pytest.mark.parametrize("name", GOLDEN_NAMES + CLUSTER_GOLDEN_NAMES + SPARSE_GOLDEN_NAMES)
def test_oracle_matches_reference_goldens(name):
    g = load_golden(name)
    clf, labels, score = _oracle_fit(name)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), g["parents"])
    np.testing.assert_array_equal(clf.communities_, g["communities"])
    np.testing.assert_array_equal(clf.synth_communities_, g["synth_communities"])
    np.testing.assert_array_equal(clf.all_scores_, g["all_scores"])
    np.testing.assert_array_equal(clf.all_log_p_values_, g["all_log_p_values"])
    np.testing.assert_array_equal(labels, g["labels"])
    np.testing.assert_array_equal(np.ma.filled(np.ma.asarray(score, dtype=np.float64), np.nan), g["doublet_score"])
    if "top_var_genes" in g:
        np.testing.assert_array_equal(clf.top_var_genes_, g["top_var_genes"])
    if "synth0_data" in g:
        st = clf.stages[0]
        np.testing.assert_array_equal(st["raw_synth"].indptr, g["synth0_indptr"])
        np.testing.assert_array_equal(st["raw_synth"].indices, g["synth0_indices"])
        np.testing.assert_array_equal(st["raw_synth"].data, g["synth0_data"])
        aug = st["aug"].toarray() if hasattr(st["aug"], "toarray") else st["aug"]  # sparse on the pseudocount == 1 branch
        np.testing.assert_array_equal(np.asarray(aug, dtype=np.float32), g["pca_input0"])
        np.testing.assert_array_equal(st["X_pca"], g["X_pca0"])
        np.testing.assert_array_equal(st["knn_indices"], g["knn_indices0"])


@pytest.mark.skipif(not refshim.reference_available(), reason="reference only exists in the build container")
def test_goldens_reproduce_from_reference_code():
    """Re-run the UNMODIFIED reference module (over the stub scanpy/anndata) and compare with the fixture."""
    counts, kw, pkw = golden_case("c1_louvain")
    mod = refshim.load_reference(louvain_fn=louvain_c.louvain)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = mod.BoostClassifier(**kw)
        clf.fit(counts)
        labels = clf.predict(**pkw)
    g = load_golden("c1_louvain")
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), g["parents"])
    np.testing.assert_array_equal(clf.all_log_p_values_, g["all_log_p_values"])
    np.testing.assert_array_equal(np.asarray(labels, dtype=np.float64), g["labels"])


def test_known_answers_from_survey():
    """SURVEY.md section 8(c) pins."""
    ch = np.random.default_rng(0).choice(500, size=(125, 2), replace=False)
    np.testing.assert_array_equal(ch[:3], [[323, 363], [42, 367], [43, 275]])
    from scipy.stats import hypergeom

    assert hypergeom.logsf(10, 625, 125, 40) == pytest.approx(-1.8748274242511567, rel=1e-12)
    assert hypergeom.logsf(0, 625, 125, 3) == pytest.approx(-0.7161786227264325, rel=1e-12)
    assert hypergeom.logsf(5, 625, 125, 5) == -np.inf


def test_louvain_c_matches_python_spec():
    rs = np.random.default_rng(5)
    for n, k in [(60, 4), (300, 6), (1000, 9)]:
        pts = rs.normal(size=(n, 5)) + rs.integers(0, 4, size=(n, 1)) * 3.0
        idx, _ = upstream.knn_brute(pts.astype(np.float32), k + 1)
        S = upstream.knn_pattern_graph(idx)
        for gamma in (1.0, 4.0):
            a = louvain_ref.louvain(S.indptr, S.indices, None, resolution=gamma, seed=3)
            b = louvain_c.louvain(S.indptr, S.indices, None, resolution=gamma, seed=3)
            np.testing.assert_array_equal(a, b)


def test_f64_pca_close_to_sklearn_f32():
    """The float64 restatement is the truth for the 1e-4 embedding tolerance; sklearn's own float32 run
    must sit within a few 1e-4 of it."""
    g = load_golden("structured_1500x300")
    X = g["pca_input0"]
    emb64, _, _ = pca_f64.randomized_pca_f64(X, 30, random_state=0)
    scale = np.abs(emb64).max()
    assert np.abs(emb64 - g["X_pca0"]).max() / scale < 2e-3


def test_f64_pca_transposed_branch_matches_sklearn():
    """Fewer samples than features: sklearn's randomized_svd transposes the problem.  The float64 restatement
    must follow it (same Omega shape, same sign convention)."""
    from sklearn.decomposition import PCA

    rs = np.random.default_rng(3)
    centres = rs.normal(size=(4, 900)) * 2.0
    X = (centres[rs.integers(0, 4, 600)] + rs.normal(size=(600, 900))).astype(np.float32)
    want = PCA(n_components=30, svd_solver="randomized", random_state=0).fit_transform(X)
    got, _, _ = pca_f64.randomized_pca_f64(X, 30, random_state=0)
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-3
