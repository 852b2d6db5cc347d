"""End-to-end parity of ``BoostClassifier.fit`` / ``predict`` with the oracle AT THE BASELINE SIZES (c2 = 10k x 3k with
25 iterations, c3 = 100k x 3k with 3 iterations) -- asserted, not printed.  Needs a B200 (`-m gpu`).

Two comparisons per case:

* **chain** -- every stage against the oracle's stage ON THE SAME INPUT, composed over the whole iteration: parents
  bit-exact; the GPU's PCA embedding within 1e-4 of the float64 truth; the oracle's kNN + pattern graph + Louvain +
  hypergeometric scoring run on that GPU embedding must reproduce the classifier's communities, scores and log p-values of
  every iteration EXACTLY (log p to 1e-9).  Together: the only place the GPU path may leave the reference path is the
  PCA embedding, inside its stated tolerance.
* **drift** -- the oracle run end to end on its own float32 sklearn PCA.  sklearn's float32 embedding is itself 1e-4
  away from the float64 truth (the GPU is 5e-6 away), a few near-tied neighbour sets flip, and Louvain at
  resolution 4 amplifies every flip -- ``scripts/parity_drift_study.py`` measures the same drift between two CPU runs
  of the ORACLE that differ only in the PCA's working precision (c2: 9 of 25 iterations with identical communities,
  adjusted Rand >= 0.93 in the others; planted doublets: 99.8 % equal labels).  The thresholds below are those
  measured rates with a margin; the numbers are printed for the record.
"""

import time
import warnings

import numpy as np
import pytest
from sklearn.metrics import adjusted_rand_score

from oracle import datasets, louvain_c, pca_f64, reference_path, upstream

pytestmark = pytest.mark.gpu

C = 30


def _rel(a, b):
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / np.abs(b).max()


def _knn_equal_up_to_ties(emb, got, want):
    """The GPU re-ranks by the exact float64 distance, sklearn by ||x||^2 - 2 x.y + ||y||^2: lists may differ only where
    two neighbours are equidistant to 1e-6 relative.  Returns the number of such rows."""
    bad_rows = np.nonzero((got != want).any(axis=1))[0]
    e64 = emb.astype(np.float64)
    for r in bad_rows:
        for c in np.nonzero(got[r] != want[r])[0]:
            d_gpu = np.linalg.norm(e64[r] - e64[got[r, c]])
            d_ref = np.linalg.norm(e64[r] - e64[want[r, c]])
            assert abs(d_gpu - d_ref) <= 1e-6 * max(d_ref, 1e-30), (r, c, d_gpu, d_ref)
    return len(bad_rows)


def _fit_both(counts, n_iters, n_jobs=8):
    from doubletdetection_b200 import BoostClassifier

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_iters=n_iters, clustering_algorithm="louvain", random_state=0, n_jobs=n_jobs).fit(counts)
        t0 = time.perf_counter()
        ora = reference_path.OracleClassifier(n_iters=n_iters, random_state=0, louvain_fn=louvain_c.louvain).fit(counts)
        t_ora = time.perf_counter() - t0
    return clf, ora, t_ora


def _chain_check(handle, counts, clf, iterations, f64_check_iters=(0,)):
    """Oracle stages downstream of the GPU embedding == what the classifier produced, iteration by iteration."""
    n = counts.shape[0]
    g_ = counts.shape[1]
    parents = np.asarray(clf._parents_array)
    omega = pca_f64.omega(g_, C, 0).astype(np.float32)
    n_aug = n + parents.shape[1]
    n_iter = pca_f64.auto_n_iter(n_aug, g_, C)
    handle.upload_counts(counts)
    tie_rows = 0
    worst_emb = 0.0
    for i in iterations:
        handle.create_doublets(parents[i])
        handle.normalise_log(handle.median_lib_size(), 0.1)
        emb, _ = handle.pca(C, omega, n_iter)
        if i in f64_check_iters:
            truth, _, _ = pca_f64.randomized_pca_f64(handle.download_dense(), C, random_state=0)
            worst_emb = max(worst_emb, _rel(emb, truth))
        idx, _ = handle.knn(10)
        want_idx, _ = upstream.knn_brute(emb, 10)
        tie_rows += _knn_equal_up_to_ties(emb, idx, want_idx)
        graph = upstream.knn_pattern_graph(idx)
        labels = louvain_c.louvain(graph.indptr, graph.indices, None, resolution=4.0, seed=0, level0="parallel")
        np.testing.assert_array_equal(clf.communities_[i], labels[:n], err_msg=f"communities of iteration {i}")
        np.testing.assert_array_equal(clf.synth_communities_[i], labels[n:])
        s, lp, _, _ = reference_path.score_communities(labels, n)
        np.testing.assert_array_equal(clf.all_scores_[i], s)
        np.testing.assert_allclose(clf.all_log_p_values_[i], lp, rtol=1e-9, atol=1e-12)
    assert worst_emb < 1e-4, f"embedding vs float64 truth: {worst_emb}"
    return worst_emb, tie_rows


def _drift(clf, ora):
    n_iters = clf.communities_.shape[0]
    same = (clf.communities_ == ora.communities_).all(axis=1) & (clf.synth_communities_ == ora.synth_communities_).all(axis=1)
    ari, calls = [], []
    for i in range(n_iters):
        fa = np.concatenate([clf.communities_[i], clf.synth_communities_[i]])
        fb = np.concatenate([ora.communities_[i], ora.synth_communities_[i]])
        ari.append(1.0 if same[i] else adjusted_rand_score(fa, fb))
        calls.append(np.mean((clf.all_log_p_values_[i] <= np.log(1e-7)) == (ora.all_log_p_values_[i] <= np.log(1e-7))))
    return same, np.asarray(ari), np.asarray(calls)


def _report(tag, clf, ora, same, ari, calls, extra=""):
    print(f"\n[{tag}] identical-community iterations {int(same.sum())}/{same.size}; adjusted Rand min {ari.min():.4f} "
          f"median {np.median(ari):.4f}; per-iteration calls equal min {calls.min():.5f}{extra}")


# ------------------------------------------------------------------------------ c2: 10k x 3k, 25 iterations
def test_c2_25_iterations_vs_oracle(handle):
    counts = datasets.structured_counts(10000, 3000, seed=1234)
    clf, ora, t_ora = _fit_both(counts, 25)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    # chain: exact in EVERY one of the 25 iterations
    worst_emb, ties = _chain_check(handle, counts, clf, range(25), f64_check_iters=(0, 12, 24))
    # drift against the oracle's own float32 PCA
    same, ari, calls = _drift(clf, ora)
    labels, want = clf.predict(), ora.predict()
    _report("c2 x 25", clf, ora, same, ari, calls,
            f"; labels equal {np.mean(labels == want):.5f} (doublets {int(np.nansum(labels))}/{int(np.nansum(want))}); "
            f"embedding vs f64 {worst_emb:.1e}; kNN near-tie rows {ties}; oracle {t_ora:.0f} s")
    assert same.sum() >= 3  # CPU float32-vs-float64 study: 9/25
    assert ari.min() >= 0.85 and np.median(ari) >= 0.97  # CPU study: min 0.928, median 0.991
    assert calls.min() >= 0.995
    # the BASELINE dataset holds no doublets: both sides call none at the default thresholds -> labels bit-identical
    np.testing.assert_array_equal(labels, want)
    sa = np.ma.filled(np.ma.asarray(clf.doublet_score(), dtype=np.float64), np.nan)
    sb = np.ma.filled(np.ma.asarray(ora.doublet_score(), dtype=np.float64), np.nan)
    assert np.corrcoef(sa, sb)[0, 1] > 0.95


def test_c2_planted_doublets_labels_vs_oracle(handle):
    """The same shape with 8 % real doublets planted among the cells, so that ``predict`` has something to call: final
    0/1 labels against the oracle."""
    counts, truth = datasets.structured_counts_with_doublets(10000, 3000, seed=1234)
    clf, ora, _ = _fit_both(counts, 10)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    worst_emb, ties = _chain_check(handle, counts, clf, range(10))
    same, ari, calls = _drift(clf, ora)
    out = []
    for p, v in ((1e-7, 0.9), (1e-16, 0.5)):
        labels, want = clf.predict(p, v), ora.predict(p, v)
        eq = float(np.mean(labels == want))
        out.append(f"predict({p}, {v}): labels equal {eq:.5f}, doublets {int(np.nansum(labels))}/{int(np.nansum(want))}, "
                   f"planted recall {np.mean(labels[truth] == 1):.3f}/{np.mean(want[truth] == 1):.3f}")
        assert eq >= 0.99  # CPU float32-vs-float64 study on the 4k x 1k version: 0.998
        assert int(np.nansum(labels)) > 100  # the comparison is not vacuous
    _report("c2 planted doublets x 10", clf, ora, same, ari, calls, "; " + "; ".join(out) + f"; kNN near-tie rows {ties}")
    assert ari.min() >= 0.85 and calls.min() >= 0.99


# ------------------------------------------------------------------------------ c3: 100k x 3k, 3 iterations
def test_c3_3_iterations_vs_oracle(handle):
    """The headline configuration.  The oracle costs ~8 s per iteration here (bench.py's cpu_baseline leg runs the same
    code), so three iterations are compared in full."""
    counts = datasets.structured_counts(100000, 3000, seed=1234)
    clf, ora, t_ora = _fit_both(counts, 3, n_jobs=8)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    worst_emb, ties = _chain_check(handle, counts, clf, range(3), f64_check_iters=())
    same, ari, calls = _drift(clf, ora)
    labels, want = clf.predict(), ora.predict()
    _report("c3 x 3", clf, ora, same, ari, calls,
            f"; labels equal {np.mean(labels == want):.5f} (doublets {int(np.nansum(labels))}/{int(np.nansum(want))}); "
            f"kNN near-tie rows {ties}; oracle {t_ora:.0f} s")
    # measured on B200 (3 iterations): adjusted Rand 0.93 / 0.94 / 0.85, per-iteration calls equal 1.0 / 0.9906 / 1.0 -- one
    # community of ~900 cells on the p = 1e-7 edge is called by one side only in one iteration
    assert ari.min() >= 0.80 and calls.min() >= 0.98
    assert np.mean(labels == want) >= 0.999


# ------------------------------------------------------------------------------ edge: int(boost_rate * n_cells) == 0
def test_zero_synthetics_like_the_reference():
    """boost_rate so small that no synthetic is made (doubletdetection.py:391): the reference runs through -- every
    cluster has 0 synthetics, score 0, log p = logsf(0; A, 0, size) = -inf, which predict masks into NaN labels."""
    from doubletdetection_b200 import BoostClassifier

    counts = datasets.structured_counts(600, 200, seed=3)  # 600 x 200: sklearn's 'auto' picks the randomized solver
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = BoostClassifier(n_iters=2, boost_rate=0.001, clustering_algorithm="louvain", random_state=0).fit(counts)
        ora = reference_path.OracleClassifier(n_iters=2, boost_rate=0.001, random_state=0, louvain_fn=louvain_c.louvain).fit(counts)
        labels, want = clf.predict(), ora.predict()
    assert clf.synth_communities_.shape == (2, 0)
    np.testing.assert_array_equal(clf.all_scores_, ora.all_scores_)
    np.testing.assert_array_equal(clf.all_log_p_values_, ora.all_log_p_values_)
    np.testing.assert_array_equal(clf.communities_, ora.communities_)
    np.testing.assert_array_equal(labels, want)
