"""Drop-in surface of BoostClassifier (constructor validation, warnings, predict / doublet_score
semantics) -- mirrors what the reference's own test exercises (tests/test_package.py) plus the quirks
listed in SURVEY.md.  CPU only: nothing here launches a kernel."""

import inspect
import warnings

import numpy as np
import pytest

from conftest import load_golden
from doubletdetection_b200 import BoostClassifier, iteration_shard
from oracle import reference_path


def test_constructor_signature_matches_reference():
    sig = inspect.signature(BoostClassifier.__init__)
    names = list(sig.parameters)[1:]
    ref = ["boost_rate", "n_components", "n_top_var_genes", "replace", "clustering_algorithm", "clustering_kwargs",
           "n_iters", "normalizer", "pseudocount", "random_state", "verbose", "standard_scaling", "n_jobs"]
    assert names[: len(ref)] == ref
    d = {k: v.default for k, v in sig.parameters.items()}
    assert (d["boost_rate"], d["n_components"], d["n_top_var_genes"], d["replace"]) == (0.25, 30, 10000, False)
    assert (d["clustering_algorithm"], d["n_iters"], d["pseudocount"], d["random_state"]) == ("phenograph", 10, 0.1, 0)
    assert (d["verbose"], d["standard_scaling"], d["n_jobs"]) == (False, False, 1)
    for extra in names[len(ref):]:
        assert sig.parameters[extra].kind is inspect.Parameter.KEYWORD_ONLY


def test_bad_clustering_algorithm_raises():
    with pytest.raises(ValueError):  # reference tests/test_package.py:45-48
        BoostClassifier(clustering_algorithm="foo")


def test_kwargs_defaults_and_guards():
    clf = BoostClassifier(clustering_algorithm="louvain")
    assert clf.clustering_kwargs == {"directed": False, "resolution": 4}
    with pytest.raises(ValueError):
        BoostClassifier(clustering_algorithm="louvain", clustering_kwargs={"key_added": "x"})
    with pytest.raises(ValueError):
        BoostClassifier(clustering_algorithm="leiden", clustering_kwargs={"random_state": 1})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert BoostClassifier(clustering_algorithm="phenograph").clustering_kwargs == {"prune": True}


def test_warnings():
    with pytest.warns(UserWarning, match="Leiden"):
        BoostClassifier(clustering_algorithm="leiden")
    with pytest.warns(UserWarning, match="boost_rate is trimmed"):
        clf = BoostClassifier(clustering_algorithm="louvain", boost_rate=0.8)
    assert clf.boost_rate == 0.5
    with pytest.warns(UserWarning, match="prune=False"):
        BoostClassifier(clustering_algorithm="phenograph", n_iters=1)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert BoostClassifier(clustering_algorithm="louvain", boost_rate=0.8, replace=True).boost_rate == 0.8


def test_n_components_capping_and_assert():
    assert BoostClassifier(clustering_algorithm="louvain", n_top_var_genes=20).n_components == 20
    assert BoostClassifier(clustering_algorithm="louvain", n_top_var_genes=-5).n_top_var_genes == 0
    with pytest.raises(AssertionError):
        BoostClassifier(clustering_algorithm="louvain", n_components=50, n_top_var_genes=40)


def test_unsupported_paths_fail_loudly():
    x = np.random.default_rng(0).poisson(1.0, (600, 120))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(NotImplementedError):
            # sc.tl.leiden arguments the native Leiden does not cover
            BoostClassifier(n_iters=2, clustering_algorithm="leiden", clustering_kwargs={"n_iterations": 2}).fit(x)
        with pytest.raises(NotImplementedError):  # phenograph arguments the native graph does not cover
            BoostClassifier(n_iters=2, clustering_kwargs={"primary_metric": "cosine"}).fit(x)
        with pytest.raises(NotImplementedError):
            BoostClassifier(n_iters=2, clustering_kwargs={"nn_method": "brute"}).fit(x)
        with pytest.raises(NotImplementedError):
            BoostClassifier(n_iters=2, clustering_algorithm="louvain", normalizer=lambda c: c).fit(x)
    with pytest.raises(ValueError):  # sklearn check_array, as in the reference (:149-155)
        BoostClassifier(n_iters=2, clustering_algorithm="louvain").fit(np.full((50, 20), np.nan))


@pytest.mark.parametrize("name,pkw", [("c1_louvain", dict(p_thresh=1e-16, voter_thresh=0.5)),
                                      ("structured_1500x300", dict(p_thresh=1e-3, voter_thresh=0.5)),
                                      ("single_iter", dict())])
def test_predict_and_score_semantics_on_golden_fit(name, pkw):
    g = load_golden(name)
    n_iters = g["all_scores"].shape[0]
    clf = BoostClassifier(n_iters=n_iters, clustering_algorithm="louvain")
    clf.all_scores_ = g["all_scores"].copy()
    clf.all_log_p_values_ = g["all_log_p_values"].copy()
    labels = clf.predict(**pkw)
    np.testing.assert_array_equal(np.asarray(labels, dtype=np.float64), g["labels"])
    sc = clf.doublet_score()
    np.testing.assert_array_equal(np.ma.filled(np.ma.asarray(sc, dtype=np.float64), np.nan), g["doublet_score"])
    if n_iters > 1:
        assert isinstance(sc, np.ma.MaskedArray)  # quirk Q5
        np.testing.assert_array_equal(clf.voting_average_, g["voting_average"])
    else:
        assert labels.dtype == bool  # quirk Q4
        assert clf.suggested_score_cutoff_ == g["suggested_score_cutoff"]


def test_predict_masks_invalid_log_p():
    clf = BoostClassifier(n_iters=3, clustering_algorithm="louvain")
    lp = np.array([[-50.0, np.nan, -np.inf, -1.0], [-40.0, np.nan, -30.0, -2.0], [-1.0, np.nan, -30.0, np.nan]])
    clf.all_log_p_values_ = lp
    clf.all_scores_ = np.zeros_like(lp)
    want = reference_path.predict(lp, clf.all_scores_, 3, p_thresh=1e-7, voter_thresh=0.6)
    got = clf.predict(p_thresh=1e-7, voter_thresh=0.6)
    np.testing.assert_array_equal(got, want["labels"])
    assert np.isnan(got[1])  # fully masked cell
    assert got[2] == 1.0  # the -inf iteration is excluded from the vote
    np.testing.assert_array_equal(np.ma.filled(clf.doublet_score(), np.nan),
                                  np.ma.filled(reference_path.doublet_score(lp, 3), np.nan))


def test_parents_property_structure():
    clf = BoostClassifier(n_iters=2, clustering_algorithm="louvain")
    with pytest.raises(AttributeError):
        clf.parents_
    clf._parents_array = np.arange(12, dtype=np.int64).reshape(2, 3, 2)
    p = clf.parents_
    assert isinstance(p, list) and isinstance(p[0], list) and p[1][2] == [10, 11]
    assert isinstance(p[0][0][0], np.int64)


def test_iteration_shard_partitions():
    for n_iters, world in [(24, 8), (25, 8), (3, 2), (5, 8)]:
        blocks = [iteration_shard(n_iters, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n_iters
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    assert [b - a for a, b in (iteration_shard(24, r, 8) for r in range(8))] == [3] * 8


@pytest.mark.parametrize("seed", range(6))
def test_fast_predict_and_score_equal_the_reference_lines(seed):
    """predict / doublet_score avoid numpy's masked-array machinery; the literal restatement of the reference's lines
    (oracle.reference_path.predict / doublet_score, pinned to the reference's real code by the goldens) must give the
    same bits on inputs with NaN, +-inf, fully invalid cells and thresholds that hit votes exactly."""
    from oracle import reference_path

    rs = np.random.default_rng(seed)
    n_iters, n = int(rs.integers(2, 12)), 4000
    lp = -rs.exponential(8.0, (n_iters, n))
    lp[rs.random(lp.shape) < 0.05] = -np.inf
    lp[rs.random(lp.shape) < 0.05] = np.nan
    lp[rs.random(lp.shape) < 0.01] = np.inf
    lp[:, :7] = np.nan  # no valid iteration at all
    lp[:, 7:11] = -np.inf
    lp[0, 11] = -3.0  # exactly one valid iteration
    lp[1:, 11] = np.nan
    clf = BoostClassifier(n_iters=n_iters, clustering_algorithm="louvain")
    clf.all_log_p_values_ = lp.copy()
    clf.all_scores_ = rs.random(lp.shape)
    for p_thresh, voter_thresh in [(1e-7, 0.9), (1e-3, 0.5), (np.exp(-3.0), 1.0), (1e-16, 0.0)]:
        want = reference_path.predict(lp, clf.all_scores_, n_iters, p_thresh, voter_thresh)
        got = clf.predict(p_thresh, voter_thresh)
        assert isinstance(got, np.ndarray) and not isinstance(got, np.ma.MaskedArray) and got.dtype == np.float64
        np.testing.assert_array_equal(got, want["labels"])
        np.testing.assert_array_equal(clf.voting_average_, want["voting_average"])
    want = reference_path.doublet_score(lp, n_iters)
    got = clf.doublet_score()
    assert isinstance(got, np.ma.MaskedArray)
    np.testing.assert_array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(want))
    np.testing.assert_array_equal(np.ma.filled(got, np.nan), np.ma.filled(want, np.nan))
    # nothing invalid: still a MaskedArray, nothing masked
    clf.all_log_p_values_ = -rs.exponential(8.0, (n_iters, 50))
    got = clf.doublet_score()
    want = reference_path.doublet_score(clf.all_log_p_values_, n_iters)
    assert isinstance(got, np.ma.MaskedArray) and not np.ma.getmaskarray(got).any()
    np.testing.assert_array_equal(np.asarray(got), np.asarray(want))


def test_integration_stub_matches_the_abi():
    """INTEGRATION.md's ctypes stub (what a maintainer of the reference would paste) must describe the same
    ``dd_fit_params`` as the binding the tests run through, and only call symbols the header declares."""
    import ctypes
    import os
    import re

    from conftest import ROOT
    from doubletdetection_b200 import _capi

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = next(b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "_FitParams" in b)
    struct_src = block[block.index("class _FitParams"):block.index("def _p(")]
    ns = {"ctypes": ctypes}
    exec(struct_src, ns)
    assert ns["_FitParams"]._fields_ == _capi.FitParams._fields_
    assert ctypes.sizeof(ns["_FitParams"]) == ctypes.sizeof(_capi.FitParams)
    used = set(re.findall(r"_lib\.(dd_[a-z0-9_]+)", block))
    assert used and used <= set(_capi.SIGNATURES), used - set(_capi.SIGNATURES)
    header = open(os.path.join(ROOT, "include", "dd_b200.h")).read()
    for name in re.findall(r"`(dd_[a-z0-9_]+)`", text):  # every entry point the document mentions exists
        assert re.search(r"\b" + name + r"\s*\(", header), name
