"""The Python shim's orchestration (doubletdetection_b200/classifier.py) on the CPU: ``_capi.Handle`` is replaced by a
stand-in whose ``fit_iterations`` evaluates every iteration with the ORACLE's stage functions on the parents the shim
drew.  What is under test is therefore everything the shim itself does -- validation, HVG selection and column order,
canonicalisation, the PCG64 parent stream (all iterations up front, continuing across fits), the mapping of constructor
arguments to ``dd_fit_params``, iteration ranges, result attributes and their types -- against ``OracleClassifier``,
which is pinned to the reference's real code by the goldens.  No CUDA, no libdd_b200 compute."""

import warnings

import numpy as np
import pytest
import scipy.sparse as sp_sparse

from conftest import golden_case, load_golden
from oracle import datasets, louvain_c, reference_path, upstream


class OracleHandle:
    """Same surface as ``_capi.Handle`` as far as ``BoostClassifier.fit`` uses it."""

    calls = []

    def __init__(self, device=0):
        self.device = device
        self.raw = None

    def upload_counts(self, csr):
        assert sp_sparse.issparse(csr) and csr.dtype == np.float32 and csr.has_canonical_format
        self.raw = csr.copy()
        self.n_cells, self.n_genes = csr.shape

    def counts_all_finite(self):
        return bool(np.isfinite(self.raw.data).all())

    def share_counts(self, src):  # a second pipeline on the same GPU reads the first one's matrix
        self.raw, self.n_cells, self.n_genes = src.raw, src.n_cells, src.n_genes

    def hvg_variances(self):  # what hvg.cu reproduces: the reference's own scipy expression (doubletdetection.py:166-169)
        return (np.array(self.raw.power(2).mean(axis=0)) - (np.array(self.raw.mean(axis=0))) ** 2)[0]

    def select_genes(self, genes):  # :173-175
        self.raw = self.raw.tocsc()[:, np.asarray(genes)].tocsr()
        self.n_cells, self.n_genes = self.raw.shape

    def fit_iterations(self, parents, omega, *, pseudocount, standard_scaling, n_comp, n_power_iter, knn_k=10, resolution=4.0,
                       seed=0, n_host_threads=1, iter_begin=0, iter_end=None, scale_max_value=15.0, clustering="louvain",
                       pheno_k=30, pheno_prune=True, pheno_min_cluster_size=10, out=None):
        OracleHandle.calls.append(dict(n_comp=n_comp, n_power_iter=n_power_iter, omega_shape=omega.shape, seed=seed,
                                       clustering=clustering, resolution=resolution, n_host_threads=n_host_threads,
                                       iter_begin=iter_begin, iter_end=iter_end, standard_scaling=standard_scaling))
        n_iters, n_synth = parents.shape[:2]
        n = self.n_cells
        iter_end = n_iters if iter_end is None else iter_end
        lib = np.asarray(self.raw.sum(axis=1)).ravel()
        normed = self.raw.copy()
        from sklearn.utils.sparsefuncs_fast import inplace_csr_row_normalize_l1

        inplace_csr_row_normalize_l1(normed)
        if out is not None:  # the pipelined wrapper's shared result arrays
            out = dict(scores=out[0], log_p=out[1], communities=out[2], synth_communities=out[3][:, :n_synth], stage_ms={"wall": 0.0})
        else:
            out = dict(scores=np.zeros((n_iters, n)), log_p=np.zeros((n_iters, n)), communities=np.zeros((n_iters, n), np.int32),
                       synth_communities=np.zeros((n_iters, n_synth), np.int32), stage_ms={"wall": 0.0})
        for i in range(iter_begin, iter_end):
            synth = reference_path.create_doublets(self.raw, parents[i])
            aug, _, _ = reference_path.normalise(synth, lib, normed, pseudocount)
            if standard_scaling:
                aug, _, _ = upstream.pp_scale(aug, max_value=scale_max_value)
            emb, _ = upstream.tl_pca(aug, n_comp, random_state=seed, svd_solver="auto")
            if clustering == "phenograph":
                full, _ = upstream.phenograph_cluster(emb, k=pheno_k, prune=pheno_prune, min_cluster_size=pheno_min_cluster_size,
                                                      seed=seed, louvain_fn=louvain_c.louvain)
            else:
                idx, dist = upstream.knn_brute(emb, knn_k)
                if clustering == "leiden":
                    from oracle import leiden_ref

                    C = upstream.fuzzy_connectivities(idx, dist)
                    full = leiden_ref.leiden(C.indptr, C.indices, C.data.astype(np.float64), resolution=resolution, seed=seed)
                else:
                    S = upstream.knn_pattern_graph(idx)
                    full = louvain_c.louvain(S.indptr, S.indices, None, resolution=resolution, seed=seed, level0="parallel")
            s, lp, comm, scomm = reference_path.score_communities(np.asarray(full), n)
            out["scores"][i], out["log_p"][i], out["communities"][i], out["synth_communities"][i] = s, lp, comm, scomm
        return out

    def close(self):
        pass


@pytest.fixture()
def shim(monkeypatch):
    from doubletdetection_b200 import _capi, classifier

    monkeypatch.setattr(_capi, "Handle", OracleHandle)
    monkeypatch.setenv("DD_PIPELINES", "1")  # the stand-in calls BLAS: keep it on one thread (see test_gpu_test_logic_dryrun.py)
    OracleHandle.calls = []
    return classifier.BoostClassifier


@pytest.mark.parametrize("name", ["c1_louvain", "c1_louvain_scaled", "hvg_replace", "single_iter", "c1_phenograph_scaled",
                                  "c1_leiden_scaled", "structured_900x200_phenograph"])
def test_shim_orchestration_reproduces_the_reference_goldens(shim, name):
    g = load_golden(name)
    counts, kw, pkw = golden_case(name)
    kw = dict(kw, clustering_kwargs=dict(kw.get("clustering_kwargs") or {}))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = shim(**kw)
        labels = np.asarray(clf.fit(counts).predict(**pkw), dtype=np.float64)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), g["parents"])
    if "top_var_genes" in g:
        np.testing.assert_array_equal(clf.top_var_genes_, g["top_var_genes"])
    np.testing.assert_array_equal(clf.communities_, g["communities"])
    np.testing.assert_array_equal(clf.synth_communities_, g["synth_communities"])
    np.testing.assert_array_equal(clf.all_scores_, g["all_scores"])
    np.testing.assert_allclose(clf.all_log_p_values_, g["all_log_p_values"], rtol=1e-12, atol=0, equal_nan=True)
    np.testing.assert_array_equal(labels, g["labels"])
    assert clf.communities_.dtype == np.float64 and clf.synth_communities_.dtype == np.float64  # :188-190
    assert isinstance(clf.parents_, list) and isinstance(clf.parents_[0], list) and len(clf.parents_[0][0]) == 2
    call = OracleHandle.calls[-1]
    n_aug = counts.shape[0] + g["parents"].shape[1]
    n_genes = kw.get("n_top_var_genes", 10000)
    n_genes = min(n_genes, counts.shape[1]) if n_genes > 0 else counts.shape[1]
    assert call["omega_shape"] == (min(n_aug, n_genes) if n_aug < n_genes else n_genes, call["n_comp"] + 10)
    assert call["n_power_iter"] == (7 if call["n_comp"] < 0.1 * min(n_aug, n_genes) else 4)
    assert call["iter_begin"] == 0 and call["iter_end"] == kw["n_iters"] and call["clustering"] == kw["clustering_algorithm"]
    if kw["clustering_algorithm"] != "phenograph":
        assert call["resolution"] == 4.0


def test_shim_rng_stream_continues_across_fits_and_inputs_are_untouched(shim):
    counts, kw, _ = golden_case("c1_louvain")
    dense = np.asarray(counts.todense() if sp_sparse.issparse(counts) else counts).astype(np.int64)
    before = dense.copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = shim(**kw)
        first = np.asarray(clf.fit(dense).parents_, dtype=np.int64).copy()
        second = np.asarray(clf.fit(dense).parents_, dtype=np.int64)
        ora = reference_path.OracleClassifier(n_iters=kw["n_iters"], random_state=0, louvain_fn=louvain_c.louvain)
        ora.fit(dense)
        ora.fit(dense)
    np.testing.assert_array_equal(dense, before)
    assert not np.array_equal(first, second)  # quirk Q2: one stream for all fits
    np.testing.assert_array_equal(second, np.asarray(ora.parents_, dtype=np.int64))


@pytest.mark.parametrize("algo,ckw", [("phenograph", {}), ("phenograph", {"prune": False, "k": 12, "min_cluster_size": 5}),
                                      ("leiden", {}), ("leiden", {"resolution": 1.5}), ("louvain", {"resolution": 2})])
def test_shim_maps_clustering_arguments(shim, algo, ckw):
    from oracle import datasets

    counts = datasets.structured_counts(600, 150, seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf = shim(n_iters=2, clustering_algorithm=algo, clustering_kwargs=dict(ckw), random_state=3, n_jobs=-1).fit(counts)
        okw = {k_: v_ for k_, v_ in ckw.items()}
        ora = reference_path.OracleClassifier(n_iters=2, random_state=3, clustering_algorithm=algo, clustering_kwargs=okw,
                                              louvain_fn=louvain_c.louvain).fit(counts)
    np.testing.assert_array_equal(np.asarray(clf.parents_, dtype=np.int64), np.asarray(ora.parents_, dtype=np.int64))
    np.testing.assert_array_equal(clf.communities_, ora.communities_)
    np.testing.assert_array_equal(clf.all_scores_, ora.all_scores_)
    np.testing.assert_allclose(clf.all_log_p_values_, ora.all_log_p_values_, rtol=1e-12, equal_nan=True)
    call = OracleHandle.calls[-1]
    assert call["clustering"] == algo and call["seed"] == 3 and call["n_host_threads"] >= 1
    if algo != "phenograph":
        assert call["resolution"] == float(ckw.get("resolution", 4))


def test_shim_upload_errors_surface_from_the_worker_thread(shim, monkeypatch):
    def boom(self, csr):
        raise MemoryError("libdd_b200: cudaMalloc: out of memory")

    monkeypatch.setattr(OracleHandle, "upload_counts", boom)
    counts, kw, _ = golden_case("c1_louvain")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(MemoryError, match="out of memory"):
            shim(**kw).fit(counts)


# ---- sklearn's exact PCA branches (covariance_eigh / full): the host half of _exact_pca over a numpy stand-in for the
# two device entry points (dd_centered_gram / dd_project)
class _NumpyDense:
    def __init__(self, dense):
        self.dense = np.asarray(dense, dtype=np.float32)
        self._dense_rows, self.n_genes = self.dense.shape
        self.emb = None

    def _centred(self):
        d = self.dense.astype(np.float64)
        return d - d.mean(axis=0)

    def centered_gram(self, transposed=False):
        c = self._centred()
        return c @ c.T if transposed else c.T @ c

    def project(self, v):
        self.emb = (self._centred() @ v).astype(np.float32)
        return self.emb

    def download_dense(self):
        return self.dense

    def upload_embedding(self, emb):
        self.emb = np.asarray(emb, dtype=np.float32)


@pytest.mark.parametrize("shape,solver", [((2600, 120), "covariance_eigh"), ((400, 90), "full"), ((60, 300), "full"),
                                          ((36, 64), "full")])
def test_exact_pca_branches_match_sklearn(shape, solver):
    from sklearn.decomposition import PCA

    from doubletdetection_b200.classifier import _exact_pca, _pca_plan, _pca_solver
    from oracle import pca_f64

    rs = np.random.default_rng(shape[0])
    centres = rs.normal(size=(4, shape[1])) * 2.0
    X = (centres[rs.integers(0, 4, shape[0])] + rs.normal(size=shape)).astype(np.float32)
    c = 30
    assert _pca_solver(shape[0], shape[1], c) == solver == pca_f64.auto_solver(shape[0], shape[1], c)
    assert _pca_plan(shape[0], shape[1], c, 0) == (None, 0)
    h = _NumpyDense(X)
    emb = _exact_pca(h, c)
    truth, _ = pca_f64.exact_pca_f64(X, c)
    assert np.abs(emb - truth).max() / np.abs(truth).max() < 1e-6
    # sklearn's own run (float32 for the eigh / svd) stays within the 1e-4 band of the same truth on the leading components
    want = PCA(n_components=c, svd_solver="auto", random_state=0).fit_transform(X)
    lead = slice(0, 3)
    assert np.abs(want[:, lead] - truth[:, lead]).max() / np.abs(truth[:, lead]).max() < 1e-3
    np.testing.assert_array_equal(h.emb, emb)


def test_fit_iterations_pipelined_merges_the_pieces():
    """``_capi.fit_iterations_pipelined``: contiguous pieces of the iteration range, one per handle, run in threads; result rows
    merged, stage times summed except wall / device_total (longest loop); errors re-raised on the caller's thread."""
    from doubletdetection_b200 import _capi

    class Piece:
        def __init__(self, tag, fail=False):
            self.tag, self.fail, self.seen = tag, fail, None

        def fit_iterations(self, parents, omega, *, iter_begin, iter_end, n_host_threads, **kw):
            if self.fail:
                raise NotImplementedError("boom")
            self.seen = (iter_begin, iter_end, n_host_threads)
            n_iters = parents.shape[0]
            out = dict(scores=np.zeros((n_iters, 5)), log_p=np.zeros((n_iters, 5)), communities=np.zeros((n_iters, 5), np.int32),
                       synth_communities=np.zeros((n_iters, 2), np.int32),
                       stage_ms=dict(pca=1.0, wall=10.0 + self.tag, device_total=9.0 + self.tag))
            for key in ("scores", "log_p", "communities", "synth_communities"):
                out[key][iter_begin:iter_end] = self.tag + 1
            return out

    parents = np.zeros((7, 2, 2), dtype=np.int64)
    hs = [Piece(0), Piece(1), Piece(2)]
    out = _capi.fit_iterations_pipelined(hs, parents, None, iter_begin=1, iter_end=7, n_host_threads=7)
    assert [h.seen for h in hs] == [(1, 3, 2), (3, 5, 2), (5, 7, 2)]
    np.testing.assert_array_equal(out["communities"][:, 0], [0, 1, 1, 2, 2, 3, 3])
    np.testing.assert_array_equal(out["scores"][:, 0], [0, 1, 1, 2, 2, 3, 3])
    assert out["stage_ms"] == dict(pca=3.0, wall=12.0, device_total=11.0, pipelines=3.0)
    # fewer iterations than handles: one loop per iteration at most; a single handle takes the plain path
    out = _capi.fit_iterations_pipelined(hs, parents, None, iter_begin=2, iter_end=3, n_host_threads=4)
    assert hs[0].seen == (2, 3, 4) and "pipelines" not in out["stage_ms"]
    with pytest.raises(NotImplementedError):
        _capi.fit_iterations_pipelined([Piece(0), Piece(1, fail=True)], parents, None, n_host_threads=2)

    class SharedPiece(Piece):  # what Handle.fit_iterations does: the wrapper's ONE set of result arrays, own rows only
        n_cells = 5

        def fit_iterations(self, parents, omega, *, iter_begin, iter_end, n_host_threads, out, **kw):
            for a in out:
                a[iter_begin:iter_end] = self.tag + 1
            return dict(scores=out[0], log_p=out[1], communities=out[2], synth_communities=out[3][:, :2],
                        stage_ms=dict(pca=1.0, wall=10.0 + self.tag, device_total=9.0 + self.tag))

    out = _capi.fit_iterations_pipelined([SharedPiece(0), SharedPiece(1)], parents, None, iter_begin=1, iter_end=7, n_host_threads=2)
    np.testing.assert_array_equal(out["communities"][:, 0], [0, 1, 1, 1, 2, 2, 2])
    np.testing.assert_array_equal(out["log_p"][:, 4], [0, 1, 1, 1, 2, 2, 2])
    assert out["synth_communities"].shape == (7, 2) and out["scores"].shape == (7, 5)


def test_sparse_input_with_nan_raises_sklearns_error(shim):
    """Sparse input skips check_array's host-side finiteness scan (the device checks while it sums the rows); a matrix with
    NaN / inf must still end in sklearn's ValueError (doubletdetection.py:149-155)."""
    import scipy.sparse as sp_sparse

    x = sp_sparse.csr_matrix(np.random.default_rng(0).poisson(1.0, (600, 120)).astype(np.float32))
    x.data[17] = np.nan
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError, match="NaN"):
            shim(n_iters=2, clustering_algorithm="louvain").fit(x)


def test_more_components_than_the_matrix_allows_raises_sklearns_value_error(shim):
    """``sc.tl.pca(n_comps=...)`` ends in sklearn's range check (PCA._fit_full / _fit_truncated): more components than
    min(augmented cells, genes) is a ValueError in the reference, and for its sparse ``pseudocount == 1`` branch
    (``svd_solver="arpack"``) the bound is strict.  The shim raises before anything is uploaded."""
    counts = datasets.poisson_counts(300, 24, seed=3) + 1
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError, match="n_components=25"):
            shim(n_iters=2, clustering_algorithm="louvain", n_components=25).fit(counts)
        with pytest.raises(ValueError, match="strictly less"):
            shim(n_iters=2, clustering_algorithm="louvain", n_components=24, pseudocount=1).fit(counts)
    assert OracleHandle.calls == []
