"""fit()'s prologue on the device (hvg.cu; doubletdetection.py:165-176): ``gene_variances`` bit for bit what scipy computes
(float32, one accumulator per gene, rows in order), ``top_var_genes_`` identical to the reference's, the column subset
identical to ``raw.tocsc()[:, top].tocsr()`` and the library sizes of the subset (:182).  Needs a B200 (`-m gpu`)."""

import warnings

import numpy as np
import pytest
import scipy.sparse as sp_sparse

from oracle import datasets

pytestmark = pytest.mark.gpu


def _reference_prologue(raw, n_top):
    """doubletdetection.py:165-176, 182 verbatim (scipy / numpy)."""
    gene_variances = (np.array(raw.power(2).mean(axis=0)) - (np.array(raw.mean(axis=0))) ** 2)[0]
    top_var_indexes = np.argsort(gene_variances)
    top = top_var_indexes[-n_top:]
    sub = raw.tocsc()[:, top].tocsr()
    return gene_variances, top, sub, np.asarray(np.sum(sub, axis=1)).ravel()


def _cases():
    rs = np.random.default_rng(5)
    yield "poisson_450x400", sp_sparse.csr_matrix(
        (datasets.poisson_counts(450, 400, seed=7, lam=0.7) * (np.arange(400) % 5 + 1)[None, :]).astype(np.float32)), 150
    yield "structured_5000x3000", datasets.structured_counts(5000, 3000, seed=3).astype(np.float32), 1000
    # 30k genes, ragged: empty rows, empty genes, one gene expressed in every cell, non-integer values
    m = sp_sparse.random(20000, 30000, density=0.02, random_state=11, format="csr", dtype=np.float32)
    m.data = np.ceil(m.data * 30).astype(np.float32)
    m = m.tolil()
    m[:, 17] = 3.0
    m[5, :] = 0
    m = m.tocsr()
    m.eliminate_zeros()
    m.data[::7] += 0.25
    m.sort_indices()
    yield "sparse_20000x30000", m, 10000
    yield "tiny_7x9", sp_sparse.csr_matrix(rs.poisson(1.0, size=(7, 9)).astype(np.float32)), 4


@pytest.mark.parametrize("name", ["poisson_450x400", "structured_5000x3000", "sparse_20000x30000", "tiny_7x9"])
def test_hvg_variances_and_subset_bit_exact(handle, name):
    raw, n_top = next((m, t) for n, m, t in _cases() if n == name)
    raw = sp_sparse.csr_matrix(raw, dtype=np.float32)
    raw.sum_duplicates()
    want_var, want_top, want_sub, want_lib = _reference_prologue(raw, n_top)
    handle.upload_counts(raw)
    got_var = handle.hvg_variances()
    assert got_var.dtype == np.float32
    np.testing.assert_array_equal(got_var.view(np.uint32), want_var.view(np.uint32))  # bit for bit
    top = np.argsort(got_var)[-n_top:]
    np.testing.assert_array_equal(top, want_top)
    handle.select_genes(top)
    got = handle.download_counts()
    assert got.shape == want_sub.shape
    np.testing.assert_array_equal(got.indptr, want_sub.indptr)
    np.testing.assert_array_equal(got.indices, want_sub.indices)
    np.testing.assert_array_equal(got.data, want_sub.data)
    np.testing.assert_array_equal(handle.lib_size(), want_lib)


def test_select_genes_rejects_bad_lists(handle):
    raw = sp_sparse.csr_matrix(np.random.default_rng(0).poisson(1.0, size=(50, 20)).astype(np.float32))
    handle.upload_counts(raw)
    with pytest.raises(Exception):
        handle.select_genes([1, 1, 2])
    handle.upload_counts(raw)
    with pytest.raises(Exception):
        handle.select_genes([0, 20])


def test_classifier_hvg_on_device_equals_host_prologue(monkeypatch):
    """The whole fit with the prologue on the device against the same fit with the reference's scipy lines on the host
    (DD_HVG_HOST=1): identical top_var_genes_ and identical results."""
    from doubletdetection_b200 import BoostClassifier

    counts = datasets.structured_counts(3000, 2500, seed=21)
    kw = dict(n_iters=2, clustering_algorithm="louvain", n_top_var_genes=1200, random_state=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dev = BoostClassifier(**kw).fit(counts)
        monkeypatch.setenv("DD_HVG_HOST", "1")
        host = BoostClassifier(**kw).fit(counts)
    np.testing.assert_array_equal(dev.top_var_genes_, host.top_var_genes_)
    np.testing.assert_array_equal(dev.communities_, host.communities_)
    np.testing.assert_array_equal(dev.all_log_p_values_, host.all_log_p_values_)


def test_sparse_input_with_nan_or_inf_raises_like_check_array():
    """The finiteness check of check_array (:149-155) runs on the device for sparse input; the error is sklearn's."""
    from doubletdetection_b200 import BoostClassifier

    for bad in (np.nan, np.inf):
        x = sp_sparse.csr_matrix(np.random.default_rng(0).poisson(1.0, (600, 120)).astype(np.float32))
        x.data[1234] = bad
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with pytest.raises(ValueError):
                BoostClassifier(n_iters=2, clustering_algorithm="louvain").fit(x)
