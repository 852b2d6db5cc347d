import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


# The seeded inputs of tests/golden/make_golden.py, re-stated so the tests can rebuild them anywhere.
def golden_case(name):
    from oracle import datasets

    c1 = datasets.poisson_counts(500, 100, seed=0)
    if name == "c1_louvain":
        return c1, dict(n_iters=3, clustering_algorithm="louvain"), dict(p_thresh=1e-16, voter_thresh=0.5)
    if name == "c1_louvain_scaled":
        return (c1, dict(n_iters=2, clustering_algorithm="louvain", standard_scaling=True),
                dict(p_thresh=1e-16, voter_thresh=0.5))
    if name == "hvg_replace":
        hv = datasets.poisson_counts(450, 400, seed=7, lam=0.7) * (np.arange(400) % 5 + 1)[None, :]
        return (hv, dict(n_iters=2, clustering_algorithm="louvain", n_top_var_genes=150, replace=True,
                         boost_rate=0.6, random_state=11), dict())
    if name == "single_iter":
        return c1, dict(n_iters=1, clustering_algorithm="louvain"), dict()
    if name == "structured_1500x300":
        st = datasets.structured_counts(1500, 300, seed=1234)
        return st, dict(n_iters=3, clustering_algorithm="louvain"), dict(p_thresh=1e-3, voter_thresh=0.5)
    # the phenograph / leiden branches of the reference's own test (tests/test_package.py:17-28) and a PhenoGraph case
    # with prune=False on structured counts
    if name == "c1_phenograph_scaled":
        return (c1, dict(n_iters=2, clustering_algorithm="phenograph", standard_scaling=True),
                dict(p_thresh=1e-16, voter_thresh=0.5))
    if name == "c1_leiden_scaled":
        return (c1, dict(n_iters=2, clustering_algorithm="leiden", standard_scaling=True, random_state=123),
                dict(p_thresh=1e-16, voter_thresh=0.5))
    if name == "structured_900x200_phenograph":
        st = datasets.structured_counts(900, 200, seed=77)
        return (st, dict(n_iters=2, clustering_algorithm="phenograph", clustering_kwargs={"prune": False}),
                dict(p_thresh=1e-3, voter_thresh=0.5))
    if name == "structured_1200x260_pc1":
        st = datasets.structured_counts(1200, 260, seed=5)
        return st, dict(n_iters=2, clustering_algorithm="louvain", pseudocount=1), dict(p_thresh=1e-3, voter_thresh=0.5)
    if name == "structured_1200x260_pc1_scaled":
        st = datasets.structured_counts(1200, 260, seed=5)
        return (st, dict(n_iters=2, clustering_algorithm="louvain", pseudocount=1, standard_scaling=True),
                dict(p_thresh=1e-3, voter_thresh=0.5))
    raise KeyError(name)


GOLDEN_NAMES = ["c1_louvain", "c1_louvain_scaled", "hvg_replace", "single_iter", "structured_1500x300"]
# goldens of the other two clustering branches (same generator: the reference's real code over the restated calls)
CLUSTER_GOLDEN_NAMES = ["c1_phenograph_scaled", "c1_leiden_scaled", "structured_900x200_phenograph"]
# the sparse branch (pseudocount == 1 -> np.log1p on the sparse matrix, svd_solver="arpack"; doubletdetection.py:296-297, 308)
SPARSE_GOLDEN_NAMES = ["structured_1200x260_pc1", "structured_1200x260_pc1_scaled"]


@pytest.fixture(scope="session")
def native():
    from doubletdetection_b200 import _capi

    _capi.load()
    return _capi


@pytest.fixture()
def handle(native):
    h = native.Handle(0)  # raises without a B200: GPU tests must not silently fall back
    yield h
    h.close()
