"""Load the UNMODIFIED reference module from /root/reference on top of stub modules for the
packages that are absent from the image (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Only usable in the build container (``/root/reference`` does not exist on the GPU box); used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the not-gpu test
that re-checks them when the reference is present.  The stubs implement exactly the calls the
reference makes (doubletdetection.py:162-163, 208, 300-303, 309-314, 320, 331-343) by delegating
to ``oracle.upstream``.
"""

import importlib.util
import os
import sys
import types

import numpy as np

from . import upstream

REFERENCE_FILE = "/root/reference/doubletdetection/doubletdetection.py"


def reference_available():
    return os.path.exists(REFERENCE_FILE)


def _make_stub_modules(louvain_fn=None, record=None, phenograph_seed=0):
    """``record``: optional dict; every stub call appends what it saw / produced."""

    def rec(key, value):
        if record is not None:
            record.setdefault(key, []).append(value)

    anndata = types.ModuleType("anndata")
    anndata.AnnData = upstream.AnnDataLite

    sc = types.ModuleType("scanpy")
    sc.settings = types.SimpleNamespace(n_jobs=1)
    pp = types.SimpleNamespace()
    tl = types.SimpleNamespace()

    def scale(adata, max_value=None):
        X = adata.X
        if hasattr(X, "toarray"):  # sc.pp.scale(zero_center=True) densifies sparse input (the pseudocount == 1 branch)
            X = np.asarray(X.toarray(), dtype=np.float32)
        adata.X, _, _ = upstream.pp_scale(X, max_value=max_value)
        rec("scaled", adata.X)

    def pca(adata, n_comps=None, random_state=0, svd_solver="auto"):
        rec("pca_input", adata.X)
        rec("n_counts", adata.obs.get("n_counts"))
        adata.obsm["X_pca"], _ = upstream.tl_pca(adata.X, n_comps, random_state=random_state, svd_solver=svd_solver)
        rec("X_pca", adata.obsm["X_pca"])

    def neighbors(adata, random_state=0, method="umap", n_neighbors=10):
        upstream.pp_neighbors(adata, random_state=random_state, method=method, n_neighbors=n_neighbors)
        rec("knn_indices", adata.uns["knn_indices"])
        rec("knn_distances", adata.uns["knn_distances"])

    def louvain(adata, key_added="louvain", random_state=0, **kw):
        upstream.tl_louvain(adata, key_added=key_added, random_state=random_state, louvain_fn=louvain_fn, **kw)

    def leiden(adata, key_added="leiden", random_state=0, **kw):
        # sc.pp.neighbors always builds the umap-weighted connectivities; only sc.tl.leiden uses the weights, so the
        # neighbors stub keeps the pattern (what tl.louvain sees) and the weights are derived here from the same lists
        adata.obsp["connectivities"] = upstream.fuzzy_connectivities(adata.uns["knn_indices"], adata.uns["knn_distances"])
        rec("connectivities", adata.obsp["connectivities"])
        upstream.tl_leiden(adata, key_added=key_added, random_state=random_state, **kw)

    pp.scale, pp.neighbors = scale, neighbors
    tl.pca, tl.louvain, tl.leiden = pca, louvain, leiden
    sc.pp, sc.tl = pp, tl

    phenograph = types.ModuleType("phenograph")

    def cluster(data, n_jobs=1, **kw):
        # phenograph.cluster(X_pca, n_jobs=..., **clustering_kwargs) -> (communities, graph, Q); unseedable upstream, the
        # restatement fixes seed 0 (what the classifier's default random_state passes on the native path)
        rec("phenograph_kwargs", dict(kw, n_jobs=n_jobs))
        labels, graph = upstream.phenograph_cluster(data, seed=phenograph_seed, louvain_fn=louvain_fn, **kw)
        return labels, graph, None

    phenograph.cluster = cluster
    return {"anndata": anndata, "scanpy": sc, "phenograph": phenograph}


def load_reference(louvain_fn=None, record=None, phenograph_seed=0):
    """Returns the reference's ``doubletdetection.doubletdetection`` module object (real code)."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_FILE)
    stubs = _make_stub_modules(louvain_fn, record, phenograph_seed)
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    try:
        spec = importlib.util.spec_from_file_location("_dd_reference_module", REFERENCE_FILE)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
