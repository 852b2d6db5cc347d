"""ctypes binding of ``libdd_b200.so`` (C ABI declared in ``include/dd_b200.h``).

There is no CPU fallback: importing this module without the built library raises ``ImportError``
and creating a :class:`Handle` without a B200 raises ``RuntimeError`` -- the product path never
routes around the CUDA extension.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdd_b200.so")

DD_OK, DD_ERR_ARG, DD_ERR_CUDA, DD_ERR_UNSUPPORTED, DD_ERR_NOMEM = 0, 1, 2, 3, 4
ABI_VERSION = 7
COMM_ID_BYTES = 128

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_f64p = ctypes.POINTER(ctypes.c_double)


class FitParams(ctypes.Structure):
    """``dd_fit_params`` of include/dd_b200.h."""

    _fields_ = [
        ("n_iters", ctypes.c_int32),
        ("n_synth", ctypes.c_int64),
        ("pseudocount", ctypes.c_float),
        ("standard_scaling", ctypes.c_int32),
        ("scale_max_value", ctypes.c_float),
        ("n_comp", ctypes.c_int32),
        ("n_random", ctypes.c_int32),
        ("n_power_iter", ctypes.c_int32),
        ("knn_k", ctypes.c_int32),
        ("resolution", ctypes.c_double),
        ("seed", ctypes.c_uint64),
        ("n_host_threads", ctypes.c_int32),
        ("iter_begin", ctypes.c_int32),
        ("iter_end", ctypes.c_int32),
        ("clustering", ctypes.c_int32),
        ("pheno_k", ctypes.c_int32),
        ("pheno_prune", ctypes.c_int32),
        ("pheno_min_cluster_size", ctypes.c_int32),
    ]


CLUSTER_LOUVAIN, CLUSTER_PHENOGRAPH, CLUSTER_LEIDEN = 0, 1, 2


# name -> (restype, argtypes); every symbol include/dd_b200.h declares
SIGNATURES = {
    "dd_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "dd_destroy": (None, [ctypes.c_void_p]),
    "dd_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "dd_abi_version": (ctypes.c_int, []),
    "dd_upload_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, c_i32p, c_i32p, c_f32p]),
    "dd_get_lib_size": (ctypes.c_int, [ctypes.c_void_p, c_f32p]),
    "dd_share_counts": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "dd_counts_all_finite": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]),
    "dd_hvg_variances": (ctypes.c_int, [ctypes.c_void_p, c_f32p]),
    "dd_select_genes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_i64p]),
    "dd_counts_nnz": (ctypes.c_int, [ctypes.c_void_p, c_i64p]),
    "dd_download_counts": (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i32p, c_f32p]),
    "dd_create_doublets": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, c_i64p]),
    "dd_synth_nnz": (ctypes.c_int, [ctypes.c_void_p, c_i64p]),
    "dd_download_synthetics": (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i32p, c_f32p]),
    "dd_get_synth_lib_size": (ctypes.c_int, [ctypes.c_void_p, c_f32p]),
    "dd_median_lib_size": (ctypes.c_int, [ctypes.c_void_p, c_f32p]),
    "dd_normalise_log": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float, ctypes.c_float]),
    "dd_standard_scale": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_float]),
    "dd_download_dense": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, c_f32p]),
    "dd_upload_dense": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, c_f32p]),
    "dd_pca": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, c_f32p, c_f32p, c_f64p],
    ),
    "dd_centered_gram": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_f64p]),
    "dd_project": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_f64p, c_f32p]),
    "dd_upload_embedding": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, c_f32p]),
    "dd_knn": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, c_i32p, c_f32p]),
    "dd_knn_listed": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_f32p]),
    "dd_set_knn_mode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "dd_knn_clustered_stats": (ctypes.c_int, [ctypes.c_void_p, c_i64p]),
    "dd_knn_uncertified": (ctypes.c_int, [ctypes.c_void_p, c_i64p]),
    "dd_knn_pruned": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_f32p, c_i64p]),
    "dd_louvain_knn": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int32, c_i32p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_phenograph_knn": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int32, c_i32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_jaccard_graph": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, c_i32p, c_i32p, c_f64p, ctypes.c_int64, c_i64p],
    ),
    "dd_louvain_csr": (
        ctypes.c_int,
        [ctypes.c_int64, c_i64p, c_i64p, c_f64p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_umap_connectivities": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int32, c_i32p, c_f32p, c_i64p, c_i32p, c_f32p, ctypes.c_int64, c_i64p],
    ),
    "dd_umap_graph": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int32, c_i64p, c_i32p, c_f32p, ctypes.c_int64, c_i64p],
    ),
    "dd_leiden_device_graph": (
        ctypes.c_int,
        [ctypes.c_int64, c_i32p, c_i32p, c_f64p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_leiden_knn": (
        ctypes.c_int,
        [ctypes.c_int64, ctypes.c_int32, c_i32p, c_f32p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_leiden_csr": (
        ctypes.c_int,
        [ctypes.c_int64, c_i64p, c_i64p, c_f64p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_louvain_level0_weighted": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.c_int64, c_i64p, c_i64p, c_f64p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_louvain_csr_level0": (
        ctypes.c_int,
        [ctypes.c_int64, c_i64p, c_i64p, c_f64p, ctypes.c_double, ctypes.c_uint64, c_i32p, c_i32p],
    ),
    "dd_score": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int64, c_i32p, c_f64p, c_f64p]),
    "dd_hypergeom_logsf": (ctypes.c_double, [ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]),
    "dd_fit_iterations": (
        ctypes.c_int,
        [ctypes.c_void_p, ctypes.POINTER(FitParams), c_i64p, c_f32p, c_f64p, c_f64p, c_i32p, c_i32p, c_f64p],
    ),
    "dd_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64]),
    "dd_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64]),
    "dd_comm_shard_cells": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "dd_comm_info": (ctypes.c_int, [ctypes.c_void_p, c_i32p, c_i32p, c_i64p]),
    "dd_kernel_launches": (ctypes.c_int64, [ctypes.c_void_p]),
    "dd_last_stage_ms": (ctypes.c_double, [ctypes.c_void_p, ctypes.c_char_p]),
    "dd_set_kernel_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int32]),
    "dd_get_kernel_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, c_f64p, c_i64p]),
    "dd_kernel_timing_report": (ctypes.c_int64, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64]),
}

_lib = None


def load():
    """Load the shared library (once) and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C doubletdetection_b200/csrc`). doubletdetection_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.dd_abi_version() != ABI_VERSION:
        raise ImportError(f"libdd_b200.so ABI {lib.dd_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctype))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdd_b200 error {code}: {msg}")
        self.code = code


def _raise(lib, h, rc):
    msg = lib.dd_last_error(h)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == DD_ERR_UNSUPPORTED:
        raise NotImplementedError(f"libdd_b200: {msg}")
    if rc == DD_ERR_ARG:
        raise ValueError(f"libdd_b200: {msg}")
    if rc == DD_ERR_NOMEM:
        raise MemoryError(f"libdd_b200: {msg}")
    raise NativeError(rc, msg)


# ---- handle-free host entry points --------------------------------------------------------------
def louvain_knn(knn_idx, resolution=4.0, seed=0):
    lib = load()
    knn_idx = np.ascontiguousarray(knn_idx, dtype=np.int32)
    n, k = knn_idx.shape
    labels = np.empty(n, dtype=np.int32)
    ncomm = ctypes.c_int32(0)
    rc = lib.dd_louvain_knn(n, k, _ptr(knn_idx, ctypes.c_int32), float(resolution), int(seed) & (2**64 - 1),
                            _ptr(labels, ctypes.c_int32), ctypes.byref(ncomm))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return labels


def phenograph_knn(knn_idx, prune=True, min_cluster_size=10, seed=0):
    """PhenoGraph labels from exact kNN lists with self in column 0 (host twin of the pipelined device path)."""
    lib = load()
    knn_idx = np.ascontiguousarray(knn_idx, dtype=np.int32)
    n, k = knn_idx.shape
    labels = np.empty(max(n, 1), dtype=np.int32)
    ncomm = ctypes.c_int32(0)
    rc = lib.dd_phenograph_knn(n, k, _ptr(knn_idx, ctypes.c_int32), int(bool(prune)), int(min_cluster_size),
                               int(seed) & (2**64 - 1), _ptr(labels, ctypes.c_int32), ctypes.byref(ncomm))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return labels[:n]


def louvain_csr(indptr, indices, weights=None, resolution=1.0, seed=0, level0="sequential"):
    """``level0="parallel"``: first level by synchronous coloured rounds (fixed-point weights), as the kNN pipeline."""
    lib = load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    n = indptr.size - 1
    labels = np.empty(max(n, 1), dtype=np.int32)
    ncomm = ctypes.c_int32(0)
    fn = lib.dd_louvain_csr_level0 if level0 == "parallel" else lib.dd_louvain_csr
    rc = fn(n, _ptr(indptr, ctypes.c_int64), _ptr(indices, ctypes.c_int64), _ptr(w, ctypes.c_double),
            float(resolution), int(seed) & (2**64 - 1), _ptr(labels, ctypes.c_int32), ctypes.byref(ncomm))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return labels[:n]


def umap_connectivities(knn_idx, knn_dist):
    """umap's fuzzy simplicial set of exact kNN lists (self in column 0) as scipy CSR (float32, sorted rows)."""
    import scipy.sparse as sp_sparse

    lib = load()
    knn_idx = np.ascontiguousarray(knn_idx, dtype=np.int32)
    knn_dist = _f32(knn_dist)
    n, k = knn_idx.shape
    nnz = ctypes.c_int64(0)
    rc = lib.dd_umap_connectivities(n, k, _ptr(knn_idx, ctypes.c_int32), _ptr(knn_dist, ctypes.c_float), None, None, None,
                                    0, ctypes.byref(nnz))
    if rc != DD_OK:
        _raise(lib, None, rc)
    indptr = np.zeros(n + 1, dtype=np.int64)
    indices = np.empty(max(nnz.value, 1), dtype=np.int32)
    weights = np.empty(max(nnz.value, 1), dtype=np.float32)
    rc = lib.dd_umap_connectivities(n, k, _ptr(knn_idx, ctypes.c_int32), _ptr(knn_dist, ctypes.c_float),
                                    _ptr(indptr, ctypes.c_int64), _ptr(indices, ctypes.c_int32),
                                    _ptr(weights, ctypes.c_float), max(nnz.value, 1), ctypes.byref(nnz))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return sp_sparse.csr_matrix((weights[: nnz.value], indices[: nnz.value], indptr), shape=(n, n))


def leiden_knn(knn_idx, knn_dist, resolution=4.0, seed=0):
    """Leiden labels of the umap-weighted neighbour graph of exact kNN lists (what the fit loop's host workers run
    for ``clustering="leiden"``)."""
    lib = load()
    knn_idx = np.ascontiguousarray(knn_idx, dtype=np.int32)
    knn_dist = _f32(knn_dist)
    n, k = knn_idx.shape
    if knn_dist.shape != knn_idx.shape:
        raise ValueError("knn_idx and knn_dist must have the same shape")
    labels = np.empty(max(n, 1), dtype=np.int32)
    ncomm = ctypes.c_int32(0)
    rc = lib.dd_leiden_knn(n, k, _ptr(knn_idx, ctypes.c_int32), _ptr(knn_dist, ctypes.c_float), float(resolution),
                           int(seed) & (2**64 - 1), _ptr(labels, ctypes.c_int32), ctypes.byref(ncomm))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return labels[:n]


def leiden_csr(indptr, indices, weights=None, resolution=1.0, seed=0):
    lib = load()
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    n = indptr.size - 1
    labels = np.empty(max(n, 1), dtype=np.int32)
    ncomm = ctypes.c_int32(0)
    rc = lib.dd_leiden_csr(n, _ptr(indptr, ctypes.c_int64), _ptr(indices, ctypes.c_int64), _ptr(w, ctypes.c_double),
                           float(resolution), int(seed) & (2**64 - 1), _ptr(labels, ctypes.c_int32), ctypes.byref(ncomm))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return labels[:n]


def leiden_device_graph(off, adj, weights, resolution=4.0, seed=0):
    """Leiden labels of a graph in the layout the device leaves for the fit loop's host workers (int32 offsets, rows in any
    order, float64 weights with 0 = no edge)."""
    lib = load()
    off = np.ascontiguousarray(off, dtype=np.int32)
    adj = np.ascontiguousarray(adj, dtype=np.int32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    n = off.size - 1
    labels = np.empty(max(n, 1), dtype=np.int32)
    ncomm = ctypes.c_int32(0)
    rc = lib.dd_leiden_device_graph(n, _ptr(off, ctypes.c_int32), _ptr(adj, ctypes.c_int32), _ptr(w, ctypes.c_double),
                                    float(resolution), int(seed) & (2**64 - 1), _ptr(labels, ctypes.c_int32), ctypes.byref(ncomm))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return labels[:n]


def score(labels, n_cells):
    lib = load()
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    n_synth = labels.size - n_cells
    scores = np.empty(n_cells, dtype=np.float64)
    logp = np.empty(n_cells, dtype=np.float64)
    rc = lib.dd_score(n_cells, n_synth, _ptr(labels, ctypes.c_int32), _ptr(scores, ctypes.c_double),
                      _ptr(logp, ctypes.c_double))
    if rc != DD_OK:
        _raise(lib, None, rc)
    return scores, logp


def hypergeom_logsf(k, M, n, N):
    return load().dd_hypergeom_logsf(int(k), int(M), int(n), int(N))


def comm_unique_id():
    """Rendezvous token of a new NCCL communicator (call on rank 0, ship the bytes to every rank)."""
    lib = load()
    buf = ctypes.create_string_buffer(COMM_ID_BYTES)
    rc = lib.dd_comm_unique_id(buf, COMM_ID_BYTES)
    if rc != DD_OK:
        _raise(lib, None, rc)
    return buf.raw


def block_of(n, rank, world):
    """The [begin, end) share of ``n`` items rank ``rank`` of ``world`` owns (same rule as the library)."""
    return n * rank // world, n * (rank + 1) // world


# ---- device handle ------------------------------------------------------------------------------
class Handle:
    """One ``dd_handle``: a CUDA device, its stream and the resident matrices of one fit."""

    def __init__(self, device=0):
        self._lib = load()
        self._h = ctypes.c_void_p()
        rc = self._lib.dd_create(int(device), ctypes.byref(self._h))
        if rc != DD_OK:
            msg = self._lib.dd_last_error(None)
            raise RuntimeError(
                "libdd_b200: " + (msg.decode("utf-8", "replace") if msg else f"dd_create failed ({rc})")
            )
        self.device = int(device)
        self.n_cells = self.n_genes = self.n_synth = 0
        self._dense_rows = self._emb_rows = 0
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dd_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc):
        if rc != DD_OK:
            _raise(self._lib, self._h, rc)

    # fit prologue
    def upload_counts(self, csr):
        # the device CSR is int32-indexed: refuse what would wrap instead of uploading garbage
        if int(csr.indptr[-1]) >= 2**31 or max(csr.shape) >= 2**31:
            raise ValueError(f"count matrix too large for the int32-indexed device CSR (nnz={int(csr.indptr[-1])}, shape={csr.shape})")
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int32)
        indices = np.ascontiguousarray(csr.indices, dtype=np.int32)
        data = _f32(csr.data)
        self.n_cells, self.n_genes = csr.shape
        self._check(self._lib.dd_upload_counts(self._h, self.n_cells, self.n_genes, _ptr(indptr, ctypes.c_int32),
                                               _ptr(indices, ctypes.c_int32), _ptr(data, ctypes.c_float)))

    def counts_all_finite(self):
        out = ctypes.c_int32(1)
        self._check(self._lib.dd_counts_all_finite(self._h, ctypes.byref(out)))
        return bool(out.value)

    def share_counts(self, src):
        """Work on ``src``'s resident count matrix (another handle on the same GPU) instead of uploading a copy."""
        self._check(self._lib.dd_share_counts(self._h, src._h))
        self.n_cells, self.n_genes = src.n_cells, src.n_genes
        self._counts_owner = src  # keeps the owner alive

    def hvg_variances(self):
        """``gene_variances`` of fit()'s prologue (:166-169), float32, computed on the device in scipy's accumulation order."""
        out = np.empty(self.n_genes, dtype=np.float32)
        self._check(self._lib.dd_hvg_variances(self._h, _ptr(out, ctypes.c_float)))
        return out

    def select_genes(self, genes):
        """``raw.tocsc()[:, genes].tocsr()`` (:173-175) on the device; the handle then holds the subset."""
        genes = np.ascontiguousarray(genes, dtype=np.int64)
        self._check(self._lib.dd_select_genes(self._h, genes.size, _ptr(genes, ctypes.c_int64)))
        self.n_genes = int(genes.size)

    def download_counts(self):
        """The CSR the handle holds (after ``select_genes``: the subset) -- inspection / tests."""
        import scipy.sparse as sp_sparse

        nnz = ctypes.c_int64(0)
        self._check(self._lib.dd_counts_nnz(self._h, ctypes.byref(nnz)))
        indptr = np.empty(self.n_cells + 1, dtype=np.int32)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float32)
        self._check(self._lib.dd_download_counts(self._h, _ptr(indptr, ctypes.c_int32), _ptr(indices, ctypes.c_int32),
                                                 _ptr(data, ctypes.c_float)))
        return sp_sparse.csr_matrix((data, indices, indptr), shape=(self.n_cells, self.n_genes))

    def lib_size(self):
        out = np.empty(self.n_cells, dtype=np.float32)
        self._check(self._lib.dd_get_lib_size(self._h, _ptr(out, ctypes.c_float)))
        return out

    # _createDoublets
    def create_doublets(self, parents):
        parents = np.ascontiguousarray(parents, dtype=np.int64).reshape(-1, 2)
        self.n_synth = parents.shape[0]
        self._check(self._lib.dd_create_doublets(self._h, self.n_synth, _ptr(parents, ctypes.c_int64)))

    def download_synthetics(self):
        import scipy.sparse as sp_sparse

        nnz = ctypes.c_int64(0)
        self._check(self._lib.dd_synth_nnz(self._h, ctypes.byref(nnz)))
        indptr = np.empty(self.n_synth + 1, dtype=np.int32)
        indices = np.empty(nnz.value, dtype=np.int32)
        data = np.empty(nnz.value, dtype=np.float32)
        self._check(self._lib.dd_download_synthetics(self._h, _ptr(indptr, ctypes.c_int32), _ptr(indices, ctypes.c_int32),
                                                     _ptr(data, ctypes.c_float)))
        return sp_sparse.csr_matrix((data, indices, indptr), shape=(self.n_synth, self.n_genes))

    def synth_lib_size(self):
        out = np.empty(self.n_synth, dtype=np.float32)
        self._check(self._lib.dd_get_synth_lib_size(self._h, _ptr(out, ctypes.c_float)))
        return out

    def median_lib_size(self):
        out = ctypes.c_float(0)
        self._check(self._lib.dd_median_lib_size(self._h, ctypes.byref(out)))
        return np.float32(out.value)

    # normalise
    def normalise_log(self, median, pseudocount):
        self._check(self._lib.dd_normalise_log(self._h, float(median), float(pseudocount)))
        self._dense_rows = self.n_cells + self.n_synth

    def standard_scale(self, max_value=15.0):
        self._check(self._lib.dd_standard_scale(self._h, float(max_value)))

    def download_dense(self, row0=0, n_rows=None):
        if n_rows is None:
            n_rows = self.n_cells + self.n_synth - row0
        out = np.empty((n_rows, self.n_genes), dtype=np.float32)
        self._check(self._lib.dd_download_dense(self._h, row0, n_rows, _ptr(out, ctypes.c_float)))
        return out

    def upload_dense(self, dense):
        dense = _f32(dense)
        if self.n_cells == 0:
            self.n_cells, self.n_genes = dense.shape
            self.n_synth = 0
        self._check(self._lib.dd_upload_dense(self._h, dense.shape[0], dense.shape[1], _ptr(dense, ctypes.c_float)))
        self._dense_rows = dense.shape[0]

    # pca / knn
    def pca(self, n_comp, omega, n_power_iter, n_rows=None):
        omega = _f32(omega)
        n_rows = n_rows or self._dense_rows
        emb = np.empty((n_rows, n_comp), dtype=np.float32)
        sv = np.empty(n_comp, dtype=np.float64)
        self._check(self._lib.dd_pca(self._h, n_comp, omega.shape[1], n_power_iter, _ptr(omega, ctypes.c_float),
                                     _ptr(emb, ctypes.c_float), _ptr(sv, ctypes.c_double)))
        self._emb_rows = n_rows
        return emb, sv

    def centered_gram(self, transposed=False):
        """Float64 Gram matrix of the centred dense matrix: G x G (default) or A x A (transposed)."""
        n = (self._dense_rows if transposed else self.n_genes)
        out = np.empty((n, n), dtype=np.float64)
        self._check(self._lib.dd_centered_gram(self._h, int(bool(transposed)), _ptr(out, ctypes.c_double)))
        return out

    def project(self, components):
        """X_pca = (D - mean) V for float64 components V (G x n_comp); the embedding stays on the device for knn()."""
        v = np.ascontiguousarray(components, dtype=np.float64)
        n_comp = v.shape[1]
        emb = np.empty((self._dense_rows, n_comp), dtype=np.float32)
        self._check(self._lib.dd_project(self._h, n_comp, _ptr(v, ctypes.c_double), _ptr(emb, ctypes.c_float)))
        self._emb_rows = self._dense_rows
        return emb

    def upload_embedding(self, emb):
        emb = _f32(emb)
        self._check(self._lib.dd_upload_embedding(self._h, emb.shape[0], emb.shape[1], _ptr(emb, ctypes.c_float)))
        self._emb_rows = emb.shape[0]

    def knn(self, k=10, with_dist=True):
        n = self._emb_rows
        idx = np.empty((n, k), dtype=np.int32)
        dist = np.empty((n, k), dtype=np.float32) if with_dist else None
        self._check(self._lib.dd_knn(self._h, k, _ptr(idx, ctypes.c_int32), _ptr(dist, ctypes.c_float)))
        return idx, dist

    def set_knn_mode(self, mode):
        """0 = cluster-ordered kNN for large embeddings (default), 1 = always the all-tiles kernel, 2 = always cluster-ordered."""
        self._check(self._lib.dd_set_knn_mode(self._h, int(mode)))

    def knn_uncertified(self):
        """Rows of the last kNN call that the filter's certificate could not clear (re-done by float64 brute force)."""
        out = ctypes.c_int64(0)
        self._check(self._lib.dd_knn_uncertified(self._h, ctypes.byref(out)))
        return int(out.value)

    def knn_clustered_stats(self):
        out = np.zeros(4, dtype=np.int64)
        self._check(self._lib.dd_knn_clustered_stats(self._h, _ptr(out, ctypes.c_int64)))
        return dict(pairs_a=int(out[0]), pairs_b=int(out[1]), blocks=int(out[2]), tiles=int(out[3]))

    def knn_listed(self, k, list_off, list_tiles, with_dist=True):
        """Test hook: exact kNN in which 256-row query block p only visits the 128-row candidate tiles
        ``list_tiles[list_off[p]:list_off[p + 1]]`` (scripts/knn_listed_experiment.py)."""
        n = self._emb_rows
        list_off = np.ascontiguousarray(list_off, dtype=np.int32)
        list_tiles = np.ascontiguousarray(list_tiles, dtype=np.int32)
        idx = np.empty((n, k), dtype=np.int32)
        dist = np.empty((n, k), dtype=np.float32) if with_dist else None
        self._check(self._lib.dd_knn_listed(self._h, k, list_off.size - 1, _ptr(list_off, ctypes.c_int32),
                                            _ptr(list_tiles, ctypes.c_int32), _ptr(idx, ctypes.c_int32),
                                            _ptr(dist, ctypes.c_float)))
        return idx, dist

    def louvain_level0_weighted(self, indptr, indices, weights, resolution=1.0, seed=0):
        """Test hook: weighted first Louvain level on the device (fixed-point weights).  Returns (comm, rounds)."""
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int64)
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        n = indptr.size - 1
        comm = np.empty(n, dtype=np.int32)
        rounds = ctypes.c_int32(0)
        self._check(self._lib.dd_louvain_level0_weighted(self._h, n, _ptr(indptr, ctypes.c_int64), _ptr(indices, ctypes.c_int64),
                                                         _ptr(weights, ctypes.c_double), float(resolution),
                                                         int(seed) & (2**64 - 1), _ptr(comm, ctypes.c_int32), ctypes.byref(rounds)))
        return comm, rounds.value

    def knn_pruned(self, k, perm, block_group, n_rows, with_dist=True):
        """Test hook: cluster-ordered exact kNN, pre-pass and both launches on the device.  ``perm``: padded position ->
        original row or -1; ``block_group``: group of every 256-row block.  Returns (idx, dist, stats)."""
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        block_group = np.ascontiguousarray(block_group, dtype=np.int32)
        idx = np.empty((n_rows, k), dtype=np.int32)
        dist = np.empty((n_rows, k), dtype=np.float32) if with_dist else None
        stats = np.zeros(4, dtype=np.int64)
        self._check(self._lib.dd_knn_pruned(self._h, k, perm.size, _ptr(perm, ctypes.c_int32), _ptr(block_group, ctypes.c_int32),
                                            _ptr(idx, ctypes.c_int32), _ptr(dist, ctypes.c_float), _ptr(stats, ctypes.c_int64)))
        return idx, dist, dict(zip(("pairs_a", "pairs_b", "blocks", "tiles"), stats.tolist()))

    def jaccard_graph(self, k, prune=True):
        """PhenoGraph graph of the last ``knn(k)`` as built on the device: scipy CSR (float64, sorted rows, pruned
        zero-weight entries dropped)."""
        import scipy.sparse as sp_sparse

        n = self._emb_rows
        indptr = np.empty(n + 1, dtype=np.int32)
        nnz = ctypes.c_int64(0)
        self._check(self._lib.dd_jaccard_graph(self._h, int(k), int(bool(prune)), _ptr(indptr, ctypes.c_int32), None, None,
                                               0, ctypes.byref(nnz)))
        indices = np.empty(max(nnz.value, 1), dtype=np.int32)
        weights = np.empty(max(nnz.value, 1), dtype=np.float64)
        self._check(self._lib.dd_jaccard_graph(self._h, int(k), int(bool(prune)), _ptr(indptr, ctypes.c_int32),
                                               _ptr(indices, ctypes.c_int32), _ptr(weights, ctypes.c_double), nnz.value,
                                               ctypes.byref(nnz)))
        g = sp_sparse.csr_matrix((weights[: nnz.value], indices[: nnz.value], indptr), shape=(n, n))
        g.eliminate_zeros()
        g.sort_indices()
        return g

    def umap_graph(self, k):
        """umap's connectivities of the last ``knn(k)`` (lists + distances) as built on the device: scipy CSR (float32,
        sorted rows, zeros dropped) -- the device twin of :func:`umap_connectivities`."""
        import scipy.sparse as sp_sparse

        n = self._emb_rows
        nnz = ctypes.c_int64(0)
        self._check(self._lib.dd_umap_graph(self._h, int(k), None, None, None, 0, ctypes.byref(nnz)))
        indptr = np.zeros(n + 1, dtype=np.int64)
        indices = np.empty(max(nnz.value, 1), dtype=np.int32)
        weights = np.empty(max(nnz.value, 1), dtype=np.float32)
        self._check(self._lib.dd_umap_graph(self._h, int(k), _ptr(indptr, ctypes.c_int64), _ptr(indices, ctypes.c_int32),
                                            _ptr(weights, ctypes.c_float), max(nnz.value, 1), ctypes.byref(nnz)))
        return sp_sparse.csr_matrix((weights[: nnz.value], indices[: nnz.value], indptr), shape=(n, n))

    # the loop
    def fit_iterations(self, parents, omega, *, pseudocount, standard_scaling, n_comp, n_power_iter, knn_k=10,
                       resolution=4.0, seed=0, n_host_threads=1, iter_begin=0, iter_end=None, scale_max_value=15.0,
                       clustering="louvain", pheno_k=30, pheno_prune=True, pheno_min_cluster_size=10, out=None):
        """``out``: result arrays of a previous / concurrent call on the same parents (``fit_iterations_pipelined``): the
        loop writes the rows of ITS iterations into them instead of allocating its own."""
        parents = np.ascontiguousarray(parents, dtype=np.int64)
        n_iters, n_synth = parents.shape[0], parents.shape[1]
        omega = _f32(omega)
        iter_end = n_iters if iter_end is None else iter_end
        p = FitParams(n_iters, n_synth, float(pseudocount), int(bool(standard_scaling)), float(scale_max_value),
                      int(n_comp), int(omega.shape[1]), int(n_power_iter), int(knn_k), float(resolution),
                      int(seed) & (2**64 - 1), int(n_host_threads), int(iter_begin), int(iter_end),
                      {"louvain": CLUSTER_LOUVAIN, "phenograph": CLUSTER_PHENOGRAPH, "leiden": CLUSTER_LEIDEN}[clustering], int(pheno_k),
                      int(bool(pheno_prune)), int(pheno_min_cluster_size))
        N = self.n_cells
        if out is None:
            out = alloc_fit_outputs(n_iters, N, n_synth)
        scores, logp, comm, synth_comm = out
        if (scores.shape != (n_iters, N) or logp.shape != (n_iters, N) or comm.shape != (n_iters, N)
                or synth_comm.shape != (n_iters, max(n_synth, 1))):
            raise ValueError("fit_iterations: `out` arrays do not match (n_iters, n_cells, n_synth)")
        stage_ms = np.zeros(8, dtype=np.float64)
        self._check(self._lib.dd_fit_iterations(
            self._h, ctypes.byref(p), _ptr(parents, ctypes.c_int64), _ptr(omega, ctypes.c_float),
            _ptr(scores, ctypes.c_double), _ptr(logp, ctypes.c_double), _ptr(comm, ctypes.c_int32),
            _ptr(synth_comm, ctypes.c_int32), _ptr(stage_ms, ctypes.c_double)))
        self.n_synth = n_synth
        names = ["host_cluster_score", "normalise", "scale", "pca", "knn", "cluster_gpu_d2h", "device_total", "wall"]
        return dict(scores=scores, log_p=logp, communities=comm, synth_communities=synth_comm[:, :n_synth],
                    stage_ms=dict(zip(names, stage_ms.tolist())))

    # cell-block sharding (one handle per rank / GPU)
    def comm_init(self, rank, world, unique_id):
        """Collective over all ranks; ``unique_id`` is rank 0's :func:`comm_unique_id` token."""
        buf = ctypes.create_string_buffer(bytes(unique_id), COMM_ID_BYTES)
        self._check(self._lib.dd_comm_init(self._h, int(rank), int(world), buf, COMM_ID_BYTES))

    def shard_cells(self, on=True):
        self._check(self._lib.dd_comm_shard_cells(self._h, int(bool(on))))

    def comm_info(self):
        rank, world = ctypes.c_int32(0), ctypes.c_int32(0)
        block = np.zeros(4, dtype=np.int64)
        self._check(self._lib.dd_comm_info(self._h, ctypes.byref(rank), ctypes.byref(world), _ptr(block, ctypes.c_int64)))
        return dict(rank=rank.value, world=world.value, first_cell=int(block[0]), n_cells=int(block[1]),
                    first_synth=int(block[2]), n_synth=int(block[3]))

    # introspection
    def kernel_launches(self):
        return int(self._lib.dd_kernel_launches(self._h))

    def last_stage_ms(self, stage):
        return float(self._lib.dd_last_stage_ms(self._h, stage.encode()))

    def set_kernel_timing(self, on):
        self._check(self._lib.dd_set_kernel_timing(self._h, int(bool(on))))

    def kernel_timing_report(self):
        """{kernel name: (total_ms, launches)} since timing was switched on."""
        need = self._lib.dd_kernel_timing_report(self._h, None, 0)
        buf = ctypes.create_string_buffer(int(need) + 16)
        self._lib.dd_kernel_timing_report(self._h, buf, len(buf))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, n = line.split()
            out[name] = (float(ms), int(n))
        return out

    def kernel_timing(self, kernel):
        ms = ctypes.c_double(0)
        n = ctypes.c_int64(0)
        self._check(self._lib.dd_get_kernel_timing(self._h, kernel.encode(), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value


def alloc_fit_outputs(n_iters, n_cells, n_synth):
    """(scores, log_p, communities, synth_communities) of ``n_iters`` iterations, zero-filled (lazily: calloc pages)."""
    return (np.zeros((n_iters, n_cells), dtype=np.float64), np.zeros((n_iters, n_cells), dtype=np.float64),
            np.zeros((n_iters, n_cells), dtype=np.int32), np.zeros((n_iters, max(n_synth, 1)), dtype=np.int32))


def fit_iterations_pipelined(handles, parents, omega, *, iter_begin=0, iter_end=None, n_host_threads=1, **kw):
    """``Handle.fit_iterations`` over several handles of ONE GPU that share the count matrix: the iteration range is cut into
    contiguous pieces, one pipelined loop (and host thread; ctypes releases the GIL) per handle.  The loops interleave on the
    device -- one loop's small latency-bound kernels run underneath another's bandwidth-bound ones (measured at c3: 1.09x with
    two loops).  Iterations are independent, so the result equals the single loop's."""
    import threading

    parents = np.ascontiguousarray(parents, dtype=np.int64)
    n_iters = parents.shape[0]
    iter_end = n_iters if iter_end is None else iter_end
    n_run = iter_end - iter_begin
    pipes = max(1, min(len(handles), n_run))
    if pipes == 1:
        return handles[0].fit_iterations(parents, omega, iter_begin=iter_begin, iter_end=iter_end, n_host_threads=n_host_threads, **kw)
    bounds = [iter_begin + (n_run * i) // pipes for i in range(pipes + 1)]
    outs, errors = [None] * pipes, [None] * pipes
    threads_each = max(1, n_host_threads // pipes)
    # ONE set of result arrays: every loop writes the rows of its own iterations (no per-loop copies to merge afterwards)
    n_cells = getattr(handles[0], "n_cells", None)
    shared = alloc_fit_outputs(n_iters, n_cells, parents.shape[1]) if n_cells else None
    extra = {} if shared is None else {"out": shared}

    def work(i):
        try:
            outs[i] = handles[i].fit_iterations(parents, omega, iter_begin=bounds[i], iter_end=bounds[i + 1],
                                                n_host_threads=threads_each, **extra, **kw)
        except Exception as e:  # noqa: BLE001 -- re-raised on the calling thread
            errors[i] = e

    ths = [threading.Thread(target=work, args=(i,)) for i in range(pipes)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    out = outs[0]
    for i in range(1, pipes):
        for key in ("scores", "log_p", "communities", "synth_communities"):
            if not np.may_share_memory(out[key], outs[i][key]):  # a handle that allocated its own arrays
                out[key][bounds[i]:bounds[i + 1]] = outs[i][key][bounds[i]:bounds[i + 1]]
    stage = {k: sum(o["stage_ms"][k] for o in outs) for k in out["stage_ms"]}
    # the loops run side by side: what the caller waited for is the longest of them
    for key in ("wall", "device_total"):
        if key in stage:
            stage[key] = max(o["stage_ms"][key] for o in outs)
    stage["pipelines"] = float(pipes)
    out["stage_ms"] = stage
    return out
