"""``BoostClassifier`` -- drop-in for ``doubletdetection.BoostClassifier``
(reference: doubletdetection/doubletdetection.py:22-426) whose per-iteration fit loop runs on a B200
through ``libdd_b200.so``.

Same constructor arguments, defaults, warnings and errors (:73-133, :404-426), same ``fit`` /
``predict`` / ``doublet_score`` semantics (:135-272) and the same fitted attributes.  What changes is
where ``_one_fit`` (:274-383) executes: synthetic doublets, normalise/log, optional scaling,
randomized PCA and the exact kNN graph are CUDA kernels, clustering (Louvain) and scoring are native
host code overlapped with the GPU.  ``clustering_algorithm="phenograph"`` (the reference's default) builds
PhenoGraph's Jaccard graph of the 30 nearest neighbours on the GPU and partitions it with the in-repo Louvain
(one seeded run; the phenograph package with its time-seeded binaries is not available); ``"leiden"`` builds
umap's fuzzy-simplicial-set weights of the exact kNN lists and distances on the GPU as well and partitions that graph
on the native host workers (in-repo Leiden; leidenalg is not available).  There is no CPU fallback: without the built library or without a B200 ``fit`` raises.

Keyword-only extensions (not in the reference): ``device`` (CUDA device index; default
``LOCAL_RANK`` or 0) and ``distributed`` (shard the work over the ranks of an initialised
``torch.distributed`` group, one process per GPU).  ``True`` deals the independent iterations to the ranks
and gathers the per-iteration results on rank 0, ``"allgather"`` on every rank (BASELINE config 4);
``"cells"`` shards the CELLS of every iteration instead (BASELINE config 5): each rank builds and factorises
its block of the augmented matrix, the PCA reductions are NCCL all-reduces and the embedding / kNN lists are
all-gathered inside ``libdd_b200.so`` -- every rank ends up with the complete results.
"""

import os
import warnings

import numpy as np
import scipy.sparse as sp_sparse
from scipy.sparse import csr_matrix
from sklearn.utils import check_array

from . import _capi


def _pca_solver(n_samples, n_features, n_components):
    """sklearn ``PCA(svd_solver="auto")`` dispatch (sklearn/decomposition/_pca.py:524-536), which is what
    ``sc.tl.pca(..., svd_solver="auto")`` (doubletdetection.py:308-314) ends up in."""
    if n_features <= 1000 and n_samples >= 10 * n_features:
        return "covariance_eigh"
    if max(n_samples, n_features) <= 500:
        return "full"
    if 1 <= n_components < 0.8 * min(n_samples, n_features):
        return "randomized"
    return "full"


def _pca_plan(n_samples, n_features, n_components, random_state, arpack=False):
    """Test matrix and power-iteration count of sklearn's randomized SVD
    (sklearn/utils/extmath.py:323-333, 584-587): Omega = RandomState(seed).normal((G, C + 10)) cast to
    float32 (one row per cell instead when there are fewer augmented cells than genes), 7 iterations when
    C < 0.1 * min(shape) else 4.  ``arpack``: the reference's sparse branch (pseudocount == 1 without scaling,
    doubletdetection.py:296-297, 308) asks for the EXACT truncated SVD whatever the shape."""
    solver = "arpack" if arpack else _pca_solver(n_samples, n_features, n_components)
    if solver != "randomized":
        return None, 0  # exact PCA (covariance_eigh / full): _exact_pca, no test matrix
    n_power_iter = 7 if n_components < 0.1 * min(n_samples, n_features) else 4
    # fewer samples than features: sklearn works on the transposed matrix, so Omega has one row per SAMPLE
    # (sklearn/utils/extmath.py:589-592: transpose = n_samples < n_features)
    rows = n_samples if n_samples < n_features else n_features
    omega = np.random.RandomState(random_state).normal(size=(rows, n_components + 10)).astype(np.float32)
    return omega, n_power_iter


def _exact_pca(h, n_components):
    """sklearn's exact branches (``covariance_eigh``: <= 1000 genes and >= 10x as many augmented cells; ``full``: tiny
    matrices or n_components >= 0.8 min(shape); sklearn/decomposition/_pca.py:524-536, 560-640) on the dense matrix the
    handle holds: X_pca = top principal components of the centred matrix, signs by ``svd_flip(u_based_decision=False)``.
    The device computes the float64 Gram matrix of the centred matrix on its smaller side and the projection of all
    augmented cells; the <= 1000 x 1000 symmetric eigenproblem in between is LAPACK on the host (O(G^3), independent of
    the number of cells -- the same routine sklearn calls).  Leaves the embedding on the device for ``knn``."""
    from scipy.linalg import eigh

    n_rows, n_genes = h._dense_rows, h.n_genes
    c = int(n_components)
    if n_genes <= n_rows:
        # the top c eigenpairs of (A - 1) * covariance (LAPACK dsyevr; ascending order)
        w, v = eigh(h.centered_gram(False), subset_by_index=[max(n_genes - c, 0), n_genes - 1], overwrite_a=True,
                    check_finite=False)
        v = np.ascontiguousarray(v[:, ::-1][:, :c])
        top = np.argmax(np.abs(v), axis=0)
        v *= np.sign(v[top, np.arange(v.shape[1])])[None, :]
        return h.project(v)
    # fewer augmented cells than genes (only reachable with <= 500 rows or n_components >= 0.8 A): eigenvectors of the
    # A x A Gram matrix are U, X_pca = U S; the sign convention needs Vt = S^-1 U^T Dc, a (c x G) product on a tiny matrix
    w, u = np.linalg.eigh(h.centered_gram(True))
    u, w = u[:, ::-1][:, :c], np.maximum(w[::-1][:c], 0.0)
    s = np.sqrt(w)
    dense = h.download_dense().astype(np.float64)
    dense -= dense.mean(axis=0)
    vt = (u.T @ dense) / np.where(s > 0, s, 1.0)[:, None]
    signs = np.sign(vt[np.arange(vt.shape[0]), np.argmax(np.abs(vt), axis=1)])
    emb = np.ascontiguousarray(u * (s * signs)[None, :], dtype=np.float32)
    h.upload_embedding(emb)
    return emb


class BoostClassifier:
    """Classifier for doublets in single-cell RNA-seq data (see the reference docstring,
    doubletdetection.py:23-71, for parameters and attributes -- they are identical)."""

    def __init__(
        self,
        boost_rate=0.25,
        n_components=30,
        n_top_var_genes=10000,
        replace=False,
        clustering_algorithm="phenograph",
        clustering_kwargs=None,
        n_iters=10,
        normalizer=None,
        pseudocount=0.1,
        random_state=0,
        verbose=False,
        standard_scaling=False,
        n_jobs=1,
        *,
        device=None,
        distributed=False,
    ):
        self.boost_rate = boost_rate
        self.replace = replace
        self.clustering_algorithm = clustering_algorithm
        self.n_iters = n_iters
        self.normalizer = normalizer
        self.random_state = random_state
        self.verbose = verbose
        self.standard_scaling = standard_scaling
        self.n_jobs = n_jobs
        self.pseudocount = pseudocount
        self.rng = np.random.default_rng(self.random_state)  # :99 -- one stream for all fits
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        # False | True (iterations sharded, results on rank 0) | "allgather" (same, results everywhere) |
        # "cells" (cells of every iteration sharded, NCCL inside the library)
        if distributed not in (False, True, "allgather", "cells"):
            raise ValueError("distributed must be one of False, True, 'allgather', 'cells'")
        self.distributed = distributed

        if self.clustering_algorithm not in ["louvain", "phenograph", "leiden"]:  # :101-104
            raise ValueError("Clustering algorithm needs to be one of ['louvain', 'phenograph', 'leiden']")
        if self.clustering_algorithm == "leiden":  # :105-106
            warnings.warn("Leiden clustering is experimental and results have not been validated.")

        if n_components == 30 and n_top_var_genes > 0:  # :108-112
            self.n_components = min(n_components, n_top_var_genes)
        else:
            self.n_components = n_components
        self.n_top_var_genes = max(0, n_top_var_genes)  # :114

        self.clustering_kwargs = {} if not isinstance(clustering_kwargs, dict) else clustering_kwargs
        self._set_clustering_kwargs()

        if not self.replace and self.boost_rate > 0.5:  # :121-127
            warnings.warn(
                "boost_rate is trimmed to 0.5 when replace=False. Set replace=True to use greater boost rates."
            )
            self.boost_rate = 0.5

        assert (self.n_top_var_genes == 0) or (
            self.n_components <= self.n_top_var_genes
        ), "n_components={0} cannot be larger than n_top_var_genes={1}".format(n_components, n_top_var_genes)

        self._handle = None
        self._extra_handles = []  # further pipelines on the same GPU (share the primary handle's count matrix)
        self._parents_array = None
        self._parents_lists = None
        self.stage_ms_ = None
        self._fitted = {}     # all_scores_ / all_log_p_values_ / communities_ / synth_communities_ once they are complete
        self._pending = None  # iteration-sharded fit whose per-iteration rows have not been collected yet

    # ------------------------------------------------------------------ kwargs (:404-426)
    def _set_clustering_kwargs(self):
        if self.clustering_algorithm == "phenograph":
            if "prune" not in self.clustering_kwargs:
                self.clustering_kwargs["prune"] = True
            if (self.n_iters == 1) and (self.clustering_kwargs.get("prune") is True):
                warnings.warn(
                    "Using phenograph parameter prune=False is strongly recommended when "
                    "running only one iteration. Otherwise, expect many NaN labels."
                )
        else:
            if "directed" not in self.clustering_kwargs:
                self.clustering_kwargs["directed"] = False
            if "resolution" not in self.clustering_kwargs:
                self.clustering_kwargs["resolution"] = 4
            if "key_added" in self.clustering_kwargs:
                raise ValueError("'key_added' param cannot be overriden")
            if "random_state" in self.clustering_kwargs:
                raise ValueError("'random_state' param cannot be overriden. Please use classifier 'random_state'.")

    # ------------------------------------------------------------------ parents_ (:395, :197)
    @property
    def parents_(self):
        """List (per iteration) of lists of ``[parent0, parent1]`` -- the reference's structure,
        materialised on first access from the int64 array the fit kept."""
        if self._parents_lists is None and self._parents_array is not None:
            self._parents_lists = [[list(p) for p in it] for it in self._parents_array]
        if self._parents_lists is None:
            raise AttributeError("parents_ is set by fit()")
        return self._parents_lists

    @parents_.setter
    def parents_(self, value):
        self._parents_lists = value

    # ------------------------------------------------------------------ fitted (n_iters, .) arrays (:186-214)
    # Plain arrays after a single-process fit.  After an ITERATION-SHARDED fit (distributed=True / "allgather") every rank
    # holds the rows of its own iterations only: predict() and doublet_score() need per-cell sums over the iterations, so
    # they all-reduce N-vectors (votes and valid counts; log-p sums for doublet_score) instead of moving the (n_iters x N) arrays; the
    # arrays themselves are gathered on first access (a collective: every rank must touch them, or none).
    def _collect(self):
        if self._pending is None:
            return
        pend, self._pending = self._pending, None
        merged = _allgather_iterations(pend["dist"], pend["out"], self.n_iters, self.device,
                                       everywhere=self.distributed == "allgather")
        self._store(merged)

    def _store(self, out):
        # the reference stores the communities in float arrays (:188); the native loop returns int32: converted when read
        # (20 MB per fit at 25 x 100k that predict / doublet_score never look at)
        self._fitted = dict(
            all_scores_=out["scores"], all_log_p_values_=out["log_p"],
            communities_=out["communities"], synth_communities_=out["synth_communities"])

    def _get_fitted(self, name):
        self._collect()
        try:
            value = self._fitted[name]
        except KeyError:
            raise AttributeError(f"{name} is set by fit()") from None
        if name in ("communities_", "synth_communities_") and isinstance(value, np.ndarray) and value.dtype != np.float64:
            value = self._fitted[name] = value.astype(np.float64)
        return value

    all_scores_ = property(lambda self: self._get_fitted("all_scores_"),
                           lambda self, v: self._fitted.__setitem__("all_scores_", v))
    all_log_p_values_ = property(lambda self: self._get_fitted("all_log_p_values_"),
                                 lambda self, v: self._fitted.__setitem__("all_log_p_values_", v))
    communities_ = property(lambda self: self._get_fitted("communities_"),
                            lambda self, v: self._fitted.__setitem__("communities_", v))
    synth_communities_ = property(lambda self: self._get_fitted("synth_communities_"),
                                  lambda self, v: self._fitted.__setitem__("synth_communities_", v))

    def _reduced_log_p_stats(self, log_p_thresh, want_total=True):
        """(votes, valid count, sum of valid log p) per cell over ALL iterations (votes / the sum only if asked for: the sum
        is two thirds of the work and predict does not need it).  Pending sharded fit: from this rank's rows, summed over
        the ranks (integers exactly; the float64 sums in rank order instead of iteration order)."""
        if self._pending is None:
            log_p = np.asarray(self.all_log_p_values_)
        else:
            log_p = np.asarray(self._pending["out"]["log_p"])[self._pending["it0"]:self._pending["it1"]]
        valid = np.isfinite(log_p)  # masked_invalid masks NaN, +inf and -inf (quirk Q5)
        with np.errstate(invalid="ignore"):
            votes = np.count_nonzero((log_p <= log_p_thresh) & valid, axis=0) if log_p_thresh is not None else None
        count = np.count_nonzero(valid, axis=0)
        total = np.where(valid, log_p, 0.0).sum(axis=0) if want_total else None
        if self._pending is not None:
            import torch

            dist = self._pending["dist"]
            dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            ints = np.stack([votes if votes is not None else np.zeros_like(count), count]).astype(np.int64)
            t_i = torch.from_numpy(ints).to(dev)
            dist.all_reduce(t_i, op=dist.ReduceOp.SUM)
            ints = t_i.cpu().numpy()
            if want_total:
                t_f = torch.from_numpy(np.ascontiguousarray(total)).to(dev)
                dist.all_reduce(t_f, op=dist.ReduceOp.SUM)
                total = t_f.cpu().numpy()
            votes, count = (ints[0] if votes is not None else None), ints[1]
        return votes, count, total

    # ------------------------------------------------------------------ fit (:135-214)
    def _native(self):
        if self._handle is None:
            self._handle = _capi.Handle(self.device)
        return self._handle

    def _pipelines(self, h, n_run):
        """Handles of the pipelined loops that share this GPU: by default TWO when there are at least four iterations to
        run (the loops' latency-bound and bandwidth-bound kernels fill each other's gaps: 1.09x at c3), DD_PIPELINES
        overrides.  The extra handles read the primary handle's resident count matrix."""
        want = int(os.environ.get("DD_PIPELINES", "2"))
        want = max(1, min(want, 4, n_run // 2))
        while len(self._extra_handles) < want - 1:
            self._extra_handles.append(_capi.Handle(self.device))
        extras = self._extra_handles[: want - 1]
        for h2 in extras:
            h2.share_counts(h)
        return [h] + extras

    def _host_threads(self):
        n = int(self.n_jobs) if self.n_jobs else 1
        if n < 0:
            n = max(1, (os.cpu_count() or 1) + 1 + n)  # joblib convention: -1 = all cores
        return max(1, n)

    def fit(self, raw_counts):
        """Fits the classifier on raw_counts (cells x genes, dense or CSR)."""
        if self.normalizer is not None:
            # the reference itself raises NameError on this path in this version (:288-291 vs :301, :372)
            raise NotImplementedError("custom `normalizer` is not supported (and is broken in the reference at this version)")
        cluster_kw = {}
        if self.clustering_algorithm == "phenograph":
            # phenograph.cluster(X_pca, n_jobs=self.n_jobs, **clustering_kwargs) (:320): the arguments that change
            # the graph or the labels; the rest of phenograph's signature must stay at its defaults
            allowed = {"prune", "k", "min_cluster_size", "jaccard", "directed", "primary_metric", "clustering_algo"}
            extra = set(self.clustering_kwargs) - allowed
            if extra:
                raise NotImplementedError(f"unsupported clustering_kwargs for the native PhenoGraph: {sorted(extra)}")
            kw = self.clustering_kwargs
            if (kw.get("jaccard", True) is not True or kw.get("directed", False) is not False
                    or kw.get("primary_metric", "euclidean") != "euclidean" or kw.get("clustering_algo", "louvain") != "louvain"):
                raise NotImplementedError("native PhenoGraph: only jaccard=True, directed=False, euclidean, louvain")
            cluster_kw = dict(clustering="phenograph", pheno_k=int(kw.get("k", 30)), pheno_prune=bool(kw.get("prune", True)),
                              pheno_min_cluster_size=int(kw.get("min_cluster_size", 10)))
        else:
            if self.clustering_kwargs.get("directed", False):
                raise NotImplementedError("clustering_kwargs['directed']=True is not supported (the reference default is False)")
            extra = set(self.clustering_kwargs) - {"directed", "resolution"}
            if extra:
                raise NotImplementedError(
                    f"unsupported clustering_kwargs for the native {self.clustering_algorithm}: {sorted(extra)}")
            # "louvain": unweighted pattern of the kNN graph (:337-338); "leiden": umap-weighted graph, iterated until
            # stable (:339-340, sc.tl.leiden's use_weights=True / n_iterations=-1) on the host workers
            cluster_kw = dict(clustering=self.clustering_algorithm, resolution=float(self.clustering_kwargs["resolution"]))
        # pseudocount == 1 is the reference's SPARSE branch (:296-297): log1p keeps the matrix sparse and :308 asks
        # sc.tl.pca for svd_solver="arpack", i.e. the exact truncated SVD of the centred matrix -- unless standard_scaling
        # densifies it first (sc.pp.scale zero-centres), which puts the dense "auto" solver back in charge.  Here the
        # matrix is dense on the device either way (log1p(0) = 0 fills the gaps); what changes is the PCA: the exact branch
        # (float64 Gram matrix on the smaller side + LAPACK, _exact_pca) instead of the randomized one.
        arpack = self.pseudocount == 1 and self.standard_scaling is not True

        import time as _time

        _t = [_time.perf_counter()]
        # :149-155.  For sparse input the finiteness scan (a full pass over the values on one host core) is left to the
        # device, which looks at every value anyway while it sums the rows; if it finds NaN / inf the reference's own call
        # is repeated below to raise sklearn's error.
        sparse_in = sp_sparse.issparse(raw_counts)
        counts_in = raw_counts
        raw_counts = check_array(
            raw_counts, accept_sparse="csr", ensure_all_finite=not sparse_in, ensure_2d=True, dtype="float32"
        )
        if sp_sparse.issparse(raw_counts) is not True:  # :157-160
            if self.verbose:
                print("Sparsifying matrix.")
            raw_counts = csr_matrix(raw_counts)

        # :165-176 highly variable genes.  Canonical CSR input: variances and the column subset are computed on the device
        # from the uploaded matrix (hvg.cu reproduces scipy's float32 accumulation order; the argsort stays numpy's, so ties
        # break as in the reference).  A matrix with duplicate or unsorted entries takes the reference's own scipy lines.
        hvg = self.n_top_var_genes > 0 and self.n_top_var_genes < raw_counts.shape[1]
        hvg_on_device = hvg and raw_counts.has_canonical_format and os.environ.get("DD_HVG_HOST") is None
        if hvg and not hvg_on_device:
            gene_variances = (
                np.array(raw_counts.power(2).mean(axis=0)) - (np.array(raw_counts.mean(axis=0))) ** 2
            )[0]
            top_var_indexes = np.argsort(gene_variances)
            self.top_var_genes_ = top_var_indexes[-self.n_top_var_genes:]
            raw_counts = raw_counts.tocsc()[:, self.top_var_genes_].tocsr()
        if not raw_counts.has_canonical_format:  # the kernels merge sorted, duplicate-free rows
            raw_counts = raw_counts.copy()
            raw_counts.sum_duplicates()

        num_cells, num_genes = raw_counts.shape
        if hvg_on_device:
            num_genes = int(self.n_top_var_genes)
        self._num_cells, self._num_genes = num_cells, num_genes
        num_synths = int(self.boost_rate * num_cells)  # :391
        n_aug = num_cells + num_synths
        omega, n_power_iter = _pca_plan(n_aug, num_genes, self.n_components, self.random_state, arpack=arpack)
        # sklearn's own argument check (PCA._fit_full / _fit_truncated), which is where the reference fails for such shapes
        n_min = min(n_aug, num_genes)
        if self.n_components > n_min or (arpack and self.n_components >= n_min):
            raise ValueError(f"n_components={self.n_components} must be {'strictly less than' if arpack else 'between 0 and'} "
                             f"min(n_samples, n_features)={n_min} for the PCA of the {n_aug} x {num_genes} augmented matrix")
        if omega is not None and num_genes <= 50 and self.clustering_algorithm != "phenograph":  # unreachable: see _pca_solver
            raise NotImplementedError("at most 50 genes (sc.pp.neighbors then works on X, not X_pca) outside the exact-PCA route")
        if omega is None and min(n_aug, num_genes) > 16384:
            raise NotImplementedError(
                f"exact PCA (pseudocount=1 selects svd_solver='arpack') of a {n_aug} x {num_genes} matrix: the float64 Gram "
                "matrix on the smaller side is limited to 16384 rows; lower n_top_var_genes or use another pseudocount")

        _t.append(_time.perf_counter())
        h = self._native()

        def upload():
            h.upload_counts(raw_counts)
            if hvg_on_device:
                gene_variances = h.hvg_variances()  # :166-169
                top_var_indexes = np.argsort(gene_variances)  # :170
                self.top_var_genes_ = top_var_indexes[-self.n_top_var_genes:]  # :171
                h.select_genes(self.top_var_genes_)  # :173-175

        # the host->device copy of the counts (ctypes releases the GIL) runs underneath the parent draws
        from concurrent.futures import ThreadPoolExecutor

        with ThreadPoolExecutor(max_workers=1) as pool:
            upload = pool.submit(upload)
            # every iteration's `choices` (:394), drawn sequentially from the classifier's stream
            parents = np.empty((self.n_iters, num_synths, 2), dtype=np.int64)
            for i in range(self.n_iters):
                parents[i] = self.rng.choice(num_cells, size=(num_synths, 2), replace=self.replace)
            upload.result()  # re-raises what the upload raised
        if sparse_in and not h.counts_all_finite():
            check_array(counts_in, accept_sparse="csr", ensure_all_finite=True, ensure_2d=True, dtype="float32")  # raises
            raise ValueError("Input contains NaN or infinity.")
        _t.append(_time.perf_counter())

        it0, it1 = 0, self.n_iters
        dist = cells_dist = None
        if self.distributed:
            import torch.distributed as dist_mod

            if dist_mod.is_available() and dist_mod.is_initialized() and dist_mod.get_world_size() > 1:
                dist = dist_mod
                rank, world = dist.get_rank(), dist.get_world_size()
                if self.distributed == "cells":
                    if not getattr(h, "_comm_ready", False):
                        token = broadcast_token(dist, _capi.comm_unique_id() if rank == 0 else None, self.device)
                        h.comm_init(rank, world, token)
                        h._comm_ready = True
                    h.shard_cells(True)
                    cells_dist, dist = dist, None  # every rank takes part in every iteration; results merged below
                else:
                    it0, it1 = iteration_shard(self.n_iters, rank, world)
        if self.distributed != "cells" and getattr(h, "_comm_ready", False):
            h.shard_cells(False)

        if self.verbose:
            print(f"Running iterations {it0 + 1}..{it1} of {self.n_iters} on cuda:{self.device}")
        if omega is None:
            if cells_dist is not None:
                raise NotImplementedError("distributed='cells' with sklearn's exact PCA branches (<= 1000 genes or a tiny matrix)")
            out = self._fit_iterations_exact_pca(h, parents, it0, it1, cluster_kw)
        else:
            failure = None
            try:
                handles = [h] if cells_dist is not None else self._pipelines(h, it1 - it0)
                out = _capi.fit_iterations_pipelined(
                    handles, parents, omega,
                    pseudocount=self.pseudocount, standard_scaling=self.standard_scaling is True,
                    n_comp=self.n_components, n_power_iter=n_power_iter, knn_k=10, seed=int(self.random_state),
                    n_host_threads=self._host_threads(), iter_begin=it0, iter_end=it1, **cluster_kw,
                )
            except Exception as e:  # noqa: BLE001 -- re-raised below, on every rank
                if cells_dist is None:
                    raise
                failure = e
            if cells_dist is not None:
                # the failure of an iteration is seen by the rank that owns it only: make it collective before the merge
                any_failed = all_ranks_any(cells_dist, failure is not None, self.device)
                if failure is not None:
                    raise failure
                if any_failed:
                    raise RuntimeError("fit failed on another rank of the cell-sharded group (see that rank's error)")
        _t.append(_time.perf_counter())
        self._pending, self._fitted = None, {}
        if cells_dist is not None:
            out = merge_owned_iterations(cells_dist, out, self.device)
        if self.verbose:
            # the reference prints these lines while it iterates (:193-194, :350-355); the pipelined loop has no such moment, so
            # the same lines are printed for the iterations this process ran once they are finished
            for i in range(it0, it1):
                full = np.concatenate([out["communities"][i], out["synth_communities"][i]])
                community_sizes = [int(np.count_nonzero(full == c)) for c in np.unique(full)]
                print("Iteration {:3}/{}".format(i + 1, self.n_iters))
                print("Found clusters [{0}, ... {2}], with sizes: {1}\n".format(full.min(), community_sizes, full.max()))
        self.stage_ms_ = out["stage_ms"]
        if dist is not None and self.n_iters > 1:
            # iteration-sharded: the (n_iters, .) rows stay where they were computed until somebody asks for them
            self._pending = dict(dist=dist, out=out, it0=it0, it1=it1)
        elif dist is not None:
            self._store(_allgather_iterations(dist, out, self.n_iters, self.device, everywhere=self.distributed == "allgather"))
        else:
            self._store(out)
        self._parents_array = parents
        self._parents_lists = None
        _t.append(_time.perf_counter())
        # host-side wall time of the phases of this fit (ms): validation + HVG, upload (with the parent draws
        # underneath), the pipelined native loop, result collection
        self.host_ms_ = dict(zip(("prologue", "upload", "fit_iterations", "collect"),
                                 [1e3 * (b - a) for a, b in zip(_t[:-1], _t[1:])]))
        return self

    def _fit_iterations_exact_pca(self, h, parents, it0, it1, cluster_kw):
        """``_one_fit`` (:274-383) iteration by iteration for the shapes where sklearn's "auto" PCA is EXACT (see
        ``_exact_pca``): same device stages as the pipelined loop (fused doublets + normalise/log, optional scaling, exact
        kNN), the exact PCA in place of the randomized one, and the native host twins of the clustering stage."""
        import time as _time

        n_cells = self._num_cells
        n_synth = parents.shape[1]
        scores = np.zeros((self.n_iters, n_cells))
        log_p = np.zeros((self.n_iters, n_cells))
        comm = np.zeros((self.n_iters, n_cells), dtype=np.int32)
        synth_comm = np.zeros((self.n_iters, n_synth), dtype=np.int32)
        algo = cluster_kw["clustering"]
        seed = int(self.random_state)
        # sc.pp.neighbors (:331-336) looks at adata.X itself, not at X_pca, when there are at most 50 genes (scanpy's
        # settings.N_PCS) -- such matrices always land on this route (sklearn's "auto" is exact for them), and the
        # device kNN then runs on the (<= 50-column) normalised matrix.  phenograph.cluster (:320) is handed X_pca regardless.
        knn_on_x = self._num_genes <= 50 and algo != "phenograph"
        t0 = _time.perf_counter()
        for i in range(it0, it1):
            h.create_doublets(parents[i])
            h.normalise_log(h.median_lib_size(), self.pseudocount)
            if self.standard_scaling is True:
                h.standard_scale(15.0)
            if knn_on_x:
                h.upload_embedding(h.download_dense())
            else:
                _exact_pca(h, self.n_components)
            if algo == "phenograph":
                idx, _ = h.knn(cluster_kw["pheno_k"] + 1, with_dist=False)
                labels = _capi.phenograph_knn(idx, prune=cluster_kw["pheno_prune"],
                                              min_cluster_size=cluster_kw["pheno_min_cluster_size"], seed=seed)
            elif algo == "leiden":
                h.knn(10, with_dist=False)  # the distances stay on the device, where umap's graph is built
                g = h.umap_graph(10)
                labels = _capi.leiden_csr(g.indptr, g.indices, g.data.astype(np.float64), resolution=cluster_kw["resolution"],
                                          seed=seed)
            else:
                idx, _ = h.knn(10, with_dist=False)
                labels = _capi.louvain_knn(idx, resolution=cluster_kw["resolution"], seed=seed)
            scores[i], log_p[i] = _capi.score(labels, n_cells)
            comm[i], synth_comm[i] = labels[:n_cells], labels[n_cells:]
        wall = 1e3 * (_time.perf_counter() - t0)
        return dict(scores=scores, log_p=log_p, communities=comm, synth_communities=synth_comm,
                    stage_ms=dict(wall=wall, device_total=wall))

    # ------------------------------------------------------------------ predict (:216-254)
    def predict(self, p_thresh=1e-7, voter_thresh=0.9):
        log_p_thresh = np.log(p_thresh)
        if self.n_iters > 1:
            # :232-241 -- np.mean(np.ma.masked_invalid(log_p) <= thresh, axis=0), the vote >= voter_thresh, both filled
            # with NaN where every iteration is masked.  Same values without the masked-array machinery (42 -> 6 ms at
            # 25 x 100k): a masked mean is (number of valid votes) * 1.0 / (number of valid entries) in float64.
            votes, count, _ = self._reduced_log_p_stats(log_p_thresh, want_total=False)
            with np.errstate(invalid="ignore", divide="ignore"):
                average = votes * 1.0 / count
                labels = (average >= voter_thresh).astype(float)
            none_valid = count == 0
            labels[none_valid] = np.nan
            average[none_valid] = np.nan
            self.voting_average_ = average
            self.labels_ = labels
        else:
            potential_cutoffs = np.unique(self.all_scores_[~np.isnan(self.all_scores_)])
            if len(potential_cutoffs) > 1:
                max_dropoff = np.argmax(potential_cutoffs[1:] - potential_cutoffs[:-1]) + 1
            else:
                max_dropoff = 0
            self.suggested_score_cutoff_ = potential_cutoffs[max_dropoff]
            with np.errstate(invalid="ignore"):
                self.labels_ = self.all_scores_[0, :] >= self.suggested_score_cutoff_
            self.labels_[np.isnan(self.all_scores_)[0, :]] = np.nan
        return self.labels_

    # ------------------------------------------------------------------ doublet_score (:256-272)
    def doublet_score(self):
        if self.n_iters > 1:
            # :268 -- np.mean(np.ma.masked_invalid(log_p), axis=0): a MaskedArray (quirk Q5) whose values are
            # filled(0).sum(axis=0) * 1.0 / count, masked where no iteration is valid; built directly
            _, count, total = self._reduced_log_p_stats(None)
            with np.errstate(invalid="ignore", divide="ignore"):
                avg = total * 1.0 / count
            none_valid = count == 0
            avg[none_valid] = 0.0
            avg_log_p = np.ma.MaskedArray(avg, mask=none_valid if none_valid.any() else np.ma.nomask)
        else:
            avg_log_p = self.all_log_p_values_[0]
        return -avg_log_p


def broadcast_token(dist, token, device):
    """Ship rank 0's NCCL rendezvous token (bytes) to every rank over the existing process group."""
    import torch

    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    buf = torch.zeros(_capi.COMM_ID_BYTES, dtype=torch.uint8)
    if token is not None:
        buf[: len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
    buf = buf.to(dev)
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def all_ranks_any(dist, flag, device):
    """True on every rank iff ``flag`` is true on at least one (an all-reduce MAX of one integer)."""
    import torch

    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))


def merge_owned_iterations(dist, out, device):
    """Cell-block sharding: the library clusters + scores iteration i on rank i % world and leaves the other ranks'
    rows zero.  An integer SUM of the bit patterns over the ranks therefore reproduces every row exactly (NaN and
    -inf included) on every rank."""
    import torch

    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    merged = {}
    for key in ("scores", "log_p", "communities", "synth_communities"):
        a = np.ascontiguousarray(out[key])
        bits = a.view(np.int64) if a.dtype == np.float64 else a.astype(np.int32)
        t = torch.from_numpy(bits.copy()).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        r = t.cpu().numpy()
        merged[key] = r.view(np.float64) if a.dtype == np.float64 else r
    stage = torch.tensor([out["stage_ms"][k] for k in sorted(out["stage_ms"])], dtype=torch.float64, device=dev)
    dist.all_reduce(stage, op=dist.ReduceOp.MAX)
    merged["stage_ms"] = dict(zip(sorted(out["stage_ms"]), stage.cpu().tolist()))
    return merged


def iteration_shard(n_iters, rank, world):
    """Contiguous block of iterations owned by ``rank`` (BASELINE config 4: 24 iterations over 8 GPUs)."""
    return (n_iters * rank) // world, (n_iters * (rank + 1)) // world


def _allgather_iterations(dist, out, n_iters, device, everywhere=False):
    """Collect the per-iteration result rows (every rank filled only its own block of the (n_iters, .) arrays).
    Default: gather on rank 0, which is where ``predict`` / ``doublet_score`` are evaluated -- the other ranks
    keep their own rows only.  ``everywhere=True`` (``distributed="allgather"``): every rank ends up with the
    complete arrays.  No data-path collective is involved -- this is result collection."""
    import torch

    backend = dist.get_backend()
    dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
    rank, world = dist.get_rank(), dist.get_world_size()
    blocks = [iteration_shard(n_iters, r, world) for r in range(world)]
    it0, it1 = blocks[rank]
    merged = {}
    for key in ("scores", "log_p", "communities", "synth_communities"):
        a = np.ascontiguousarray(out[key])
        as_bits = key in ("scores", "log_p")  # NaN / -inf entries must survive: ship float64 as bit patterns
        src = a.view(np.int64) if as_bits else a.astype(np.int32)
        tdtype = torch.int64 if as_bits else torch.int32
        rows = max(b1 - b0 for b0, b1 in blocks)  # collectives want equal shapes: pad the shorter blocks
        block = np.zeros((rows,) + src.shape[1:], dtype=src.dtype)
        block[: it1 - it0] = src[it0:it1]
        mine = torch.from_numpy(block).to(dev)
        parts = None
        if everywhere or rank == 0:
            parts = [torch.empty_like(mine) for _ in blocks]
        if everywhere:
            dist.all_gather(parts, mine)
        else:
            dist.gather(mine, gather_list=parts, dst=0)
        if parts is not None:
            r = torch.cat([t[: b1 - b0] for t, (b0, b1) in zip(parts, blocks)], dim=0).cpu().numpy()
            merged[key] = r.view(np.float64) if as_bits else r
        else:
            merged[key] = out[key]
    stage = torch.tensor([out["stage_ms"][k] for k in sorted(out["stage_ms"])], dtype=torch.float64, device=dev)
    dist.all_reduce(stage, op=dist.ReduceOp.MAX)
    merged["stage_ms"] = dict(zip(sorted(out["stage_ms"]), stage.cpu().tolist()))
    return merged
