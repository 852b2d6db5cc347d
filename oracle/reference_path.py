"""numpy/scipy/sklearn restatement of ``BoostClassifier.fit -> _one_fit`` (TEST INFRASTRUCTURE --
see oracle/__init__.py).  Line numbers cite ``/root/reference/doubletdetection/doubletdetection.py``.

The restatement is stage-wise so that each CUDA stage can be checked against the matching
function on the same inputs; ``OracleClassifier`` strings them together exactly like the
reference does.  ``tests/golden/make_golden.py`` pins this file against the reference's real code.
"""

import collections

import numpy as np
import scipy.sparse as sp_sparse
from scipy.sparse import csr_matrix
from scipy.stats import hypergeom
from sklearn.utils import check_array
from sklearn.utils.sparsefuncs_fast import inplace_csr_row_normalize_l1

from . import upstream


# --------------------------------------------------------------------------- fit prologue
def prologue(raw_counts, n_top_var_genes):
    """doubletdetection.py:149-184.  Returns dict(raw, top_var_genes|None, lib_size, normed)."""
    raw_counts = check_array(  # :149-155
        raw_counts, accept_sparse="csr", ensure_all_finite=True, ensure_2d=True, dtype="float32"
    )
    if sp_sparse.issparse(raw_counts) is not True:  # :157-160
        raw_counts = csr_matrix(raw_counts)
    top_var_genes = None
    if n_top_var_genes > 0 and n_top_var_genes < raw_counts.shape[1]:  # :165-176
        gene_variances = (
            np.array(raw_counts.power(2).mean(axis=0)) - (np.array(raw_counts.mean(axis=0))) ** 2
        )[0]
        top_var_indexes = np.argsort(gene_variances)
        top_var_genes = top_var_indexes[-n_top_var_genes:]
        raw_counts = raw_counts.tocsc()
        raw_counts = raw_counts[:, top_var_genes]
        raw_counts = raw_counts.tocsr()
    lib_size = np.asarray(np.sum(raw_counts, axis=1)).ravel()  # :182
    normed = raw_counts.copy()  # :183-184
    inplace_csr_row_normalize_l1(normed)
    return dict(raw=raw_counts, top_var_genes=top_var_genes, lib_size=lib_size, normed=normed)


# --------------------------------------------------------------------------- _createDoublets
def draw_parents(rng, num_cells, boost_rate, replace):
    """doubletdetection.py:391-394 -- consumes the classifier's PCG64 stream."""
    num_synths = int(boost_rate * num_cells)
    return rng.choice(num_cells, size=(num_synths, 2), replace=replace)


def create_doublets(raw, choices):
    """doubletdetection.py:397-399: gather both parents' rows and add them (canonical CSR)."""
    parent0 = raw[choices[:, 0], :]
    parent1 = raw[choices[:, 1], :]
    return parent0 + parent1


# --------------------------------------------------------------------------- normalise
def normalise(raw_synth, lib_size, normed_raw, pseudocount):
    """doubletdetection.py:288-297.  Returns (aug_counts, aug_lib_size, median)."""
    synth_lib_size = np.asarray(np.sum(raw_synth, axis=1)).ravel()
    aug_lib_size = np.concatenate([lib_size, synth_lib_size])
    normed_synths = raw_synth.copy()
    inplace_csr_row_normalize_l1(normed_synths)
    aug_counts = sp_sparse.vstack((normed_raw, normed_synths))
    median = np.median(aug_lib_size)
    scaled_aug_counts = aug_counts * median
    if pseudocount != 1:
        aug_counts = np.log(scaled_aug_counts.toarray() + pseudocount)
    else:
        aug_counts = np.log1p(scaled_aug_counts)
    return aug_counts, aug_lib_size, median


# --------------------------------------------------------------------------- scoring
def score_communities(fullcommunities, num_cells):
    """doubletdetection.py:344-383.  Returns (scores, log_p_values, communities, synth_communities)."""
    fullcommunities = np.asarray(fullcommunities)
    n_aug = fullcommunities.shape[0]
    n_synth = n_aug - num_cells
    min_ID = min(fullcommunities)
    communities = fullcommunities[:num_cells]
    synth_communities = fullcommunities[num_cells:]
    synth_cells_per_comm = collections.Counter(synth_communities)
    orig_cells_per_comm = collections.Counter(communities)
    community_IDs = orig_cells_per_comm.keys()
    community_scores = {
        i: float(synth_cells_per_comm[i]) / (synth_cells_per_comm[i] + orig_cells_per_comm[i])
        for i in community_IDs
    }
    scores = np.array([community_scores[i] for i in communities])
    community_log_p_values = {
        i: hypergeom.logsf(
            synth_cells_per_comm[i], n_aug, n_synth, synth_cells_per_comm[i] + orig_cells_per_comm[i]
        )
        for i in community_IDs
    }
    log_p_values = np.array([community_log_p_values[i] for i in communities])
    if min_ID < 0:
        scores[communities == -1] = np.nan
        log_p_values[communities == -1] = np.nan
    return scores, log_p_values, communities, synth_communities


# --------------------------------------------------------------------------- predict / score
def predict(all_log_p_values, all_scores, n_iters, p_thresh=1e-7, voter_thresh=0.9):
    """doubletdetection.py:216-254.  Returns dict(labels, voting_average | suggested_score_cutoff)."""
    log_p_thresh = np.log(p_thresh)
    out = {}
    if n_iters > 1:
        with np.errstate(invalid="ignore"):
            voting_average = np.mean(np.ma.masked_invalid(all_log_p_values) <= log_p_thresh, axis=0)
            labels = np.ma.filled((voting_average >= voter_thresh).astype(float), np.nan)
            voting_average = np.ma.filled(voting_average, np.nan)
        out["voting_average"] = voting_average
    else:
        potential_cutoffs = np.unique(all_scores[~np.isnan(all_scores)])
        if len(potential_cutoffs) > 1:
            max_dropoff = np.argmax(potential_cutoffs[1:] - potential_cutoffs[:-1]) + 1
        else:
            max_dropoff = 0
        cutoff = potential_cutoffs[max_dropoff]
        with np.errstate(invalid="ignore"):
            labels = all_scores[0, :] >= cutoff
        labels[np.isnan(all_scores)[0, :]] = np.nan
        out["suggested_score_cutoff"] = cutoff
    out["labels"] = labels
    return out


def doublet_score(all_log_p_values, n_iters):
    """doubletdetection.py:256-272."""
    if n_iters > 1:
        with np.errstate(invalid="ignore"):
            avg_log_p = np.mean(np.ma.masked_invalid(all_log_p_values), axis=0)
    else:
        avg_log_p = all_log_p_values[0]
    return -avg_log_p


# --------------------------------------------------------------------------- whole classifier
class OracleClassifier:
    """The reference's louvain / leiden / phenograph paths end to end (doubletdetection.py:73-214, 274-383), with the
    absent upstream calls replaced by ``oracle.upstream``.  ``hooks`` lets a test capture the
    intermediate of every stage of every iteration."""

    def __init__(
        self,
        boost_rate=0.25,
        n_components=30,
        n_top_var_genes=10000,
        replace=False,
        clustering_kwargs=None,
        n_iters=10,
        pseudocount=0.1,
        random_state=0,
        standard_scaling=False,
        louvain_fn=None,
        keep_stages=False,
        clustering_algorithm="louvain",
        leiden_fn=None,
    ):
        self.clustering_algorithm = clustering_algorithm
        self.boost_rate = boost_rate
        self.replace = replace
        self.n_iters = n_iters
        self.random_state = random_state
        self.standard_scaling = standard_scaling
        self.pseudocount = pseudocount
        self.rng = np.random.default_rng(self.random_state)  # :99
        if n_components == 30 and n_top_var_genes > 0:  # :108-112
            self.n_components = min(n_components, n_top_var_genes)
        else:
            self.n_components = n_components
        self.n_top_var_genes = max(0, n_top_var_genes)
        kw = dict(clustering_kwargs or {})
        if clustering_algorithm == "phenograph":
            kw.setdefault("prune", True)  # :408-409
        else:
            kw.setdefault("directed", False)  # :417-420
            kw.setdefault("resolution", 4)
        self.clustering_kwargs = kw
        if not self.replace and self.boost_rate > 0.5:  # :121-127
            self.boost_rate = 0.5
        self.louvain_fn = louvain_fn
        self.leiden_fn = leiden_fn
        self.keep_stages = keep_stages
        self.stages = []

    def fit(self, raw_counts):
        pro = prologue(raw_counts, self.n_top_var_genes)
        if pro["top_var_genes"] is not None:
            self.top_var_genes_ = pro["top_var_genes"]
        raw = pro["raw"]
        num_cells = raw.shape[0]
        self._num_cells = num_cells
        self.all_scores_ = np.zeros((self.n_iters, num_cells))
        self.all_log_p_values_ = np.zeros((self.n_iters, num_cells))
        all_communities = np.zeros((self.n_iters, num_cells))
        all_parents = []
        all_synth_communities = np.zeros((self.n_iters, int(self.boost_rate * num_cells)))
        for i in range(self.n_iters):
            st = self.one_fit(pro)
            self.all_scores_[i], self.all_log_p_values_[i] = st["scores"], st["log_p"]
            all_communities[i] = st["communities"]
            all_parents.append([list(p) for p in st["choices"]])
            all_synth_communities[i] = st["synth_communities"]
            if self.keep_stages:
                self.stages.append(st)
        self.communities_ = all_communities
        self.parents_ = all_parents
        self.synth_communities_ = all_synth_communities
        return self

    def one_fit(self, pro):
        raw, num_cells = pro["raw"], pro["raw"].shape[0]
        st = {}
        st["choices"] = draw_parents(self.rng, num_cells, self.boost_rate, self.replace)
        st["raw_synth"] = create_doublets(raw, st["choices"])
        aug, aug_lib, median = normalise(st["raw_synth"], pro["lib_size"], pro["normed"], self.pseudocount)
        st["median"] = median
        st["aug_lib_size"] = aug_lib
        if self.standard_scaling is True:
            if sp_sparse.issparse(aug):
                # sc.pp.scale zero-centres, and zero-centring a sparse matrix densifies it first (scanpy >= 1.10:
                # "... as `zero_center=True`, sparse input is densified"): from here on the pseudocount == 1 branch is dense,
                # and :308 picks svd_solver="auto" for it.  [upstream, absent -- restated from scanpy's behaviour]
                aug = np.asarray(aug.toarray(), dtype=np.float32)
            aug, _, _ = upstream.pp_scale(aug, max_value=15)
        if self.keep_stages:
            st["aug"] = aug
        adata = upstream.AnnDataLite(aug)
        solver = "arpack" if sp_sparse.issparse(aug) else "auto"
        emb, _ = upstream.tl_pca(aug, self.n_components, random_state=self.random_state, svd_solver=solver)
        adata.obsm["X_pca"] = emb
        st["X_pca"] = emb
        if self.clustering_algorithm == "phenograph":  # :318-327
            full, graph = upstream.phenograph_cluster(emb, seed=self.random_state, louvain_fn=self.louvain_fn,
                                                      **self.clustering_kwargs)
            st["fullcommunities"] = full
            if self.keep_stages:
                st["jaccard_graph"] = graph
            st["scores"], st["log_p"], st["communities"], st["synth_communities"] = score_communities(full, num_cells)
            return st
        leiden = self.clustering_algorithm == "leiden"  # :337-340
        upstream.pp_neighbors(adata, random_state=self.random_state, method="umap", n_neighbors=10, with_weights=leiden)
        st["knn_indices"] = adata.uns["knn_indices"]
        st["knn_distances"] = adata.uns["knn_distances"]
        if leiden:
            if self.keep_stages:
                st["connectivities"] = adata.obsp["connectivities"]
            upstream.tl_leiden(adata, key_added="clusters", random_state=self.random_state, leiden_fn=self.leiden_fn,
                               **self.clustering_kwargs)
        else:
            upstream.tl_louvain(
                adata, key_added="clusters", random_state=self.random_state, louvain_fn=self.louvain_fn,
                **self.clustering_kwargs,
            )
        full = np.array(adata.obs["clusters"], dtype=int)
        st["fullcommunities"] = full
        st["scores"], st["log_p"], st["communities"], st["synth_communities"] = score_communities(full, num_cells)
        return st

    def predict(self, p_thresh=1e-7, voter_thresh=0.9):
        out = predict(self.all_log_p_values_, self.all_scores_, self.n_iters, p_thresh, voter_thresh)
        self.labels_ = out["labels"]
        if "voting_average" in out:
            self.voting_average_ = out["voting_average"]
        else:
            self.suggested_score_cutoff_ = out["suggested_score_cutoff"]
        return self.labels_

    def doublet_score(self):
        return doublet_score(self.all_log_p_values_, self.n_iters)
