/* dd_b200.h -- C ABI of libdd_b200.so: the B200-native replacement for the per-iteration hot path
 * of DoubletDetection's BoostClassifier.fit (reference: doubletdetection/doubletdetection.py).
 *
 * The reference is pure Python and has no FFI of its own; the seam this ABI replaces is the
 * private pair BoostClassifier._one_fit() / _createDoublets() (doubletdetection.py:274-402) and
 * the loop around it (doubletdetection.py:186-198).  Each entry point below names the reference
 * lines whose arithmetic it performs.  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success, non-zero on error
 *    (DD_ERR_*); the message is available from dd_last_error(h) (or dd_last_error(NULL) when
 *    no handle exists yet).  No exceptions cross the ABI.
 *  - host pointers unless the name says "_dev"; the caller owns every buffer it passes in and
 *    every output buffer (the library copies what it needs to the device).
 *  - a handle is bound to one CUDA device and is NOT thread-safe; different handles are
 *    independent.  Long calls do not touch Python, so ctypes releases the GIL around them.
 *  - there is no CPU fallback: without a CUDA device dd_create fails with DD_ERR_CUDA.
 *  - matrices are row-major.  The augmented matrix has A = N + M rows: the N original cells
 *    first, then the M synthetic doublets (doubletdetection.py:292).
 */
#ifndef DD_B200_H
#define DD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DD_OK 0
#define DD_ERR_ARG 1     /* bad argument / call order            */
#define DD_ERR_CUDA 2    /* CUDA runtime or driver error         */
#define DD_ERR_UNSUPPORTED 3 /* shape / option outside the hot path */
#define DD_ERR_NOMEM 4

typedef struct dd_handle dd_handle;

/* ---- lifetime ------------------------------------------------------------------------- */
int dd_create(int device, dd_handle **out);
void dd_destroy(dd_handle *h);
const char *dd_last_error(const dd_handle *h);
/* ABI version of the loaded library (bumped on any signature change). */
int dd_abi_version(void);

/* ---- fit() prologue, doubletdetection.py:178-184 ---------------------------------------
 * Upload the (already float32, canonical: sorted, duplicate-free) CSR count matrix
 * `_raw_counts` (N cells x G genes).  Computes on the device `_lib_size` (:182) and keeps the
 * raw rows resident; the memoised L1-normalised copy (:183-184) is never materialised -- the
 * normalise kernel divides by the row sum on the fly with the same rounding. */
int dd_upload_counts(dd_handle *h, int64_t n_cells, int64_t n_genes, const int32_t *indptr,
                     const int32_t *indices, const float *data);
/* A second pipeline on the same GPU (the Python shim runs two fit loops per device, each on half of the iterations, so that
 * one loop's latency-bound PCA kernels run underneath the other's HBM-bound products): `dst` reads `src`'s resident count
 * matrix and library sizes instead of holding a copy.  `src` must outlive the use and must not be re-uploaded meanwhile. */
int dd_share_counts(dd_handle *dst, const dd_handle *src);
/* 1 if every uploaded value is finite (checked on the device while the library sizes are summed): the shim then skips the
 * host-side scan of check_array(ensure_all_finite=True) (:149-155) and only runs it to raise sklearn's own error. */
int dd_counts_all_finite(dd_handle *h, int32_t *all_finite_out);
/* `_lib_size` (float32[N]) as computed on the device. */
int dd_get_lib_size(dd_handle *h, float *lib_size_out);

/* ---- fit() prologue, highly variable genes, doubletdetection.py:165-176 -----------------
 * dd_hvg_variances: `gene_variances` (:166-169) of the uploaded matrix, float32[G], bit for bit what scipy's
 * `raw.power(2).mean(axis=0) - raw.mean(axis=0) ** 2` returns (one float32 accumulator per gene, entries added in row
 * order after the multiplication by float32(1/N)).  The caller runs `np.argsort(...)[-n_top:]` (:170-171: numpy's own
 * sort decides ties) and hands the result to
 * dd_select_genes: `raw = raw.tocsc()[:, top_var_genes_].tocsr()` (:173-175) on the device -- column j of the new
 * matrix is gene genes[j], rows keep sorted column ids -- followed by `_lib_size` (:182) of the subset. */
int dd_hvg_variances(dd_handle *h, float *variances_out);
int dd_select_genes(dd_handle *h, int64_t n_selected, const int64_t *genes);
/* The count matrix the handle holds (after dd_select_genes: the subset), for inspection: nnz, then indptr int32[N+1],
 * indices int32[nnz], data float[nnz]. */
int dd_counts_nnz(dd_handle *h, int64_t *nnz_out);
int dd_download_counts(dd_handle *h, int32_t *indptr_out, int32_t *indices_out, float *data_out);

/* ---- _createDoublets(), doubletdetection.py:397-399 ------------------------------------
 * parents: int64[M*2] (the `choices` array of :394, drawn by the caller from its PCG64 stream).
 * Builds `_raw_synthetics` = raw[parents[:,0]] + raw[parents[:,1]] as canonical CSR on the
 * device (sorted merge, values added where both parents express a gene). */
int dd_create_doublets(dd_handle *h, int64_t n_synth, const int64_t *parents);
int dd_synth_nnz(dd_handle *h, int64_t *nnz_out);
/* Copy the synthetic CSR back: indptr int32[M+1], indices int32[nnz], data float[nnz]. */
int dd_download_synthetics(dd_handle *h, int32_t *indptr_out, int32_t *indices_out, float *data_out);
/* Synthetic library sizes (:288), float32[M]. */
int dd_get_synth_lib_size(dd_handle *h, float *lib_size_out);

/* ---- _one_fit() normalisation, doubletdetection.py:286-295 -----------------------------
 * Dense float32 A x G matrix  log(x / rowsum * median + pseudocount)  for originals followed by
 * synthetics.  `median` = np.median(aug_lib_size) (:293) computed by the caller (or by
 * dd_median_lib_size below, same value).  pseudocount == 1 (the sparse log1p branch, :296-297)
 * is DD_ERR_UNSUPPORTED. */
int dd_median_lib_size(dd_handle *h, float *median_out);
int dd_normalise_log(dd_handle *h, float median, float pseudocount);
/* optional sc.pp.scale(max_value) on the dense matrix, doubletdetection.py:302-303.
 * max_value <= 0 means no clipping. */
int dd_standard_scale(dd_handle *h, float max_value);
/* Copy rows [row0, row0+n_rows) of the dense matrix back (float32, n_rows x G, row-major). */
int dd_download_dense(dd_handle *h, int64_t row0, int64_t n_rows, float *out);
/* Test hook: replace the dense matrix by a caller-supplied one (A x G float32). */
int dd_upload_dense(dd_handle *h, int64_t n_rows, int64_t n_genes, const float *dense);

/* ---- sc.tl.pca(..., svd_solver="auto") == sklearn randomized PCA, doubletdetection.py:309-314
 * omega: float32[R * n_random] row-major, the Gaussian test matrix sklearn draws
 * (RandomState(random_state).normal(size=(R, n_comp+10)) cast to float32) with R = G, or R = A when
 * there are fewer augmented cells than genes (sklearn then factorises the transposed matrix).
 * n_power_iter = 7 or 4 (sklearn's "auto").  Leaves the float32 A x n_comp embedding on the device; emb_out may be
 * NULL.  singular_values_out (float64[n_comp]) may be NULL. */
int dd_pca(dd_handle *h, int32_t n_comp, int32_t n_random, int32_t n_power_iter, const float *omega,
           float *emb_out, double *singular_values_out);
/* ---- sklearn's EXACT PCA branches (svd_solver "covariance_eigh" / "full", picked by "auto" for <= 1000 genes with
 * >= 10x as many augmented cells, for matrices with max(shape) <= 500 and for n_components >= 0.8 min(shape):
 * sklearn/decomposition/_pca.py:524-536, 560-640; reached from doubletdetection.py:309-314) and svd_solver="arpack", which
 * doubletdetection.py:308 selects when pseudocount == 1 keeps the matrix sparse (:296-297; dd_normalise_log then computes
 * log1p, and the zeros of the dense matrix are the gaps of the sparse one).  The device does what scales with the number of
 * cells; the caller solves the small symmetric eigenproblem in between (scipy.linalg.eigh, top n_comp pairs, in the shim).
 * The Gram side is limited to 16384 rows.
 *   dd_centered_gram  float64 Gram matrix of the centred dense matrix: transposed == 0 -> G x G (sum over cells, i.e.
 *                     (A - 1) x the covariance matrix), else A x A (sum over genes); out: n x n row-major, symmetric
 *   dd_project        X_pca = (D - mean) V for sign-fixed components V (float64[G x n_comp], row-major); leaves the
 *                     float32 embedding on the device for dd_knn like dd_pca does; emb_out (A x n_comp) may be NULL */
int dd_centered_gram(dd_handle *h, int32_t transposed, double *out);
int dd_project(dd_handle *h, int32_t n_comp, const double *components, float *emb_out);
/* Test hook: replace the embedding by a caller-supplied one (A x n_comp float32). */
int dd_upload_embedding(dd_handle *h, int64_t n_rows, int32_t n_comp, const float *emb);

/* ---- sc.pp.neighbors(n_neighbors=k) exact kNN, doubletdetection.py:331-336 -------------
 * Exact Euclidean k nearest neighbours of every augmented cell in the embedding; column 0 is
 * the cell itself (distance 0), then the k-1 nearest others by (distance, index).  2 <= k <= 31 (PhenoGraph: 30 + self).
 * A 3xBF16 tcgen05 distance GEMM filters 16 (k <= 13) or 40 candidates per row, which are re-ranked in float64.
 * idx_out int32[A*k], dist_out float32[A*k] (may be NULL). */
int dd_knn(dd_handle *h, int32_t k, int32_t *idx_out, float *dist_out);

/* Test hook of the list-driven kernel (no counterpart in the reference): the same exact kNN, but the 256-row query block p only visits
 * the 128-row candidate tiles list_tiles[list_off[p] .. list_off[p + 1]) (host arrays, n_blocks = ceil(ceil(A / 128) / 2)
 * lists).  The caller guarantees that the lists are sufficient -- scripts/knn_listed_experiment.py derives them from
 * bounding boxes of a cluster-ordered embedding; neighbours missing from a list are simply not
 * found (-1 where fewer than k - 1 exist).  Outputs as dd_knn. */
int dd_knn_listed(dd_handle *h, int32_t k, int64_t n_blocks, const int32_t *list_off, const int32_t *list_tiles,
                  int32_t *idx_out, float *dist_out);

/* Test hook, one step further: both launches of the cluster-ordered kNN on the device with a caller-made ordering.  perm int32[n_pad] maps a padded
 * position to an original row or -1 (n_pad a multiple of 256; rows of a group contiguous, every group padded to whole
 * 256-row blocks), block_group int32[n_pad / 256] names the group of every block.  Launch A (own group) bounds every
 * query's k-th distance, launch B visits the tiles whose bounding boxes the bound cannot exclude, the result is re-ranked
 * in the original numbering: identical to dd_knn (which must have been called once on this embedding: it sizes the output
 * buffers).  stats_out int64[4] (may be NULL): block-tile pairs of launch A, of launch B, blocks, tiles.  k <= 13. */
int dd_knn_pruned(dd_handle *h, int32_t k, int64_t n_pad, const int32_t *perm, const int32_t *block_group, int32_t *idx_out,
                  float *dist_out, int64_t *stats_out);
/* The fit loop's kNN for large embeddings (>= 50 000 rows, k <= 13; doubletdetection.py:331-336 at BASELINE sizes) is the
 * cluster-ordered form of the same exact search: rows grouped by a device k-means, groups padded to 256-row blocks, every
 * block first against its own group, then against the tiles of other groups its bounding box cannot exclude (valid lower
 * bounds: identical neighbours to the all-tiles kernel).  dd_set_knn_mode: 0 = choose by size, 1 = always all tiles, 2 =
 * always cluster-ordered.  dd_knn_clustered_stats: block-tile pairs visited by the last such call: [launch A, launch B,
 * blocks in use, tiles in use]. */
int dd_set_knn_mode(dd_handle *h, int32_t mode);
/* Exactness: the tcgen05 filter's scores carry an error of at most 2^-16 |q| |c|; the re-ranking kernel certifies every row
 * (no excluded candidate can be closer than the reported k-th neighbour) and the rows it cannot clear are re-done by float64
 * brute force on the device.  dd_knn_uncertified: how many rows of the last kNN call took that route. */
int dd_knn_uncertified(dd_handle *h, int64_t *count_out);
int dd_knn_clustered_stats(dd_handle *h, int64_t *stats_out);


/* ---- clustering call, doubletdetection.py:337-343 -----------------------------------------
 * Louvain (RB configuration null model, resolution gamma, unweighted, seeded) on the symmetrised
 * kNN pattern -- what sc.tl.louvain(resolution, random_state, directed=False) optimises.  The
 * louvain package is absent from the image; the algorithm is specified in oracle/louvain_ref.py:
 * the first level by synchronous coloured rounds (on the GPU inside dd_fit_iterations; this host
 * entry computes the identical partition on the CPU), the levels above sequentially.
 * labels_out int32[n], 0 = largest community.  No handle: pure host code, thread-safe. */
int dd_louvain_knn(int64_t n, int32_t k, const int32_t *knn_idx, double resolution, uint64_t seed,
                   int32_t *labels_out, int32_t *n_communities_out);
/* phenograph.cluster(X_pca, prune=...)[0], doubletdetection.py:320 (defaults k = 30, jaccard=True,
 * min_cluster_size = 10), from the exact kNN lists knn_idx int32[n * k] (self in column 0, so k = 31 for
 * PhenoGraph's k = 30): Jaccard graph w_ij = s / (2(k-1) - s), s = shared neighbours; prune keeps mutual edges
 * with the product of the two directed weights, otherwise the average; Louvain at resolution 1 on the weighted
 * graph; labels by decreasing size, communities of at most min_cluster_size cells get -1 (phenograph keeps sizes > min_size).  The phenograph package
 * (absent from the image) runs its bundled Louvain binaries repeatedly with time-based seeds; this is ONE seeded
 * run of the in-repo Louvain (specification: oracle/upstream.py phenograph_cluster).  Pure host code: the host
 * twin of what dd_fit_iterations does with the graph built on the device. */
int dd_phenograph_knn(int64_t n, int32_t k, const int32_t *knn_idx, int32_t prune, int32_t min_cluster_size,
                      uint64_t seed, int32_t *labels_out, int32_t *n_communities_out);
/* The PhenoGraph graph of the lists of the last dd_knn(h, k, ...) as built on the device: symmetric pattern CSR
 * (rows unordered) + weights (0 = pruned).  Call with capacity 0 to get nnz_out, then with buffers of that size. */
int dd_jaccard_graph(dd_handle *h, int32_t k, int32_t prune, int32_t *indptr_out, int32_t *indices_out,
                     double *weights_out, int64_t capacity, int64_t *nnz_out);
/* Fully sequential Louvain on an explicit symmetric CSR graph (weights may be NULL = unweighted). */
int dd_louvain_csr(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                   double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_communities_out);
/* The same graph partitioned the way the kNN pipeline does it: first level by synchronous coloured rounds (the host twin
 * of the device level), the levels above sequentially.  With weights the first level works in fixed point (multiples of
 * 2^-32 summed in int64: exact, hence order-independent sums) -- the specification of the device level that
 * dd_fit_iterations runs for PhenoGraph's weighted graph (oracle/louvain_ref.py). */
int dd_louvain_csr_level0(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                          double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_communities_out);

/* Test hook: the weighted first level of dd_louvain_csr_level0 on the DEVICE (what dd_fit_iterations runs for PhenoGraph) -- fixed-point weights, 64-bit integer atomics -- for an explicit symmetric CSR graph
 * (host arrays, positive weights, no self-loops).  comm_out int32[n] = community (a node id) of every node after the level,
 * to be compared with oracle/louvain_ref.py:level0_parallel (tests/gpu_weighted_level_check.py). */
int dd_louvain_level0_weighted(dd_handle *h, int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                               double gamma, uint64_t seed, int32_t *comm_out, int32_t *rounds_out);

/* ---- clustering_algorithm="leiden", doubletdetection.py:331-342 (host) -------------------------------------
 * sc.pp.neighbors(method="umap", n_neighbors=k) weights + sc.tl.leiden(resolution, random_state, directed=False).
 * dd_umap_connectivities: umap's fuzzy simplicial set of the exact kNN lists (knn_idx int32[n * k] with the cell
 * itself in column 0, knn_dist float32[n * k]): smooth_knn_dist (rho = nearest positive distance, sigma by
 * bisection so that the memberships of the k - 1 neighbours sum to log2(k)), membership strengths
 * exp(-(d - rho) / sigma), fuzzy union W + W^T - W o W^T in float32, zeros dropped; CSR with ascending rows.
 * Call with capacity 0 to get nnz_out, then with buffers of that size (indptr_out int64[n + 1]).
 * dd_leiden_knn: that graph partitioned by the Leiden algorithm (RB-configuration quality, resolution gamma,
 * edge weights used, iterated until stable like leidenalg's n_iterations=-1), labels by decreasing community
 * size.  umap-learn / leidenalg are absent from the image: the arithmetic of the weights is pinned by
 * oracle/upstream.py, the move order by oracle/leiden_ref.py (parity with leidenalg itself is unpinned).  This is
 * what dd_fit_iterations runs on its host workers for DD_CLUSTER_LEIDEN; the GRAPH of every iteration is built on the
 * device (dd_umap_graph below is that stage's test hook), dd_umap_connectivities is its host twin.
 * dd_leiden_csr: the same partitioning of an explicit symmetric CSR graph (weights may be NULL = unweighted).
 * No handle: pure host code, thread-safe. */
int dd_umap_connectivities(int64_t n, int32_t k, const int32_t *knn_idx, const float *knn_dist,
                           int64_t *indptr_out, int32_t *indices_out, float *weights_out, int64_t capacity,
                           int64_t *nnz_out);
int dd_leiden_knn(int64_t n, int32_t k, const int32_t *knn_idx, const float *knn_dist, double resolution,
                  uint64_t seed, int32_t *labels_out, int32_t *n_communities_out);
int dd_leiden_csr(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                  double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_communities_out);
/* umap's connectivities of the lists + distances of the last dd_knn(h, k, ...) as built on the DEVICE (one thread per
 * cell runs smooth_knn_dist's bisection and the membership strengths, one warp per cell the fuzzy union on the symmetric
 * pattern), returned in dd_umap_connectivities' format (ascending rows, zeros dropped, indptr_out int64[n + 1]).  Call
 * with capacity 0 to get nnz_out, then with buffers of that size. */
int dd_umap_graph(dd_handle *h, int32_t k, int64_t *indptr_out, int32_t *indices_out, float *weights_out,
                  int64_t capacity, int64_t *nnz_out);
/* The second half of the fit loop's Leiden stage on its own (host): the graph in the layout the device leaves in the
 * pinned slot -- off int32[n + 1], adj int32[off[n]] with the rows in ANY order, weights float64 with 0 = no edge -- is
 * put into canonical form (rows ascending, zeros dropped) and partitioned like dd_leiden_knn. */
int dd_leiden_device_graph(int64_t n, const int32_t *off, const int32_t *adj, const double *weights, double resolution,
                           uint64_t seed, int32_t *labels_out, int32_t *n_communities_out);

/* ---- scoring, doubletdetection.py:344-383 (host) ----------------------------------------
 * labels int32[n_cells + n_synth]; scores_out / log_p_out float64[n_cells]:
 * synth fraction of the cell's community and hypergeom.logsf(k, A, M, size); NaN where the
 * label is negative. */
int dd_score(int64_t n_cells, int64_t n_synth, const int32_t *labels, double *scores_out, double *log_p_out);
/* scipy.stats.hypergeom.logsf(k, M, n, N) for one argument tuple. */
double dd_hypergeom_logsf(int64_t k, int64_t M, int64_t n, int64_t N);

/* ---- the loop of fit(), doubletdetection.py:192-198 -------------------------------------
 * Runs n_iters iterations of _one_fit for the uploaded counts: GPU stages back to back on the
 * handle's stream, clustering + scoring of finished iterations on `n_host_threads` host workers
 * concurrently with the GPU work of the next ones.
 *   parents      int64[n_iters * M * 2]   all iterations' `choices`, drawn sequentially by the caller
 *   omega        float32[G * n_random]
 * outputs (caller allocated):
 *   scores_out, log_p_out   float64[n_iters * N]
 *   communities_out         int32[n_iters * N]
 *   synth_communities_out   int32[n_iters * M]
 *   stage_ms_out            float64[8] or NULL: milliseconds summed over iterations for
 *                           {host aggregation + upper Louvain levels + scoring (summed over workers),
 *                            doublets+normalise, scale, pca, knn, first Louvain level on the device + d2h,
 *                            device time first launch -> last copy, wall time of the call}
 */
typedef struct dd_fit_params {
    int32_t n_iters;
    int64_t n_synth;
    float pseudocount;
    int32_t standard_scaling; /* 0 / 1 */
    float scale_max_value;    /* 15 in the reference */
    int32_t n_comp;
    int32_t n_random;
    int32_t n_power_iter;
    int32_t knn_k;            /* 10 in the reference */
    double resolution;        /* 4 in the reference  */
    uint64_t seed;            /* random_state        */
    int32_t n_host_threads;
    int32_t iter_begin;       /* run iterations [iter_begin, iter_end) of the n_iters drawn */
    int32_t iter_end;
    int32_t clustering;       /* DD_CLUSTER_LOUVAIN / DD_CLUSTER_LEIDEN (:329-343) or DD_CLUSTER_PHENOGRAPH (:318-327) */
    int32_t pheno_k;          /* phenograph.cluster k (30): neighbours per cell, self excluded */
    int32_t pheno_prune;      /* 1: keep mutual edges, weight product (the reference's default); 0: average */
    int32_t pheno_min_cluster_size; /* 10: communities of <= 10 cells are labelled -1 (NaN scores, :379-381) */
} dd_fit_params;
#define DD_CLUSTER_LOUVAIN 0
#define DD_CLUSTER_PHENOGRAPH 1
#define DD_CLUSTER_LEIDEN 2

int dd_fit_iterations(dd_handle *h, const dd_fit_params *p, const int64_t *parents, const float *omega,
                      double *scores_out, double *log_p_out, int32_t *communities_out,
                      int32_t *synth_communities_out, double *stage_ms_out);

/* ---- cell-block sharding of one fit across the GPUs of a box (SURVEY 8e level 2; BASELINE config 5) ------
 * The reference has no multi-GPU path; this is the partition doubletdetection.py:292-336 admits: the rows of
 * the augmented matrix (vstack of originals and synthetics, :292) are dealt to `world` ranks, one process and
 * one handle per GPU.  Every rank uploads the SAME counts and passes the SAME parents / omega; it builds and
 * factorises only its block of rows.  Sums over cells (column means, D^T Y, the Gram matrix of the tall panel)
 * are NCCL all-reduces, the n_comp-dimensional embedding is all-gathered (the single exchange kNN needs), each
 * rank answers the kNN queries of its share of rows against all cells and the lists are all-gathered; the
 * clustering + scoring of iteration i then runs on rank i % world only: dd_fit_iterations fills the result rows of
 * the iterations a rank owns and leaves the others zero (the caller merges them, e.g. by an integer sum of the
 * bit patterns over the ranks).
 * NCCL is resolved at run time from the process (libnccl.so.2; env DD_NCCL_LIB overrides the path).
 *   dd_comm_unique_id   rank 0 creates the rendezvous token (DD_COMM_ID_BYTES bytes) and ships it to the
 *                       other ranks by any host channel (the Python shim uses torch.distributed.broadcast)
 *   dd_comm_init        collective: all ranks call it with the same token
 *   dd_comm_shard_cells 1 = shard cells across the communicator from the next dd_create_doublets /
 *                       dd_fit_iterations on; 0 = every rank works on the whole matrix again
 *   dd_comm_info        rank, world and this rank's block {first original, n originals, first synthetic,
 *                       n synthetics} (int64[4]) for the parents set last
 * On a sharded handle dd_download_dense addresses the LOCAL rows (originals of the block, then its
 * synthetics); dd_pca / dd_knn return the global embedding / lists on every rank. */
#define DD_COMM_ID_BYTES 128
int dd_comm_unique_id(void *id_out, int64_t id_bytes);
int dd_comm_init(dd_handle *h, int32_t rank, int32_t world, const void *id, int64_t id_bytes);
int dd_comm_shard_cells(dd_handle *h, int32_t on);
int dd_comm_info(const dd_handle *h, int32_t *rank_out, int32_t *world_out, int64_t *block_out);

/* ---- introspection used by bench.py -------------------------------------------------------
 * Number of kernels this library has launched on the handle since creation. */
int64_t dd_kernel_launches(const dd_handle *h);
/* Milliseconds of the most recent call of the named stage measured with CUDA events on the
 * handle's stream ("doublets", "normalise", "scale", "pca", "knn"); < 0 if unknown. */
double dd_last_stage_ms(const dd_handle *h, const char *stage);
/* Per-kernel accumulated device time when profiling is on: a CUDA event pair is recorded around each
 * launch on the handle's stream without synchronising, and resolved when the numbers are read. */
int dd_set_kernel_timing(dd_handle *h, int32_t on);
int dd_get_kernel_timing(dd_handle *h, const char *kernel, double *total_ms_out, int64_t *launches_out);
/* All kernels at once as text lines "name total_ms launches\n"; returns the buffer size needed
 * (pass buf == NULL to query it), < 0 on error. */
int64_t dd_kernel_timing_report(dd_handle *h, char *buf, int64_t buflen);

#ifdef __cplusplus
}
#endif
#endif /* DD_B200_H */
