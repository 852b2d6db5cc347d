// dd_internal.h -- handle layout, error plumbing and launch accounting shared by the translation
// units of libdd_b200.so.  Nothing here is part of the ABI (see include/dd_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dd_b200.h"

#define DD_ABI_VERSION 7

// padded leading dimension of the dense A x G matrix: rows start on 128-byte boundaries
static inline int64_t dd_round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

struct dd_kernel_stat {
    double total_ms = 0.0;
    int64_t launches = 0;
};

// one timed launch: events are recorded without synchronising and resolved when the timing is queried
struct dd_timed_launch {
    const char *name;  // string literal
    cudaEvent_t start, stop;
};

struct dd_tc_state;

// One clustering LANE of the fit loop: the device state of a first Louvain level (graph CSR, communities, totals, captured
// round sequence) with its own stream.  The level is a chain of ~300 dependent, latency-bound steps (10 ms at 125 k cells
// when it shares the GPU with the main stream's kernels, 4 ms alone) and uses a few percent of the SMs, so the levels of
// CONSECUTIVE ITERATIONS run concurrently, one lane each.  Lane 0 lives in the handle's own fields (stage-wise calls use
// it); dd_lv_swap() exchanges a lane with the handle's fields around the calls that work on it.
struct dd_lv_lane {
    int32_t *d_lv_off = nullptr, *d_lv_adj = nullptr, *d_lv_comm = nullptr, *d_lv_i32 = nullptr;
    double *d_lv_tot = nullptr, *d_lv_w = nullptr;
    int64_t cap_lv_n = 0, cap_lv_nnz = 0, cap_lv_w = 0;
    long long *d_lvw_wq = nullptr, *d_lvw_i64 = nullptr;
    int32_t *d_lvw_i32 = nullptr;
    int64_t cap_lvw_nnz = 0, cap_lvw_n = 0, lvw_bucket_n = -1;
    uint64_t lvw_bucket_seed = 0;
    void *lvw_graph_exec = nullptr;
    int64_t lvw_graph_n = -1, lvw_graph_launches = 0;
    double lvw_graph_gamma = 0.0;
    uint64_t lvw_graph_seed = 0;
    const void *lvw_graph_key[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t lv_bucket_n = -1;
    uint64_t lv_bucket_seed = 0;
    int32_t *h_lv_rounds = nullptr;
    void *lv_graph_exec = nullptr;
    int64_t lv_graph_n = -1, lv_graph_launches = 0;
    bool lv_graph_is_loop = false;
    double lv_graph_gamma = 0.0;
    uint64_t lv_graph_seed = 0;
    cudaStream_t stream = nullptr;
};

struct dd_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // clustering (first Louvain level + result copies) overlaps the next iteration
    cudaStream_t stream3 = nullptr;  // dense build of the next iteration, underneath this iteration's PCA tail and kNN
    cudaStream_t stream4 = nullptr;  // kNN of iteration i, concurrent with the PCA of iteration i + 1 (fit loop, unsharded)
    cudaEvent_t ev_pca_done[2] = {nullptr, nullptr};  // per embedding buffer: the PCA has written it
    cudaEvent_t ev_emb_free[2] = {nullptr, nullptr};  // per embedding buffer: the kNN that read it has finished
    cudaEvent_t ev_dense_done = nullptr, ev_gemms_done = nullptr;
    bool gemms_done_recorded = false;  // run_pca recorded ev_gemms_done right after its last pass over the matrix
    cudaEvent_t ev_knn_done = nullptr, ev_lv_done = nullptr, ev_lv_done2 = nullptr;  // lv_done per kNN list buffer
    int num_sms = 148;
    std::string err;

    // ---- raw counts (fit prologue) ----
    int64_t N = 0, G = 0, nnz = 0, cap_rows = 0, cap_nnz = 0;
    int32_t *d_indptr = nullptr, *d_indices = nullptr;
    float *d_data = nullptr;
    float *d_lib = nullptr;   // float32 row sums (_lib_size)
    double *d_l1 = nullptr;   // double sums of |x| (the L1 normaliser of sklearn)
    std::vector<float> h_lib;
    bool all_finite = true;  // no NaN / inf among the uploaded values (k_row_sums)
    bool counts_borrowed = false;  // dd_share_counts: the CSR / library-size buffers belong to another handle of this device
    bool nonneg = true;  // no negative value in the uploaded matrix (then L1 norms of row sums are additive)

    // ---- synthetics ----
    int64_t M = 0, cap_M = 0;
    int64_t *d_parents = nullptr;
    int32_t *d_sindptr = nullptr, *d_scount = nullptr;
    int32_t *d_sindices = nullptr;
    float *d_sdata = nullptr;
    int64_t cap_snnz = 0, snnz = -1;
    float *d_slib = nullptr;
    bool synth_csr_valid = false;

    // ---- cell-block sharding (comm.cu; config c5).  With world == 1 or sharding off the block is everything.
    // This rank builds and factorises the dense rows of originals [blk_n0, blk_n0 + blk_n) followed by synthetics
    // [blk_m0, blk_m0 + blk_m): A = blk_n + blk_m LOCAL rows, A_glob = N + M rows over all ranks.
    int world = 1, rank = 0;
    bool shard_cells = false;
    void *nccl_comm = nullptr;  // ncclComm_t
    int64_t blk_n0 = 0, blk_n = 0, blk_m0 = 0, blk_m = 0, A_glob = 0;

    // ---- dense log-normalised augmented matrix ----
    int64_t A = 0, ld = 0, cap_dense = 0;  // A (local) rows, leading dimension ld >= G (multiple of 32)
    float *d_dense = nullptr;
    bool dense_valid = false;
    double *d_colsum = nullptr, *d_colsumsq = nullptr;  // G doubles each
    int64_t cap_cols = 0;

    // ---- PCA workspace ----
    int32_t L = 0, LP = 0, C = 0;
    int64_t cap_pca_rows = 0, cap_pca_cols = 0;
    int32_t cap_LP = 0;
    float *d_Qt = nullptr;     // LP x ld  (Q transposed, K-major for both GEMM flavours)
    float *d_Y = nullptr;      // A x LP
    double *d_Zacc = nullptr;  // G x LP accumulators of D^T Y
    double *d_small = nullptr; // scratch for L x L matrices, sums, flags (see pca.cu)
    // tcgen05 path: small GEMM operands as canonical hi/lo UMMA tiles (see pca_tc.h) + TMA tensor maps
    uint8_t *d_qb = nullptr, *d_yb = nullptr, *d_omega_b = nullptr;
    int64_t cap_qb = 0, cap_yb = 0, cap_omega_b = 0;
    float *d_mu = nullptr;  // float32 column means of the dense matrix (ld entries, 0 in the pad columns)
    int64_t cap_mu = 0;
    dd_tc_state *tc = nullptr;
    float *d_emb = nullptr;    // A x KP embedding (KP = 32 or 64, zero padded): the buffer in use, d_emb_base + {0, emb_stride}
    float *d_emb_base = nullptr;  // two buffers: the kNN of iteration i reads one while the PCA of i + 1 writes the other
    int64_t emb_stride = 0;
    int32_t KP = 0;
    int64_t emb_rows = 0;
    bool emb_valid = false;
    int64_t cap_emb = 0;

    // ---- kNN outputs ----
    int32_t *d_knn_idx = nullptr;       // the list buffer in use: d_knn_idx_base + {0, knn_idx_stride}
    int32_t *d_knn_idx_base = nullptr;  // two buffers (dd_knn_flip): clustering of iteration i overlaps the kNN of i + 1
    int64_t knn_idx_stride = 0;
    float *d_knn_dist = nullptr;
    int64_t cap_knn = 0;
    int32_t knn_last_k = 0;  // row stride of the lists / distances written last (the graph hooks refuse another k)
    uint8_t *d_knn_ops = nullptr;  // tcgen05 operand tiles (queries, candidates) in UMMA canonical layout
    int64_t cap_knn_ops = 0;
    // list-driven kernel (dd_knn_listed, cluster-ordered kNN): candidate-tile lists per 256-row query block; knn_list_pairs > 0 while such a call runs
    int32_t *d_knn_list_off = nullptr, *d_knn_list_tiles = nullptr;
    // cluster-ordered kNN (knn_prune.cu: dd_dev_knn_clustered): one carved allocation; the k-means centroids inside it are
    // carried from call to call while the number of rows stays the same
    uint8_t *d_knn_cl = nullptr;
    int64_t cap_knn_cl = 0, knn_cl_rows = -1;
    int knn_cl_tl = 0;
    int32_t *d_knn_cert = nullptr;      // [0] rows the filter certificate could not clear, [4..] their indices (knn.cu)
    int64_t cap_knn_cert = 0;
    int32_t *h_knn_uncert = nullptr;    // pinned copy of the count of the last kNN call
    int32_t *h_knn_cl_pairs = nullptr;  // pinned: block-tile pairs of launch A / B of the last completed cluster-ordered kNN
    int64_t cap_knn_list_off = 0, cap_knn_list_tiles = 0;
    int knn_list_pairs = 0;
    int knn_mode = 0;  // dd_set_knn_mode: 0 = by size, 1 = always all tiles, 2 = always the cluster-ordered path (tests)
    bool knn_narrow = false;  // 64-row pipeline steps / 256 TMEM columns: co-resident with a PCA product CTA (fit loop)

    // ---- GPU Louvain level 0 (louvain_gpu.cu): symmetric kNN pattern as CSR + community state ----
    int32_t *d_lv_off = nullptr, *d_lv_adj = nullptr, *d_lv_comm = nullptr, *d_lv_i32 = nullptr;
    double *d_lv_tot = nullptr;
    double *d_lv_w = nullptr;  // Jaccard edge weights of the PhenoGraph graph (one per adjacency entry)
    int64_t cap_lv_n = 0, cap_lv_nnz = 0, cap_lv_w = 0;
    float *d_umap_w = nullptr;  // Leiden branch: umap's directed membership strengths, one per kNN list entry
    int64_t cap_umap_w = 0;
    // weighted first level (louvain_gpu_w.cu): fixed-point weights, int64 degrees / totals, buckets, state
    long long *d_lvw_wq = nullptr, *d_lvw_i64 = nullptr;  // wq[nnz]; k | tot | two_m (2 n + 1)
    int32_t *d_lvw_i32 = nullptr;                          // csize | desired | bucket (n each) + counters
    int64_t cap_lvw_nnz = 0, cap_lvw_n = 0, lvw_bucket_n = -1;
    uint64_t lvw_bucket_seed = 0;
    std::vector<int32_t> lvw_colour_off;
    void *lvw_graph_exec = nullptr;  // cudaGraphExec_t of the captured weighted round sequence (louvain_gpu_w.cu)
    int64_t lvw_graph_n = -1, lvw_graph_launches = 0;
    double lvw_graph_gamma = 0.0;
    uint64_t lvw_graph_seed = 0;
    const void *lvw_graph_key[4] = {nullptr, nullptr, nullptr, nullptr};  // buffer addresses baked into the capture
    int64_t lv_bucket_n = -1;
    uint64_t lv_bucket_seed = 0;
    std::vector<int32_t> lv_colour_off;
    int32_t *h_lv_rounds = nullptr;  // pinned: rounds the last first level ran (reported as stage "lv_rounds")
    void *lv_graph_exec = nullptr;  // cudaGraphExec_t of the captured round sequence
    int64_t lv_graph_n = -1, lv_graph_launches = 0;
    bool lv_graph_is_loop = false;
    double lv_graph_gamma = 0.0;
    uint64_t lv_graph_seed = 0;

    std::vector<dd_lv_lane> lv_lanes;             // lanes 1.. (lane 0 = the fields above + stream2)
    cudaEvent_t ev_after_graph_build = nullptr;   // if set, dd_dev_louvain_level0 records it once the kNN lists have been read

    // ---- pinned host slots of the pipelined fit loop (kNN graph + PCA flag per in-flight iteration) ----
    std::vector<int32_t *> slot_knn;
    std::vector<double *> slot_flag;
    int64_t slot_knn_elems = 0;

    // ---- accounting ----
    int64_t launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t stage_ev0 = nullptr, stage_ev1 = nullptr;
    std::map<std::string, dd_kernel_stat> kstats;
    std::vector<dd_timed_launch> pending;    // recorded, not yet resolved
    std::vector<cudaEvent_t> event_pool;     // recycled events
    std::map<std::string, double> stage_ms;
};

// cudaFuncSetAttribute applies to the CURRENT device only: a call site remembers per device (not per process) that it has
// configured its kernels, so handles on several GPUs of one process all get their opt-in shared-memory sizes.  The lock is
// held while the attributes are set (another thread must not launch on that device before they are in place).
struct dd_once_per_device {
    std::mutex mu;
    uint64_t done[4] = {0, 0, 0, 0};  // device ordinals 0..255
    template <class F>
    void run(int device, F &&f) {
        std::lock_guard<std::mutex> lk(mu);
        const unsigned d = (unsigned)device & 255u;
        if (done[d >> 6] & (1ull << (d & 63))) return;
        f();
        done[d >> 6] |= 1ull << (d & 63);
    }
};

// thread-local message for failures that happen without a handle
void dd_set_global_error(const std::string &msg);
int dd_fail(dd_handle *h, int code, const std::string &msg);

#define DD_CUDA(h, expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return dd_fail((h), _e == cudaErrorMemoryAllocation ? DD_ERR_NOMEM : DD_ERR_CUDA,  \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

#define DD_TRY(expr)                 \
    do {                             \
        int _rc = (expr);            \
        if (_rc != DD_OK) return _rc; \
    } while (0)

// Launch accounting: every kernel goes through DD_LAUNCH so that dd_kernel_launches() is exact and
// per-kernel device time can be collected (dd_set_kernel_timing).
void dd_launch_begin(dd_handle *h);
int dd_launch_end(dd_handle *h, const char *name);

#define DD_LAUNCH(h, name, kernel, grid, block, smem, ...)                  \
    do {                                                                    \
        dd_launch_begin(h);                                                 \
        kernel<<<(grid), (block), (smem), (h)->stream>>>(__VA_ARGS__);      \
        DD_TRY(dd_launch_end((h), (name)));                                 \
    } while (0)

// stage timers (CUDA events on the handle's stream)
int dd_stage_begin(dd_handle *h);
int dd_stage_end(dd_handle *h, const char *stage);

// grow-only device buffers
template <typename T>
static inline int dd_reserve(dd_handle *h, T **p, int64_t *cap, int64_t need) {
    if (need <= *cap && *p) return DD_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (size_t)(need > 0 ? need : 1));
    if (e != cudaSuccess) return dd_fail(h, DD_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    *cap = need;
    return DD_OK;
}

// ---- cell-block sharding helpers (comm.cu).  All are no-ops when the handle is not sharded; otherwise they
// enqueue NCCL collectives on h->stream (every rank must issue the same sequence).
static inline bool dd_sharded(const dd_handle *h) { return h->shard_cells && h->world > 1; }
void dd_set_block(dd_handle *h);  // derive blk_* / A / A_glob from N, M, rank, world
int dd_comm_allreduce_f64(dd_handle *h, double *buf, int64_t count);
int dd_comm_bcast(dd_handle *h, void *buf, int64_t bytes, int root);
int dd_comm_allreduce_i32(dd_handle *h, int32_t *buf, int64_t count);
// "all-gather" of row ranges that already sit at their final place in a replicated buffer: range r of `begin` /
// `count` (in rows of row_bytes bytes) is owned by rank owner[r] and broadcast from there in one NCCL group
int dd_comm_gather_ranges(dd_handle *h, void *base, int64_t row_bytes, int n_ranges, const int64_t *begin,
                          const int64_t *count, const int *owner);
void dd_comm_destroy(dd_handle *h);

// ---- stage entry points implemented across the .cu files (all asynchronous on h->stream) ----
int dd_dev_create_doublets_csr(dd_handle *h);                       // csr.cu
int dd_dev_build_dense(dd_handle *h, float median, float pseudocount);  // csr.cu (fused pair-add + normalise)
int dd_dev_colstats(dd_handle *h, bool with_sq);                    // scale.cu
int dd_dev_standard_scale(dd_handle *h, float max_value);           // scale.cu
int dd_dev_pca(dd_handle *h, int32_t n_comp, int32_t n_random, int32_t n_power_iter,
               const float *omega_host);                            // pca.cu
void dd_drop_borrowed_counts(dd_handle *h);                         // csr.cu
int dd_dev_knn(dd_handle *h, int32_t k);                            // knn.cu
bool dd_knn_clustered_applies(const dd_handle *h, int32_t k);        // knn_prune.cu
int dd_dev_knn_clustered(dd_handle *h, int32_t k);                  // knn_prune.cu
int dd_emb_reserve(dd_handle *h, int64_t rows, int32_t KP);         // pca.cu: (re)allocate the two embedding buffers
int dd_dev_louvain_level0(dd_handle *h, int32_t k, double gamma, uint64_t seed);  // louvain_gpu.cu
int dd_dev_jaccard_graph(dd_handle *h, int32_t k, int prune);                      // louvain_gpu.cu (PhenoGraph)
void dd_lv_swap(dd_handle *h, dd_lv_lane &lane);                                   // handle.cu
void dd_lv_lane_free(dd_lv_lane &lane);                                            // handle.cu

// host pieces (louvain.cpp / score.cpp)
int dd_host_louvain_knn(int64_t n, int32_t k, const int32_t *knn_idx, double resolution, uint64_t seed,
                        int32_t *labels_out, int32_t *n_comm_out);
// upper Louvain levels from a first-level partition: off/adj = symmetric pattern CSR, comm0 = community per node
int dd_host_louvain_from_level0(int64_t n, const int32_t *off, const int32_t *adj, const int32_t *comm0,
                                double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_comm_out);
// PhenoGraph on the host from the device-built weighted graph (rows in any order, zero weights = pruned edges):
// Louvain at resolution 1 on the weighted graph, labels by decreasing size, communities <= min_cluster_size -> -1
// comm0 (may be NULL): the first level already done on the device (default; DD_PHENO_LEVEL0=0 turns it off)
int dd_host_phenograph_from_graph(int64_t n, const int32_t *off, const int32_t *adj, const double *w, uint64_t seed,
                                  int32_t min_cluster_size, int32_t *labels_out, int32_t *n_comm_out, const int32_t *comm0);
int dd_dev_louvain_level0_weighted(dd_handle *h, double gamma, uint64_t seed);  // louvain_gpu_w.cu
// Leiden on the umap-weighted neighbour graph (leiden.cpp): kNN lists with self in column 0 + float32 distances
int dd_host_leiden_knn(int64_t n, int32_t k, const int32_t *knn_idx, const float *knn_dist, double resolution,
                       uint64_t seed, int32_t *labels_out, int32_t *n_comm_out);
// the same from the umap graph built on the device (dd_dev_umap_graph: rows in any order, zero weights = no edge)
int dd_dev_umap_graph(dd_handle *h, int32_t k);  // louvain_gpu.cu
int dd_host_leiden_from_graph(int64_t n, const int32_t *off, const int32_t *adj, const double *w, double resolution,
                              uint64_t seed, int32_t *labels_out, int32_t *n_comm_out);
int dd_host_umap_canonical(int64_t n, const int32_t *off, const int32_t *adj, const double *w, int64_t *indptr_out,
                           int32_t *indices_out, float *weights_out, int64_t capacity, int64_t *nnz_out);
float dd_host_median(std::vector<float> &v);
