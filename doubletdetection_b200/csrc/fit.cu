// fit.cu -- the loop of BoostClassifier.fit (doubletdetection.py:192-198) as one pipelined call:
// the GPU stages of _one_fit run back to back on the handle's stream (no host synchronisation between
// iterations), each iteration's kNN graph lands in a pinned host buffer, and a pool of host threads
// clusters + scores finished iterations while the GPU works on the next ones.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "dd_internal.h"

int dd_set_parents(dd_handle *h, int64_t n_synth, const int64_t *parents);  // csr.cu
int dd_pca_check(dd_handle *h);                                             // pca.cu
int dd_pca_flag_copy(dd_handle *h, double *host_flag);                      // pca.cu

namespace {

struct Slot {
    int32_t *graph = nullptr;  // pinned: [off (A + 1) | comm (A) | adj (<= A * 2 (k - 1)) | PhenoGraph / Leiden: weights (f64)]
    double *flag = nullptr;    // pinned, PCA breakdown flag
    cudaEvent_t done = nullptr;
};

struct Job {
    int iter;
    int slot;
};

constexpr int kStages = 7;  // events: build begin/end (build stream), pca begin/end, knn end, cluster end, scale begin

}  // namespace

extern "C" int dd_fit_iterations(dd_handle *h, const dd_fit_params *p, const int64_t *parents, const float *omega,
                                 double *scores_out, double *log_p_out, int32_t *communities_out,
                                 int32_t *synth_communities_out, double *stage_ms_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_fit_iterations: null handle");
    if (!p || !omega || !scores_out || !log_p_out || !communities_out)
        return dd_fail(h, DD_ERR_ARG, "dd_fit_iterations: null argument");
    if (!h->d_indptr) return dd_fail(h, DD_ERR_ARG, "dd_fit_iterations: call dd_upload_counts first");
    if (p->n_iters < 1 || p->iter_begin < 0 || p->iter_end > p->n_iters || p->iter_begin > p->iter_end)
        return dd_fail(h, DD_ERR_ARG, "dd_fit_iterations: bad iteration range");
    if (p->n_synth < 0 || (p->n_synth > 0 && (!parents || !synth_communities_out)))
        return dd_fail(h, DD_ERR_ARG, "dd_fit_iterations: bad n_synth / parents");
    if (p->knn_k < 2) return dd_fail(h, DD_ERR_ARG, "dd_fit_iterations: knn_k < 2");
    if (p->clustering != DD_CLUSTER_LOUVAIN && p->clustering != DD_CLUSTER_PHENOGRAPH && p->clustering != DD_CLUSTER_LEIDEN)
        return dd_fail(h, DD_ERR_ARG, "dd_fit_iterations: unknown clustering");
    const bool pheno = p->clustering == DD_CLUSTER_PHENOGRAPH;
    // Leiden works on umap's weighted graph, whose weights come from the kNN DISTANCES: the fuzzy simplicial set of every
    // iteration is built on the device right behind its kNN (dd_dev_umap_graph) and the host workers partition it (leiden.cpp)
    const bool leiden = p->clustering == DD_CLUSTER_LEIDEN;
    // DD_UMAP_HOST=1 (A/B timing only): the round-1 split -- lists + distances go to the workers, which build the graph too
    static const bool umap_host = getenv("DD_UMAP_HOST") && atoi(getenv("DD_UMAP_HOST")) != 0;
    // PhenoGraph's first Louvain level on the device, in fixed point (louvain_gpu_w.cu); DD_PHENO_LEVEL0=0: the host twin
    // of that level runs on the workers instead (same communities, A/B timing only)
    static const bool pheno_level0 = !(getenv("DD_PHENO_LEVEL0") && atoi(getenv("DD_PHENO_LEVEL0")) == 0);
    if (pheno && (p->pheno_k < 1 || p->pheno_k > 30))
        return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_fit_iterations: phenograph k must be in [1, 30]");
    DD_CUDA(h, cudaSetDevice(h->device));

    const int64_t N = h->N, M = p->n_synth, A = N + M;
    // PhenoGraph looks at its own neighbourhood size (k nearest + self), not sc.pp.neighbors' 10
    const int k = pheno ? p->pheno_k + 1 : p->knn_k;
    const int64_t max_nnz = A * 2 * (k - 1);
    const int64_t w_off = ((A + 1) + A + max_nnz + 1) / 2 * 2;  // weights start 8-byte aligned
    const int64_t slot_elems = (pheno || leiden) ? std::max<int64_t>(w_off + 2 * max_nnz, 2 * A * k) : (A + 1) + A + max_nnz;
    // Host workers: as many as the caller allows (n_jobs).  The host side of a Louvain iteration (aggregate ~10^2 communities,
    // upper levels, scoring) is a few milliseconds, so a few workers keep up with the GPU; PhenoGraph and Leiden partition
    // much more on the host and are host-bound.  Every worker owns a pinned result slot (10 MB for Louvain, 90 MB for
    // PhenoGraph at 125 k cells, 0.9 GB at 1.25 M), so the count is bounded by 8 GB of pinned memory per loop, and by 64.
    const int64_t slot_budget = (8ll << 30) / (int64_t)sizeof(int32_t);
    const int by_memory = (int)std::max<int64_t>(1, std::min<int64_t>(64, slot_budget / std::max<int64_t>(slot_elems, 1) - 2));
    const int n_threads = std::max(1, std::min(std::min(p->n_host_threads, 64), by_memory));
    const int n_slots = n_threads + 2;
    const int n_run = p->iter_end - p->iter_begin;
    if (stage_ms_out) std::fill(stage_ms_out, stage_ms_out + 8, 0.0);
    if (n_run == 0) return DD_OK;

    std::vector<Slot> slots(n_slots);
    std::vector<cudaEvent_t> evs((size_t)n_run * kStages, nullptr);
    int rc = DD_OK;
    std::string err;
    auto cleanup = [&]() {
        for (Slot &s : slots)
            if (s.done) cudaEventDestroy(s.done);
        for (cudaEvent_t e : evs)
            if (e) cudaEventDestroy(e);
    };
    // pinned result slots live in the handle (grow-only): cudaMallocHost costs milliseconds per call
    if (slot_elems > h->slot_knn_elems) {
        for (int32_t *p : h->slot_knn) cudaFreeHost(p);
        h->slot_knn.clear();
        h->slot_knn_elems = slot_elems;
    }
    while ((int)h->slot_knn.size() < n_slots) {
        int32_t *p = nullptr;
        if (cudaMallocHost(&p, sizeof(int32_t) * h->slot_knn_elems) != cudaSuccess)
            return dd_fail(h, DD_ERR_NOMEM, "dd_fit_iterations: pinned host buffers");
        h->slot_knn.push_back(p);
    }
    while ((int)h->slot_flag.size() < n_slots) {
        double *p = nullptr;
        if (cudaMallocHost(&p, sizeof(double)) != cudaSuccess)
            return dd_fail(h, DD_ERR_NOMEM, "dd_fit_iterations: pinned host buffers");
        h->slot_flag.push_back(p);
    }
    for (int s = 0; s < n_slots; s++) {
        slots[s].graph = h->slot_knn[s];
        slots[s].flag = h->slot_flag[s];
        // blocking sync: a host worker that waits for its iteration sleeps instead of spinning (8 ranks x 2 pipelines share
        // the box's cores with the issuing threads)
        if (cudaEventCreateWithFlags(&slots[s].done, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) {
            cleanup();
            return dd_fail(h, DD_ERR_CUDA, "dd_fit_iterations: event creation");
        }
    }
    for (cudaEvent_t &e : evs)
        if (cudaEventCreate(&e) != cudaSuccess) {
            cleanup();
            return dd_fail(h, DD_ERR_CUDA, "dd_fit_iterations: event creation");
        }

    // ---- host workers: wait for an iteration's kNN graph, cluster, score
    std::mutex mu;
    std::condition_variable cv_job, cv_slot;
    std::deque<Job> jobs;
    std::deque<int> free_slots;
    for (int s = 0; s < n_slots; s++) free_slots.push_back(s);
    bool closing = false;
    std::atomic<int> worker_rc{DD_OK};
    std::string worker_err;
    std::atomic<int64_t> host_us{0};
    auto now = [] { return std::chrono::steady_clock::now(); };
    const auto t_call = now();

    auto worker = [&]() {
        std::vector<int32_t> labels((size_t)A);
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return closing || !jobs.empty(); });
                if (jobs.empty()) return;
                job = jobs.front();
                jobs.pop_front();
            }
            Slot &s = slots[job.slot];
            int wrc = DD_OK;
            std::string werr;
            if (cudaEventSynchronize(s.done) != cudaSuccess) {
                wrc = DD_ERR_CUDA;
                werr = "dd_fit_iterations: device error while waiting for an iteration";
            } else if (*s.flag != 0.0) {
                wrc = DD_ERR_UNSUPPORTED;
                werr = "pca: rank-deficient range (Cholesky breakdown)";
            } else {
                const auto t0 = now();
                int32_t n_comm = 0;
                const int32_t *off = s.graph, *comm0 = s.graph + (A + 1), *adj = s.graph + (A + 1) + A;
                if (leiden && umap_host)
                    wrc = dd_host_leiden_knn(A, k, s.graph, reinterpret_cast<const float *>(s.graph + A * k), p->resolution,
                                             p->seed, labels.data(), &n_comm);
                else if (leiden)
                    wrc = dd_host_leiden_from_graph(A, off, adj, reinterpret_cast<const double *>(s.graph + w_off), p->resolution,
                                                    p->seed, labels.data(), &n_comm);
                else if (pheno)
                    wrc = dd_host_phenograph_from_graph(A, off, adj, reinterpret_cast<const double *>(s.graph + w_off), p->seed,
                                                        p->pheno_min_cluster_size, labels.data(), &n_comm,
                                                        pheno_level0 ? comm0 : nullptr);
                else
                    wrc = dd_host_louvain_from_level0(A, off, adj, comm0, p->resolution, p->seed, labels.data(), &n_comm);
                if (wrc != DD_OK) werr = "dd_fit_iterations: clustering rejected the device graph";
                if (wrc == DD_OK) {
                    wrc = dd_score(N, M, labels.data(), scores_out + (size_t)job.iter * N, log_p_out + (size_t)job.iter * N);
                    std::copy(labels.begin(), labels.begin() + N, communities_out + (size_t)job.iter * N);
                    if (M > 0) std::copy(labels.begin() + N, labels.end(), synth_communities_out + (size_t)job.iter * M);
                }
                host_us += std::chrono::duration_cast<std::chrono::microseconds>(now() - t0).count();
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                if (wrc != DD_OK && worker_rc.load() == DD_OK) {
                    worker_rc.store(wrc);
                    worker_err = werr;
                }
                free_slots.push_back(job.slot);
            }
            cv_slot.notify_one();
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(worker);

    // ---- GPU producer
    // Three streams: the dense build of iteration i + 1 (HBM-bound, no shared memory / TMEM) runs on the build stream as
    // soon as the last PCA product of iteration i has read the matrix, i.e. underneath iteration i's small PCA kernels and
    // its TMEM-bound kNN; the clustering of iteration i runs on the clustering stream underneath iteration i + 1.
    std::vector<float> aug((size_t)A), tmp;
    bool omega_sent = false;
    int issued = 0;
    auto issue_dense = [&](int it, cudaEvent_t *ev, bool wait_gemms) -> int {
        const int64_t *par = M > 0 ? parents + (size_t)it * M * 2 : nullptr;
        // np.median(aug_lib_size): synthetic library sizes are the parents' sums (exact for counts)
        std::copy(h->h_lib.begin(), h->h_lib.end(), aug.begin());
        for (int64_t r = 0; r < M; r++) aug[N + r] = h->h_lib[par[2 * r]] + h->h_lib[par[2 * r + 1]];
        tmp = aug;
        const float median = dd_host_median(tmp);
        cudaStream_t main_stream = h->stream;
        h->stream = h->stream3;
        if (wait_gemms) cudaStreamWaitEvent(h->stream3, h->ev_gemms_done, 0);  // the previous PCA still reads the matrix
        int r = dd_set_parents(h, M, par);
        if (r == DD_OK) {
            cudaEventRecord(ev[0], h->stream3);
            r = dd_dev_build_dense(h, median, p->pseudocount);
            cudaEventRecord(ev[1], h->stream3);
            cudaEventRecord(h->ev_dense_done, h->stream3);
        }
        h->stream = main_stream;
        return r;
    };
    rc = issue_dense(p->iter_begin, &evs[0], false);
    // Four streams (unsharded): main = scale + PCA, build = dense matrix of the next iteration, kNN = the neighbour search of
    // iteration i CONCURRENT with the PCA of iteration i + 1 (the kNN is bound by the TMEM read path and leaves HBM idle,
    // the PCA products are HBM-bound and its ~2 ms of small float64 kernels leave the SMs idle), clustering = graph + first
    // Louvain level of iteration i.  The embedding and the kNN lists are double-buffered.  Cell-block sharding keeps the kNN
    // on the main stream: both stages enqueue NCCL collectives on the same communicator, whose order must be the same on
    // every rank.
    static const bool knn_inline = getenv("DD_KNN_INLINE") != nullptr;  // A/B: the round-1 order (kNN on the main stream)
    const bool knn_own_stream = !dd_sharded(h) && !knn_inline;
    static const bool knn_narrow_ok = getenv("DD_KNN_NARROW") != nullptr;  // A/B: the 256-column kernel (measured slower: 6.2 vs 4.95 ms, profiles/r2g)
    // clustering lanes (dd_lv_lane): the first Louvain levels of up to n_lanes consecutive iterations in flight at once
    int n_lanes = 2;  // measured 1..4 (profiles/r2f_lanes.log): wall per iteration within 0.5 ms of each other
    if (const char *e = getenv("DD_LV_LANES")) n_lanes = std::max(1, std::min(8, atoi(e)));
    if (leiden) n_lanes = 1;
    while ((int)h->lv_lanes.size() < n_lanes - 1) {
        dd_lv_lane l;
        int prio_least = 0, prio_greatest = 0;
        cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
        if (cudaStreamCreateWithPriority(&l.stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
            rc = dd_fail(h, DD_ERR_CUDA, "dd_fit_iterations: clustering lane stream");
            break;
        }
        h->lv_lanes.push_back(l);
    }
    cudaStream_t const main_stream = h->stream;
    cudaStream_t const knn_stream = knn_own_stream ? h->stream4 : h->stream;
    // A host worker's failure (rank-deficient PCA, clustering rejected) stops the loop early -- except on a cell-sharded
    // handle: there only the OWNING rank sees it, and a rank that left the loop would leave the others waiting in NCCL
    // collectives it never joins.  Sharded ranks therefore issue every iteration and report the failure at the end (the
    // Python shim all-reduces the status so that every rank raises).
    for (int it = p->iter_begin; it < p->iter_end && rc == DD_OK && (dd_sharded(h) || worker_rc.load() == DD_OK); it++, issued++) {
        // Cell-block sharding: every rank holds the all-gathered kNN lists of every iteration, so the clustering + scoring
        // of the iterations is dealt round-robin to the ranks (rank it % world finishes iteration it; the caller merges the
        // per-iteration result rows, which stay zero on the other ranks).
        const bool cluster_here = !dd_sharded(h) || (it % h->world) == h->rank;
        int slot = -1;
        if (cluster_here) {
            std::unique_lock<std::mutex> lk(mu);
            cv_slot.wait(lk, [&] { return !free_slots.empty(); });
            slot = free_slots.front();
            free_slots.pop_front();
        }
        const int eb = issued & 1;  // embedding / list buffer of this iteration
        cudaEvent_t *ev = &evs[(size_t)issued * kStages];
        cudaStreamWaitEvent(h->stream, h->ev_dense_done, 0);
        cudaEventRecord(ev[6], h->stream);
        if (p->standard_scaling && (rc = dd_dev_standard_scale(h, p->scale_max_value)) != DD_OK) break;
        cudaEventRecord(ev[2], h->stream);
        h->gemms_done_recorded = false;
        // this PCA overwrites the embedding buffer the kNN of two iterations ago read
        if (knn_own_stream && issued > 1) cudaStreamWaitEvent(h->stream, h->ev_emb_free[eb], 0);
        if (h->d_emb_base) h->d_emb = h->d_emb_base + eb * h->emb_stride;
        if ((rc = dd_dev_pca(h, p->n_comp, p->n_random, p->n_power_iter, omega_sent ? nullptr : omega)) != DD_OK) break;
        if (h->d_emb != h->d_emb_base + eb * h->emb_stride) {  // the first call allocated the buffers (and wrote buffer 0)
            if (eb != 0) { rc = dd_fail(h, DD_ERR_CUDA, "dd_fit_iterations: embedding buffers re-allocated mid-fit"); break; }
        }
        if (!h->gemms_done_recorded) cudaEventRecord(h->ev_gemms_done, h->stream);
        omega_sent = true;
        if (cluster_here) dd_pca_flag_copy(h, slots[slot].flag);  // before the next PCA resets the flag
        cudaEventRecord(ev[3], h->stream);
        cudaEventRecord(h->ev_pca_done[eb], h->stream);
        // the dense build of the next iteration goes out now (build stream; it waits for this PCA's last pass over the matrix)
        if (it + 1 < p->iter_end && (rc = issue_dense(it + 1, &evs[(size_t)(issued + 1) * kStages], true)) != DD_OK) break;

        // ---- kNN (own stream).  The lists are double-buffered too: this iteration writes buffer eb, which the clustering
        // stream finished reading two iterations ago -- so the kNN never waits for the previous iteration's Louvain level.
        h->stream = knn_stream;
        if (knn_own_stream) cudaStreamWaitEvent(knn_stream, h->ev_pca_done[eb], 0);
        cudaEvent_t lv_done = eb ? h->ev_lv_done2 : h->ev_lv_done;
        if (issued > 1) cudaStreamWaitEvent(knn_stream, lv_done, 0);
        if (h->d_knn_idx_base) h->d_knn_idx = h->d_knn_idx_base + eb * h->knn_idx_stride;
        h->emb_valid = true;  // issuing the next dense build (new parents) marked this iteration's embedding stale
        h->knn_narrow = knn_own_stream && knn_narrow_ok;
        rc = dd_dev_knn(h, k);
        h->knn_narrow = false;
        if (rc == DD_OK && h->d_knn_idx != h->d_knn_idx_base + eb * h->knn_idx_stride)  // first call allocated the buffers
            h->d_knn_idx = h->d_knn_idx_base + eb * h->knn_idx_stride;
        cudaEventRecord(ev[4], knn_stream);
        if (rc == DD_OK && leiden && cluster_here) {
            // no clustering lane: umap's graph is built on the kNN stream right behind the search (the distance buffer and
            // the graph buffers are not double-buffered: the next kNN / graph build is ordered behind these kernels and
            // copies on the same stream)
            Slot &s = slots[slot];
            if (umap_host) {
                cudaMemcpyAsync(s.graph, h->d_knn_idx, sizeof(int32_t) * A * k, cudaMemcpyDeviceToHost, knn_stream);
                cudaMemcpyAsync(s.graph + A * k, h->d_knn_dist, sizeof(float) * A * k, cudaMemcpyDeviceToHost, knn_stream);
            } else if ((rc = dd_dev_umap_graph(h, k)) == DD_OK) {
                cudaMemcpyAsync(s.graph, h->d_lv_off, sizeof(int32_t) * (A + 1), cudaMemcpyDeviceToHost, knn_stream);
                cudaMemcpyAsync(s.graph + (A + 1) + A, h->d_lv_adj, sizeof(int32_t) * max_nnz, cudaMemcpyDeviceToHost, knn_stream);
                cudaMemcpyAsync(s.graph + w_off, h->d_lv_w, sizeof(double) * max_nnz, cudaMemcpyDeviceToHost, knn_stream);
            }
        }
        cudaEventRecord(h->ev_emb_free[eb], knn_stream);
        cudaEventRecord(h->ev_knn_done, knn_stream);
        h->stream = main_stream;
        if (rc != DD_OK) break;
        if (!cluster_here) {
            cudaEventRecord(ev[5], knn_stream);
            continue;
        }
        Slot &s = slots[slot];
        if (leiden) {
            cudaEventRecord(ev[5], knn_stream);
            cudaEventRecord(lv_done, knn_stream);
            cudaStreamWaitEvent(knn_stream, h->ev_pca_done[eb], 0);  // s.done also covers the flag copy on the main stream
            cudaEventRecord(s.done, knn_stream);
            {
                std::lock_guard<std::mutex> lk(mu);
                jobs.push_back(Job{it, slot});
            }
            cv_job.notify_one();
            continue;
        }
        // clustering, first level: symmetric kNN pattern + synchronous coloured Louvain rounds on the device.
        // These are hundreds of small latency-bound kernels: they run on a clustering lane (own stream, own state) and
        // overlap the dense build / PCA / kNN of the NEXT iterations -- and the levels of the neighbouring iterations on
        // the other lanes.
        const int lane = issued % n_lanes;
        cudaStream_t cl_stream = lane == 0 ? h->stream2 : h->lv_lanes[lane - 1].stream;
        cudaStreamWaitEvent(cl_stream, h->ev_knn_done, 0);
        cudaStreamWaitEvent(cl_stream, h->ev_pca_done[eb], 0);  // s.done must also cover the flag copy (main stream)
        if (lane > 0) dd_lv_swap(h, h->lv_lanes[lane - 1]);
        h->stream = cl_stream;
        if (pheno) {
            rc = dd_dev_jaccard_graph(h, k, p->pheno_prune);
            cudaEventRecord(lv_done, cl_stream);  // the lists have been read
            if (rc == DD_OK && pheno_level0) rc = dd_dev_louvain_level0_weighted(h, 1.0, p->seed);
        } else {
            h->ev_after_graph_build = lv_done;  // recorded as soon as the pattern graph exists: the lists are free again
            rc = dd_dev_louvain_level0(h, k, p->resolution, p->seed);
            h->ev_after_graph_build = nullptr;
        }
        h->stream = main_stream;
        if (rc == DD_OK) {
            cudaMemcpyAsync(s.graph, h->d_lv_off, sizeof(int32_t) * (A + 1), cudaMemcpyDeviceToHost, cl_stream);
            cudaMemcpyAsync(s.graph + (A + 1), h->d_lv_comm, sizeof(int32_t) * A, cudaMemcpyDeviceToHost, cl_stream);
            cudaMemcpyAsync(s.graph + (A + 1) + A, h->d_lv_adj, sizeof(int32_t) * max_nnz, cudaMemcpyDeviceToHost, cl_stream);
            if (pheno)
                cudaMemcpyAsync(s.graph + w_off, h->d_lv_w, sizeof(double) * max_nnz, cudaMemcpyDeviceToHost, cl_stream);
        }
        if (lane > 0) dd_lv_swap(h, h->lv_lanes[lane - 1]);
        if (rc != DD_OK) break;
        cudaEventRecord(ev[5], cl_stream);
        cudaEventRecord(s.done, cl_stream);
        {
            std::lock_guard<std::mutex> lk(mu);
            jobs.push_back(Job{it, slot});
        }
        cv_job.notify_one();
    }
    h->stream = main_stream;
    {
        std::lock_guard<std::mutex> lk(mu);
        closing = true;
    }
    cv_job.notify_all();
    for (std::thread &t : pool) t.join();
    cudaError_t ce = cudaStreamSynchronize(h->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream2);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream3);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(h->stream4);
    for (dd_lv_lane &l : h->lv_lanes)
        if (ce == cudaSuccess && l.stream) ce = cudaStreamSynchronize(l.stream);
    if (rc == DD_OK && ce != cudaSuccess) rc = dd_fail(h, DD_ERR_CUDA, std::string("dd_fit_iterations: ") + cudaGetErrorString(ce));
    if (rc == DD_OK && worker_rc.load() != DD_OK) rc = dd_fail(h, worker_rc.load(), worker_err);
    if (rc == DD_OK && h->h_lv_rounds && !pheno && !leiden) h->stage_ms["lv_rounds"] = (double)*h->h_lv_rounds;
    if (rc == DD_OK && stage_ms_out) {
        for (int i = 0; i < issued; i++) {
            cudaEvent_t *ev = &evs[(size_t)i * kStages];
            float ms;
            // {doublets, normalise, scale, pca, knn, d2h}: the fused build is booked under "normalise"
            if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) stage_ms_out[1] += ms;
            if (cudaEventElapsedTime(&ms, ev[6], ev[2]) == cudaSuccess) stage_ms_out[2] += ms;
            if (cudaEventElapsedTime(&ms, ev[2], ev[3]) == cudaSuccess) stage_ms_out[3] += ms;
            if (cudaEventElapsedTime(&ms, ev[3], ev[4]) == cudaSuccess) stage_ms_out[4] += ms;
            if (cudaEventElapsedTime(&ms, ev[4], ev[5]) == cudaSuccess) stage_ms_out[5] += ms;
        }
        float total = 0.f;
        if (cudaEventElapsedTime(&total, evs[0], evs[(size_t)(issued - 1) * kStages + 5]) == cudaSuccess)
            stage_ms_out[6] = total;  // first launch -> last copy, device time of the whole call
        stage_ms_out[0] = host_us.load() / 1e3;  // clustering + scoring, summed over the host workers
        stage_ms_out[7] = std::chrono::duration_cast<std::chrono::microseconds>(now() - t_call).count() / 1e3;
    }
    cleanup();
    return rc;
}
