// handle.cu -- lifetime, error reporting, launch accounting and stage timers of libdd_b200.so.
#include <utility>

#include "dd_internal.h"
#include "pca_tc.h"

#include <algorithm>
#include <cstring>

static thread_local std::string g_error;

void dd_set_global_error(const std::string &msg) { g_error = msg; }

int dd_fail(dd_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    g_error = msg;
    return code;
}

static cudaEvent_t pool_get(dd_handle *h) {
    if (!h->event_pool.empty()) {
        cudaEvent_t e = h->event_pool.back();
        h->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Per-kernel timing records an event pair around every launch WITHOUT synchronising, so the pipeline
// runs as it does untimed; the pairs are resolved when the numbers are read.
void dd_launch_begin(dd_handle *h) {
    if (!h->timing) return;
    dd_timed_launch t{nullptr, pool_get(h), pool_get(h)};
    cudaEventRecord(t.start, h->stream);
    h->pending.push_back(t);
}

int dd_launch_end(dd_handle *h, const char *name) {
    cudaError_t e = cudaGetLastError();
    if (h->timing && !h->pending.empty() && h->pending.back().name == nullptr) {
        h->pending.back().name = name;
        cudaEventRecord(h->pending.back().stop, h->stream);
    }
    if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("launch of ") + name + ": " + cudaGetErrorString(e));
    h->launches++;
    return DD_OK;
}

static void resolve_pending(dd_handle *h) {
    if (h->pending.empty()) return;
    cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    if (h->stream3) cudaStreamSynchronize(h->stream3);
    for (dd_timed_launch &t : h->pending) {
        float ms = 0.f;
        if (t.name && cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
            dd_kernel_stat &st = h->kstats[t.name];
            st.total_ms += ms;
            st.launches++;
        }
        h->event_pool.push_back(t.start);
        h->event_pool.push_back(t.stop);
    }
    h->pending.clear();
}

int dd_stage_begin(dd_handle *h) {
    DD_CUDA(h, cudaEventRecord(h->stage_ev0, h->stream));
    return DD_OK;
}

int dd_stage_end(dd_handle *h, const char *stage) {
    DD_CUDA(h, cudaEventRecord(h->stage_ev1, h->stream));
    DD_CUDA(h, cudaEventSynchronize(h->stage_ev1));
    float ms = 0.f;
    DD_CUDA(h, cudaEventElapsedTime(&ms, h->stage_ev0, h->stage_ev1));
    h->stage_ms[stage] = ms;
    return DD_OK;
}

extern "C" int dd_abi_version(void) { return DD_ABI_VERSION; }

extern "C" int dd_create(int device, dd_handle **out) {
    if (!out) return dd_fail(nullptr, DD_ERR_ARG, "dd_create: null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return dd_fail(nullptr, DD_ERR_CUDA,
                       std::string("dd_create: no CUDA device (there is no CPU fallback): ") +
                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    if (device < 0 || device >= count) return dd_fail(nullptr, DD_ERR_ARG, "dd_create: device index out of range");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return dd_fail(nullptr, DD_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return dd_fail(nullptr, DD_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return dd_fail(nullptr, DD_ERR_UNSUPPORTED,
                       std::string("dd_create: device '") + prop.name + "' is not sm_100 (this library is B200-only)");
    dd_handle *h = new dd_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    // the clustering stream runs hundreds of tiny latency-bound kernels next to the main stream's HBM-bound ones: at the
    // highest priority its CTAs are placed first whenever an SM frees resources, so it keeps pace with the main stream
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    // A/B switches (levels above the lowest priority, clamped to the device's range): DD_PRIO_MAIN (PCA), DD_PRIO_BUILD (dense
    // build), DD_PRIO_KNN
    auto prio_of = [&](const char *name) {
        const char *e = getenv(name);
        const int up = e ? atoi(e) : 0;
        return std::max(prio_greatest, prio_least - std::max(0, up));
    };
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_of("DD_PRIO_MAIN")) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, prio_of("DD_PRIO_BUILD")) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->stream4, cudaStreamNonBlocking, prio_of("DD_PRIO_KNN")) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_pca_done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_pca_done[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_emb_free[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_emb_free[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_dense_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_gemms_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_knn_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_lv_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_lv_done2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess ||
        cudaEventCreate(&h->stage_ev0) != cudaSuccess || cudaEventCreate(&h->stage_ev1) != cudaSuccess) {
        delete h;
        return dd_fail(nullptr, DD_ERR_CUDA, "dd_create: stream/event creation failed");
    }
    *out = h;
    return DD_OK;
}

// exchange the Louvain-level state of the handle (= lane 0) with `lane`
void dd_lv_swap(dd_handle *h, dd_lv_lane &l) {
    std::swap(h->d_lv_off, l.d_lv_off); std::swap(h->d_lv_adj, l.d_lv_adj); std::swap(h->d_lv_comm, l.d_lv_comm);
    std::swap(h->d_lv_i32, l.d_lv_i32); std::swap(h->d_lv_tot, l.d_lv_tot); std::swap(h->d_lv_w, l.d_lv_w);
    std::swap(h->cap_lv_n, l.cap_lv_n); std::swap(h->cap_lv_nnz, l.cap_lv_nnz); std::swap(h->cap_lv_w, l.cap_lv_w);
    std::swap(h->d_lvw_wq, l.d_lvw_wq); std::swap(h->d_lvw_i64, l.d_lvw_i64); std::swap(h->d_lvw_i32, l.d_lvw_i32);
    std::swap(h->cap_lvw_nnz, l.cap_lvw_nnz); std::swap(h->cap_lvw_n, l.cap_lvw_n);
    std::swap(h->lvw_bucket_n, l.lvw_bucket_n); std::swap(h->lvw_bucket_seed, l.lvw_bucket_seed);
    std::swap(h->lv_bucket_n, l.lv_bucket_n); std::swap(h->lv_bucket_seed, l.lv_bucket_seed);
    std::swap(h->h_lv_rounds, l.h_lv_rounds); std::swap(h->lv_graph_exec, l.lv_graph_exec);
    std::swap(h->lv_graph_n, l.lv_graph_n); std::swap(h->lv_graph_launches, l.lv_graph_launches);
    std::swap(h->lv_graph_is_loop, l.lv_graph_is_loop); std::swap(h->lv_graph_gamma, l.lv_graph_gamma);
    std::swap(h->lv_graph_seed, l.lv_graph_seed);
    std::swap(h->lvw_graph_exec, l.lvw_graph_exec); std::swap(h->lvw_graph_n, l.lvw_graph_n);
    std::swap(h->lvw_graph_launches, l.lvw_graph_launches); std::swap(h->lvw_graph_gamma, l.lvw_graph_gamma);
    std::swap(h->lvw_graph_seed, l.lvw_graph_seed);
    for (int i = 0; i < 4; i++) std::swap(h->lvw_graph_key[i], l.lvw_graph_key[i]);
}

void dd_lv_lane_free(dd_lv_lane &l) {
    if (l.stream) cudaStreamSynchronize(l.stream);
    for (void *p : {(void *)l.d_lv_off, (void *)l.d_lv_adj, (void *)l.d_lv_comm, (void *)l.d_lv_i32, (void *)l.d_lv_tot,
                    (void *)l.d_lv_w, (void *)l.d_lvw_wq, (void *)l.d_lvw_i64, (void *)l.d_lvw_i32})
        if (p) cudaFree(p);
    if (l.lv_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)l.lv_graph_exec);
    if (l.lvw_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)l.lvw_graph_exec);
    if (l.h_lv_rounds) cudaFreeHost(l.h_lv_rounds);
    if (l.stream) cudaStreamDestroy(l.stream);
    l = dd_lv_lane();
}

extern "C" void dd_destroy(dd_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    dd_drop_borrowed_counts(h);  // borrowed buffers are freed by their owner
    cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    if (h->stream3) cudaStreamSynchronize(h->stream3);
    if (h->stream4) cudaStreamSynchronize(h->stream4);
    void *bufs[] = {h->d_indptr, h->d_indices, h->d_data,   h->d_lib,   h->d_l1,      h->d_parents, h->d_sindptr,
                    h->d_scount, h->d_sindices, h->d_sdata, h->d_slib,  h->d_dense,   h->d_colsum,  h->d_colsumsq,
                    h->d_Qt,     h->d_Y,        h->d_Zacc,  h->d_small, h->d_emb_base, h->d_knn_idx_base, h->d_knn_dist, h->d_knn_ops, h->d_knn_list_off, h->d_knn_list_tiles, h->d_knn_cl, h->d_knn_cert, h->d_lvw_wq, h->d_lvw_i64, h->d_lvw_i32, h->d_qb, h->d_yb, h->d_omega_b, h->d_mu, h->d_lv_off, h->d_lv_adj, h->d_lv_comm, h->d_lv_tot, h->d_lv_i32, h->d_lv_w, h->d_umap_w};
    for (void *p : bufs)
        if (p) cudaFree(p);
    for (dd_lv_lane &l : h->lv_lanes) dd_lv_lane_free(l);
    dd_tc_free(h);
    dd_comm_destroy(h);
    if (h->lv_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->lv_graph_exec);
    if (h->lvw_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->lvw_graph_exec);
    if (h->h_knn_cl_pairs) cudaFreeHost(h->h_knn_cl_pairs);
    if (h->h_knn_uncert) cudaFreeHost(h->h_knn_uncert);
    for (int32_t *p : h->slot_knn) cudaFreeHost(p);
    for (double *p : h->slot_flag) cudaFreeHost(p);
    if (h->h_lv_rounds) cudaFreeHost(h->h_lv_rounds);
    resolve_pending(h);
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stage_ev0) cudaEventDestroy(h->stage_ev0);
    if (h->stage_ev1) cudaEventDestroy(h->stage_ev1);
    if (h->ev_knn_done) cudaEventDestroy(h->ev_knn_done);
    if (h->ev_lv_done) cudaEventDestroy(h->ev_lv_done);
    if (h->ev_lv_done2) cudaEventDestroy(h->ev_lv_done2);
    if (h->ev_dense_done) cudaEventDestroy(h->ev_dense_done);
    if (h->ev_gemms_done) cudaEventDestroy(h->ev_gemms_done);
    for (int b = 0; b < 2; b++) {
        if (h->ev_pca_done[b]) cudaEventDestroy(h->ev_pca_done[b]);
        if (h->ev_emb_free[b]) cudaEventDestroy(h->ev_emb_free[b]);
    }
    if (h->stream4) cudaStreamDestroy(h->stream4);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *dd_last_error(const dd_handle *h) { return h ? h->err.c_str() : g_error.c_str(); }

extern "C" int64_t dd_kernel_launches(const dd_handle *h) { return h ? h->launches : -1; }

extern "C" double dd_last_stage_ms(const dd_handle *h, const char *stage) {
    if (!h || !stage) return -1.0;
    auto it = h->stage_ms.find(stage);
    return it == h->stage_ms.end() ? -1.0 : it->second;
}

extern "C" int dd_set_kernel_timing(dd_handle *h, int32_t on) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_set_kernel_timing: null handle");
    resolve_pending(h);
    h->timing = on != 0;
    if (on) h->kstats.clear();
    return DD_OK;
}

// "name total_ms launches\n" for every kernel seen since timing was switched on; returns the number of
// bytes needed (call with buf == NULL to size the buffer).
extern "C" int64_t dd_kernel_timing_report(dd_handle *h, char *buf, int64_t buflen) {
    if (!h) return -1;
    resolve_pending(h);
    std::string out;
    for (auto &kv : h->kstats)
        out += kv.first + " " + std::to_string(kv.second.total_ms) + " " + std::to_string(kv.second.launches) + "\n";
    if (buf && buflen > 0) {
        const int64_t n = std::min<int64_t>(buflen - 1, (int64_t)out.size());
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int64_t)out.size() + 1;
}

extern "C" int dd_get_kernel_timing(dd_handle *h, const char *kernel, double *total_ms_out, int64_t *launches_out) {
    if (!h || !kernel) return dd_fail(h, DD_ERR_ARG, "dd_get_kernel_timing: bad arguments");
    resolve_pending(h);
    auto it = h->kstats.find(kernel);
    if (total_ms_out) *total_ms_out = it == h->kstats.end() ? 0.0 : it->second.total_ms;
    if (launches_out) *launches_out = it == h->kstats.end() ? 0 : it->second.launches;
    return DD_OK;
}
