// leiden.cpp -- clustering_algorithm="leiden" of _one_fit (doubletdetection.py:337-343):
// sc.pp.neighbors(method="umap", n_neighbors=10) turns the exact kNN lists into umap's fuzzy-simplicial-set
// connectivities, and sc.tl.leiden(resolution=4, random_state, directed=False) partitions that WEIGHTED graph
// (use_weights=True, RB-configuration quality, n_iterations=-1).  umap-learn and leidenalg are absent from the
// image; the arithmetic of the weights is pinned by oracle/upstream.py (smooth_knn_dist, membership_strengths,
// fuzzy_connectivities) and the move order of the partitioning by oracle/leiden_ref.py -- this file reproduces
// both bit for bit / label for label.  In the fit loop the graph itself is built on the GPU (louvain_gpu.cu:
// dd_dev_umap_graph, the device twin of umap_weights / umap_graph below) and arrives in a pinned slot; the host workers
// partition it (dd_host_leiden_from_graph) concurrently while the GPU runs ahead.  The host graph builder serves the
// stage-wise entry points and is what the device graph is tested against.
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <vector>

#include "dd_internal.h"
#include "louvain_host.h"

namespace {

using ddlv::Graph;
using ddlv::SplitMix64;

constexpr double kTheta = 0.01;
constexpr int kMaxIterations = 32;

void permutation(int32_t n, SplitMix64 &rng, std::vector<int32_t> &p) {
    p.resize(n);
    std::iota(p.begin(), p.end(), 0);
    for (int64_t i = (int64_t)n - 1; i >= 1; i--) {
        const int64_t j = (int64_t)(rng.next() % (uint64_t)(i + 1));
        std::swap(p[i], p[j]);
    }
}

void degrees(const Graph &g, std::vector<double> &k) {
    k.resize(g.n);
    for (int32_t i = 0; i < g.n; i++) {
        double s = g.selfw[i];
        for (int64_t e = g.indptr[i]; e < g.indptr[i + 1]; e++) s += g.w(e);
        k[i] = s;
    }
}

// community ids by first appearance over the index; returns the number of communities
int32_t first_appearance(std::vector<int32_t> &ids) {
    const int32_t n = (int32_t)ids.size();
    int32_t hi = 0;
    for (int32_t c : ids) hi = std::max(hi, c + 1);
    std::vector<int32_t> new_id(hi, -1);
    int32_t nc = 0;
    for (int32_t i = 0; i < n; i++) {
        if (new_id[ids[i]] < 0) new_id[ids[i]] = nc++;
        ids[i] = new_id[ids[i]];
    }
    return nc;
}

struct Work {
    std::vector<double> k, tot, neigh_w, ptot, rtot, ext, gains, probs;
    std::vector<int32_t> size, free_ids, queue, cands, order, refined, rsize, elig;
    std::vector<uint8_t> seen, inq;
    int64_t n_moves = 0;  // of the last move_nodes call (DD_LEIDEN_TRACE)
};

// fast local moving from the partition `part` (ids in [0, n)); oracle/leiden_ref.py:_move_nodes
bool move_nodes(const Graph &g, double gamma, double two_m, SplitMix64 &rng, std::vector<int32_t> &part, Work &w) {
    const int32_t n = g.n;
    w.tot.assign(n, 0.0);
    w.size.assign(n, 0);
    for (int32_t i = 0; i < n; i++) {
        w.tot[part[i]] += w.k[i];
        w.size[part[i]]++;
    }
    w.free_ids.clear();
    for (int32_t c = n - 1; c >= 0; c--)
        if (w.size[c] == 0) w.free_ids.push_back(c);  // back() is the smallest free id
    permutation(n, rng, w.queue);
    w.inq.assign(n, 1);
    w.neigh_w.assign(n, 0.0);
    w.seen.assign(n, 0);
    w.cands.resize(std::max<int32_t>(n, 1));
    int32_t head = 0, count = n;
    bool moved_any = false;
    int64_t n_moves = 0;
    while (count > 0) {
        const int32_t i = w.queue[head];
        if (count > 8) {  // entries further down the ring are valid node ids even where they are stale
            int32_t a = head + 8, b = head + 4;
            if (a >= n) a -= n;
            if (b >= n) b -= n;
            g.prefetch_offset(w.queue[a]);
            g.prefetch_row(w.queue[b]);
        }
        head = (head + 1 == n) ? 0 : head + 1;
        count--;
        w.inq[i] = 0;
        const int32_t ci = part[i];
        const double ki = w.k[i];
        const int64_t e0 = g.indptr[i], e1 = g.indptr[i + 1];
        int32_t nc = 0;
        w.cands[nc++] = ci;
        w.seen[ci] = 1;
        w.neigh_w[ci] = 0.0;
        for (int64_t e = e0; e < e1; e++) {
            const int32_t c = part[g.indices[e]];
            if (!w.seen[c]) {
                w.seen[c] = 1;
                w.neigh_w[c] = 0.0;
                w.cands[nc++] = c;
            }
            w.neigh_w[c] += g.w(e);
        }
        w.tot[ci] -= ki;
        const double gk = gamma * ki;
        int32_t best = ci;
        double best_gain = w.neigh_w[ci] - (gk * w.tot[ci]) / two_m;
        for (int32_t t = 1; t < nc; t++) {
            const int32_t c = w.cands[t];
            const double gn = w.neigh_w[c] - (gk * w.tot[c]) / two_m;
            if (gn > best_gain) {
                best = c;
                best_gain = gn;
            }
        }
        for (int32_t t = 0; t < nc; t++) w.seen[w.cands[t]] = 0;
        if (w.size[ci] > 1 && 0.0 > best_gain) {  // an empty community: gain 0
            best = w.free_ids.back();
            w.free_ids.pop_back();
        }
        w.tot[best] += ki;
        if (best != ci) {
            part[i] = best;
            w.size[ci]--;
            w.size[best]++;
            if (w.size[ci] == 0) w.free_ids.push_back(ci);
            moved_any = true;
            n_moves++;
            for (int64_t e = e0; e < e1; e++) {
                const int32_t j = g.indices[e];
                if (part[j] != best && !w.inq[j]) {
                    w.inq[j] = 1;
                    int32_t tail = head + count;
                    if (tail >= n) tail -= n;
                    w.queue[tail] = j;
                    count++;
                }
            }
        }
    }
    w.n_moves = n_moves;
    return moved_any;
}

// refined partition (ids = node ids) inside the communities of `part`; oracle/leiden_ref.py:_refine
void refine(const Graph &g, double gamma, double two_m, SplitMix64 &rng, const std::vector<int32_t> &part, Work &w) {
    const int32_t n = g.n;
    w.ptot.assign(n, 0.0);
    for (int32_t i = 0; i < n; i++) w.ptot[part[i]] += w.k[i];
    w.refined.resize(n);
    std::iota(w.refined.begin(), w.refined.end(), 0);
    w.rtot = w.k;
    w.rsize.assign(n, 1);
    w.ext.assign(n, 0.0);
    for (int32_t i = 0; i < n; i++) {
        double s = 0.0;
        for (int64_t e = g.indptr[i]; e < g.indptr[i + 1]; e++)
            if (part[g.indices[e]] == part[i]) s += g.w(e);
        w.ext[i] = s;
    }
    w.neigh_w.assign(n, 0.0);
    w.seen.assign(n, 0);
    w.cands.resize(std::max<int32_t>(n, 1));
    permutation(n, rng, w.order);
    for (int32_t t = 0; t < n; t++) {
        const int32_t v = w.order[t];
        if (t + 8 < n) {
            g.prefetch_offset(w.order[t + 8]);
            g.prefetch_row(w.order[t + 4]);
            const int32_t v2 = w.order[t + 2];  // its row has arrived: the per-neighbour state it will look at
            for (int64_t e = g.indptr[v2]; e < g.indptr[v2 + 1]; e++) {
                __builtin_prefetch(part.data() + g.indices[e]);
                __builtin_prefetch(w.refined.data() + g.indices[e]);
            }
        }
        if (w.rsize[w.refined[v]] != 1) continue;
        const int32_t C = part[v];
        const double kv = w.k[v];
        if (w.ext[v] < (gamma * kv * (w.ptot[C] - kv)) / two_m) continue;  // v is not well connected inside C
        int32_t nc = 0;
        w.cands[nc++] = v;
        w.seen[v] = 1;
        w.neigh_w[v] = 0.0;
        for (int64_t e = g.indptr[v]; e < g.indptr[v + 1]; e++) {
            const int32_t u = g.indices[e];
            if (part[u] != C) continue;
            const int32_t R = w.refined[u];
            if (!w.seen[R]) {
                w.seen[R] = 1;
                w.neigh_w[R] = 0.0;
                w.cands[nc++] = R;
            }
            w.neigh_w[R] += g.w(e);
        }
        w.elig.clear();
        w.gains.clear();
        w.elig.push_back(v);
        w.gains.push_back(0.0);
        for (int32_t t = 1; t < nc; t++) {
            const int32_t R = w.cands[t];
            if (w.ext[R] < (gamma * w.rtot[R] * (w.ptot[C] - w.rtot[R])) / two_m) continue;
            const double gn = w.neigh_w[R] - ((gamma * kv) * w.rtot[R]) / two_m;
            if (gn < 0.0) continue;
            w.elig.push_back(R);
            w.gains.push_back(gn);
        }
        for (int32_t t = 0; t < nc; t++) w.seen[w.cands[t]] = 0;
        if (w.elig.size() == 1) continue;
        double gmax = w.gains[0];
        for (double x : w.gains) gmax = std::max(gmax, x);
        w.probs.resize(w.gains.size());
        double total = 0.0;
        for (size_t t = 0; t < w.gains.size(); t++) {
            w.probs[t] = std::exp((w.gains[t] - gmax) / kTheta);
            total += w.probs[t];
        }
        const double r = (double)(rng.next() >> 11) * 0x1.0p-53 * total;
        int32_t chosen = w.elig.back();
        double acc = 0.0;
        for (size_t t = 0; t < w.elig.size(); t++) {
            acc += w.probs[t];
            if (r < acc) {
                chosen = w.elig[t];
                break;
            }
        }
        if (chosen != v) {
            w.refined[v] = chosen;
            w.rtot[chosen] += kv;
            w.rtot[v] -= kv;
            w.rsize[chosen]++;
            w.rsize[v] = 0;
            w.ext[chosen] = (w.ext[chosen] + w.ext[v]) - 2.0 * w.neigh_w[chosen];
        }
    }
}

// one Leiden iteration from `membership` (any ids in [0, n0)); oracle/leiden_ref.py:_iteration
bool iteration(const Graph &g0, double gamma, double two_m, SplitMix64 &rng, std::vector<int32_t> &membership, Work &w) {
    const int32_t n0 = g0.n;
    Graph owned;               // the current aggregate (level >= 1)
    const Graph *g = &g0;
    std::vector<int32_t> part = membership, node_of(n0), node2new, part2;
    first_appearance(part);
    std::iota(node_of.begin(), node_of.end(), 0);
    bool improved = false;
    static const bool trace = getenv("DD_LEIDEN_TRACE") != nullptr;
    if (trace) fprintf(stderr, "leiden: iteration\n");
    for (;;) {
        const int32_t n = g->n;
        const auto t0 = std::chrono::steady_clock::now();
        degrees(*g, w.k);
        if (move_nodes(*g, gamma, two_m, rng, part, w)) improved = true;
        const auto t1 = std::chrono::steady_clock::now();
        for (int32_t v = 0; v < n0; v++) membership[v] = part[node_of[v]];
        {
            std::vector<int32_t> ids = part;
            if (first_appearance(ids) == n) break;  // every node is its own community
        }
        refine(*g, gamma, two_m, rng, part, w);
        const auto t2 = std::chrono::steady_clock::now();
        Graph ng;
        ddlv::aggregate(*g, w.refined, ng, node2new);
        if (trace) {
            const auto t3 = std::chrono::steady_clock::now();
            auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            fprintf(stderr, "leiden:   level n=%d nnz=%zu moves=%lld: move %.1f ms, refine %.1f ms, aggregate %.1f ms -> n=%d\n", n,
                    g->indices.size(), (long long)w.n_moves, ms(t0, t1), ms(t1, t2), ms(t2, t3), ng.n);
        }
        if (ng.n == n) break;  // the refinement merged nothing
        part2.assign(ng.n, 0);
        for (int32_t i = 0; i < n; i++) part2[node2new[i]] = part[i];
        first_appearance(part2);
        part = part2;
        for (int32_t v = 0; v < n0; v++) node_of[v] = node2new[node_of[v]];
        owned = std::move(ng);
        g = &owned;
    }
    return improved;
}

int run_leiden(const Graph &g, double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_comm_out) {
    const int32_t n = g.n;
    double two_m = 0.0;
    if (g.weights.empty())
        two_m = (double)g.indices.size();
    else
        for (size_t e = 0; e < g.indices.size(); e++) two_m += g.weights[e];
    std::vector<int32_t> membership(n);
    std::iota(membership.begin(), membership.end(), 0);
    SplitMix64 rng{seed};
    Work w;
    if (two_m > 0.0)
        for (int it = 0; it < kMaxIterations; it++)
            if (!iteration(g, resolution, two_m, rng, membership, w)) break;
    ddlv::labels_by_size(membership, labels_out, n_comm_out);
    return DD_OK;
}

// umap smooth_knn_dist + compute_membership_strengths (local_connectivity = 1, bandwidth = 1): the directed
// float32 weights of the kNN lists; arithmetic pinned by oracle/upstream.py (float32 differences and stores,
// float64 bisection with libm exp, sequential sums).
int umap_weights(int64_t n, int32_t k, const int32_t *idx, const float *dist, std::vector<float> &wts) {
    constexpr double kTol = 1e-5, kMinScale = 1e-3;
    const double target = std::log2((double)k);
    std::vector<double> row_sum((size_t)n);
    double total = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double s = 0.0;
        for (int32_t c = 0; c < k; c++) s += (double)dist[i * k + c];
        row_sum[i] = s;
    }
    for (int64_t i = 0; i < n; i++) total += row_sum[i];
    const double mean_all = n * (int64_t)k > 0 ? total / (double)(n * (int64_t)k) : 0.0;
    wts.assign((size_t)n * k, 0.f);
    std::vector<double> d((size_t)k);
    for (int64_t i = 0; i < n; i++) {
        const float *di = dist + i * k;
        float rho = 0.f;
        for (int32_t c = 0; c < k; c++)
            if (di[c] > 0.f) {
                rho = di[c];
                break;
            }
        for (int32_t c = 1; c < k; c++) d[c] = (double)(float)(di[c] - rho);
        double lo = 0.0, hi = std::numeric_limits<double>::infinity(), mid = 1.0;
        for (int it = 0; it < 64; it++) {
            double psum = 0.0;
            for (int32_t c = 1; c < k; c++) psum += d[c] > 0.0 ? std::exp(-(d[c] / mid)) : 1.0;
            if (std::fabs(psum - target) < kTol) break;
            if (psum > target) {
                hi = mid;
                mid = (lo + hi) / 2.0;
            } else {
                lo = mid;
                if (hi == std::numeric_limits<double>::infinity())
                    mid *= 2.0;
                else
                    mid = (lo + hi) / 2.0;
            }
        }
        float sigma = (float)mid;
        const double floor_ = kMinScale * (rho > 0.f ? row_sum[i] / (double)k : mean_all);
        if ((double)sigma < floor_) sigma = (float)floor_;
        for (int32_t c = 0; c < k; c++) {
            const int32_t j = idx[i * k + c];
            if (j < 0 || j >= n) return DD_ERR_ARG;
            const float diff = di[c] - rho;
            float v;
            if (j == i)
                v = 0.f;
            else if (diff <= 0.f || sigma == 0.f)
                v = 1.f;
            else
                v = (float)std::exp(-((double)diff / (double)sigma));
            wts[(size_t)i * k + c] = v;
        }
    }
    return DD_OK;
}

// fuzzy union W + W^T - W o W^T in float32 (what scipy computes for umap), zeros dropped, rows ascending
int umap_graph(int64_t n, int32_t k, const int32_t *idx, const float *dist, Graph &g) {
    std::vector<float> wts;
    const int rc = umap_weights(n, k, idx, dist, wts);
    if (rc != DD_OK) return rc;
    // incoming entries per node (counting sort by target)
    std::vector<int64_t> in_off((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; i++)
        for (int32_t c = 0; c < k; c++)
            if (wts[(size_t)i * k + c] != 0.f) in_off[idx[i * k + c] + 1]++;
    for (int64_t i = 0; i < n; i++) in_off[i + 1] += in_off[i];
    std::vector<int32_t> in_src((size_t)std::max<int64_t>(in_off[n], 1));
    std::vector<float> in_w(in_src.size());
    {
        std::vector<int64_t> fill(in_off.begin(), in_off.end() - 1);
        for (int64_t i = 0; i < n; i++)
            for (int32_t c = 0; c < k; c++) {
                const float v = wts[(size_t)i * k + c];
                if (v == 0.f) continue;
                const int64_t p = fill[idx[i * k + c]]++;
                in_src[p] = (int32_t)i;
                in_w[p] = v;
            }
    }
    struct Entry {
        int32_t j;
        float out, in;
    };
    std::vector<Entry> row;
    g.n = (int32_t)n;
    g.indptr.assign((size_t)n + 1, 0);
    g.indices.clear();
    g.weights.clear();
    g.indices.reserve((size_t)n * k * 3 / 2);
    g.weights.reserve((size_t)n * k * 3 / 2);
    for (int64_t i = 0; i < n; i++) {
        row.clear();
        for (int32_t c = 0; c < k; c++) {
            const float v = wts[(size_t)i * k + c];
            if (v != 0.f) row.push_back(Entry{idx[i * k + c], v, 0.f});
        }
        for (int64_t p = in_off[i]; p < in_off[i + 1]; p++) row.push_back(Entry{in_src[p], 0.f, in_w[p]});
        std::sort(row.begin(), row.end(), [](const Entry &a, const Entry &b) { return a.j < b.j; });
        for (size_t t = 0; t < row.size();) {
            float a = row[t].out, b = row[t].in;
            size_t u = t + 1;
            for (; u < row.size() && row[u].j == row[t].j; u++) {  // the same neighbour from both directions
                a += row[u].out;                                   // (lists hold a neighbour at most once: one of the
                b += row[u].in;                                    //  two terms is 0, the sum is exact)
            }
            const float s = a + b;
            const float p = a * b;
            const float cw = s - p;
            if (cw != 0.f) {
                g.indices.push_back(row[t].j);
                g.weights.push_back((double)cw);
            }
            t = u;
        }
        g.indptr[i + 1] = (int64_t)g.indices.size();
    }
    g.selfw.assign((size_t)n, 0.0);
    if (g.weights.empty()) g.weights.push_back(0.0);  // keep "empty == unit weights" unambiguous
    return DD_OK;
}

// The umap graph as the device builds it (louvain_gpu.cu: symmetric pattern with rows in arbitrary order, float32 union
// weights widened to double, 0 = no edge) in the form umap_graph() produces: rows ascending, zeros dropped.
int graph_from_device(int64_t n, const int32_t *off, const int32_t *adj, const double *w, Graph &g) {
    if (n > 0 && (!off || off[0] != 0)) return DD_ERR_ARG;
    const int64_t raw = n > 0 ? off[n] : 0;
    if (raw < 0 || (raw > 0 && (!adj || !w))) return DD_ERR_ARG;
    g.n = (int32_t)n;
    g.indptr.assign((size_t)n + 1, 0);
    g.indices.clear();
    g.weights.clear();
    g.indices.reserve((size_t)raw);
    g.weights.reserve((size_t)raw);
    std::vector<std::pair<int32_t, double>> row;
    for (int64_t i = 0; i < n; i++) {
        if (off[i + 1] < off[i] || off[i + 1] > raw) return DD_ERR_ARG;
        row.clear();
        for (int64_t e = off[i]; e < off[i + 1]; e++) {
            if (adj[e] < 0 || adj[e] >= n) return DD_ERR_ARG;
            if (w[e] != 0.0) row.emplace_back(adj[e], w[e]);
        }
        std::sort(row.begin(), row.end());
        for (const auto &e : row) {
            g.indices.push_back(e.first);
            g.weights.push_back(e.second);
        }
        g.indptr[i + 1] = (int64_t)g.indices.size();
    }
    g.selfw.assign((size_t)n, 0.0);
    if (g.weights.empty()) g.weights.push_back(0.0);  // keep "empty == unit weights" unambiguous
    return DD_OK;
}

}  // namespace

// The pipeline's path: the umap graph built on the device -> labels.
int dd_host_leiden_from_graph(int64_t n, const int32_t *off, const int32_t *adj, const double *w, double resolution,
                              uint64_t seed, int32_t *labels_out, int32_t *n_comm_out) {
    if (n < 0 || (n > 0 && !labels_out) || n >= (1ll << 31) - 1) return DD_ERR_ARG;
    Graph g;
    const int rc = graph_from_device(n, off, adj, w, g);
    if (rc != DD_OK) return rc;
    return run_leiden(g, resolution, seed, labels_out, n_comm_out);
}

int dd_host_umap_canonical(int64_t n, const int32_t *off, const int32_t *adj, const double *w, int64_t *indptr_out,
                           int32_t *indices_out, float *weights_out, int64_t capacity, int64_t *nnz_out) {
    Graph g;
    const int rc = graph_from_device(n, off, adj, w, g);
    if (rc != DD_OK) return rc;
    const int64_t nnz = (int64_t)g.indices.size();
    *nnz_out = nnz;
    if (capacity < nnz) return DD_OK;  // size query
    if (!indptr_out || (nnz > 0 && (!indices_out || !weights_out))) return DD_ERR_ARG;
    std::copy(g.indptr.begin(), g.indptr.end(), indptr_out);
    for (int64_t e = 0; e < nnz; e++) {
        indices_out[e] = g.indices[e];
        weights_out[e] = (float)g.weights[e];
    }
    return DD_OK;
}

// Host-callable form of the pipeline's second half (tests: the slot layout of dd_fit_iterations without a GPU).
extern "C" int dd_leiden_device_graph(int64_t n, const int32_t *off, const int32_t *adj, const double *weights, double resolution,
                                      uint64_t seed, int32_t *labels_out, int32_t *n_communities_out) {
    const int rc = dd_host_leiden_from_graph(n, off, adj, weights, resolution, seed, labels_out, n_communities_out);
    if (rc != DD_OK) dd_set_global_error("dd_leiden_device_graph: bad arguments or malformed graph (offsets, neighbour index out of range)");
    return rc;
}

// The host entry: kNN lists (self in column 0) + float32 distances -> labels.
int dd_host_leiden_knn(int64_t n, int32_t k, const int32_t *knn_idx, const float *knn_dist, double resolution,
                       uint64_t seed, int32_t *labels_out, int32_t *n_comm_out) {
    if (n < 0 || k < 2 || (n > 0 && (!knn_idx || !knn_dist || !labels_out))) return DD_ERR_ARG;
    if (n >= (1ll << 31) - 1) return DD_ERR_UNSUPPORTED;
    Graph g;
    const int rc = umap_graph(n, k, knn_idx, knn_dist, g);
    if (rc != DD_OK) return rc;
    return run_leiden(g, resolution, seed, labels_out, n_comm_out);
}

extern "C" int dd_leiden_knn(int64_t n, int32_t k, const int32_t *knn_idx, const float *knn_dist, double resolution,
                             uint64_t seed, int32_t *labels_out, int32_t *n_communities_out) {
    const int rc = dd_host_leiden_knn(n, k, knn_idx, knn_dist, resolution, seed, labels_out, n_communities_out);
    if (rc != DD_OK) dd_set_global_error("dd_leiden_knn: bad arguments (null pointer, k < 2 or neighbour index out of range)");
    return rc;
}

extern "C" int dd_umap_connectivities(int64_t n, int32_t k, const int32_t *knn_idx, const float *knn_dist,
                                      int64_t *indptr_out, int32_t *indices_out, float *weights_out, int64_t capacity,
                                      int64_t *nnz_out) {
    if (n < 0 || k < 2 || (n > 0 && (!knn_idx || !knn_dist)) || !nnz_out || n >= (1ll << 31) - 1) {
        dd_set_global_error("dd_umap_connectivities: bad arguments");
        return DD_ERR_ARG;
    }
    Graph g;
    const int rc = umap_graph(n, k, knn_idx, knn_dist, g);
    if (rc != DD_OK) {
        dd_set_global_error("dd_umap_connectivities: neighbour index out of range");
        return rc;
    }
    const int64_t nnz = (int64_t)g.indices.size();
    *nnz_out = nnz;
    if (capacity < nnz) return DD_OK;  // size query
    if (!indptr_out || (nnz > 0 && (!indices_out || !weights_out))) {
        dd_set_global_error("dd_umap_connectivities: null output");
        return DD_ERR_ARG;
    }
    std::copy(g.indptr.begin(), g.indptr.end(), indptr_out);
    for (int64_t e = 0; e < nnz; e++) {
        indices_out[e] = g.indices[e];
        weights_out[e] = (float)g.weights[e];
    }
    return DD_OK;
}

extern "C" int dd_leiden_csr(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                             double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_communities_out) {
    if (n < 0 || !indptr || (n > 0 && !labels_out) || n >= (1ll << 31) - 1) {
        dd_set_global_error("dd_leiden_csr: bad arguments");
        return DD_ERR_ARG;
    }
    const int64_t nnz = indptr[n];
    if (nnz > 0 && !indices) {
        dd_set_global_error("dd_leiden_csr: null indices");
        return DD_ERR_ARG;
    }
    Graph g;
    g.n = (int32_t)n;
    g.indptr.assign(indptr, indptr + n + 1);
    g.indices.resize(nnz);
    for (int64_t e = 0; e < nnz; e++) {
        if (indices[e] < 0 || indices[e] >= n) {
            dd_set_global_error("dd_leiden_csr: neighbour index out of range");
            return DD_ERR_ARG;
        }
        g.indices[e] = (int32_t)indices[e];
    }
    if (weights && nnz > 0) g.weights.assign(weights, weights + nnz);
    g.selfw.assign(n, 0.0);
    return run_leiden(g, resolution, seed, labels_out, n_communities_out);
}
