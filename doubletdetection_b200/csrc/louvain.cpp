// louvain.cpp -- the clustering call of _one_fit (doubletdetection.py:337-343): what
// sc.tl.louvain(adata, resolution=4, random_state, directed=False) optimises on the symmetrised kNN
// pattern -- RB-configuration modularity, resolution gamma, unweighted, seeded.  The louvain-igraph
// package is not available; the move order is specified by oracle/louvain_ref.py (FIFO queue in a
// SplitMix64-shuffled order, stay-wins-ties, first-best candidate in adjacency order, aggregation with
// ascending neighbour lists, labels by decreasing community size) and this file reproduces it label for
// label.  Host code on purpose: the graph has ~15 edges per node and the algorithm is a sequential
// sweep; iterations are clustered concurrently on host threads while the GPU runs ahead
// (dd_fit_iterations).
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <string>
#include <vector>

#include "dd_internal.h"
#include "louvain_host.h"

namespace {

using ddlv::Graph;
using ddlv::SplitMix64;

struct Scratch {
    std::vector<double> k, tot, neigh_w;
    std::vector<uint8_t> seen, inq;
    std::vector<int32_t> queue, cands;
};

// one level of local moving; returns true if any node moved
bool one_level(const Graph &g, double gamma, double two_m, SplitMix64 &rng, std::vector<int32_t> &comm, Scratch &sc) {
    const int32_t n = g.n;
    sc.k.assign(n, 0.0);
    sc.tot.assign(n, 0.0);
    sc.neigh_w.assign(n, 0.0);
    sc.seen.assign(n, 0);
    sc.inq.assign(n, 1);
    sc.queue.resize(n);
    sc.cands.resize(std::max<int32_t>(n, 1));
    comm.resize(n);
    const bool unit = g.weights.empty();
    for (int32_t i = 0; i < n; i++) {
        double s = g.selfw[i];
        if (unit) {
            s += (double)(g.indptr[i + 1] - g.indptr[i]);
        } else {
            for (int64_t e = g.indptr[i]; e < g.indptr[i + 1]; e++) s += g.weights[e];
        }
        sc.k[i] = s;
        sc.tot[i] = s;
        comm[i] = i;
        sc.queue[i] = i;
    }
    for (int64_t i = (int64_t)n - 1; i >= 1; i--) {
        const int64_t j = (int64_t)(rng.next() % (uint64_t)(i + 1));
        std::swap(sc.queue[i], sc.queue[j]);
    }
    double *tot = sc.tot.data(), *neigh_w = sc.neigh_w.data();
    uint8_t *seen = sc.seen.data(), *inq = sc.inq.data();
    int32_t *queue = sc.queue.data(), *cands = sc.cands.data();
    const int64_t *indptr = g.indptr.data();
    const int32_t *indices = g.indices.data();
    int32_t head = 0, count = n;
    bool moved_any = false;
    int32_t *comm_p = comm.data();
    const double *kk = sc.k.data();
    constexpr int kLocal = 48;  // low-degree nodes (every node of a kNN graph) gather their candidates on the stack
    int32_t lc[kLocal];
    double lw[kLocal];
    while (count > 0) {
        const int32_t i = queue[head];
        // The queue order is a random permutation, so every pop would miss the caches on the node's adjacency
        // list and on its neighbours' community ids; the FIFO content ahead is known -> software prefetch.
        if (count > 24) {
            int32_t p = head + 24;
            if (p >= n) p -= n;
            __builtin_prefetch(indptr + queue[p]);
            p = head + 12;
            if (p >= n) p -= n;
            const int64_t pe = indptr[queue[p]];
            __builtin_prefetch(indices + pe);
            __builtin_prefetch(indices + pe + 16);
            p = head + 4;
            if (p >= n) p -= n;
            const int32_t q4 = queue[p];
            const int64_t qe0 = indptr[q4], qe1 = std::min<int64_t>(indptr[q4 + 1], qe0 + 32);
            for (int64_t e = qe0; e < qe1; e++) __builtin_prefetch(comm_p + indices[e]);
            __builtin_prefetch(comm_p + q4);
            __builtin_prefetch(tot + q4);
        }
        head = (head + 1 == n) ? 0 : head + 1;
        count--;
        inq[i] = 0;
        const int32_t ci = comm_p[i];
        const double ki = kk[i];
        const int64_t e0 = indptr[i], e1 = indptr[i + 1];
        int32_t best = ci;
        if (e1 - e0 < kLocal) {
            // same candidate order (own community, then first appearance in the adjacency list) and the same
            // arithmetic as the array-based path below, without touching the per-community scratch arrays
            int32_t nc = 1;
            lc[0] = ci;
            lw[0] = 0.0;
            for (int64_t e = e0; e < e1; e++) {
                const int32_t c = comm_p[indices[e]];
                int32_t t = 0;
                while (t < nc && lc[t] != c) t++;
                if (t == nc) {
                    lc[nc] = c;
                    lw[nc] = 0.0;
                    nc++;
                }
                lw[t] += unit ? 1.0 : g.weights[e];
            }
            tot[ci] -= ki;
            const double gk = gamma * ki;
            double best_gain = lw[0] - (gk * tot[ci]) / two_m;
            for (int32_t t = 1; t < nc; t++) {
                const double gn = lw[t] - (gk * tot[lc[t]]) / two_m;
                if (gn > best_gain) {
                    best = lc[t];
                    best_gain = gn;
                }
            }
        } else {
            int32_t nc = 0;
            cands[nc++] = ci;
            seen[ci] = 1;
            neigh_w[ci] = 0.0;
            for (int64_t e = e0; e < e1; e++) {
                const int32_t c = comm_p[indices[e]];
                if (!seen[c]) {
                    seen[c] = 1;
                    neigh_w[c] = 0.0;
                    cands[nc++] = c;
                }
                neigh_w[c] += unit ? 1.0 : g.weights[e];
            }
            tot[ci] -= ki;
            const double gk = gamma * ki;
            double best_gain = neigh_w[ci] - (gk * tot[ci]) / two_m;
            for (int32_t t = 1; t < nc; t++) {
                const int32_t c = cands[t];
                const double gn = neigh_w[c] - (gk * tot[c]) / two_m;
                if (gn > best_gain) {
                    best = c;
                    best_gain = gn;
                }
            }
            for (int32_t t = 0; t < nc; t++) seen[cands[t]] = 0;
        }
        tot[best] += ki;
        if (best != ci) {
            comm_p[i] = best;
            moved_any = true;
            for (int64_t e = e0; e < e1; e++) {
                const int32_t j = indices[e];
                if (comm_p[j] != best && !inq[j]) {
                    inq[j] = 1;
                    int32_t tail = head + count;
                    if (tail >= n) tail -= n;
                    queue[tail] = j;
                    count++;
                }
            }
        }
    }
    return moved_any;
}

}  // namespace

// aggregate g by comm; node2new renumbers communities by first appearance over node index
void ddlv::aggregate(const Graph &g, const std::vector<int32_t> &comm, Graph &out, std::vector<int32_t> &node2new) {
    const int32_t n = g.n;
    std::vector<int32_t> new_id(n, -1);
    node2new.resize(n);
    int32_t nc = 0;
    for (int32_t i = 0; i < n; i++) {
        if (new_id[comm[i]] < 0) new_id[comm[i]] = nc++;
        node2new[i] = new_id[comm[i]];
    }
    std::vector<int64_t> mstart(nc + 1, 0);
    for (int32_t i = 0; i < n; i++) mstart[node2new[i] + 1]++;
    for (int32_t a = 0; a < nc; a++) mstart[a + 1] += mstart[a];
    std::vector<int32_t> members(std::max<int32_t>(n, 1));
    {
        std::vector<int64_t> fill(mstart.begin(), mstart.end() - 1);
        for (int32_t i = 0; i < n; i++) members[fill[node2new[i]]++] = i;  // ascending node order
    }
    out.n = nc;
    out.indptr.assign(nc + 1, 0);
    out.selfw.assign(nc, 0.0);
    // the aggregate has at most as many entries as g: written by position, trimmed at the end
    out.indices.resize(g.indices.size());
    out.weights.resize(g.indices.size());
    int32_t *const oi = out.indices.data();
    double *const ow = out.weights.data();
    int64_t pos = 0;
    std::vector<double> acc(nc, 0.0);
    // The neighbour communities of `a` must come out in ASCENDING order (the sweeps above add weights in adjacency order).
    // A two-level bit set (one bit per community, one summary bit per 64-bit word) replaces "collect + std::sort": marking
    // is one OR, the ordered walk visits only the set words -- the comparison sort of ~40 random ids per community was
    // the largest single item of the aggregation (50 of 120 ms at 125 k nodes).
    const int32_t n_words = (nc + 63) >> 6, n_summary = (n_words + 63) >> 6;
    std::vector<uint64_t> bits(std::max(n_words, 1), 0), summary(std::max(n_summary, 1), 0);
    for (int32_t a = 0; a < nc; a++) {
        double s = 0.0;
        for (int64_t p = mstart[a]; p < mstart[a + 1]; p++) {
            const int32_t i = members[p];
            if (p + 8 < n) {  // members[] is one array over all communities: the look-ahead crosses their boundaries
                g.prefetch_offset(members[p + 8]);
                g.prefetch_row(members[p + 4]);
            }
            s += g.selfw[i];
            for (int64_t e = g.indptr[i]; e < g.indptr[i + 1]; e++) {
                const int32_t b = node2new[g.indices[e]];
                const double w = g.w(e);
                if (b == a) {
                    s += w;
                } else {
                    uint64_t &word = bits[b >> 6];
                    const uint64_t m = 1ull << (b & 63);
                    if (!(word & m)) {
                        word |= m;
                        summary[b >> 12] |= 1ull << ((b >> 6) & 63);
                        acc[b] = 0.0;
                    }
                    acc[b] += w;
                }
            }
        }
        out.selfw[a] = s;
        for (int32_t si = 0; si < n_summary; si++) {
            uint64_t sw = summary[si];
            if (!sw) continue;
            summary[si] = 0;
            while (sw) {
                const int32_t wi = (si << 6) + __builtin_ctzll(sw);
                sw &= sw - 1;
                uint64_t bw = bits[wi];
                bits[wi] = 0;
                while (bw) {
                    const int32_t b = (wi << 6) + __builtin_ctzll(bw);
                    bw &= bw - 1;
                    oi[pos] = b;
                    ow[pos] = acc[b];
                    pos++;
                }
            }
        }
        out.indptr[a + 1] = pos;
    }
    out.indices.resize((size_t)pos);
    out.weights.resize((size_t)pos);
    if (out.weights.empty()) out.weights.push_back(0.0);  // keep "empty == unit weights" unambiguous
}

void ddlv::labels_by_size(std::vector<int32_t> &membership, int32_t *labels_out, int32_t *n_comm_out) {
    const int32_t n = (int32_t)membership.size();
    std::vector<int32_t> fa_of(n, -1);
    int32_t nc = 0;
    for (int32_t i = 0; i < n; i++) {
        if (fa_of[membership[i]] < 0) fa_of[membership[i]] = nc++;
        membership[i] = fa_of[membership[i]];
    }
    std::vector<int64_t> size(nc, 0);
    for (int32_t i = 0; i < n; i++) size[membership[i]]++;
    std::vector<int32_t> order(nc);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return size[a] > size[b]; });
    std::vector<int32_t> newlab(nc);
    for (int32_t r = 0; r < nc; r++) newlab[order[r]] = r;
    for (int32_t i = 0; i < n; i++) labels_out[i] = newlab[membership[i]];
    if (n_comm_out) *n_comm_out = nc;
}

namespace {

// First level by synchronous coloured rounds on the host (the twin of louvain_gpu.cu; specification:
// oracle/louvain_ref.py:level0_parallel).  Unweighted graphs only.
constexpr int kColours = 8, kMaxRounds = 32;
inline int colour_of(uint64_t seed, int32_t i) {
    uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (int)(z % kColours);
}

void level0_parallel_host(const Graph &g, double gamma, double two_m, uint64_t seed, std::vector<int32_t> &comm) {
    const int32_t n = g.n;
    comm.resize(n);
    std::vector<double> tot(n), cnt(n, 0.0);
    std::vector<int32_t> size(n, 1), desired(n, -1);
    std::vector<std::vector<int32_t>> bucket(kColours);
    for (int32_t i = 0; i < n; i++) {
        comm[i] = i;
        tot[i] = (double)(g.indptr[i + 1] - g.indptr[i]);
        bucket[colour_of(seed, i)].push_back(i);
    }
    for (int round = 0; round < kMaxRounds; round++) {
        int64_t moved = 0;
        for (int c = 0; c < kColours; c++) {
            for (int32_t i : bucket[c]) {  // decide from the frozen state
                desired[i] = -1;
                const int64_t e0 = g.indptr[i], e1 = g.indptr[i + 1];
                if (e1 == e0) continue;
                const int32_t ci = comm[i];
                const double ki = (double)(e1 - e0), gk = gamma * ki;
                for (int64_t e = e0; e < e1; e++) cnt[comm[g.indices[e]]] += 1.0;
                const double gain_stay = cnt[ci] - (gk * (tot[ci] - ki)) / two_m;
                int32_t best = -1;
                double best_gain = 0.0;
                for (int64_t e = e0; e < e1; e++) {
                    const int32_t cc = comm[g.indices[e]];
                    if (cc == ci) continue;
                    const double gn = cnt[cc] - (gk * tot[cc]) / two_m;
                    if (best < 0 || gn > best_gain || (gn == best_gain && cc < best)) {
                        best = cc;
                        best_gain = gn;
                    }
                }
                for (int64_t e = e0; e < e1; e++) cnt[comm[g.indices[e]]] = 0.0;
                if (best >= 0 && best_gain > gain_stay && !(size[ci] == 1 && size[best] == 1 && best > ci)) desired[i] = best;
            }
            for (int32_t i : bucket[c]) {  // apply simultaneously
                const int32_t b = desired[i];
                if (b < 0) continue;
                const int32_t ci = comm[i];
                const double ki = (double)(g.indptr[i + 1] - g.indptr[i]);
                comm[i] = b;
                tot[ci] -= ki;
                tot[b] += ki;
                size[ci]--;
                size[b]++;
                moved++;
            }
        }
        if (moved <= (int64_t)(n >> 9)) break;  // at most n / 512 moves: the level is settled
    }
}

// The same level for WEIGHTED graphs (specification: oracle/louvain_ref.py, "WEIGHTED graphs"; the device counterpart is
// round 2's work): weights in fixed point, wq = rint(w * 2^32) in int64, so that w(i, c), k_i, tot[c]
// and two_m are exact integer sums -- identical in any summation order, which is what lets a GPU apply simultaneous moves
// with atomics and still be bit-reproducible.  Gains: the unit formula on those integers converted to double.
void level0_parallel_host_w(const Graph &g, double gamma, uint64_t seed, std::vector<int32_t> &comm) {
    const int32_t n = g.n;
    comm.resize(n);
    const size_t nnz = g.indices.size();
    std::vector<int64_t> wq(nnz), k(n, 0), tot(n), cnt(n, 0);
    for (size_t e = 0; e < nnz; e++) wq[e] = (int64_t)std::nearbyint(g.weights[e] * 4294967296.0);  // ties to even, like numpy
    int64_t two_m_q = 0;
    std::vector<int32_t> size(n, 1), desired(n, -1);
    std::vector<std::vector<int32_t>> bucket(kColours);
    for (int32_t i = 0; i < n; i++) {
        comm[i] = i;
        for (int64_t e = g.indptr[i]; e < g.indptr[i + 1]; e++) k[i] += wq[e];
        tot[i] = k[i];
        two_m_q += k[i];
        if (g.indptr[i + 1] > g.indptr[i]) bucket[colour_of(seed, i)].push_back(i);
    }
    if (two_m_q == 0) return;
    const double two_m = (double)two_m_q;
    for (int round = 0; round < kMaxRounds; round++) {
        int64_t moved = 0;
        for (int c = 0; c < kColours; c++) {
            for (int32_t i : bucket[c]) {  // decide from the frozen state
                desired[i] = -1;
                const int64_t e0 = g.indptr[i], e1 = g.indptr[i + 1];
                const int32_t ci = comm[i];
                const double gk = gamma * (double)k[i];
                for (int64_t e = e0; e < e1; e++) cnt[comm[g.indices[e]]] += wq[e];
                const double gain_stay = (double)cnt[ci] - (gk * (double)(tot[ci] - k[i])) / two_m;
                int32_t best = -1;
                double best_gain = 0.0;
                for (int64_t e = e0; e < e1; e++) {
                    const int32_t cc = comm[g.indices[e]];
                    if (cc == ci) continue;
                    const double gn = (double)cnt[cc] - (gk * (double)tot[cc]) / two_m;
                    if (best < 0 || gn > best_gain || (gn == best_gain && cc < best)) {
                        best = cc;
                        best_gain = gn;
                    }
                }
                for (int64_t e = e0; e < e1; e++) cnt[comm[g.indices[e]]] = 0;
                if (best >= 0 && best_gain > gain_stay && !(size[ci] == 1 && size[best] == 1 && best > ci)) desired[i] = best;
            }
            for (int32_t i : bucket[c]) {  // apply simultaneously
                const int32_t b = desired[i];
                if (b < 0) continue;
                const int32_t ci = comm[i];
                comm[i] = b;
                tot[ci] -= k[i];
                tot[b] += k[i];
                size[ci]--;
                size[b]++;
                moved++;
            }
        }
        if (moved <= (int64_t)(n >> 9)) break;  // at most n / 512 moves: the level is settled
    }
}

// The first aggregation of the kNN pipeline: a unit-weight graph without self-loops whose ~10^2 first-level communities fit
// a dense count table.  Produces exactly what ddlv::aggregate produces (ids by first appearance, ascending neighbour lists,
// integer-valued weights -- counts are exact in any order) in one tight pass over the edges instead of per-community
// member lists.  Returns false (nothing written) when the graph does not qualify.
bool aggregate_unit_dense(const Graph &g, const std::vector<int32_t> &comm, Graph &out, std::vector<int32_t> &node2new) {
    constexpr int32_t kMaxDense = 1024;
    const int32_t n = g.n;
    if (!g.weights.empty()) return false;
    std::vector<int32_t> new_id(n, -1);
    node2new.resize(n);
    int32_t nc = 0;
    for (int32_t i = 0; i < n; i++) {
        if (g.selfw[i] != 0.0) return false;
        if (new_id[comm[i]] < 0) {
            if (nc == kMaxDense) return false;
            new_id[comm[i]] = nc++;
        }
        node2new[i] = new_id[comm[i]];
    }
    std::vector<int64_t> tab((size_t)nc * nc, 0);
    const int32_t *n2n = node2new.data();
    for (int32_t i = 0; i < n; i++) {
        int64_t *row = tab.data() + (size_t)n2n[i] * nc;
        for (int64_t e = g.indptr[i]; e < g.indptr[i + 1]; e++) row[n2n[g.indices[e]]]++;
    }
    out.n = nc;
    out.indptr.assign(nc + 1, 0);
    out.indices.clear();
    out.weights.clear();
    out.selfw.assign(nc, 0.0);
    for (int32_t a = 0; a < nc; a++) {
        const int64_t *row = tab.data() + (size_t)a * nc;
        out.selfw[a] = (double)row[a];
        for (int32_t b = 0; b < nc; b++)
            if (b != a && row[b] != 0) {
                out.indices.push_back(b);
                out.weights.push_back((double)row[b]);
            }
        out.indptr[a + 1] = (int64_t)out.indices.size();
    }
    if (out.weights.empty()) out.weights.push_back(0.0);  // keep "empty == unit weights" unambiguous
    return true;
}

// comm0: optional first-level partition (community id per node, any ids in [0, n)); parallel0: compute it here.
int run_louvain(Graph &g, double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_comm_out,
                const int32_t *comm0 = nullptr, bool parallel0 = false) {
    const int32_t n = g.n;
    double two_m = 0.0;
    if (g.weights.empty())
        two_m = (double)g.indices.size();
    else
        for (size_t e = 0; e < g.indices.size(); e++) two_m += g.weights[e];
    std::vector<int32_t> membership(n);
    std::iota(membership.begin(), membership.end(), 0);
    SplitMix64 rng{seed};
    Scratch sc;
    if (two_m > 0.0 && (comm0 != nullptr || parallel0)) {
        std::vector<int32_t> comm, node2new;
        if (comm0 != nullptr)
            comm.assign(comm0, comm0 + n);
        else if (g.weights.empty())
            level0_parallel_host(g, resolution, two_m, seed, comm);
        else
            level0_parallel_host_w(g, resolution, seed, comm);
        Graph ng;
        static const bool trace0 = getenv("DD_LOUVAIN_TRACE") != nullptr;
        const auto t0 = std::chrono::steady_clock::now();
        static const bool no_dense = getenv("DD_LOUVAIN_NO_DENSE_AGG") != nullptr;  // A/B switch
        if (no_dense || !aggregate_unit_dense(g, comm, ng, node2new)) ddlv::aggregate(g, comm, ng, node2new);
        if (trace0)
            fprintf(stderr, "louvain first aggregation: n=%d nnz=%zu -> n=%d in %.1f ms\n", g.n, g.indices.size(), ng.n,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        for (int32_t i = 0; i < n; i++) membership[i] = node2new[membership[i]];
        g = std::move(ng);
    }
    if (two_m > 0.0) {
        std::vector<int32_t> comm, node2new;
        static const bool trace = getenv("DD_LOUVAIN_TRACE") != nullptr;
        for (int level = 0; level < 64; level++) {
            const auto t0 = std::chrono::steady_clock::now();
            const bool moved = one_level(g, resolution, two_m, rng, comm, sc);
            const auto t1 = std::chrono::steady_clock::now();
            if (!moved) break;
            Graph ng;
            ddlv::aggregate(g, comm, ng, node2new);
            if (trace)
                fprintf(stderr, "louvain level %d: n=%d nnz=%zu move %.1f ms aggregate %.1f ms -> n=%d\n", level, g.n,
                        g.indices.size(), std::chrono::duration<double, std::milli>(t1 - t0).count(),
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), ng.n);
            for (int32_t i = 0; i < n; i++) membership[i] = node2new[membership[i]];
            g = std::move(ng);
        }
    }
    ddlv::labels_by_size(membership, labels_out, n_comm_out);
    return DD_OK;
}

}  // namespace

// Symmetric 0/1 pattern of the neighbour graph: i ~ j iff j in kNN(i)\{i} or i in kNN(j)\{j}
// (sc.tl.louvain ignores the connectivities' weights: use_weights=False).
int dd_host_louvain_knn(int64_t n, int32_t k, const int32_t *knn_idx, double resolution, uint64_t seed,
                        int32_t *labels_out, int32_t *n_comm_out) {
    if (n < 0 || k < 1 || (n > 0 && (!knn_idx || !labels_out))) return DD_ERR_ARG;
    if (n >= (1ll << 31) - 1) return DD_ERR_UNSUPPORTED;
    Graph g;
    g.n = (int32_t)n;
    std::vector<int64_t> deg(n + 1, 0);
    for (int64_t i = 0; i < n; i++)
        for (int32_t c = 0; c < k; c++) {
            const int32_t j = knn_idx[i * k + c];
            if (j < 0 || j >= n) return DD_ERR_ARG;
            if (j == i) continue;
            deg[i + 1]++;
            deg[j + 1]++;
        }
    for (int64_t i = 0; i < n; i++) deg[i + 1] += deg[i];
    std::vector<int32_t> adj(deg[n] > 0 ? deg[n] : 1);
    {
        std::vector<int64_t> fill(deg.begin(), deg.end() - 1);
        for (int64_t i = 0; i < n; i++)
            for (int32_t c = 0; c < k; c++) {
                const int32_t j = knn_idx[i * k + c];
                if (j == i) continue;
                adj[fill[i]++] = j;
                adj[fill[j]++] = (int32_t)i;
            }
    }
    g.indptr.assign(n + 1, 0);
    g.indices.reserve(deg[n]);
    for (int64_t i = 0; i < n; i++) {
        int32_t *b = adj.data() + deg[i], *e = adj.data() + deg[i + 1];
        std::sort(b, e);
        e = std::unique(b, e);
        g.indices.insert(g.indices.end(), b, e);
        g.indptr[i + 1] = (int64_t)g.indices.size();
    }
    g.selfw.assign(n, 0.0);
    return run_louvain(g, resolution, seed, labels_out, n_comm_out, nullptr, /*parallel0=*/true);
}

// The pipeline's path: the GPU built the pattern graph and optimised the first level; aggregate and finish here.
int dd_host_louvain_from_level0(int64_t n, const int32_t *off, const int32_t *adj, const int32_t *comm0,
                                double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_comm_out) {
    if (n < 0 || (n > 0 && (!off || !comm0 || !labels_out))) return DD_ERR_ARG;
    Graph g;
    g.n = (int32_t)n;
    g.indptr.resize(n + 1);
    for (int64_t i = 0; i <= n; i++) g.indptr[i] = off[i];
    const int64_t nnz = n > 0 ? off[n] : 0;
    if (nnz > 0 && !adj) return DD_ERR_ARG;
    g.indices.assign(adj, adj + nnz);
    for (int64_t i = 0; i < n; i++)
        if (comm0[i] < 0 || comm0[i] >= n) return DD_ERR_ARG;
    g.selfw.assign(n, 0.0);
    return run_louvain(g, resolution, seed, labels_out, n_comm_out, comm0, false);
}

// ---- PhenoGraph (doubletdetection.py:318-327 -> phenograph.cluster; specification: oracle/upstream.py
// phenograph_cluster).  The graph arrives as CSR with rows in any order and zero weights for pruned entries.
namespace {
int phenograph_finish(Graph &g, uint64_t seed, int32_t min_cluster_size, int32_t *labels_out, int32_t *n_comm_out,
                      const int32_t *comm0 = nullptr) {
    const int32_t n = g.n;
    int32_t nc = 0;
    // standard modularity; labels by decreasing size.  The first level is the synchronous coloured one on fixed-point
    // weights (oracle/louvain_ref.py:level0_parallel): comm0 = already done on the device, else its host twin runs here
    const int rc = run_louvain(g, 1.0, seed, labels_out, &nc, comm0, /*parallel0=*/comm0 == nullptr);
    if (rc != DD_OK) return rc;
    std::vector<int64_t> size(std::max(nc, 1), 0);
    for (int32_t i = 0; i < n; i++) size[labels_out[i]]++;
    for (int32_t i = 0; i < n; i++)
        if (size[labels_out[i]] <= min_cluster_size) labels_out[i] = -1;  // phenograph.core.sort_by_size keeps sizes > min_size
    if (n_comm_out) *n_comm_out = nc;
    return DD_OK;
}
}  // namespace

int dd_host_phenograph_from_graph(int64_t n, const int32_t *off, const int32_t *adj, const double *w, uint64_t seed,
                                  int32_t min_cluster_size, int32_t *labels_out, int32_t *n_comm_out, const int32_t *comm0) {
    if (n < 0 || (n > 0 && (!off || !labels_out))) return DD_ERR_ARG;
    Graph g;
    g.n = (int32_t)n;
    g.indptr.assign(n + 1, 0);
    const int64_t nnz_in = n > 0 ? off[n] : 0;
    if (nnz_in > 0 && (!adj || !w)) return DD_ERR_ARG;
    g.indices.reserve(nnz_in);
    g.weights.reserve(nnz_in);
    std::vector<std::pair<int32_t, double>> row;
    for (int64_t i = 0; i < n; i++) {
        row.clear();
        for (int64_t p = off[i]; p < off[i + 1]; p++) {
            if (adj[p] < 0 || adj[p] >= n) return DD_ERR_ARG;
            if (w[p] != 0.0) row.emplace_back(adj[p], w[p]);
        }
        std::sort(row.begin(), row.end(), [](const std::pair<int32_t, double> &a, const std::pair<int32_t, double> &b) {
            return a.first < b.first;
        });
        for (const auto &e : row) {
            g.indices.push_back(e.first);
            g.weights.push_back(e.second);
        }
        g.indptr[i + 1] = (int64_t)g.indices.size();
    }
    g.selfw.assign(n, 0.0);
    if (comm0)
        for (int64_t i = 0; i < n; i++)
            if (comm0[i] < 0 || comm0[i] >= n) return DD_ERR_ARG;
    return phenograph_finish(g, seed, min_cluster_size, labels_out, n_comm_out, comm0);
}

// Host twin of the device path: the same graph from the kNN lists (n x k, self in column 0) on the CPU.
extern "C" int dd_phenograph_knn(int64_t n, int32_t k, const int32_t *knn_idx, int32_t prune, int32_t min_cluster_size,
                                 uint64_t seed, int32_t *labels_out, int32_t *n_communities_out) {
    if (n < 0 || k < 2 || (n > 0 && (!knn_idx || !labels_out)) || n >= (1ll << 31) - 1) {
        dd_set_global_error("dd_phenograph_knn: bad arguments");
        return DD_ERR_ARG;
    }
    const int kk = k - 1;
    std::vector<int32_t> sorted((size_t)n * kk);
    for (int64_t i = 0; i < n; i++) {
        for (int c = 0; c < kk; c++) {
            const int32_t j = knn_idx[i * k + 1 + c];
            if (j < 0 || j >= n) {
                dd_set_global_error("dd_phenograph_knn: neighbour index out of range");
                return DD_ERR_ARG;
            }
            sorted[i * kk + c] = j;
        }
        std::sort(sorted.begin() + i * kk, sorted.begin() + (i + 1) * kk);
    }
    auto has = [&](int64_t i, int32_t j) {
        return std::binary_search(sorted.begin() + i * kk, sorted.begin() + (i + 1) * kk, j);
    };
    // symmetric pattern: out-neighbours and in-neighbours
    std::vector<std::vector<int32_t>> nbr(n);
    for (int64_t i = 0; i < n; i++)
        for (int c = 0; c < kk; c++) {
            const int32_t j = sorted[i * kk + c];
            if (j == i) continue;
            nbr[i].push_back(j);
            if (!has(j, (int32_t)i)) nbr[j].push_back((int32_t)i);
        }
    std::vector<int32_t> off(n + 1, 0), adj;
    std::vector<double> w;
    for (int64_t i = 0; i < n; i++) {
        std::sort(nbr[i].begin(), nbr[i].end());
        nbr[i].erase(std::unique(nbr[i].begin(), nbr[i].end()), nbr[i].end());
        for (int32_t j : nbr[i]) {
            int s = 0;
            const int32_t *a = sorted.data() + i * kk, *b = sorted.data() + (int64_t)j * kk;
            for (int x = 0, y = 0; x < kk && y < kk;) {
                if (a[x] < b[y]) x++;
                else if (a[x] > b[y]) y++;
                else { s++; x++; y++; }
            }
            const double wij = (double)s / (double)(2 * kk - s);
            const bool mutual = has(i, j) && has(j, (int32_t)i);
            adj.push_back(j);
            w.push_back(prune ? (mutual ? wij * wij : 0.0) : (mutual ? wij : wij * 0.5));
        }
        off[i + 1] = (int32_t)adj.size();
    }
    const int rc = dd_host_phenograph_from_graph(n, off.data(), adj.data(), w.data(), seed, min_cluster_size, labels_out,
                                                 n_communities_out, nullptr);
    if (rc != DD_OK) dd_set_global_error("dd_phenograph_knn: clustering failed");
    return rc;
}

extern "C" int dd_louvain_knn(int64_t n, int32_t k, const int32_t *knn_idx, double resolution, uint64_t seed,
                              int32_t *labels_out, int32_t *n_communities_out) {
    const int rc = dd_host_louvain_knn(n, k, knn_idx, resolution, seed, labels_out, n_communities_out);
    if (rc != DD_OK) dd_set_global_error("dd_louvain_knn: bad arguments (null pointer, k < 1 or neighbour index out of range)");
    return rc;
}

namespace {
int louvain_csr_impl(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights, double resolution,
                     uint64_t seed, int32_t *labels_out, int32_t *n_communities_out, bool parallel0);
}

extern "C" int dd_louvain_csr(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                              double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_communities_out) {
    return louvain_csr_impl(n, indptr, indices, weights, resolution, seed, labels_out, n_communities_out, false);
}

extern "C" int dd_louvain_csr_level0(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                                     double resolution, uint64_t seed, int32_t *labels_out, int32_t *n_communities_out) {
    return louvain_csr_impl(n, indptr, indices, weights, resolution, seed, labels_out, n_communities_out, true);
}

namespace {
int louvain_csr_impl(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights, double resolution,
                     uint64_t seed, int32_t *labels_out, int32_t *n_communities_out, bool parallel0) {
    if (n < 0 || !indptr || (n > 0 && !labels_out) || n >= (1ll << 31) - 1) {
        dd_set_global_error("dd_louvain_csr: bad arguments");
        return DD_ERR_ARG;
    }
    const int64_t nnz = indptr[n];
    if (nnz > 0 && !indices) {
        dd_set_global_error("dd_louvain_csr: null indices");
        return DD_ERR_ARG;
    }
    Graph g;
    g.n = (int32_t)n;
    g.indptr.assign(indptr, indptr + n + 1);
    g.indices.resize(nnz);
    for (int64_t e = 0; e < nnz; e++) {
        if (indices[e] < 0 || indices[e] >= n) {
            dd_set_global_error("dd_louvain_csr: neighbour index out of range");
            return DD_ERR_ARG;
        }
        g.indices[e] = (int32_t)indices[e];
    }
    if (weights && nnz > 0) g.weights.assign(weights, weights + nnz);
    g.selfw.assign(n, 0.0);
    return run_louvain(g, resolution, seed, labels_out, n_communities_out, nullptr, parallel0);
}
}  // namespace
