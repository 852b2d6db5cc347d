// louvain_gpu_w.cu -- the first Louvain level by synchronous coloured rounds for WEIGHTED graphs: what dd_fit_iterations
// runs for PhenoGraph's Jaccard graph (the reference's default clustering, doubletdetection.py:317-325).  Measured and
// parity-green on B200 (tests/test_gpu_pheno_level0.py, profiles/r2a_weighted_level.log, r2z_ncu_clustering_summary.md).
//
// Specification: oracle/louvain_ref.py:level0_parallel(..., weights); host twin: louvain.cpp:level0_parallel_host_w.  The
// level works on fixed-point weights wq = rint(w * 2^32) held in int64: w(i, c), k_i, tot[c] and two_m are exact integer
// sums, identical in any order, so the simultaneous moves of a sub-round can be applied with 64-bit integer atomics and
// the result is still bit-reproducible (the float64 atomics of the unweighted kernel are exact only because degrees are
// integers).  The gain is the unweighted formula evaluated in double on those integers (explicit _rn intrinsics: no
// contraction), candidates are compared by (gain, smaller id), singletons never move into a larger-id singleton.
// The device-built rows are compacted in place first (pruned entries to the tail) and the 32 x 17 steps are one CUDA-graph
// replay on a clustering lane.
#include "dd_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

constexpr int kColoursW = 8;
constexpr int kMaxRoundsW = 32;
constexpr int kTableW = 256, kTableMaxDegW = 192, kTableShiftW = 24;  // per-warp open addressing, hash = top 8 bits

inline int colour_of_w(uint64_t seed, int i) {
    uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (int)(z % kColoursW);
}

__global__ void k_lvw_init(const int32_t *__restrict__ off, const long long *__restrict__ wq, int n, int32_t *__restrict__ comm,
                           long long *__restrict__ k, long long *__restrict__ tot, int32_t *__restrict__ csize,
                           int32_t *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) counters[0] = counters[1] = counters[2] = 0;  // moved this round, done flag, rounds executed
    if (i >= n) return;
    long long s = 0;
    for (int e = off[i]; e < off[i + 1]; e++) s += wq[e];
    k[i] = s;
    tot[i] = s;
    comm[i] = i;
    csize[i] = 1;
}

__device__ __forceinline__ double gain_of(long long w_ic, double gk, long long tot_c, double two_m) {
    return __dsub_rn((double)w_ic, __ddiv_rn(__dmul_rn(gk, (double)tot_c), two_m));
}

// one warp per node: the node's desired community, or -1 (every lane returns it).  d = number of (leading) entries of the row
// that count; tkey/tsum: this warp's kTableW-slot table in shared memory.
__device__ __forceinline__ int propose_one_w(const int32_t *__restrict__ off, const int32_t *__restrict__ adj,
                                             const long long *__restrict__ wq, const int32_t *comm,
                                             const long long *__restrict__ k, const long long *tot, const int32_t *csize, int i,
                                             int d, double gamma, double two_m, int32_t *tkey, unsigned long long *tsum,
                                             int lane) {
    const int s = off[i];
    if (d <= 0) return -1;
    const int ci = __ldcg(comm + i);
    const long long ki = k[i];
    const double gk = __dmul_rn(gamma, (double)ki);
    double best_gain = 0.0;
    int best = 0x7fffffff;
    long long w_stay = 0;
    if (d <= kTableMaxDegW) {
        __syncwarp();
        for (int t = lane; t < kTableW; t += 32) {
            tkey[t] = -1;
            tsum[t] = 0ull;
        }
        __syncwarp();
        for (int f = lane; f < d; f += 32) {
            if (wq[s + f] == 0) continue;  // pruned entry of the device-built graph: not an edge
            const int c = __ldcg(comm + adj[s + f]);
            unsigned slot = ((unsigned)c * 2654435761u) >> kTableShiftW;
            for (;;) {
                const int prev = atomicCAS(tkey + slot, -1, c);
                if (prev == -1 || prev == c) break;
                slot = (slot + 1) & (kTableW - 1);
            }
            atomicAdd(tsum + slot, (unsigned long long)wq[s + f]);
        }
        __syncwarp();
        for (int t = lane; t < kTableW; t += 32) {
            const int c = tkey[t];
            if (c < 0) continue;
            const long long wc = (long long)tsum[t];
            if (c == ci) {
                w_stay = wc;
                continue;
            }
            const double gn = gain_of(wc, gk, __ldcg(tot + c), two_m);
            if (best == 0x7fffffff || gn > best_gain || (gn == best_gain && c < best)) {
                best = c;
                best_gain = gn;
            }
        }
        __syncwarp();
    } else {
        for (int f = lane; f < d; f += 32)
            if (__ldcg(comm + adj[s + f]) == ci) w_stay += wq[s + f];
        for (int e = lane; e < d; e += 32) {
            if (wq[s + e] == 0) continue;
            const int c = __ldcg(comm + adj[s + e]);
            if (c == ci) continue;
            long long wc = 0;
            for (int f = 0; f < d; f++)
                if (__ldcg(comm + adj[s + f]) == c) wc += wq[s + f];
            const double gn = gain_of(wc, gk, __ldcg(tot + c), two_m);
            if (best == 0x7fffffff || gn > best_gain || (gn == best_gain && c < best)) {
                best = c;
                best_gain = gn;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w_stay += __shfl_xor_sync(0xffffffffu, w_stay, o);  // table path: one lane holds it; fallback: partial sums
        const double og = __shfl_xor_sync(0xffffffffu, best_gain, o);
        const int ob = __shfl_xor_sync(0xffffffffu, best, o);
        if (ob != 0x7fffffff && (best == 0x7fffffff || og > best_gain || (og == best_gain && ob < best))) {
            best = ob;
            best_gain = og;
        }
    }
    const double gain_stay = gain_of(w_stay, gk, __ldcg(tot + ci) - ki, two_m);
    if (best != 0x7fffffff && best_gain > gain_stay &&
        !(__ldcg(csize + ci) == 1 && __ldcg(csize + best) == 1 && best > ci))
        return best;
    return -1;
}

// one warp per node of the current colour (test hook dd_louvain_level0_weighted: explicit graphs of any degree)
__global__ void __launch_bounds__(256) k_lvw_propose(const int32_t *__restrict__ off, const int32_t *__restrict__ adj,
                                                     const long long *__restrict__ wq, const int32_t *comm,
                                                     const long long *__restrict__ k, const long long *tot,
                                                     const int32_t *csize, const int32_t *__restrict__ bucket, int b0, int b1,
                                                     double gamma, double two_m_arg, const long long *__restrict__ two_m_dev,
                                                     int32_t *__restrict__ desired, const int32_t *__restrict__ counters) {
    __shared__ int32_t s_key[8 * kTableW];
    __shared__ unsigned long long s_sum[8 * kTableW];
    if (counters[1]) return;  // the level settled in an earlier round
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wl;
    if (b0 + w >= b1) return;  // whole warp
    const int i = bucket[b0 + w];
    // pipeline flavour: 2m was summed on the device (k_lvw_prepare); test hook: the host passes it
    const double two_m = two_m_dev ? (double)*two_m_dev : two_m_arg;
    const int res = propose_one_w(off, adj, wq, comm, k, tot, csize, i, off[i + 1] - off[i], gamma, two_m, s_key + wl * kTableW,
                                  s_sum + wl * kTableW, lane);
    if (lane == 0) desired[i] = res;
}

// EIGHT LANES per node, four nodes per warp -- the weighted twin of louvain_gpu.cu:k_lv_propose_g (same reasoning: a
// PhenoGraph node has <= 30 mutual neighbours, a warp per node wastes the launch on table clears).  deg[i] = number of
// LEADING entries of row i that are edges (k_lvw_prepare compacts the rows in place: pruned entries go to the tail).  A lane
// holds up to four (neighbour, fixed-point weight) pairs; w(i, c) comes from group-wide shuffles in exact int64 arithmetic,
// so the result equals propose_one_w's.  Nodes with more than 32 edges (prune=False hubs) are handed to the whole warp.
constexpr int kGroupLanesW = 8, kPerLaneW = 4, kPropWarpsW = 4, kNodesPerCtaW = kPropWarpsW * (32 / kGroupLanesW);
__global__ void __launch_bounds__(kPropWarpsW * 32) k_lvw_propose_g(const int32_t *__restrict__ off, const int32_t *__restrict__ deg,
                                                                     const int32_t *__restrict__ adj, const long long *__restrict__ wq,
                                                                     const int32_t *comm, const long long *__restrict__ k,
                                                                     const long long *tot, const int32_t *csize,
                                                                     const int32_t *__restrict__ bucket, int b0, int b1, double gamma,
                                                                     const long long *__restrict__ two_m_dev,
                                                                     int32_t *__restrict__ desired,
                                                                     const int32_t *__restrict__ counters) {
    __shared__ int32_t s_key[kPropWarpsW * kTableW];
    __shared__ unsigned long long s_sum[kPropWarpsW * kTableW];
    if (__ldcg(counters + 1)) return;  // the level settled in an earlier round
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int gl = lane & (kGroupLanesW - 1), grp = lane / kGroupLanesW;
    const unsigned gmask = ((1u << kGroupLanesW) - 1u) << (grp * kGroupLanesW);
    const int t = (blockIdx.x * kPropWarpsW + wl) * (32 / kGroupLanesW) + grp;
    const double two_m = (double)__ldcg(two_m_dev);
    int i = -1, s = 0, d = 0;
    if (b0 + t < b1) {
        i = bucket[b0 + t];
        s = off[i];
        d = deg[i];
    }
    const bool big = d > kGroupLanesW * kPerLaneW;
    if (i >= 0 && !big) {  // uniform inside the group
        int res = -1;
        if (d > 0) {
            int a[kPerLaneW];
            long long q[kPerLaneW];
#pragma unroll
            for (int j = 0; j < kPerLaneW; j++) {
                const int e = j * kGroupLanesW + gl;
                a[j] = e < d ? adj[s + e] : -1;
                q[j] = e < d ? wq[s + e] : 0ll;
            }
            const int ci = __ldcg(comm + i);
            const long long ki = k[i];
            int c[kPerLaneW];
#pragma unroll
            for (int j = 0; j < kPerLaneW; j++) c[j] = a[j] >= 0 ? __ldcg(comm + a[j]) : -1 - (j * kGroupLanesW + gl);  // unique sentinels
            long long tt[kPerLaneW];
#pragma unroll
            for (int j = 0; j < kPerLaneW; j++) tt[j] = (c[j] >= 0 && c[j] != ci) ? __ldcg(tot + c[j]) : 0ll;
            const long long tot_ci = __ldcg(tot + ci);
            const int cs_ci = __ldcg(csize + ci);
            long long wc[kPerLaneW] = {0, 0, 0, 0};
            long long w_stay = 0;
#pragma unroll
            for (int src = 0; src < kGroupLanesW; src++) {
#pragma unroll
                for (int jj = 0; jj < kPerLaneW; jj++) {
                    const int o = __shfl_sync(gmask, c[jj], grp * kGroupLanesW + src);
                    const long long ow = __shfl_sync(gmask, q[jj], grp * kGroupLanesW + src);
#pragma unroll
                    for (int j = 0; j < kPerLaneW; j++) wc[j] += (o == c[j]) ? ow : 0ll;
                    w_stay += (o == ci) ? ow : 0ll;
                }
            }
            const double gk = __dmul_rn(gamma, (double)ki);
            double best_gain = 0.0;
            int best = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < kPerLaneW; j++) {
                if (c[j] < 0 || c[j] == ci) continue;
                const double gn = gain_of(wc[j], gk, tt[j], two_m);
                if (best == 0x7fffffff || gn > best_gain || (gn == best_gain && c[j] < best)) {
                    best = c[j];
                    best_gain = gn;
                }
            }
#pragma unroll
            for (int o = kGroupLanesW / 2; o > 0; o >>= 1) {
                const double og = __shfl_xor_sync(gmask, best_gain, o);
                const int ob = __shfl_xor_sync(gmask, best, o);
                if (ob != 0x7fffffff && (best == 0x7fffffff || og > best_gain || (og == best_gain && ob < best))) {
                    best = ob;
                    best_gain = og;
                }
            }
            const double gain_stay = gain_of(w_stay, gk, tot_ci - ki, two_m);
            if (best != 0x7fffffff && best_gain > gain_stay && !(cs_ci == 1 && __ldcg(csize + best) == 1 && best > ci))
                res = best;
        }
        if (gl == 0) desired[i] = res;
    }
    // hubs: one at a time with the whole warp
    unsigned bigmask = __ballot_sync(0xffffffffu, big && gl == 0);
    while (bigmask) {
        const int src = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const int node = __shfl_sync(0xffffffffu, i, src);
        const int dn = __shfl_sync(0xffffffffu, d, src);
        const int res = propose_one_w(off, adj, wq, comm, k, tot, csize, node, dn, gamma, two_m, s_key + wl * kTableW,
                                      s_sum + wl * kTableW, lane);
        if (lane == 0) desired[node] = res;
    }
}

__global__ void k_lvw_apply(int32_t *__restrict__ comm, const long long *__restrict__ k, long long *__restrict__ tot,
                            int32_t *__restrict__ csize, const int32_t *__restrict__ bucket, int b0, int b1,
                            const int32_t *__restrict__ desired, int32_t *__restrict__ counters) {
    if (counters[1]) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (b0 + t >= b1) return;
    const int i = bucket[b0 + t];
    const int b = desired[i];
    if (b < 0) return;
    const int ci = comm[i];
    const unsigned long long ki = (unsigned long long)k[i];
    comm[i] = b;
    atomicAdd(reinterpret_cast<unsigned long long *>(tot + ci), 0ull - ki);  // two's complement: exact in any order
    atomicAdd(reinterpret_cast<unsigned long long *>(tot + b), ki);
    atomicSub(csize + ci, 1);
    atomicAdd(csize + b, 1);
    atomicAdd(counters, 1);
}

__global__ void k_lvw_round_end(int32_t *__restrict__ counters, int n) {
    if (counters[1]) return;
    counters[2]++;
    if (counters[0] <= (n >> 9)) counters[1] = 1;
    counters[0] = 0;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, sizeof(T) * (count ? count : 1)); }
};

}  // namespace

// Test hook: the weighted first level on an explicit symmetric CSR graph without self-loops and without zero weights
// (host arrays).  comm_out int32[n]: the community (a node id) of every node after the level; rounds_out: rounds executed.
extern "C" int dd_louvain_level0_weighted(dd_handle *h, int64_t n, const int64_t *indptr, const int64_t *indices,
                                          const double *weights, double gamma, uint64_t seed, int32_t *comm_out,
                                          int32_t *rounds_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_louvain_level0_weighted: null handle");
    if (n < 1 || !indptr || !comm_out) return dd_fail(h, DD_ERR_ARG, "dd_louvain_level0_weighted: null argument");
    const int64_t nnz = indptr[n];
    if (nnz < 0 || (nnz > 0 && (!indices || !weights)) || n >= (1ll << 31) - 1 || nnz >= (1ll << 31) - 1)
        return dd_fail(h, DD_ERR_ARG, "dd_louvain_level0_weighted: bad graph");
    DD_CUDA(h, cudaSetDevice(h->device));
    std::vector<int32_t> off(n + 1), adj((size_t)std::max<int64_t>(nnz, 1));
    std::vector<long long> wq((size_t)std::max<int64_t>(nnz, 1));
    long long two_m_q = 0;
    for (int64_t i = 0; i <= n; i++) off[i] = (int32_t)indptr[i];
    for (int64_t e = 0; e < nnz; e++) {
        if (indices[e] < 0 || indices[e] >= n) return dd_fail(h, DD_ERR_ARG, "dd_louvain_level0_weighted: index out of range");
        adj[e] = (int32_t)indices[e];
        wq[e] = (long long)std::nearbyint(weights[e] * 4294967296.0);
        if (wq[e] <= 0) return dd_fail(h, DD_ERR_ARG, "dd_louvain_level0_weighted: weights must be positive");
        two_m_q += wq[e];
    }
    // colour classes of the nodes that have neighbours
    std::vector<int32_t> cnt(kColoursW + 1, 0), nodes;
    for (int64_t i = 0; i < n; i++)
        if (off[i + 1] > off[i]) cnt[colour_of_w(seed, (int)i) + 1]++;
    for (int c = 0; c < kColoursW; c++) cnt[c + 1] += cnt[c];
    nodes.resize((size_t)std::max(cnt[kColoursW], 1));
    {
        std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
        for (int64_t i = 0; i < n; i++)
            if (off[i + 1] > off[i]) nodes[fill[colour_of_w(seed, (int)i)]++] = (int32_t)i;
    }
    DevBuf<int32_t> d_off, d_adj, d_comm, d_csize, d_desired, d_bucket, d_counters;
    DevBuf<long long> d_wq, d_k, d_tot;
    if (d_off.alloc(n + 1) || d_adj.alloc(nnz) || d_comm.alloc(n) || d_csize.alloc(n) || d_desired.alloc(n) ||
        d_bucket.alloc(nodes.size()) || d_counters.alloc(4) || d_wq.alloc(nnz) || d_k.alloc(n) || d_tot.alloc(n))
        return dd_fail(h, DD_ERR_NOMEM, "dd_louvain_level0_weighted: device buffers");
    DD_CUDA(h, cudaMemcpyAsync(d_off.p, off.data(), sizeof(int32_t) * (n + 1), cudaMemcpyHostToDevice, h->stream));
    if (nnz > 0) {
        DD_CUDA(h, cudaMemcpyAsync(d_adj.p, adj.data(), sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, h->stream));
        DD_CUDA(h, cudaMemcpyAsync(d_wq.p, wq.data(), sizeof(long long) * nnz, cudaMemcpyHostToDevice, h->stream));
    }
    DD_CUDA(h, cudaMemcpyAsync(d_bucket.p, nodes.data(), sizeof(int32_t) * nodes.size(), cudaMemcpyHostToDevice, h->stream));
    DD_CUDA(h, cudaMemsetAsync(d_desired.p, 0xff, sizeof(int32_t) * n, h->stream));
    const int ni = (int)n;
    DD_LAUNCH(h, "lvw_init", k_lvw_init, (unsigned)((n + 255) / 256), 256, 0, d_off.p, d_wq.p, ni, d_comm.p, d_k.p, d_tot.p,
              d_csize.p, d_counters.p);
    if (two_m_q > 0) {
        const double two_m = (double)two_m_q;
        for (int round = 0; round < kMaxRoundsW; round++) {
            for (int c = 0; c < kColoursW; c++) {
                const int b0 = cnt[c], b1 = cnt[c + 1];
                if (b1 == b0) continue;
                DD_LAUNCH(h, "lvw_propose", k_lvw_propose, (unsigned)((b1 - b0 + 7) / 8), 256, 0, d_off.p, d_adj.p, d_wq.p, d_comm.p,
                          d_k.p, d_tot.p, d_csize.p, d_bucket.p, b0, b1, gamma, two_m, (const long long *)nullptr, d_desired.p,
                          d_counters.p);
                DD_LAUNCH(h, "lvw_apply", k_lvw_apply, (unsigned)((b1 - b0 + 255) / 256), 256, 0, d_comm.p, d_k.p, d_tot.p, d_csize.p,
                          d_bucket.p, b0, b1, d_desired.p, d_counters.p);
            }
            DD_LAUNCH(h, "lvw_round_end", k_lvw_round_end, 1, 1, 0, d_counters.p, ni);
        }
    }
    int32_t counters[4] = {0, 0, 0, 0};
    DD_CUDA(h, cudaMemcpyAsync(comm_out, d_comm.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaMemcpyAsync(counters, d_counters.p, sizeof(int32_t) * 3, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (rounds_out) *rounds_out = counters[2];
    return DD_OK;
}

// ---- pipeline flavour (experimental, DD_PHENO_LEVEL0): the level on the device-built PhenoGraph graph -------------------
namespace {
// fixed-point weights, weighted degrees, 2m and the initial state in one pass over the device graph.  The rows are COMPACTED
// in place (stable): the entries that are edges come first, the pruned ones (weight 0) go to the tail with weight 0 -- deg[i]
// leading entries count.  The host side drops zero-weight entries and sorts the rows anyway.
__global__ void k_lvw_prepare(const int32_t *__restrict__ off, int32_t *__restrict__ adj, double *__restrict__ w, int n,
                              long long *__restrict__ wq, int32_t *__restrict__ deg, long long *__restrict__ k,
                              long long *__restrict__ tot, long long *__restrict__ two_m, int32_t *__restrict__ comm,
                              int32_t *__restrict__ csize, int32_t *__restrict__ counters) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) counters[0] = counters[1] = counters[2] = 0;
    if (i >= n) return;
    long long s = 0;
    const int b = off[i], e1 = off[i + 1];
    int p = b;
    for (int e = b; e < e1; e++) {
        const double we = w[e];
        const long long q = __double2ll_rn(__dmul_rn(we, 4294967296.0));  // round to nearest even, like the specification
        if (we != 0.0) {  // an edge (the specification drops exact zeros, whatever they would quantise to)
            const int a = adj[e];
            adj[p] = a;
            w[p] = we;
            wq[p] = q;
            s += q;
            p++;
        }
    }
    deg[i] = p - b;
    for (int e = p; e < e1; e++) {
        w[e] = 0.0;
        wq[e] = 0;
    }
    k[i] = s;
    tot[i] = s;
    comm[i] = i;
    csize[i] = 1;
    if (s) atomicAdd(reinterpret_cast<unsigned long long *>(two_m), (unsigned long long)s);
}
}  // namespace

// Graph: h->d_lv_off / d_lv_adj / d_lv_w as left by dd_dev_jaccard_graph (rows in any order, 0 = pruned).  Result:
// h->d_lv_comm (community = node id), which the fit loop already copies to the host slot.  Asynchronous on h->stream.
int dd_dev_louvain_level0_weighted(dd_handle *h, double gamma, uint64_t seed) {
    const int64_t n64 = h->emb_rows;
    if (!h->d_lv_off || !h->d_lv_adj || !h->d_lv_w || n64 > h->cap_lv_n)
        return dd_fail(h, DD_ERR_ARG, "weighted louvain level: build the graph first");
    const int n = (int)n64;
    if (h->cap_lv_nnz > h->cap_lvw_nnz || n > h->cap_lvw_n) {
        for (void *p : {(void *)h->d_lvw_wq, (void *)h->d_lvw_i64, (void *)h->d_lvw_i32})
            if (p) cudaFree(p);
        h->d_lvw_wq = h->d_lvw_i64 = nullptr;
        h->d_lvw_i32 = nullptr;
        h->cap_lvw_nnz = h->cap_lvw_n = 0;
        DD_CUDA(h, cudaMalloc(&h->d_lvw_wq, sizeof(long long) * (size_t)std::max<int64_t>(h->cap_lv_nnz, 1)));
        DD_CUDA(h, cudaMalloc(&h->d_lvw_i64, sizeof(long long) * (2 * (size_t)h->cap_lv_n + 1)));
        DD_CUDA(h, cudaMalloc(&h->d_lvw_i32, sizeof(int32_t) * (4 * (size_t)h->cap_lv_n + 16)));
        h->cap_lvw_nnz = h->cap_lv_nnz;
        h->cap_lvw_n = h->cap_lv_n;
        h->lvw_bucket_n = -1;
    }
    long long *k = h->d_lvw_i64, *tot = k + h->cap_lvw_n, *two_m = tot + h->cap_lvw_n;
    int32_t *csize = h->d_lvw_i32, *desired = csize + h->cap_lvw_n, *bucket = desired + h->cap_lvw_n,
            *deg = bucket + h->cap_lvw_n, *counters = deg + h->cap_lvw_n;
    if (h->lvw_bucket_n != n || h->lvw_bucket_seed != seed) {  // colour classes: a pure function of (n, seed)
        std::vector<int32_t> cnt(kColoursW + 1, 0), nodes((size_t)std::max(n, 1));
        for (int i = 0; i < n; i++) cnt[colour_of_w(seed, i) + 1]++;
        for (int c = 0; c < kColoursW; c++) cnt[c + 1] += cnt[c];
        h->lvw_colour_off.assign(cnt.begin(), cnt.end());
        std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
        for (int i = 0; i < n; i++) nodes[fill[colour_of_w(seed, i)]++] = i;
        DD_CUDA(h, cudaMemcpyAsync(bucket, nodes.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
        DD_CUDA(h, cudaStreamSynchronize(h->stream));  // `nodes` is a temporary (once per fit)
        h->lvw_bucket_n = n;
        h->lvw_bucket_seed = seed;
    }
    DD_CUDA(h, cudaMemsetAsync(two_m, 0, sizeof(long long), h->stream));
    DD_LAUNCH(h, "lvw_prepare", k_lvw_prepare, (unsigned)((n + 127) / 128), 128, 0, h->d_lv_off, h->d_lv_adj, h->d_lv_w, n,
              h->d_lvw_wq, deg, k, tot, two_m, h->d_lv_comm, csize, counters);
    auto issue_rounds = [&]() -> int {
        for (int round = 0; round < kMaxRoundsW; round++) {
            for (int c = 0; c < kColoursW; c++) {
                const int b0 = h->lvw_colour_off[c], b1 = h->lvw_colour_off[c + 1];
                if (b1 == b0) continue;
                DD_LAUNCH(h, "lvw_propose", k_lvw_propose_g, (unsigned)((b1 - b0 + kNodesPerCtaW - 1) / kNodesPerCtaW),
                          kPropWarpsW * 32, 0, h->d_lv_off, (const int32_t *)deg, h->d_lv_adj, h->d_lvw_wq, h->d_lv_comm, k, tot,
                          csize, bucket, b0, b1, gamma, (const long long *)two_m, desired, counters);
                DD_LAUNCH(h, "lvw_apply", k_lvw_apply, (unsigned)((b1 - b0 + 255) / 256), 256, 0, h->d_lv_comm, k, tot, csize,
                          bucket, b0, b1, desired, counters);
            }
            DD_LAUNCH(h, "lvw_round_end", k_lvw_round_end, 1, 1, 0, counters, n);
        }
        return DD_OK;
    };
    // The 32 x 17 steps are replayed from a CUDA graph (captured once per lane and problem size; the kernels of the rounds
    // after convergence return at once).  DD_LVW_NO_GRAPH=1: plain launches (A/B).
    static const bool no_graph = getenv("DD_LVW_NO_GRAPH") != nullptr;
    if (no_graph) return issue_rounds();
    const void *key[4] = {h->d_lv_off, h->d_lv_adj, h->d_lvw_wq, h->d_lv_comm};
    if (h->lvw_graph_exec == nullptr || h->lvw_graph_n != n || h->lvw_graph_gamma != gamma || h->lvw_graph_seed != seed ||
        !std::equal(key, key + 4, h->lvw_graph_key)) {
        if (h->lvw_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->lvw_graph_exec);
        h->lvw_graph_exec = nullptr;
        const bool timing = h->timing;
        const int64_t launches_before = h->launches;
        h->timing = false;  // no event records inside the capture
        DD_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = issue_rounds();
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        h->timing = timing;
        h->lvw_graph_launches = h->launches - launches_before;
        h->launches = launches_before;
        if (rc != DD_OK) return rc;
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("weighted louvain graph capture: ") + cudaGetErrorString(e));
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("weighted louvain graph instantiate: ") + cudaGetErrorString(e));
        h->lvw_graph_exec = exec;
        h->lvw_graph_n = n;
        h->lvw_graph_gamma = gamma;
        h->lvw_graph_seed = seed;
        std::copy(key, key + 4, h->lvw_graph_key);
    }
    dd_launch_begin(h);
    {
        cudaError_t e = cudaGraphLaunch((cudaGraphExec_t)h->lvw_graph_exec, h->stream);
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("weighted louvain graph launch: ") + cudaGetErrorString(e));
    }
    DD_TRY(dd_launch_end(h, "lvw_rounds_graph"));
    h->launches += h->lvw_graph_launches - 1;
    return DD_OK;
}
