// louvain_gpu.cu -- the first (and by far largest) Louvain level of the clustering call on the GPU.
//
// The kNN pipeline optimises the first level by synchronous coloured rounds (specification:
// oracle/louvain_ref.py:level0_parallel; host twin: louvain.cpp:level0_parallel_host): 8 colour classes,
// all nodes of a class decide simultaneously from the frozen state, moves are applied at once, up to 32
// rounds.  Everything the decision needs is integer-valued (unweighted graph), so float64 atomics are exact
// and the result does not depend on scheduling.  This file also builds the symmetrised kNN pattern
// (i ~ j iff j in kNN(i) or i in kNN(j)) as CSR on the device; the host only aggregates the ~10^2
// communities that come out and runs the (tiny) upper levels.
#include "dd_internal.h"

#include <cooperative_groups.h>
#include <math_constants.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace {

constexpr int kColours = 8;
constexpr int kMaxRounds = 32;

__host__ __device__ inline int colour_of(uint64_t seed, int i) {
    uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (int)(z % kColours);
}

// Programmatic dependent launch (PDL): a kernel of the round sequence lets its successor's CTAs become resident as soon as
// it is itself unblocked (launch_dependents right after its own wait) and the successor fetches everything that is STATIC during the level (colour bucket, CSR
// offsets, adjacency) before it waits for its predecessor to complete and flush (pdl_wait) -- the launch latency and three
// of the five dependent loads of a step leave the critical path.  Only data written inside the level (communities, totals,
// sizes, desired moves, counters) may be touched after pdl_wait, and nothing written inside the level before it.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
cudaError_t launch_step(void (*kernel)(KArgs...), unsigned grid, unsigned block, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ bool in_out_list(const int32_t *__restrict__ knn, int k, int j, int i) {
    for (int c = 0; c < k; c++)
        if (knn[(int64_t)j * k + c] == i) return true;
    return false;
}

// deg[j] = out-neighbours of j + in-neighbours of j that are not among its out-neighbours
__global__ void k_graph_count(const int32_t *__restrict__ knn, int n, int k, int32_t *__restrict__ deg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int out = 0;
    for (int c = 0; c < k; c++) {
        const int j = knn[(int64_t)i * k + c];
        if (j == i || j < 0) continue;
        out++;
        if (!in_out_list(knn, k, j, i)) atomicAdd(deg + j, 1);
    }
    atomicAdd(deg + i, out);
}

// Exclusive scan of deg -> off[0..n] in three small launches (the single-CTA version it replaces took 0.7-1.3 ms at 125 k
// nodes, on the clustering stream's critical path): chunks of kScanChunk elements scanned by one CTA each, the chunk totals
// scanned by one CTA (which also resets the level's counters), the chunk bases added back.
constexpr int kScanChunk = 4096;  // 1024 threads x 4 elements
__global__ void __launch_bounds__(1024) k_graph_scan_local(const int32_t *__restrict__ deg, int n, int32_t *__restrict__ off,
                                                           int32_t *__restrict__ part) {
    __shared__ int s_warp[32];
    const int t = threadIdx.x, lane = t & 31, wl = t >> 5;
    const int base = blockIdx.x * kScanChunk + t * 4;
    int v[4];
#pragma unroll
    for (int e = 0; e < 4; e++) v[e] = base + e < n ? deg[base + e] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[wl] = incl;
    __syncthreads();
    if (wl == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int run = incl - mine + (wl > 0 ? s_warp[wl - 1] : 0);
#pragma unroll
    for (int e = 0; e < 4; e++) {
        if (base + e < n) off[base + e] = run;
        run += v[e];
    }
    if (t == 1023) part[blockIdx.x] = run;  // chunk total
}

__global__ void __launch_bounds__(1024) k_graph_scan_parts(int32_t *__restrict__ part, int n_parts, int n,
                                                           int32_t *__restrict__ off, int32_t *__restrict__ counters) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int t = threadIdx.x, lane = t & 31, wl = t >> 5;
    if (t == 0) s_carry = 0;
    __syncthreads();
    for (int p0 = 0; p0 < n_parts; p0 += 1024) {
        const int mine = p0 + t < n_parts ? part[p0 + t] : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[wl] = incl;
        __syncthreads();
        if (wl == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        if (p0 + t < n_parts) part[p0 + t] = carry + incl - mine + (wl > 0 ? s_warp[wl - 1] : 0);  // exclusive chunk base
        __syncthreads();
        if (t == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (t == 0) {
        off[n] = s_carry;
        counters[0] = 0;  // moved in this round
        counters[1] = 0;  // done flag
        counters[2] = 0;  // rounds executed
        counters[3] = 0;  // grid barrier arrivals (cooperative kernel)
    }
}

__global__ void __launch_bounds__(1024) k_graph_scan_add(const int32_t *__restrict__ part, int n, int32_t *__restrict__ off) {
    const int base = blockIdx.x * kScanChunk + threadIdx.x * 4;
    const int add = part[blockIdx.x];
#pragma unroll
    for (int e = 0; e < 4; e++)
        if (base + e < n) off[base + e] += add;
}

// adjacency rows: the node's own out-neighbours first, then the in-neighbours that are not out-neighbours
// (claimed with an atomic cursor: their order inside the row is irrelevant to the algorithm)
__global__ void k_graph_fill(const int32_t *__restrict__ knn, int n, int k, const int32_t *__restrict__ off,
                             int32_t *__restrict__ cursor, int32_t *__restrict__ adj, int32_t *__restrict__ comm,
                             double *__restrict__ tot, int32_t *__restrict__ csize) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int out = 0;
    for (int c = 0; c < k; c++) {
        const int j = knn[(int64_t)i * k + c];
        if (j == i || j < 0) continue;
        adj[off[i] + out] = j;
        out++;
    }
    for (int c = 0; c < k; c++) {
        const int j = knn[(int64_t)i * k + c];
        if (j == i || j < 0) continue;
        if (!in_out_list(knn, k, j, i)) {
            int outj = 0;
            for (int cc = 0; cc < k; cc++) {
                const int jj = knn[(int64_t)j * k + cc];
                outj += (jj != j && jj >= 0);
            }
            const int pos = atomicAdd(cursor + j, 1);
            adj[off[j] + outj + pos] = i;
        }
    }
    comm[i] = i;
    tot[i] = (double)(off[i + 1] - off[i]);
    csize[i] = 1;
}

// per-warp open-addressing table (community -> number of neighbours in it) for nodes with more than 32
// neighbours; kTable slots, linear probing.  Degrees above kTableMaxDeg fall back to a quadratic scan.
// 256 slots = 2 KB per warp, 16 KB per propose CTA: three of them fit next to a kNN CTA's 173 KB of shared memory.
constexpr int kTable = 256, kTableMaxDeg = 192, kTableShift = 24;  // hash = top 8 bits
__device__ __forceinline__ int propose_one(const int32_t *__restrict__ off, const int32_t *__restrict__ adj,
                                           const int32_t *comm, const double *tot, const int32_t *csize, int i,
                                           double gamma, double two_m, int32_t *tkey, int32_t *tcnt, int lane) {
    const int s = off[i], d = off[i + 1] - s;
    if (d <= 0) return -1;
    const int ci = __ldcg(comm + i);
    const double ki = (double)d;
    const double gk = __dmul_rn(gamma, ki);
    double best_gain = 0.0;
    int best = 0x7fffffff;
    int w_stay = 0;
    if (d <= 32) {
        const int c = lane < d ? __ldcg(comm + adj[s + lane]) : -1 - lane;  // unique sentinels never match
        const int wc = __popc(__match_any_sync(0xffffffffu, c));
        w_stay = __popc(__ballot_sync(0xffffffffu, c == ci));
        if (lane < d && c != ci) {
            best = c;
            best_gain = __dsub_rn((double)wc, __ddiv_rn(__dmul_rn(gk, __ldcg(tot + c)), two_m));
        }
    } else if (d <= kTableMaxDeg) {
        __syncwarp();
        for (int t = lane; t < kTable; t += 32) {
            tkey[t] = -1;
            tcnt[t] = 0;
        }
        __syncwarp();
        for (int f = lane; f < d; f += 32) {
            const int c = __ldcg(comm + adj[s + f]);
            unsigned slot = ((unsigned)c * 2654435761u) >> kTableShift;
            for (;;) {
                const int prev = atomicCAS(tkey + slot, -1, c);
                if (prev == -1 || prev == c) break;
                slot = (slot + 1) & (kTable - 1);
            }
            atomicAdd(tcnt + slot, 1);
        }
        __syncwarp();
        for (int t = lane; t < kTable; t += 32) {
            const int c = tkey[t];
            if (c < 0) continue;
            const int wc = tcnt[t];
            if (c == ci) {
                w_stay = wc;
                continue;
            }
            const double gn = __dsub_rn((double)wc, __ddiv_rn(__dmul_rn(gk, __ldcg(tot + c)), two_m));
            if (best == 0x7fffffff || gn > best_gain || (gn == best_gain && c < best)) {
                best = c;
                best_gain = gn;
            }
        }
        w_stay = __reduce_add_sync(0xffffffffu, w_stay);  // at most one lane holds the own-community count
    } else {
        for (int f = lane; f < d; f += 32) w_stay += (__ldcg(comm + adj[s + f]) == ci);
        w_stay = __reduce_add_sync(0xffffffffu, w_stay);
        for (int e = lane; e < d; e += 32) {
            const int c = __ldcg(comm + adj[s + e]);
            if (c == ci) continue;
            int wc = 0;
            for (int f = 0; f < d; f++) wc += (__ldcg(comm + adj[s + f]) == c);
            const double gn = __dsub_rn((double)wc, __ddiv_rn(__dmul_rn(gk, __ldcg(tot + c)), two_m));
            if (best == 0x7fffffff || gn > best_gain || (gn == best_gain && c < best)) {
                best = c;
                best_gain = gn;
            }
        }
    }
    const double gain_stay = __dsub_rn((double)w_stay, __ddiv_rn(__dmul_rn(gk, __dsub_rn(__ldcg(tot + ci), ki)), two_m));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double og = __shfl_xor_sync(0xffffffffu, best_gain, o);
        const int ob = __shfl_xor_sync(0xffffffffu, best, o);
        if (ob != 0x7fffffff && (best == 0x7fffffff || og > best_gain || (og == best_gain && ob < best))) {
            best = ob;
            best_gain = og;
        }
    }
    if (best != 0x7fffffff && best_gain > gain_stay &&
        !(__ldcg(csize + ci) == 1 && __ldcg(csize + best) == 1 && best > ci))
        return best;
    return -1;
}

// one warp per node of the current colour (multi-launch / CUDA-graph variant of the rounds)
__global__ void __launch_bounds__(256) k_lv_propose(const int32_t *__restrict__ off, const int32_t *__restrict__ adj,
                                                    const int32_t *__restrict__ comm, const double *__restrict__ tot,
                                                    const int32_t *__restrict__ csize,
                                                    const int32_t *__restrict__ bucket, int b0, int b1, int n,
                                                    double gamma, int32_t *__restrict__ desired,
                                                    const int32_t *__restrict__ counters) {
    __shared__ int32_t s_tab[8 * 2 * kTable];
    pdl_wait();
    if (counters[1]) return;  // converged in an earlier round
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int w = blockIdx.x * (blockDim.x >> 5) + wl;
    if (b0 + w >= b1) return;
    const int i = bucket[b0 + w];
    const double two_m = (double)off[n];
    const int res = propose_one(off, adj, comm, tot, csize, i, gamma, two_m, s_tab + wl * 2 * kTable,
                                s_tab + wl * 2 * kTable + kTable, lane);
    if (lane == 0) desired[i] = res;
}

// EIGHT LANES per node of the current colour, four nodes per warp (the default).  The warp-per-node kernel above spends 32
// threads on the ~15 neighbours of a kNN-graph node: a colour class of 15.6 k nodes (c3) is 1950 CTAs, and next to the main
// stream's kernels (a 173 KB kNN CTA leaves room for two of them per SM) that is ~7 waves per step, 500 steps per level --
// the level took 11.5 ms per iteration overlapped against 3 ms alone and was the critical path of the fit loop.  (One
// THREAD per node was measured too: a single partial wave, but the serial per-thread work made a step slower, 15 ms.)
// Here a lane holds up to four neighbour communities in registers, multiplicities come from 32 group-wide shuffles, the
// community totals are fetched with independent loads; the class is 488 CTAs = under two waves of the same short latency
// chain.  Nodes with more than 32 neighbours (hubs of the in-degree tail) are handed to the whole warp afterwards
// (propose_one).  Same arithmetic and tie-breaking as propose_one: identical results.
constexpr int kGroupLanes = 8, kPerLane = 4, kPropWarps = 4, kNodesPerCta = kPropWarps * (32 / kGroupLanes);
__global__ void __launch_bounds__(kPropWarps * 32, 10) k_lv_propose_g(const int32_t *__restrict__ off, const int32_t *__restrict__ adj,
                                                      const int32_t *__restrict__ comm, const double *__restrict__ tot,
                                                      const int32_t *__restrict__ csize,
                                                      const int32_t *__restrict__ bucket, int b0, int b1, int n,
                                                      double gamma, int32_t *__restrict__ desired,
                                                      const int32_t *__restrict__ counters) {
    __shared__ int32_t s_tab[kPropWarps * 2 * kTable];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int gl = lane & (kGroupLanes - 1), grp = lane / kGroupLanes;
    const unsigned gmask = ((1u << kGroupLanes) - 1u) << (grp * kGroupLanes);
    const int t = (blockIdx.x * kPropWarps + wl) * (32 / kGroupLanes) + grp;
    // ---- static during the level: fetched while the previous step is still running
    const double two_m = (double)off[n];
    int i = -1, s = 0, d = 0;
    if (b0 + t < b1) {
        i = bucket[b0 + t];
        s = off[i];
        d = off[i + 1] - s;
    }
    const bool big = d > kGroupLanes * kPerLane;
    int a[kPerLane];
#pragma unroll
    for (int j = 0; j < kPerLane; j++) {
        const int e = j * kGroupLanes + gl;
        a[j] = (!big && e < d) ? adj[s + e] : -1;
    }
    pdl_wait();
    // only now: the successor may become resident (and prefetch ITS static data) while this step works -- triggering before
    // the wait would let the whole chain of future steps pile up on the SMs as waiting CTAs and starve the main stream
    pdl_launch_dependents();
    if (__ldcg(counters + 1)) return;  // converged in an earlier round
    if (i >= 0 && !big) {  // uniform inside the group
        int res = -1;
        if (d > 0) {
            const int ci = __ldcg(comm + i);
            int c[kPerLane];
#pragma unroll
            for (int j = 0; j < kPerLane; j++) c[j] = a[j] >= 0 ? __ldcg(comm + a[j]) : -1 - (j * kGroupLanes + gl);  // unique sentinels
            double tt[kPerLane];
#pragma unroll
            for (int j = 0; j < kPerLane; j++) tt[j] = (c[j] >= 0 && c[j] != ci) ? __ldcg(tot + c[j]) : 0.0;
            const double tot_ci = __ldcg(tot + ci);
            const int cs_ci = __ldcg(csize + ci);
            int wc[kPerLane] = {0, 0, 0, 0};
            int w_stay = 0;
#pragma unroll
            for (int src = 0; src < kGroupLanes; src++) {
#pragma unroll
                for (int jj = 0; jj < kPerLane; jj++) {
                    const int o = __shfl_sync(gmask, c[jj], grp * kGroupLanes + src);
#pragma unroll
                    for (int j = 0; j < kPerLane; j++) wc[j] += (o == c[j]);
                    w_stay += (o == ci);
                }
            }
            const double ki = (double)d;
            const double gk = __dmul_rn(gamma, ki);
            double best_gain = 0.0;
            int best = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < kPerLane; j++) {
                if (c[j] < 0 || c[j] == ci) continue;
                const double gn = __dsub_rn((double)wc[j], __ddiv_rn(__dmul_rn(gk, tt[j]), two_m));
                if (best == 0x7fffffff || gn > best_gain || (gn == best_gain && c[j] < best)) {
                    best = c[j];
                    best_gain = gn;
                }
            }
#pragma unroll
            for (int o = kGroupLanes / 2; o > 0; o >>= 1) {
                const double og = __shfl_xor_sync(gmask, best_gain, o);
                const int ob = __shfl_xor_sync(gmask, best, o);
                if (ob != 0x7fffffff && (best == 0x7fffffff || og > best_gain || (og == best_gain && ob < best))) {
                    best = ob;
                    best_gain = og;
                }
            }
            const double gain_stay = __dsub_rn((double)w_stay, __ddiv_rn(__dmul_rn(gk, __dsub_rn(tot_ci, ki)), two_m));
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
        const int res = propose_one(off, adj, comm, tot, csize, node, gamma, two_m, s_tab + wl * 2 * kTable,
                                    s_tab + wl * 2 * kTable + kTable, lane);
        if (lane == 0) desired[node] = res;
    }
}

__global__ void k_lv_apply(const int32_t *__restrict__ off, int32_t *__restrict__ comm, double *__restrict__ tot,
                           int32_t *__restrict__ csize, const int32_t *__restrict__ bucket, int b0, int b1,
                           const int32_t *__restrict__ desired, int32_t *__restrict__ counters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    int i = -1;
    double ki = 0.0;
    if (b0 + t < b1) {  // static during the level
        i = bucket[b0 + t];
        ki = (double)(off[i + 1] - off[i]);
    }
    pdl_wait();
    pdl_launch_dependents();
    if (i < 0 || __ldcg(counters + 1)) return;
    const int b = __ldcg(desired + i);
    if (b < 0) return;
    const int ci = __ldcg(comm + i);
    comm[i] = b;
    atomicAdd(tot + ci, -ki);
    atomicAdd(tot + b, ki);
    atomicSub(csize + ci, 1);
    atomicAdd(csize + b, 1);
    atomicAdd(counters, 1);
}

// a round that moved at most n / 512 nodes ends the level (the stragglers oscillate or trickle; the levels
// above merge whole communities anyway)
// loop != 0: the round is the body of a conditional WHILE node of the CUDA graph (cond = its handle): the loop ends with the
// level instead of replaying the remaining rounds as no-op launches
__global__ void k_lv_round_end(int32_t *__restrict__ counters, int n, int loop, cudaGraphConditionalHandle cond, int max_rounds) {
    pdl_wait();
    pdl_launch_dependents();
    if (!counters[1]) {
        counters[2]++;
        if (counters[0] <= (n >> 9)) counters[1] = 1;
        counters[0] = 0;
    }
    if (loop && (counters[1] || counters[2] >= max_rounds)) cudaGraphSetConditional(cond, 0);
}

// All rounds in ONE cooperative launch: the 500+ sub-round steps are separated by grid-wide barriers instead of
// kernel boundaries (a launch per step made the level launch-bound).  Mutable state is read with ld.global.cg
// (L1 is not coherent across SMs inside a kernel).
struct ColourOffsets {
    int v[kColours + 1];
};

// grid-wide barrier for a co-resident grid (cooperative launch): one arrival per CTA on a monotonically
// increasing counter; ~2 us, where cooperative_groups' grid.sync() measured ~15 us on 592 CTAs.
__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int n_blocks, unsigned int &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += n_blocks;
        __threadfence();
        atomicAdd(bar, 1u);
        while (*(volatile unsigned int *)bar < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// kRoundThreads = 1024 (64 KB of tables): one CTA per SM on an otherwise idle GPU.  256 (16 KB): a light resident
// grid that fits NEXT TO the main stream's kernels on the same SMs (the kNN CTA leaves 54 KB / 27 k registers free).
template <int kRoundThreads>
constexpr size_t round_smem() { return sizeof(int32_t) * (kRoundThreads / 32) * 2 * kTable; }
template <int kRoundThreads>
__global__ void __launch_bounds__(kRoundThreads, 1) k_lv_rounds(const int32_t *__restrict__ off,
                                                               const int32_t *__restrict__ adj, int32_t *comm,
                                                               double *tot, int32_t *csize,
                                                               const int32_t *__restrict__ bucket, ColourOffsets co,
                                                               int n, double gamma, int32_t *desired,
                                                               int32_t *counters) {
    extern __shared__ int32_t s_tab[];  // per warp: kTable keys + kTable counts
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    constexpr int WPB = kRoundThreads / 32;
    int32_t *tkey = s_tab + (size_t)wl * 2 * kTable, *tcnt = tkey + kTable;
    const int n_warps = gridDim.x * WPB, gw = blockIdx.x * WPB + wl;
    const int n_thr = gridDim.x * kRoundThreads, gt = blockIdx.x * kRoundThreads + threadIdx.x;
    const double two_m = (double)off[n];
    unsigned int *bar = reinterpret_cast<unsigned int *>(counters + 3);
    unsigned int target = 0;
    for (int round = 0; round < kMaxRounds; round++) {
        for (int c = 0; c < kColours; c++) {
            const int b0 = co.v[c], b1 = co.v[c + 1];
            for (int t = b0 + gw; t < b1; t += n_warps) {
                const int i = bucket[t];
                const int res = propose_one(off, adj, comm, tot, csize, i, gamma, two_m, tkey, tcnt, lane);
                if (lane == 0) desired[i] = res;
            }
            grid_barrier(bar, gridDim.x, target);
            int moved = 0;
            for (int t = b0 + gt; t < b1; t += n_thr) {
                const int i = bucket[t];
                const int b = __ldcg(desired + i);
                if (b < 0) continue;
                const int ci = __ldcg(comm + i);
                const double ki = (double)(off[i + 1] - off[i]);
                comm[i] = b;
                atomicAdd(tot + ci, -ki);
                atomicAdd(tot + b, ki);
                atomicSub(csize + ci, 1);
                atomicAdd(csize + b, 1);
                moved++;
            }
            if (moved) atomicAdd(counters, moved);
            grid_barrier(bar, gridDim.x, target);
        }
        const int total = __ldcg(counters);
        grid_barrier(bar, gridDim.x, target);  // everyone has read the round's move count before it is reset
        if (gt == 0) {
            counters[0] = 0;
            counters[2] = round + 1;
        }
        if (total <= (n >> 9)) break;  // uniform across the grid
    }
}

}  // namespace

// Build the symmetric kNN pattern from h->d_knn_idx (n x k, self in column 0) and run the parallel first
// Louvain level.  Results (device): h->d_lv_off (n + 1), h->d_lv_adj (off[n] entries), h->d_lv_comm (n).
// Fully asynchronous on h->stream (no host read-back: the colour classes depend only on (n, seed) and are
// bucketed on the host once).
namespace {
struct LvBuffers {
    int n;
    int32_t *deg, *cursor, *csize, *desired, *bucket, *counters;
};

// (re)allocate the graph / state buffers for n nodes of out-degree <= k - 1 and build the symmetric pattern CSR
int lv_build_graph(dd_handle *h, int32_t k, LvBuffers &b) {
    const int64_t n64 = h->emb_rows;
    if (n64 >= (1ll << 31) / (2 * k)) return dd_fail(h, DD_ERR_UNSUPPORTED, "louvain: graph too large for int32 offsets");
    const int n = (int)n64;
    const int64_t max_nnz = (int64_t)n * 2 * (k - 1);
    if (n > h->cap_lv_n || max_nnz > h->cap_lv_nnz) {
        for (void *p : {(void *)h->d_lv_off, (void *)h->d_lv_adj, (void *)h->d_lv_comm, (void *)h->d_lv_tot, (void *)h->d_lv_i32})
            if (p) cudaFree(p);
        h->d_lv_off = h->d_lv_adj = h->d_lv_comm = h->d_lv_i32 = nullptr;
        h->d_lv_tot = nullptr;
        h->cap_lv_n = 0; h->cap_lv_nnz = 0;
        DD_CUDA(h, cudaMalloc(&h->d_lv_off, sizeof(int32_t) * (n + 1)));
        DD_CUDA(h, cudaMalloc(&h->d_lv_adj, sizeof(int32_t) * std::max<int64_t>(max_nnz, 1)));
        DD_CUDA(h, cudaMalloc(&h->d_lv_comm, sizeof(int32_t) * n));
        DD_CUDA(h, cudaMalloc(&h->d_lv_tot, sizeof(double) * n));
        // deg | cursor | csize | desired | bucket (n each) + counters (8) + scan chunk totals
        DD_CUDA(h, cudaMalloc(&h->d_lv_i32, sizeof(int32_t) * (5 * (size_t)n + 16 + (size_t)n / kScanChunk + 2)));
        h->cap_lv_n = n; h->cap_lv_nnz = max_nnz;
        h->lv_bucket_n = -1;
        h->lv_graph_n = -1;  // the captured launches hold the old pointers
    }
    b.n = n;
    b.deg = h->d_lv_i32; b.cursor = b.deg + h->cap_lv_n; b.csize = b.cursor + h->cap_lv_n;
    b.desired = b.csize + h->cap_lv_n; b.bucket = b.desired + h->cap_lv_n; b.counters = b.bucket + h->cap_lv_n;
    DD_CUDA(h, cudaMemsetAsync(b.deg, 0, sizeof(int32_t) * 2 * (size_t)h->cap_lv_n, h->stream));  // deg + cursor
    const unsigned nb = (unsigned)((n + 255) / 256);
    DD_LAUNCH(h, "lv_graph_count", k_graph_count, nb, 256, 0, h->d_knn_idx, n, (int)k, b.deg);
    const int n_parts = (n + kScanChunk - 1) / kScanChunk;
    int32_t *part = b.counters + 8;  // chunk totals / bases live behind the counters
    DD_LAUNCH(h, "lv_graph_scan", k_graph_scan_local, n_parts, 1024, 0, b.deg, n, h->d_lv_off, part);
    DD_LAUNCH(h, "lv_graph_scan", k_graph_scan_parts, 1, 1024, 0, part, n_parts, n, h->d_lv_off, b.counters);
    DD_LAUNCH(h, "lv_graph_scan", k_graph_scan_add, n_parts, 1024, 0, part, n, h->d_lv_off);
    DD_LAUNCH(h, "lv_graph_fill", k_graph_fill, nb, 256, 0, h->d_knn_idx, n, (int)k, h->d_lv_off, b.cursor, h->d_lv_adj,
              h->d_lv_comm, h->d_lv_tot, b.csize);
    return DD_OK;
}

// PhenoGraph's edge weights (phenograph/core.py jaccard_kernel; doubletdetection.py:320) on the symmetric pattern:
// for the entry (i, j):  s = |N(i) & N(j)|,  w = s / (2 kk - s)  with N(x) the kk = k - 1 neighbours of x (self, in
// column 0 of the kNN lists, excluded).  prune: mutual edges carry w_ij * w_ji = w^2, the others 0 (the host drops
// them); otherwise (w_ij + w_ji) / 2 = w for mutual edges, w / 2 for one-directional ones.  One warp per node, the
// node's list in registers (lane c holds neighbour c), every other list compared by shuffles.
__global__ void __launch_bounds__(256) k_jaccard_weights(const int32_t *__restrict__ knn, int n, int k,
                                                         const int32_t *__restrict__ off,
                                                         const int32_t *__restrict__ adj, int prune,
                                                         double *__restrict__ w_out) {
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const int kk = k - 1;
    const int mine = lane < kk ? knn[(int64_t)i * k + 1 + lane] : -1;
    const int b = off[i], e = off[i + 1];
    for (int p = b; p < e; p++) {
        const int j = adj[p];
        const int theirs = lane < kk ? knn[(int64_t)j * k + 1 + lane] : -2;
        bool shared = false;
        for (int t = 0; t < kk; t++) shared |= theirs == __shfl_sync(0xffffffffu, mine, t);
        const int s = __popc(__ballot_sync(0xffffffffu, shared && lane < kk && theirs >= 0));
        const bool ij = __any_sync(0xffffffffu, mine == j);   // j in N(i)
        const bool ji = __any_sync(0xffffffffu, theirs == i);  // i in N(j)
        if (lane == 0) {
            const double w = (double)s / (double)(2 * kk - s);
            double out;
            if (prune)
                out = (ij && ji) ? w * w : 0.0;
            else
                out = (ij && ji) ? w : w * 0.5;
            w_out[p] = out;
        }
    }
}
// umap's smooth_knn_dist + compute_membership_strengths (sc.pp.neighbors(method="umap"), doubletdetection.py:331-336;
// local_connectivity = 1, bandwidth = 1) for the Leiden branch: one thread per cell.  rho = the nearest positive distance,
// sigma by bisection in float64 so that the memberships of the k - 1 neighbours sum to log2(k), directed weight
// exp(-(d - rho) / sigma) stored as float32.  Same operations in the same order as the host twin (leiden.cpp:umap_weights,
// specification oracle/upstream.py): float32 differences, float64 quotient / exp / sequential sum; explicit _rn
// intrinsics keep the compiler from contracting anything.  (The host's global mean distance only floors sigma of rows
// whose distances are all zero, and those rows' weights are 1 whatever sigma is: it is not needed here.)
__global__ void __launch_bounds__(128) k_umap_rows(const int32_t *__restrict__ knn, const float *__restrict__ dist, int n,
                                                   int k, double target, float *__restrict__ wdir) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *di = dist + (int64_t)i * k;
    float rho = 0.f;
    for (int c = 0; c < k; c++)
        if (di[c] > 0.f) {
            rho = di[c];
            break;
        }
    double row_sum = 0.0;
    for (int c = 0; c < k; c++) row_sum = __dadd_rn(row_sum, (double)di[c]);
    double lo = 0.0, hi = CUDART_INF, mid = 1.0;
    for (int it = 0; it < 64; it++) {
        double psum = 0.0;
        for (int c = 1; c < k; c++) {
            const double d = (double)__fsub_rn(di[c], rho);
            psum = __dadd_rn(psum, d > 0.0 ? exp(-__ddiv_rn(d, mid)) : 1.0);
        }
        if (fabs(__dsub_rn(psum, target)) < 1e-5) break;
        if (psum > target) {
            hi = mid;
            mid = __ddiv_rn(__dadd_rn(lo, hi), 2.0);
        } else {
            lo = mid;
            if (hi == CUDART_INF)
                mid = __dmul_rn(mid, 2.0);
            else
                mid = __ddiv_rn(__dadd_rn(lo, hi), 2.0);
        }
    }
    float sigma = __double2float_rn(mid);
    const double floor_ = rho > 0.f ? __dmul_rn(1e-3, __ddiv_rn(row_sum, (double)k)) : 0.0;
    if ((double)sigma < floor_) sigma = __double2float_rn(floor_);
    for (int c = 0; c < k; c++) {
        const int j = knn[(int64_t)i * k + c];
        const float diff = __fsub_rn(di[c], rho);
        float v;
        if (j == i || j < 0)
            v = 0.f;
        else if (diff <= 0.f || sigma == 0.f)
            v = 1.f;
        else
            v = __double2float_rn(exp(-__ddiv_rn((double)diff, (double)sigma)));
        wdir[(int64_t)i * k + c] = v;
    }
}

// umap's fuzzy union on the symmetric pattern: entry (i, j) carries a + b - a b in float32, a = w(i -> j), b = w(j -> i)
// (0 where the list does not hold the other cell).  One warp per cell, lanes over its entries.  Entries whose union is 0
// (underflowed memberships) stay in the pattern with weight 0; the host drops them.
__global__ void __launch_bounds__(256) k_umap_union(const int32_t *__restrict__ knn, const float *__restrict__ wdir, int n, int k,
                                                    const int32_t *__restrict__ off, const int32_t *__restrict__ adj,
                                                    double *__restrict__ w_out) {
    const int lane = threadIdx.x & 31;
    const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    for (int p = off[i] + lane; p < off[i + 1]; p += 32) {
        const int j = adj[p];
        float a = 0.f, b = 0.f;
        for (int c = 0; c < k; c++) {
            if (knn[(int64_t)i * k + c] == j) a = wdir[(int64_t)i * k + c];
            if (knn[(int64_t)j * k + c] == i) b = wdir[(int64_t)j * k + c];
        }
        w_out[p] = (double)__fsub_rn(__fadd_rn(a, b), __fmul_rn(a, b));
    }
}
}  // namespace

// Symmetric pattern of the kNN lists (k columns, self in column 0) + umap's connectivities (Leiden branch):
// h->d_lv_off / d_lv_adj / d_lv_w on the device, from h->d_knn_idx and h->d_knn_dist (asynchronous on h->stream).
int dd_dev_umap_graph(dd_handle *h, int32_t k) {
    if (k < 2) return dd_fail(h, DD_ERR_ARG, "umap graph: k < 2");
    if (!h->d_knn_dist) return dd_fail(h, DD_ERR_ARG, "umap graph: no kNN distances on the device");
    LvBuffers b;
    DD_TRY(lv_build_graph(h, k, b));
    DD_TRY(dd_reserve(h, &h->d_lv_w, &h->cap_lv_w, h->cap_lv_nnz));
    DD_TRY(dd_reserve(h, &h->d_umap_w, &h->cap_umap_w, (int64_t)b.n * k));
    DD_LAUNCH(h, "umap_rows", k_umap_rows, (unsigned)((b.n + 127) / 128), 128, 0, h->d_knn_idx, h->d_knn_dist, b.n, (int)k,
              std::log2((double)k), h->d_umap_w);
    DD_LAUNCH(h, "umap_union", k_umap_union, (unsigned)(((int64_t)b.n * 32 + 255) / 256), 256, 0, h->d_knn_idx, h->d_umap_w, b.n,
              (int)k, h->d_lv_off, h->d_lv_adj, h->d_lv_w);
    return DD_OK;
}

// Symmetric pattern of the kNN lists (k columns, self in column 0) + Jaccard weights: h->d_lv_off / d_lv_adj /
// d_lv_w on the device (asynchronous on h->stream).
int dd_dev_jaccard_graph(dd_handle *h, int32_t k, int prune) {
    if (k < 2 || k > 32) return dd_fail(h, DD_ERR_UNSUPPORTED, "jaccard graph: k + 1 must be in [2, 32]");
    LvBuffers b;
    DD_TRY(lv_build_graph(h, k, b));
    DD_TRY(dd_reserve(h, &h->d_lv_w, &h->cap_lv_w, h->cap_lv_nnz));
    const unsigned nb = (unsigned)(((int64_t)b.n * 32 + 255) / 256);
    DD_LAUNCH(h, "jaccard_weights", k_jaccard_weights, nb, 256, 0, h->d_knn_idx, b.n, (int)k, h->d_lv_off, h->d_lv_adj, prune,
              h->d_lv_w);
    return DD_OK;
}

int dd_dev_louvain_level0(dd_handle *h, int32_t k, double gamma, uint64_t seed) {
    LvBuffers lb;
    {   // colour classes: a pure function of (n, seed); the bucket array lives in the (possibly re-allocated) state block
        const int64_t n64 = h->emb_rows;
        if (n64 >= (1ll << 31) / (2 * k)) return dd_fail(h, DD_ERR_UNSUPPORTED, "louvain: graph too large for int32 offsets");
    }
    DD_TRY(lv_build_graph(h, k, lb));
    if (h->ev_after_graph_build) DD_CUDA(h, cudaEventRecord(h->ev_after_graph_build, h->stream));  // the kNN lists are free again
    const int n = lb.n;
    int32_t *csize = lb.csize, *desired = lb.desired, *bucket = lb.bucket, *counters = lb.counters;
    if (h->lv_bucket_n != n || h->lv_bucket_seed != seed) {
        std::vector<int32_t> cnt(kColours + 1, 0), nodes(n);
        for (int i = 0; i < n; i++) cnt[colour_of(seed, i) + 1]++;
        for (int c = 0; c < kColours; c++) cnt[c + 1] += cnt[c];
        h->lv_colour_off.assign(cnt.begin(), cnt.end());
        std::vector<int32_t> fill(cnt.begin(), cnt.end() - 1);
        for (int i = 0; i < n; i++) nodes[fill[colour_of(seed, i)]++] = i;
        DD_CUDA(h, cudaMemcpyAsync(bucket, nodes.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->stream));
        DD_CUDA(h, cudaStreamSynchronize(h->stream));  // `nodes` is a temporary (once per fit)
        h->lv_bucket_n = n;
        h->lv_bucket_seed = seed;
    }
    // default: one small launch per step, replayed from a CUDA graph -- the steps are latency-bound (~10 us each),
    // but as ordinary kernels on the clustering stream they share the SMs with the HBM-bound kernels of the next
    // iteration.  DD_LOUVAIN_COOP=1: all rounds in ONE cooperative launch (one CTA per SM, hand-written grid
    // barrier) -- fewer launches, but it monopolises the SMs.
    // DD_LOUVAIN_COOP=2: the same single launch as a LIGHT resident grid (256 threads, 32 KB per CTA, DD_LOUVAIN_COOP_CTAS
    // CTAs, default 128) that fits next to the main stream's CTAs on the same SMs instead of queueing ~300 tiny kernels
    // behind them.
    static const int coop = getenv("DD_LOUVAIN_COOP") ? atoi(getenv("DD_LOUVAIN_COOP")) : 0;
    if (coop != 0) {
        static const int light_ctas = getenv("DD_LOUVAIN_COOP_CTAS") ? atoi(getenv("DD_LOUVAIN_COOP_CTAS")) : 128;
        static dd_once_per_device attr_set;  // function attributes are per device
        attr_set.run(h->device, [&] {
            cudaFuncSetAttribute(k_lv_rounds<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem<1024>());
            cudaFuncSetAttribute(k_lv_rounds<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)round_smem<256>());
        });
        ColourOffsets co;
        for (int c = 0; c <= kColours; c++) co.v[c] = h->lv_colour_off[c];
        int n_arg = n;
        double g_arg = gamma;
        const int32_t *off_p = h->d_lv_off, *adj_p = h->d_lv_adj, *bucket_p = bucket;
        int32_t *comm_p = h->d_lv_comm, *csize_p = csize, *desired_p = desired, *counters_p = counters;
        double *tot_p = h->d_lv_tot;
        void *args[] = {&off_p, &adj_p, &comm_p, &tot_p, &csize_p, &bucket_p, &co, &n_arg, &g_arg, &desired_p, &counters_p};
        dd_launch_begin(h);
        cudaError_t e;
        if (coop == 1)
            e = cudaLaunchCooperativeKernel((const void *)k_lv_rounds<1024>, dim3(h->num_sms), dim3(1024), args,
                                            round_smem<1024>(), h->stream);
        else
            e = cudaLaunchCooperativeKernel((const void *)k_lv_rounds<256>, dim3(std::max(1, std::min(light_ctas, h->num_sms))),
                                            dim3(256), args, round_smem<256>(), h->stream);
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("cooperative launch of lv_rounds: ") + cudaGetErrorString(e));
        DD_TRY(dd_launch_end(h, "lv_rounds"));
        return DD_OK;
    }
    if (h->lv_graph_exec == nullptr || h->lv_graph_n != n || h->lv_graph_gamma != gamma || h->lv_graph_seed != seed) {
        if (h->lv_graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)h->lv_graph_exec);
        h->lv_graph_exec = nullptr;
        const bool timing = h->timing;
        const int64_t launches_before = h->launches;
        static const bool warp_per_node = getenv("DD_LOUVAIN_WARP") != nullptr;  // A/B: the round-1 propose kernel
        static const bool pdl = !warp_per_node && getenv("DD_LOUVAIN_NO_PDL") == nullptr;  // programmatic dependent launches
        // experiment: the shared-memory carve-out the small kernels ask for (percent of the maximum; an SM cannot run CTAs of
        // kernels with different carve-outs at the same time, and the main stream's kernels use the maximum)
        if (const char *cv = getenv("DD_LV_CARVEOUT")) {
            const int pct = atoi(cv);
            cudaFuncSetAttribute(k_lv_propose_g, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(k_lv_propose, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(k_lv_apply, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
            cudaFuncSetAttribute(k_lv_round_end, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        }
        h->timing = false;  // no event records inside the capture
        // DD_LOUVAIN_WHILE=1: ONE captured round as the body of a conditional WHILE node (ends when the level has converged);
        // default: kMaxRounds unrolled rounds whose kernels return at once after convergence
        static const bool use_while = getenv("DD_LOUVAIN_WHILE") != nullptr;
        cudaGraph_t graph = nullptr;
        cudaGraphConditionalHandle cond = 0;
        if (use_while) {
            DD_CUDA(h, cudaGraphCreate(&graph, 0));
            DD_CUDA(h, cudaGraphConditionalHandleCreate(&cond, graph, 1, cudaGraphCondAssignDefault));
            cudaGraphNodeParams np = {};
            np.type = cudaGraphNodeTypeConditional;
            np.conditional.handle = cond;
            np.conditional.type = cudaGraphCondTypeWhile;
            np.conditional.size = 1;
            cudaGraphNode_t node;
            DD_CUDA(h, cudaGraphAddNode(&node, graph, nullptr, 0, &np));
            DD_CUDA(h, cudaStreamBeginCaptureToGraph(h->stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                                     cudaStreamCaptureModeThreadLocal));
        } else {
            DD_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        }
        int rc = DD_OK;
        const int rounds_captured = use_while ? 1 : kMaxRounds;
        for (int round = 0; round < rounds_captured && rc == DD_OK; round++) {
            for (int c = 0; c < kColours && rc == DD_OK; c++) {
                const int b0 = h->lv_colour_off[c], b1 = h->lv_colour_off[c + 1];
                if (b1 == b0) continue;
                dd_launch_begin(h);
                if (warp_per_node)
                    k_lv_propose<<<(unsigned)((b1 - b0 + 7) / 8), 256, 0, h->stream>>>(h->d_lv_off, h->d_lv_adj, h->d_lv_comm,
                                                                                        h->d_lv_tot, csize, bucket, b0, b1, n,
                                                                                        gamma, desired, counters);
                else
                    launch_step(k_lv_propose_g, (unsigned)((b1 - b0 + kNodesPerCta - 1) / kNodesPerCta), kPropWarps * 32, h->stream,
                                pdl, h->d_lv_off, h->d_lv_adj, h->d_lv_comm, h->d_lv_tot, csize, bucket, b0, b1, n, gamma, desired,
                                counters);
                rc = dd_launch_end(h, "lv_propose");
                if (rc != DD_OK) break;
                dd_launch_begin(h);
                launch_step(k_lv_apply, (unsigned)((b1 - b0 + 255) / 256), 256, h->stream, pdl, h->d_lv_off, h->d_lv_comm, h->d_lv_tot,
                            csize, bucket, b0, b1, desired, counters);
                rc = dd_launch_end(h, "lv_apply");
            }
            if (rc != DD_OK) break;
            dd_launch_begin(h);
            launch_step(k_lv_round_end, 1, 1, h->stream, pdl, counters, n, use_while ? 1 : 0, cond, (int)kMaxRounds);
            rc = dd_launch_end(h, "lv_round_end");
        }
        cudaGraph_t captured = nullptr;
        cudaError_t e = cudaStreamEndCapture(h->stream, &captured);
        if (!use_while) graph = captured;
        h->timing = timing;
        h->lv_graph_launches = h->launches - launches_before;
        h->lv_graph_is_loop = use_while;
        h->launches = launches_before;
        if (rc != DD_OK) return rc;
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("louvain graph capture: ") + cudaGetErrorString(e));
        cudaGraphExec_t exec = nullptr;
        e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("louvain graph instantiate: ") + cudaGetErrorString(e));
        h->lv_graph_exec = exec;
        h->lv_graph_n = n;
        h->lv_graph_gamma = gamma;
        h->lv_graph_seed = seed;
    }
    dd_launch_begin(h);
    {
        cudaError_t e = cudaGraphLaunch((cudaGraphExec_t)h->lv_graph_exec, h->stream);
        if (e != cudaSuccess) return dd_fail(h, DD_ERR_CUDA, std::string("louvain graph launch: ") + cudaGetErrorString(e));
    }
    DD_TRY(dd_launch_end(h, "lv_rounds_graph"));
    // the replay runs every captured kernel; the loop variant runs its one captured round once per round of the level (the
    // count of the previous replay stands in: the host does not wait for this one)
    h->launches += h->lv_graph_launches * (h->lv_graph_is_loop ? std::max(1, h->h_lv_rounds ? *h->h_lv_rounds : 1) : 1) - 1;
    if (!h->h_lv_rounds && cudaMallocHost(&h->h_lv_rounds, sizeof(int32_t)) != cudaSuccess) h->h_lv_rounds = nullptr;
    if (h->h_lv_rounds)  // rounds actually executed (the replay launches all kMaxRounds; converged rounds return at once)
        cudaMemcpyAsync(h->h_lv_rounds, counters + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
    return DD_OK;
}

// Test / inspection hook: the PhenoGraph graph of the kNN lists computed last (dd_knn with k neighbours incl. self).
extern "C" int dd_jaccard_graph(dd_handle *h, int32_t k, int32_t prune, int32_t *indptr_out, int32_t *indices_out,
                                double *weights_out, int64_t capacity, int64_t *nnz_out) {
    if (!h || !indptr_out || !nnz_out) return dd_fail(h, DD_ERR_ARG, "dd_jaccard_graph: null argument");
    if (!h->emb_valid || !h->d_knn_idx) return dd_fail(h, DD_ERR_ARG, "dd_jaccard_graph: call dd_knn first");
    if (k != h->knn_last_k) return dd_fail(h, DD_ERR_ARG, "dd_jaccard_graph: k differs from the k of the last kNN on this handle");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_dev_jaccard_graph(h, k, prune));
    const int64_t n = h->emb_rows;
    DD_CUDA(h, cudaMemcpyAsync(indptr_out, h->d_lv_off, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    const int64_t nnz = indptr_out[n];
    *nnz_out = nnz;
    if (nnz > capacity || !indices_out || !weights_out) return DD_OK;  // caller sizes its buffers from nnz_out and calls again
    DD_CUDA(h, cudaMemcpyAsync(indices_out, h->d_lv_adj, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaMemcpyAsync(weights_out, h->d_lv_w, sizeof(double) * nnz, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

// Test / inspection hook: umap's connectivities of the kNN lists + distances computed last (dd_knn with k neighbours incl.
// self) as built on the device, in the canonical form the Leiden workers use (rows ascending, zero weights dropped) -- the
// same format as the host twin dd_umap_connectivities.  Call with capacity 0 to get nnz_out.
extern "C" int dd_umap_graph(dd_handle *h, int32_t k, int64_t *indptr_out, int32_t *indices_out, float *weights_out,
                             int64_t capacity, int64_t *nnz_out) {
    if (!h || !nnz_out) return dd_fail(h, DD_ERR_ARG, "dd_umap_graph: null argument");
    if (!h->emb_valid || !h->d_knn_idx || !h->d_knn_dist) return dd_fail(h, DD_ERR_ARG, "dd_umap_graph: call dd_knn first");
    if (k != h->knn_last_k) return dd_fail(h, DD_ERR_ARG, "dd_umap_graph: k differs from the k of the last kNN on this handle");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_dev_umap_graph(h, k));
    const int64_t n = h->emb_rows;
    std::vector<int32_t> off((size_t)n + 1);
    DD_CUDA(h, cudaMemcpyAsync(off.data(), h->d_lv_off, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    const int64_t raw = off[n];
    std::vector<int32_t> adj((size_t)std::max<int64_t>(raw, 1));
    std::vector<double> w((size_t)std::max<int64_t>(raw, 1));
    DD_CUDA(h, cudaMemcpyAsync(adj.data(), h->d_lv_adj, sizeof(int32_t) * raw, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaMemcpyAsync(w.data(), h->d_lv_w, sizeof(double) * raw, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    const int rc = dd_host_umap_canonical(n, off.data(), adj.data(), w.data(), indptr_out, indices_out, weights_out, capacity, nnz_out);
    if (rc != DD_OK) return dd_fail(h, rc, "dd_umap_graph: the device graph is malformed or an output pointer is null");
    return DD_OK;
}
