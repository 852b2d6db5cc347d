// knn.cu -- exact Euclidean k nearest neighbours on the PCA embedding: the kNN search inside
// sc.pp.neighbors(n_neighbors=10) (doubletdetection.py:331-336; below 8192 observations scanpy runs
// sklearn's brute-force KNeighborsTransformer, which returns the point itself in column 0 followed by
// the k-1 nearest others -- restated in oracle/upstream.py:knn_brute).
//
// Design: one CTA owns 128 query rows and sweeps all candidate tiles (128 rows, cp.async double
// buffered; the whole embedding is L2 resident).  Squared distances come from the expanded form
// |q|^2 + |c|^2 - 2 q.c with an 8x8 register tile per thread.  Selection is threshold-filtered: each
// query keeps a sorted list of its TL best (distance, index) pairs in shared memory; a thread that sees
// a candidate no worse than the query's current TL-th best appends it to a small per-query buffer, and
// after the tile one warp per query merges the (rare) survivors with shuffles.  The float32 expanded
// form can mis-order near ties, so TL exceeds k-1 by a margin and the final k-1 are chosen by exact
// float64 distances (order: distance, then index).
#include "dd_internal.h"

#include <vector>

#include <cuda_bf16.h>

#include <cfloat>
#include <cstdlib>

namespace {

constexpr int BQ = 128, BC = 128, CAP = 16;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int KP>
__global__ void k_row_norms(const float *__restrict__ emb, int64_t n, float *__restrict__ norms) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < KP; c += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(emb + i * KP + c);
        s = fmaf(v.x, v.x, s);
        s = fmaf(v.y, v.y, s);
        s = fmaf(v.z, v.z, s);
        s = fmaf(v.w, v.w, s);
    }
    norms[i] = s;
}

// insert (d, idx) into the sorted list held by lanes [0, TL) of the warp
template <int TL>
__device__ __forceinline__ void list_insert(float &ld, int &li, float d, int idx, int lane) {
    const bool less = lane < TL && (ld < d || (ld == d && li < idx));
    const int pos = __popc(__ballot_sync(0xffffffffu, less));
    const float ud = __shfl_up_sync(0xffffffffu, ld, 1);
    const int ui = __shfl_up_sync(0xffffffffu, li, 1);
    if (lane == pos) {
        ld = d;
        li = idx;
    } else if (lane > pos) {
        ld = ud;
        li = ui;
    }
}

template <int KP, int TL>
struct KnnSmem {
    static constexpr int ST = KP + 4;
    float q[BQ][ST];
    float c[2][BC][ST];
    float qn[BQ];
    float cn[2][BC];
    float tau[BQ];
    float list_d[BQ][TL];
    int list_i[BQ][TL];
    float buf_d[BQ][CAP];
    int buf_i[BQ][CAP];
    int cnt[BQ];
};

template <int KP, int TL>
__global__ void __launch_bounds__(256) k_knn_scan(const float *__restrict__ emb, const float *__restrict__ norms,
                                                  int64_t n, float *__restrict__ cand_d, int *__restrict__ cand_i) {
    using S = KnnSmem<KP, TL>;
    constexpr int ST = S::ST;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S &sm = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tq = tid >> 4, tc = tid & 15;
    const int64_t q0 = (int64_t)blockIdx.x * BQ;
    const int n_tiles = (int)((n + BC - 1) / BC);

    auto load_cand = [&](int s, int t) {
        const int64_t c0 = (int64_t)t * BC;
        for (int e = tid; e < BC * (KP / 4); e += 256) {
            const int r = e / (KP / 4), kc = e % (KP / 4);
            const bool ok = c0 + r < n;
            cp_async16(&sm.c[s][r][kc * 4], emb + (ok ? (c0 + r) * KP + kc * 4 : 0), ok);
        }
        if (tid < BC) sm.cn[s][tid] = (c0 + tid < n) ? norms[c0 + tid] : INFINITY;
    };

    // queries + list initialisation
    for (int e = tid; e < BQ * (KP / 4); e += 256) {
        const int r = e / (KP / 4), kc = e % (KP / 4);
        const bool ok = q0 + r < n;
        cp_async16(&sm.q[r][kc * 4], emb + (ok ? (q0 + r) * KP + kc * 4 : 0), ok);
    }
    if (tid < BQ) {
        sm.qn[tid] = (q0 + tid < n) ? norms[q0 + tid] : 0.f;
        sm.tau[tid] = FLT_MAX;
        sm.cnt[tid] = 0;
    }
    for (int e = tid; e < BQ * TL; e += 256) {
        sm.list_d[e / TL][e % TL] = FLT_MAX;
        sm.list_i[e / TL][e % TL] = 0x7fffffff;
    }
    load_cand(0, 0);
    cp_async_commit();

    for (int t = 0; t < n_tiles; t++) {
        const int s = t & 1;
        if (t + 1 < n_tiles) load_cand(s ^ 1, t + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        // ---- 8 x 8 dot products per thread: queries tq + 16 i, candidates tc + 16 j
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
#pragma unroll 2
        for (int kk = 0; kk < KP; kk += 4) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const float4 *>(&sm.q[tq + 16 * i][kk]);
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 b = *reinterpret_cast<const float4 *>(&sm.c[s][tc + 16 * j][kk]);
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    acc[i][j] = fmaf(a[i].x, b.x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b.y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b.z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b.w, acc[i][j]);
                }
            }
        }
        // ---- threshold filter
        const int64_t c0 = (int64_t)t * BC;
        const bool diag = (c0 < q0 + BQ) && (q0 < c0 + BC);
        float cn[8];
#pragma unroll
        for (int j = 0; j < 8; j++) cn[j] = sm.cn[s][tc + 16 * j];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int q = tq + 16 * i;
            const float qn = sm.qn[q], tau = sm.tau[q];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float d2 = fmaf(-2.f, acc[i][j], qn + cn[j]);
                if (d2 <= tau) {
                    const int c = tc + 16 * j;
                    if (!(diag && q0 + q == c0 + c)) {
                        const int slot = atomicAdd(&sm.cnt[q], 1);
                        if (slot < CAP) {
                            sm.buf_d[q][slot] = d2;
                            sm.buf_i[q][slot] = (int)(c0 + c);
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- merge survivors: warp w owns queries [16 w, 16 w + 16)
        {
            const int qb = warp * 16;
            const int my_cnt = lane < 16 ? sm.cnt[qb + lane] : 0;
            unsigned pending = __ballot_sync(0xffffffffu, my_cnt > 0);
            while (pending) {
                const int ql = __ffs(pending) - 1;
                pending &= pending - 1;
                const int q = qb + ql;
                const int cnt = __shfl_sync(0xffffffffu, my_cnt, ql);
                float ld = lane < TL ? sm.list_d[q][lane] : FLT_MAX;
                int li = lane < TL ? sm.list_i[q][lane] : 0x7fffffff;
                if (cnt <= CAP) {
                    for (int e = 0; e < cnt; e++) list_insert<TL>(ld, li, sm.buf_d[q][e], sm.buf_i[q][e], lane);
                } else {
                    // buffer overflow (first tiles): recompute this query against the whole tile
                    const float qn = sm.qn[q];
                    for (int cb = 0; cb < BC; cb += 32) {
                        const int c = cb + lane;
                        float dot = 0.f;
#pragma unroll
                        for (int kk = 0; kk < KP; kk += 4) {
                            const float4 a = *reinterpret_cast<const float4 *>(&sm.q[q][kk]);
                            const float4 b = *reinterpret_cast<const float4 *>(&sm.c[s][c][kk]);
                            dot = fmaf(a.x, b.x, dot);
                            dot = fmaf(a.y, b.y, dot);
                            dot = fmaf(a.z, b.z, dot);
                            dot = fmaf(a.w, b.w, dot);
                        }
                        float d2 = fmaf(-2.f, dot, qn + sm.cn[s][c]);
                        if (q0 + q == c0 + c) d2 = INFINITY;
                        const float tau_now = __shfl_sync(0xffffffffu, ld, TL - 1);
                        unsigned pass = __ballot_sync(0xffffffffu, d2 <= tau_now);
                        while (pass) {
                            const int src = __ffs(pass) - 1;
                            pass &= pass - 1;
                            const float d = __shfl_sync(0xffffffffu, d2, src);
                            list_insert<TL>(ld, li, d, (int)(c0 + cb + src), lane);
                        }
                    }
                }
                if (lane < TL) {
                    sm.list_d[q][lane] = ld;
                    sm.list_i[q][lane] = li;
                }
                if (lane == TL - 1) sm.tau[q] = ld;
                if (lane == 0) sm.cnt[q] = 0;
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
    for (int e = tid; e < BQ * TL; e += 256) {
        const int q = e / TL, l = e % TL;
        if (q0 + q < n) {
            cand_d[(q0 + q) * TL + l] = sm.list_d[q][l];
            cand_i[(q0 + q) * TL + l] = sm.list_i[q][l];
        }
    }
}

// Exact float64 re-ranking of the TL candidates of every query; writes self + (k-1) neighbours.
template <int KP, int TL>
__global__ void k_knn_refine(const float *__restrict__ emb, const int *__restrict__ cand_i, int64_t q0, int64_t n, int k,
                             int32_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    const int lane = threadIdx.x & 31;
    const int64_t q = q0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // queries [q0, n)
    if (q >= n) return;
    int ci = lane < TL ? cand_i[q * TL + lane] : 0x7fffffff;
    double d = INFINITY;
    if (ci != 0x7fffffff) {
        d = 0.0;
#pragma unroll
        for (int c = 0; c < KP; c += 4) {
            const float4 a = *reinterpret_cast<const float4 *>(emb + q * KP + c);
            const float4 b = *reinterpret_cast<const float4 *>(emb + (int64_t)ci * KP + c);
            const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
            const double dz = (double)a.z - (double)b.z, dw = (double)a.w - (double)b.w;
            d += dx * dx + dy * dy + dz * dz + dw * dw;
        }
    }
    int rank = 0;
    for (int l = 0; l < TL; l++) {
        const double od = __shfl_sync(0xffffffffu, d, l);
        const int oi = __shfl_sync(0xffffffffu, ci, l);
        rank += (od < d) || (od == d && oi < ci);
    }
    if (lane == 0) {
        idx_out[q * k] = (int32_t)q;
        dist_out[q * k] = 0.f;
    }
    if (lane < TL && rank < k - 1) {
        idx_out[q * k + 1 + rank] = ci == 0x7fffffff ? -1 : ci;
        dist_out[q * k + 1 + rank] = (float)sqrt(d);
    }
}


// ---- certificate of the approximate filter ------------------------------------------------------------------------------
// The tcgen05 kernels rank candidates by an APPROXIMATE score (3 of the 9 bf16 cross products of q.c, fp32 accumulation):
//     |approx - exact| <= e(q, c) = 2^-16 |q| |c| + 2^-26 |c|^2        (dropped terms 3 x 2^-18, accumulation, norm term)
// A row's result is exact iff no candidate the filter EXCLUDED is closer than the reported k-th neighbour.  An excluded y
// scored at most the filter's last kept candidate, hence at most the kept candidate z with the largest exact distance
// sqrt(D) plus e(z); if y were closer than the k-th neighbour (distance d_k), then |y| <= |q| + d_k, and
//     D - d_k^2  >  2 [ e(|q| + d_k) + e(|q| + sqrt(D)) ]
// contradicts it.  Rows that fail the test (embeddings whose offset from the origin is ~1000 x their local spacing) are
// appended to `rows` and re-done by brute force in float64 (k_knn_exact_rows): the result is exact either way.
// Two lists (cluster-ordered kNN): list A = own group, list B = other groups above A's threshold; an excluded candidate is
// bounded by A's last kept (own group, or B not full) or by B's last kept (B full): D = min(D_A, B full ? D_B : D_A).
struct KnnCert {
    int list_w = 0;          // 0: no certification
    int n_lists = 1;
    int *count = nullptr;    // number of uncertified rows
    int *rows = nullptr;     // their indices (capacity `cap`)
    int cap = 0;
};

__device__ __forceinline__ double knn_filter_err(double qn, double r) {
    const double c = qn + r;
    return 1.52587890625e-05 * qn * c + 1.4901161193847656e-08 * c * c;  // 2^-16, 2^-26
}

// exact re-ranking of `width` <= 32 PER candidates per query (one or two lists side by side): every lane owns PER of them
template <int PER>
__global__ void k_knn_refine_w(const float *__restrict__ emb, const int *__restrict__ cand_i, int width, int64_t q0, int64_t n,
                               int k, int32_t *__restrict__ idx_out, float *__restrict__ dist_out, KnnCert cert = KnnCert(),
                               const int32_t *__restrict__ row_map = nullptr, int shard_world = 1, int shard_rank = 0) {
    const int lane = threadIdx.x & 31;
    int64_t q = q0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // queries [q0, n)
    if (q >= n) return;
    // cell-block sharding of the cluster-ordered kNN: the 256-row blocks of the PERMUTED order are dealt round-robin; row_map
    // (the permutation) turns a permuted position into the original row whose candidates and output these are
    if (shard_world > 1 && ((q >> 8) % shard_world) != shard_rank) return;
    if (row_map) {
        q = row_map[q];
        if (q < 0) return;  // padding position
    }
    int ci[PER];
    double d[PER];
#pragma unroll
    for (int s = 0; s < PER; s++) {
        const int col = 32 * s + lane;
        ci[s] = col < width ? cand_i[q * width + col] : 0x7fffffff;
        d[s] = INFINITY;
        if (ci[s] != 0x7fffffff) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 a = *reinterpret_cast<const float4 *>(emb + q * 32 + c);
                const float4 b = *reinterpret_cast<const float4 *>(emb + (int64_t)ci[s] * 32 + c);
                const double dx = (double)a.x - (double)b.x, dy = (double)a.y - (double)b.y;
                const double dz = (double)a.z - (double)b.z, dw = (double)a.w - (double)b.w;
                acc += dx * dx + dy * dy + dz * dz + dw * dw;
            }
            d[s] = acc;
        }
    }
    int rank[PER];
#pragma unroll
    for (int t = 0; t < PER; t++) rank[t] = 0;
    for (int l = 0; l < 32; l++) {
#pragma unroll
        for (int s = 0; s < PER; s++) {
            const double od = __shfl_sync(0xffffffffu, d[s], l);
            const int oi = __shfl_sync(0xffffffffu, ci[s], l);
#pragma unroll
            for (int t = 0; t < PER; t++) rank[t] += (od < d[t]) || (od == d[t] && oi < ci[t]);
        }
    }
    if (lane == 0) {
        idx_out[q * k] = (int32_t)q;
        dist_out[q * k] = 0.f;
    }
#pragma unroll
    for (int s = 0; s < PER; s++)
        if (rank[s] < k - 1) {
            idx_out[q * k + 1 + rank[s]] = ci[s] == 0x7fffffff ? -1 : ci[s];
            dist_out[q * k + 1 + rank[s]] = (float)sqrt(d[s]);
        }
    if (cert.list_w > 0) {  // uniform
        // per list: is it full (no "none" entry), and the largest exact distance^2 among its entries
        double dmax[2] = {0.0, 0.0}, dk2 = -1.0;
        bool hole[2] = {false, false};
#pragma unroll
        for (int s = 0; s < PER; s++) {
            const int col = 32 * s + lane;
            if (col < width) {
                const int li = col >= cert.list_w ? 1 : 0;
                if (ci[s] == 0x7fffffff) hole[li] = true; else dmax[li] = fmax(dmax[li], d[s]);
                if (rank[s] == k - 2 && ci[s] != 0x7fffffff) dk2 = d[s];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dmax[0] = fmax(dmax[0], __shfl_xor_sync(0xffffffffu, dmax[0], o));
            dmax[1] = fmax(dmax[1], __shfl_xor_sync(0xffffffffu, dmax[1], o));
            dk2 = fmax(dk2, __shfl_xor_sync(0xffffffffu, dk2, o));
        }
        const bool hole0 = __any_sync(0xffffffffu, hole[0]), hole1 = __any_sync(0xffffffffu, hole[1]);
        const float x = emb[q * 32 + lane];
        double qn2 = (double)x * (double)x;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qn2 += __shfl_xor_sync(0xffffffffu, qn2, o);
        const double da = hole0 ? INFINITY : dmax[0];
        const double db = (cert.n_lists > 1 && !hole1) ? dmax[1] : da;
        const double D = fmin(da, db);
        bool ok = true;
        if (D < INFINITY && dk2 >= 0.0) {  // something was excluded, and k - 1 neighbours were reported
            const double qn = sqrt(qn2);
            ok = D - dk2 > 2.0 * (knn_filter_err(qn, sqrt(dk2)) + knn_filter_err(qn, sqrt(D)));
        }
        if (!ok && lane == 0) {
            const int slot = atomicAdd(cert.count, 1);
            if (slot < cert.cap) cert.rows[slot] = (int)q;
        }
    }
}

// Brute force in float64 for the rows the certificate could not clear: one warp per row, every lane keeps the k - 1 best of
// its share of the candidates (sorted by (distance, index), the order of the re-ranking kernels), the lanes' lists are merged
// by repeated warp-wide minimum.  Same distance arithmetic as k_knn_refine_w, so the two paths agree to the bit.
__global__ void __launch_bounds__(256) k_knn_exact_rows(const float *__restrict__ emb, int64_t n, const int *__restrict__ rows,
                                                        const int *__restrict__ count, int cap, int k,
                                                        int32_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = gridDim.x * (blockDim.x >> 5);
    const int total = min(*count, cap);
    const int kk = k - 1;
    for (int it = warp; it < total; it += n_warps) {
        const int64_t q = rows[it];
        float4 qa[8];
#pragma unroll
        for (int c = 0; c < 8; c++) qa[c] = *reinterpret_cast<const float4 *>(emb + q * 32 + 4 * c);
        double ld[31];
        int li[31];
        int cnt = 0;
        for (int64_t cnd = lane; cnd < n; cnd += 32) {
            if (cnd == q) continue;
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 b = *reinterpret_cast<const float4 *>(emb + cnd * 32 + 4 * c);
                const double dx = (double)qa[c].x - (double)b.x, dy = (double)qa[c].y - (double)b.y;
                const double dz = (double)qa[c].z - (double)b.z, dw = (double)qa[c].w - (double)b.w;
                acc += dx * dx + dy * dy + dz * dz + dw * dw;
            }
            if (cnt == kk && !(acc < ld[kk - 1])) continue;  // candidates arrive in increasing index: a tie loses
            int pos = cnt < kk ? cnt : kk - 1;
            while (pos > 0 && acc < ld[pos - 1]) {
                ld[pos] = ld[pos - 1];
                li[pos] = li[pos - 1];
                pos--;
            }
            ld[pos] = acc;
            li[pos] = (int)cnd;
            if (cnt < kk) cnt++;
        }
        int head = 0;
        for (int r = 0; r < kk; r++) {
            double bd = head < cnt ? ld[head] : INFINITY;
            int bi = head < cnt ? li[head] : 0x7fffffff;
            int bl = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
                if (od < bd || (od == bd && oi < bi)) {
                    bd = od;
                    bi = oi;
                    bl = ol;
                }
            }
            if (bl == lane && bi != 0x7fffffff) head++;
            if (lane == 0) {
                idx_out[q * k + 1 + r] = bi == 0x7fffffff ? -1 : bi;
                dist_out[q * k + 1 + r] = (float)sqrt(bd);
            }
        }
        if (lane == 0) {
            idx_out[q * k] = (int32_t)q;
            dist_out[q * k] = 0.f;
        }
    }
}

// =================================================================================================
// Tensor-core path (KP == 32, TL == 16): the distance GEMM on tcgen05 / TMEM.
//
//   t(q, c) = q . c - |c|^2 / 2     (larger = closer; the query norm does not change a query's ranking)
//
// is one K = 112 BF16 GEMM ("3xBF16"): every coordinate is split into bf16 parts x = x1 + x2 (+ x3, dropped),
// and the three significant cross products plus the norm term are concatenated along K:
//   A' (query row)     = [ q1(32) | q1(32) | q2(32) | 1, 1, 1, 0...0 ]
//   B' (candidate row) = [ c1(32) | c2(32) | c1(32) | n1, n2, n3, 0...0 ],  n = -|c|^2/2 in three bf16 parts
// The dropped terms are 2^-17 relative: enough for a FILTER whose top 16 are re-ranked exactly in float64
// (k_knn_refine), and half the tensor time and operand bytes of the 3xTF32 variant it replaces.
// k_knn_prep writes both operands tile by tile (128 rows) in the canonical no-swizzle K-major UMMA
// layout (8 x 16-byte core matrices, LBO = 128 B along K, SBO = 1792 B along rows), so that a stage is
// ONE 28 KB bulk copy (cp.async.bulk, TMA engine) and the shared-memory descriptors are constants.
//
// k_knn_tc: one CTA = 256 query rows (two M=128 accumulators) x all candidate tiles.
//   warp 0    bulk-copy producer (4-stage ring of candidate tiles, mbarrier complete_tx)
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer (7 K-steps x 2 query tiles per stage)
//   warps 2-9 epilogue: tcgen05.ld 32 columns at a time, one query row per thread; a value survives only
//             if it beats the row's current 16th best, and the survivors are inserted into the row's sorted
//             candidate list, which lives in the thread's registers
// TMEM: 512 columns = 2 (double buffer) x 2 (query tiles) x 128 fp32 accumulator columns.
// Cell-block sharding: a launch covers the query-tile pairs [pair0, pair0 + gridDim.x) against ALL candidates.
namespace tc {

constexpr int KC = 14;                         // 16-byte chunks (8 bf16) per operand row (K = 112)
constexpr int TILE = 128;                      // rows per operand tile
constexpr int TILE_BYTES = TILE * KC * 16;     // 28672
constexpr int LBO = 128, SBO = KC * 128;       // bytes
constexpr int QT = 2;                          // query tiles per CTA
constexpr int NS = 4;                          // candidate stages
constexpr int KSTEPS = KC / 2;                 // 7 MMAs of K = 16
constexpr float kEmptyT = -1e29f;              // list filler; padded candidates score -1e30 and never pass
// TN = candidate rows per pipeline step.  128: one whole operand tile per step, 512 TMEM columns (2 buffers x 2 query tiles
// x 128), 169 KB of shared memory -- the SM is this kernel's alone.  64 (the fit loop's choice when the kNN runs on its own
// stream next to the following iteration's PCA): half a tile per step, 256 TMEM columns, 113 KB -- one CTA of the HBM-bound
// PCA product kernel (87 KB, 256 columns) fits on the SM beside it, so the two streams overlap on every SM instead of
// taking turns.
template <int TN>
constexpr size_t smem_bytes() { return (size_t)QT * TILE_BYTES + (size_t)NS * TN * KC * 16 + 1024; }
constexpr size_t SMEM_BYTES = smem_bytes<128>();

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle: start address, LBO (K-adjacent core matrices), SBO (row-adjacent), version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(LBO >> 4) << 16;
    d |= (uint64_t)(SBO >> 4) << 32;
    d |= 1ull << 46;
    return d;
}
// kind::f16 with BF16 operands, fp32 accumulate, A and B K-major, M = 128, N = 128 / 64
template <int TN>
struct Idesc {
    static constexpr uint32_t value = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)TN >> 3) << 17) | ((128u >> 4) << 24);
};

// x = b1 + b2 + b3 with bf16 parts (round to nearest at every step; the remainders are exact in float32)
__device__ __forceinline__ void bf16_split3(float x, __nv_bfloat16 &b1, __nv_bfloat16 &b2, __nv_bfloat16 &b3) {
    b1 = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(b1);
    b2 = __float2bfloat16_rn(r1);
    b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
}

// Operand tiles in the canonical layout.  Block = 8 rows x 14 chunks (one chunk = 8 bf16 = 16 bytes).
//   chunks 0-3: part 1 of dims 0-31 (A and B);  4-7: A part 1 / B part 2;  8-11: A part 2 / B part 1;
//   chunk 12: A = (1, 1, 1, 0...) / B = (n1, n2, n3, 0...);  chunk 13: zero
__global__ void __launch_bounds__(112) k_knn_prep(const float *__restrict__ emb, int64_t n, int64_t n_pad,
                                                  uint4 *__restrict__ qa, uint4 *__restrict__ cb) {
    const int j = threadIdx.x >> 3, rr = threadIdx.x & 7;
    const int64_t row = (int64_t)blockIdx.x * 8 + rr;
    if (row >= n_pad) return;
    const bool real = row < n;
    __nv_bfloat16 a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; e++) a[e] = b[e] = __float2bfloat16_rn(0.f);
    if (j < 12) {
        if (real) {
            const int d0 = 8 * (j & 3);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                __nv_bfloat16 p1, p2, p3;
                bf16_split3(emb[row * 32 + d0 + e], p1, p2, p3);
                a[e] = j < 8 ? p1 : p2;
                b[e] = (j >= 4 && j < 8) ? p2 : p1;
            }
        }
    } else if (j == 12) {
        if (real) {
            double nn = 0.0;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 x = *reinterpret_cast<const float4 *>(emb + row * 32 + c);
                nn += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
            }
            const double half = -0.5 * nn;
            const __nv_bfloat16 n1 = __float2bfloat16_rn((float)half);
            const double r1 = half - (double)__bfloat162float(n1);
            const __nv_bfloat16 n2 = __float2bfloat16_rn((float)r1);
            const __nv_bfloat16 n3 = __float2bfloat16_rn((float)(r1 - (double)__bfloat162float(n2)));
            a[0] = a[1] = a[2] = __float2bfloat16_rn(1.f);
            b[0] = n1;
            b[1] = n2;
            b[2] = n3;
        } else {
            b[0] = __float2bfloat16_rn(-1e30f);  // padded candidates can never be selected
        }
    }
    const int64_t tile = row / TILE;
    const int r = (int)(row % TILE);
    const int64_t off16 = tile * (TILE_BYTES / 16) + (int64_t)(r >> 3) * (SBO / 16) + (int64_t)j * (LBO / 16) + (r & 7);
    qa[off16] = *reinterpret_cast<const uint4 *>(a);
    cb[off16] = *reinterpret_cast<const uint4 *>(b);
}

// The candidate list of a query row lives in the registers of its epilogue thread, sorted by decreasing t.
// Insertion is position-parallel: slot l takes its upper neighbour if the new score beats that neighbour
// (everything from the insertion point on shifts down), the new entry if it beats only slot l itself.  All slots
// are independent, so the latency is one compare + two selects instead of an N-long dependent chain.
template <int N>
struct RegList {
    float t[N];
    int i[N];
};
template <int N>
__device__ __forceinline__ void reglist_insert(RegList<N> &L, float t, int idx) {
    bool gt[N];
#pragma unroll
    for (int l = 0; l < N; l++) gt[l] = t > L.t[l];  // strict: an equal score keeps the earlier entry in front
#pragma unroll
    for (int l = N - 1; l >= 1; l--) {
        L.t[l] = gt[l - 1] ? L.t[l - 1] : (gt[l] ? t : L.t[l]);
        L.i[l] = gt[l - 1] ? L.i[l - 1] : (gt[l] ? idx : L.i[l]);
    }
    L.t[0] = gt[0] ? t : L.t[0];
    L.i[0] = gt[0] ? idx : L.i[0];
}

// What bounds this kernel (measured, profiles/r1j_*): the epilogue has to LOOK at every fp32 score, and the TMEM ->
// register path moves 64 B per cycle per SM (B300_MICROARCH "LDTM throughput"): 128 KB of accumulators per step =
// 2048 cycles, against ~900 cycles of tensor-core time for the step's 14 MMAs.  Neither more epilogue warps (16 warps
// with per-half lists: 7.1 ms) nor cheaper list maintenance (pending slots merged in lockstep: 5.9-7.4 ms) beat the
// 8-warp peel-and-insert epilogue below (5.25 ms at c3, 3.4 ms TMEM-read floor, 4.1 ms with the 3.3 -> 4 wave rounding).
constexpr int THREADS = 320;  // producer warp + MMA warp + 8 epilogue warps

// LISTED (experimental, dd_knn_listed): instead of ALL candidate tiles a query-tile pair visits the tiles of its own list
// (list_tiles[list_off[pair] .. list_off[pair + 1])) -- the three roles below loop over the same step counter, so the list
// only changes which tile a step loads and which candidate indices it stands for.
template <int LIST, bool LISTED = false, int TN = 128>  // LIST: candidates kept per query row, 16 (k <= 13) or 32 (k <= 31, PhenoGraph)
__global__ void __launch_bounds__(THREADS, 1)
    k_knn_tc(const uint8_t *__restrict__ qa, const uint8_t *__restrict__ cb, int64_t n, int n_tiles, int pair0,
             int n_full, int *__restrict__ cand_i, const int *__restrict__ list_off = nullptr,
             const int *__restrict__ list_tiles = nullptr, const int *__restrict__ list_len = nullptr,
             const int *__restrict__ block_order = nullptr, const float *__restrict__ tau_init = nullptr,
             float *__restrict__ tau_out = nullptr, int shard_world = 1, int shard_rank = 0) {
    constexpr int SUB = TILE / TN;                     // pipeline steps per 128-row candidate tile
    constexpr int STAGE_BYTES = TN * KC * 16;
    constexpr uint32_t TMEM_COLS = 4 * TN;             // 2 buffers x 2 query tiles x TN accumulator columns
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sA = smem;                                // QT tiles
    uint8_t *sB = smem + (size_t)QT * TILE_BYTES;      // NS stages
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)QT * TILE_BYTES + (size_t)NS * STAGE_BYTES);
    uint64_t *a_full = bars;            // 1
    uint64_t *full = bars + 1;          // NS
    uint64_t *empty = full + NS;        // NS
    uint64_t *tfull = empty + NS;       // 2
    uint64_t *tempty = tfull + 2;       // 2
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pair0: first query-tile pair of this launch (cell-block sharding).  The first n_full CTAs own a pair of query tiles;
    // the CTAs behind them own ONE tile each (half the work): the launch's last, partial wave is cut into half-size pieces
    // so that it occupies twice as many SMs for half as long.
    const bool half_cta = (int)blockIdx.x >= n_full;
    const int n_qt = half_cta ? 1 : QT;
    int tile0 = half_cta ? (pair0 + n_full) * QT + ((int)blockIdx.x - n_full) : ((int)blockIdx.x + pair0) * QT;
    const int *my_list = nullptr;
    if constexpr (LISTED) {  // launched without half CTAs: one list per pair
        // block_order: the pairs in order of decreasing list length (longest first: the hardware hands CTAs out in index
        // order, so the tail of the launch is made of the short lists).  list_len: lists at fixed strides (list_off[p]) with
        // explicit lengths, instead of packed lists delimited by list_off[p + 1]
        const int blk = pair0 + (block_order ? block_order[blockIdx.x] : (int)blockIdx.x);
        // cell-block sharding: the query blocks are dealt to the ranks round-robin (whole CTA leaves, before any barrier)
        if (shard_world > 1 && (blk % shard_world) != shard_rank) return;
        const int off = list_off[blk];
        n_tiles = list_len ? list_len[blk] : list_off[blk + 1] - off;
        my_list = list_tiles + off;
        tile0 = blk * QT;
    }
    const int n_steps = n_tiles * SUB;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < NS; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(tfull + b, 1);
            mbar_init(tempty + b, 128 * n_qt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(a_full, n_qt * TILE_BYTES);
            for (int qt = 0; qt < n_qt; qt++)
                bulk_g2s(sA + (size_t)qt * TILE_BYTES, qa + (size_t)(tile0 + qt) * TILE_BYTES, TILE_BYTES, a_full);
            for (int step = 0; step < n_steps; step++) {
                const int s = step % NS;
                const uint32_t ph = (step / NS) & 1;
                mbar_wait(empty + s, ph ^ 1);
                mbar_expect_tx(full + s, STAGE_BYTES);
                int tile = step / SUB;
                if constexpr (LISTED) tile = my_list[step / SUB];
                // a TN-row slice of a tile is contiguous in the canonical layout (whole 8-row groups of SBO bytes)
                bulk_g2s(sB + (size_t)s * STAGE_BYTES, cb + (size_t)tile * TILE_BYTES + (size_t)(step % SUB) * STAGE_BYTES,
                         STAGE_BYTES, full + s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(a_full, 0);
            fence_after();
            uint64_t a_desc[QT];
            for (int qt = 0; qt < QT; qt++) a_desc[qt] = make_desc(smem_u32(sA + (size_t)qt * TILE_BYTES));
            for (int step = 0; step < n_steps; step++) {
                const int s = step % NS;
                const uint32_t ph = (step / NS) & 1;
                const int buf = step & 1;
                const uint32_t bph = (step >> 1) & 1;
                mbar_wait(full + s, ph);
                mbar_wait(tempty + buf, bph ^ 1);
                fence_after();
                const uint64_t b_desc = make_desc(smem_u32(sB + (size_t)s * STAGE_BYTES));
#pragma unroll
                for (int qt = 0; qt < QT; qt++) {
                    if (qt >= n_qt) break;
                    const uint32_t d = tmem_base + buf * (2 * TN) + qt * TN;
#pragma unroll
                    for (int k = 0; k < KSTEPS; k++)
                        mma_bf16(d, a_desc[qt] + (uint64_t)(k * 2 * LBO / 16), b_desc + (uint64_t)(k * 2 * LBO / 16), Idesc<TN>::value,
                                 k > 0);
                }
                mma_commit(empty + s);     // the stage may be refilled once these MMAs have read it
                mma_commit(tfull + buf);   // accumulators ready for the epilogue
            }
        }
    } else {
        const int e = warp - 2;          // 0..7
        const int qt = e >> 2;           // query tile of the pair
        const int quad = warp & 3;       // TMEM lane quadrant this warp may access (four consecutive e cover all)
        const int64_t qrow = (int64_t)(tile0 + qt) * TILE + quad * 32 + lane;
        const bool active = qrow < n;
        const int my_tiles = qt < n_qt ? n_steps : 0;  // the second tile's warps of a half CTA have nothing to do
        RegList<LIST> L;
#pragma unroll
        for (int l = 0; l < LIST; l++) {
            L.t[l] = kEmptyT;
            L.i[l] = 0x7fffffff;
        }
        // tau_init (cluster-ordered kNN, second launch): the row's LIST-th best score of the first launch -- only candidates
        // that beat it can still belong to the row's LIST best overall (the two lists are merged by the re-ranking)
        float tau0 = kEmptyT;
        if constexpr (LISTED)
            if (tau_init && active) tau0 = tau_init[qrow];
        float tau = active ? tau0 : INFINITY;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16) + qt * TN;
        for (int step = 0; step < my_tiles; step++) {
            const int buf = step & 1;
            const uint32_t bph = (step >> 1) & 1;
            mbar_wait(tfull + buf, bph);
            fence_after();
            // software pipeline: the tcgen05.ld of the next 32 columns is in flight while this group is scanned
            uint32_t va[32], vb[32];
            const uint32_t col0 = lane_base + buf * (2 * TN);
            int cand0 = (step / SUB) * TILE + (step % SUB) * TN;
            if constexpr (LISTED) cand0 = my_list[step / SUB] * TILE + (step % SUB) * TN;
            auto scan = [&](uint32_t (&v)[32], int c) {
                // balanced max tree (depth 5) instead of a 31-long dependent chain
                float m8[8];
#pragma unroll
                for (int i = 0; i < 8; i++)
                    m8[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                                  fmaxf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
                float m = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])),
                                fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                // some value of this row beats its LIST-th best: peel maxima until none does
                while (m > tau) {
                    int pos = 0;
                    bool done = false;
#pragma unroll
                    for (int i = 0; i < 32; i++) {  // peel the FIRST column holding the maximum
                        const bool hit = !done && __uint_as_float(v[i]) == m;
                        pos = hit ? i : pos;
                        v[i] = hit ? 0xff800000u : v[i];  // -inf
                        done = done || hit;
                    }
                    const int cidx = cand0 + c + pos;
                    if ((int64_t)cidx != qrow) {
                        reglist_insert(L, m, cidx);
                        tau = LISTED ? fmaxf(tau0, L.t[LIST - 1]) : L.t[LIST - 1];
                    }
                    m = __uint_as_float(v[0]);
#pragma unroll
                    for (int i = 1; i < 32; i++) m = fmaxf(m, __uint_as_float(v[i]));
                }
            };
            tmem_ld32(col0, va);
#pragma unroll 1
            for (int half = 0; half < TN / 64; half++) {  // rolled: two copies of the scan code, not four
                tmem_ld_wait();
                tmem_ld32(col0 + 64 * half + 32, vb);
                scan(va, 64 * half);
                tmem_ld_wait();
                if (half + 1 < TN / 64) tmem_ld32(col0 + 64 * (half + 1), va);
                scan(vb, 64 * half + 32);
            }
            fence_before();
            mbar_arrive(tempty + buf);
        }
        if (active && (LISTED || my_tiles > 0)) {  // a listed pair with an empty list still owns its rows: all "none"
#pragma unroll
            for (int l = 0; l < LIST; l++) cand_i[qrow * LIST + l] = L.i[l];
            if constexpr (LISTED)
                if (tau_out) tau_out[qrow] = fmaxf(tau0, L.t[LIST - 1]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc

template <int KP, int TL>
int run_knn(dd_handle *h, int k, float *norms, float *cand_d, int *cand_i) {
    const int64_t n = h->emb_rows;
    using S = KnnSmem<KP, TL>;
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(k_knn_scan<KP, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S));
    });
    DD_LAUNCH(h, "knn_norms", k_row_norms<KP>, (unsigned)((n + 255) / 256), 256, 0, h->d_emb, n, norms);
    DD_LAUNCH(h, "knn_scan", (k_knn_scan<KP, TL>), (unsigned)((n + BQ - 1) / BQ), 256, sizeof(S), h->d_emb, norms, n,
              cand_d, cand_i);
    DD_LAUNCH(h, "knn_refine", (k_knn_refine<KP, TL>), (unsigned)((n + 7) / 8), 256, 0, h->d_emb, cand_i, (int64_t)0, n, k,
              h->d_knn_idx, h->d_knn_dist);
    return DD_OK;
}

}  // namespace

// Final re-ranking of a kNN call: exact float64 order of the filter's candidates (`width` per row: n_lists lists of list_w),
// the filter's certificate, and the float64 brute-force fix-up of the rows it could not clear.  Rows [q0, q1).
int dd_knn_refine_final(dd_handle *h, const float *emb, const int *cand_i, int width, int list_w, int n_lists, int64_t q0,
                        int64_t q1, int64_t n, int k, const int32_t *row_map, int shard_world, int shard_rank) {
    DD_TRY(dd_reserve(h, &h->d_knn_cert, &h->cap_knn_cert, n + 4));
    DD_CUDA(h, cudaMemsetAsync(h->d_knn_cert, 0, sizeof(int32_t), h->stream));
    KnnCert cert;
    cert.list_w = list_w;
    cert.n_lists = n_lists;
    cert.count = h->d_knn_cert;
    cert.rows = h->d_knn_cert + 4;
    cert.cap = (int)n;
    const unsigned grid = (unsigned)((q1 - q0 + 7) / 8);
    if (width <= 32)
        DD_LAUNCH(h, "knn_refine", k_knn_refine_w<1>, grid, 256, 0, emb, cand_i, width, q0, q1, k, h->d_knn_idx, h->d_knn_dist, cert,
                  row_map, shard_world, shard_rank);
    else if (width <= 64)
        DD_LAUNCH(h, "knn_refine", k_knn_refine_w<2>, grid, 256, 0, emb, cand_i, width, q0, q1, k, h->d_knn_idx, h->d_knn_dist, cert,
                  row_map, shard_world, shard_rank);
    else
        DD_LAUNCH(h, "knn_refine", k_knn_refine_w<3>, grid, 256, 0, emb, cand_i, width, q0, q1, k, h->d_knn_idx, h->d_knn_dist, cert,
                  row_map, shard_world, shard_rank);
    DD_LAUNCH(h, "knn_exact_rows", k_knn_exact_rows, (unsigned)(h->num_sms * 2), 256, 0, emb, n, (const int *)cert.rows,
              (const int *)cert.count, cert.cap, k, h->d_knn_idx, h->d_knn_dist);
    if (!h->h_knn_uncert && cudaMallocHost(&h->h_knn_uncert, sizeof(int32_t)) != cudaSuccess) h->h_knn_uncert = nullptr;
    if (h->h_knn_uncert)
        DD_CUDA(h, cudaMemcpyAsync(h->h_knn_uncert, h->d_knn_cert, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    return DD_OK;
}

namespace {

int run_knn_tc(dd_handle *h, int k, int TL, float *cand_t, int *cand_i) {
    const int64_t n = h->emb_rows;
    const int n_tiles = (int)((n + tc::TILE - 1) / tc::TILE);
    const int n_tiles_pad = (n_tiles + tc::QT - 1) / tc::QT * tc::QT;
    const int64_t n_pad = (int64_t)n_tiles_pad * tc::TILE;
    const int64_t op_bytes = (int64_t)n_tiles_pad * tc::TILE_BYTES;
    DD_TRY(dd_reserve(h, &h->d_knn_ops, &h->cap_knn_ops, 2 * op_bytes));
    uint8_t *qa = h->d_knn_ops, *cb = h->d_knn_ops + op_bytes;
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(tc::k_knn_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
        cudaFuncSetAttribute(tc::k_knn_tc<40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
        cudaFuncSetAttribute(tc::k_knn_tc<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
        cudaFuncSetAttribute(tc::k_knn_tc<40, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
        cudaFuncSetAttribute(tc::k_knn_tc<16, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<64>());
        cudaFuncSetAttribute(tc::k_knn_tc<40, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::smem_bytes<64>());
    });
    DD_LAUNCH(h, "knn_prep", tc::k_knn_prep, (unsigned)(n_pad / 8), 112, 0, h->d_emb, n, n_pad,
              reinterpret_cast<uint4 *>(qa), reinterpret_cast<uint4 *>(cb));
    // Cell-block sharding: every rank holds the whole (all-gathered) embedding and answers the queries of its
    // share of the 256-row query blocks against ALL candidates; the lists are then all-gathered.
    const int n_pairs = n_tiles_pad / tc::QT;
    const int W = dd_sharded(h) ? h->world : 1, R = dd_sharded(h) ? h->rank : 0;
    const int pair0 = (int)((int64_t)n_pairs * R / W), pair1 = (int)((int64_t)n_pairs * (R + 1) / W);
    const int64_t q0 = std::min<int64_t>(n, (int64_t)pair0 * tc::QT * tc::TILE);
    const int64_t q1 = std::min<int64_t>(n, (int64_t)pair1 * tc::QT * tc::TILE);
    if (h->knn_list_pairs > 0) {  // experimental: per-pair candidate-tile lists (dd_knn_listed)
        if (W > 1 || h->knn_list_pairs != n_pairs)
            return dd_fail(h, DD_ERR_ARG, "knn: candidate lists need an unsharded handle and one list per 256-row query block");
        if (TL == 16)
            DD_LAUNCH(h, "knn_tc_listed", (tc::k_knn_tc<16, true>), (unsigned)n_pairs, tc::THREADS, tc::SMEM_BYTES, qa, cb, n, n_tiles,
                      0, n_pairs, cand_i, h->d_knn_list_off, h->d_knn_list_tiles);
        else
            DD_LAUNCH(h, "knn_tc_listed", (tc::k_knn_tc<40, true>), (unsigned)n_pairs, tc::THREADS, tc::SMEM_BYTES, qa, cb, n, n_tiles,
                      0, n_pairs, cand_i, h->d_knn_list_off, h->d_knn_list_tiles);
        // (test hook with caller-made lists: candidates outside the lists are not the filter's doing -- no certificate)
        if (TL == 16)
            DD_LAUNCH(h, "knn_refine", (k_knn_refine<32, 16>), (unsigned)((n + 7) / 8), 256, 0, h->d_emb, cand_i, (int64_t)0, n, k,
                      h->d_knn_idx, h->d_knn_dist);
        else
            DD_LAUNCH(h, "knn_refine", k_knn_refine_w<2>, (unsigned)((n + 7) / 8), 256, 0, h->d_emb, cand_i, 40, (int64_t)0, n, k,
                      h->d_knn_idx, h->d_knn_dist);
        return DD_OK;
    }
    if (pair1 > pair0) {
        // whole waves of pair CTAs, then the remainder as single-tile CTAs (twice as many, half as long)
        const int pairs = pair1 - pair0;
        static const bool split_tail = getenv("DD_KNN_NO_TAIL_SPLIT") == nullptr;
        const int n_full = split_tail ? pairs / h->num_sms * h->num_sms : pairs;
        const unsigned grid = (unsigned)(n_full + (pairs - n_full) * tc::QT);
        if (h->knn_narrow && TL == 16) {
            DD_LAUNCH(h, "knn_tc", (tc::k_knn_tc<16, false, 64>), grid, tc::THREADS, tc::smem_bytes<64>(), qa, cb, n, n_tiles, pair0,
                      n_full, cand_i, (const int *)nullptr, (const int *)nullptr);
            DD_TRY(dd_knn_refine_final(h, h->d_emb, cand_i, 16, 16, 1, q0, q1, n, k, nullptr, 1, 0));
        } else if (h->knn_narrow) {
            DD_LAUNCH(h, "knn_tc", (tc::k_knn_tc<40, false, 64>), grid, tc::THREADS, tc::smem_bytes<64>(), qa, cb, n, n_tiles, pair0,
                      n_full, cand_i, (const int *)nullptr, (const int *)nullptr);
            DD_TRY(dd_knn_refine_final(h, h->d_emb, cand_i, 40, 40, 1, q0, q1, n, k, nullptr, 1, 0));
        } else if (TL == 16) {
            DD_LAUNCH(h, "knn_tc", tc::k_knn_tc<16>, grid, tc::THREADS, tc::SMEM_BYTES, qa, cb, n, n_tiles, pair0, n_full, cand_i);
            DD_TRY(dd_knn_refine_final(h, h->d_emb, cand_i, 16, 16, 1, q0, q1, n, k, nullptr, 1, 0));
        } else {
            DD_LAUNCH(h, "knn_tc", tc::k_knn_tc<40>, grid, tc::THREADS, tc::SMEM_BYTES, qa, cb, n, n_tiles, pair0, n_full, cand_i);
            DD_TRY(dd_knn_refine_final(h, h->d_emb, cand_i, 40, 40, 1, q0, q1, n, k, nullptr, 1, 0));
        }
    }
    if (W > 1) {
        std::vector<int64_t> begin(W), count(W);
        std::vector<int> owner(W);
        for (int r = 0; r < W; r++) {
            const int64_t b = std::min<int64_t>(n, (int64_t)n_pairs * r / W * tc::QT * tc::TILE);
            const int64_t e = std::min<int64_t>(n, (int64_t)n_pairs * (r + 1) / W * tc::QT * tc::TILE);
            begin[r] = b; count[r] = e - b; owner[r] = r;
        }
        DD_TRY(dd_comm_gather_ranges(h, h->d_knn_idx, (int64_t)sizeof(int32_t) * k, W, begin.data(), count.data(), owner.data()));
        DD_TRY(dd_comm_gather_ranges(h, h->d_knn_dist, (int64_t)sizeof(float) * k, W, begin.data(), count.data(), owner.data()));
    }
    return DD_OK;
}

}  // namespace

// scratch layout inside d_knn_dist's allocation: [n*k dist][n norms][n*TL cand_d][n*TL cand_i]
int dd_dev_knn(dd_handle *h, int32_t k) {
    if (!h->emb_valid) return dd_fail(h, DD_ERR_ARG, "knn: no embedding (call dd_pca first)");
    const int64_t n = h->emb_rows;
    if (k < 2 || k > 31) return dd_fail(h, DD_ERR_UNSUPPORTED, "knn: k must be in [2, 31]");
    if (k > n) return dd_fail(h, DD_ERR_ARG, "knn: k exceeds the number of rows");
    if (n >= (1ll << 31) - 1) return dd_fail(h, DD_ERR_UNSUPPORTED, "knn: too many rows for int32 indices");
    h->knn_last_k = k;
    const int TL = (k - 1 <= 12) ? 16 : 32;
    const int64_t n_padded = (n + 255) / 256 * 256;  // the tensor-core path keeps lists for whole 256-row CTAs
    // candidate lists: 32 per row on the FFMA fallback, up to 40 per row on the tcgen05 path (k - 1 > 12: lists of 40)
    const int64_t need = n * k + n + n_padded * 32 + n_padded * 40;
    if (need > h->cap_knn) {
        if (h->d_knn_idx_base) cudaFree(h->d_knn_idx_base);
        if (h->d_knn_dist) cudaFree(h->d_knn_dist);
        h->d_knn_idx_base = h->d_knn_idx = nullptr; h->d_knn_dist = nullptr; h->cap_knn = 0;
        // two list buffers: the clustering stream still reads iteration i's lists while iteration i + 1 writes its own
        DD_CUDA(h, cudaMalloc(&h->d_knn_idx_base, sizeof(int32_t) * 2 * n * 32));
        DD_CUDA(h, cudaMalloc(&h->d_knn_dist, sizeof(float) * need));
        h->cap_knn = need;
        h->knn_idx_stride = n * 32;
        h->d_knn_idx = h->d_knn_idx_base;
    }
    float *norms = h->d_knn_dist + n * k;
    float *cand_d = norms + n;
    int *cand_i = reinterpret_cast<int *>(cand_d + n_padded * 32);
    // default: tcgen05 distance GEMM; DD_KNN_FFMA=1 keeps the CUDA-core kernel (A/B comparison, KP=64, k>13)
    static const bool force_ffma = getenv("DD_KNN_FFMA") != nullptr;
    // large embeddings: cluster-ordered candidate tiles (knn_prune.cu) -- the same exact result from a fraction of the tile pairs
    if (!force_ffma && h->knn_list_pairs == 0 && dd_knn_clustered_applies(h, k)) return dd_dev_knn_clustered(h, k);
    if (h->KP == 32 && !force_ffma) return run_knn_tc(h, k, TL, cand_d, cand_i);
    if (dd_sharded(h)) return dd_fail(h, DD_ERR_UNSUPPORTED, "knn: cell-block sharding needs the tcgen05 path (<= 32 components)");
    if (h->KP == 32)
        return TL == 16 ? run_knn<32, 16>(h, k, norms, cand_d, cand_i) : run_knn<32, 32>(h, k, norms, cand_d, cand_i);
    return TL == 16 ? run_knn<64, 16>(h, k, norms, cand_d, cand_i) : run_knn<64, 32>(h, k, norms, cand_d, cand_i);
}

extern "C" int dd_knn(dd_handle *h, int32_t k, int32_t *idx_out, float *dist_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_knn: null handle");
    if (!idx_out) return dd_fail(h, DD_ERR_ARG, "dd_knn: null output");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_knn(h, k));
    DD_TRY(dd_stage_end(h, "knn"));
    const int64_t n = h->emb_rows;
    DD_CUDA(h, cudaMemcpyAsync(idx_out, h->d_knn_idx, sizeof(int32_t) * n * k, cudaMemcpyDeviceToHost, h->stream));
    if (dist_out)
        DD_CUDA(h, cudaMemcpyAsync(dist_out, h->d_knn_dist, sizeof(float) * n * k, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

// How many rows of the last dd_knn / fit-loop kNN on this handle the filter's certificate could not clear (they were re-done by
// float64 brute force: the result is exact either way; a large number means the embedding sits far from the origin
// relative to its local spacing and the kNN runs at brute-force speed).
extern "C" int dd_knn_uncertified(dd_handle *h, int64_t *count_out) {
    if (!h || !count_out) return dd_fail(h, DD_ERR_ARG, "dd_knn_uncertified: null argument");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    *count_out = h->h_knn_uncert ? (int64_t)*h->h_knn_uncert : 0;
    return DD_OK;
}

// Test hook of the list-driven kernel: exact kNN in which the 256-row query block p only visits the
// candidate tiles (128 rows each) list_tiles[list_off[p] .. list_off[p + 1]).  The caller is responsible for the lists being
// sufficient (scripts/knn_listed_experiment.py derives them from bounding boxes); rows whose lists hold fewer than k - 1
// other points get -1 entries.  Same outputs as dd_knn.
extern "C" int dd_knn_listed(dd_handle *h, int32_t k, int64_t n_blocks, const int32_t *list_off, const int32_t *list_tiles,
                             int32_t *idx_out, float *dist_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_knn_listed: null handle");
    if (!idx_out || !list_off || n_blocks < 1) return dd_fail(h, DD_ERR_ARG, "dd_knn_listed: null argument");
    if (!h->emb_valid || h->KP != 32) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_knn_listed: needs an embedding of <= 32 components");
    const int64_t n = h->emb_rows;
    const int64_t n_tiles = (n + 127) / 128;
    if (n_blocks != (n_tiles + 1) / 2) return dd_fail(h, DD_ERR_ARG, "dd_knn_listed: one list per 256-row query block");
    const int64_t total = list_off[n_blocks];
    if (list_off[0] != 0 || total < 0 || (total > 0 && !list_tiles)) return dd_fail(h, DD_ERR_ARG, "dd_knn_listed: bad list offsets");
    for (int64_t p = 0; p < n_blocks; p++)
        if (list_off[p + 1] < list_off[p]) return dd_fail(h, DD_ERR_ARG, "dd_knn_listed: bad list offsets");
    for (int64_t e = 0; e < total; e++)
        if (list_tiles[e] < 0 || list_tiles[e] >= n_tiles) return dd_fail(h, DD_ERR_ARG, "dd_knn_listed: tile index out of range");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_reserve(h, &h->d_knn_list_off, &h->cap_knn_list_off, n_blocks + 1));
    DD_TRY(dd_reserve(h, &h->d_knn_list_tiles, &h->cap_knn_list_tiles, std::max<int64_t>(total, 1)));
    DD_CUDA(h, cudaMemcpyAsync(h->d_knn_list_off, list_off, sizeof(int32_t) * (n_blocks + 1), cudaMemcpyHostToDevice, h->stream));
    if (total > 0)
        DD_CUDA(h, cudaMemcpyAsync(h->d_knn_list_tiles, list_tiles, sizeof(int32_t) * total, cudaMemcpyHostToDevice, h->stream));
    h->knn_list_pairs = (int)n_blocks;
    int rc = dd_stage_begin(h);
    if (rc == DD_OK) rc = dd_dev_knn(h, k);
    if (rc == DD_OK) rc = dd_stage_end(h, "knn");
    h->knn_list_pairs = 0;
    DD_TRY(rc);
    DD_CUDA(h, cudaMemcpyAsync(idx_out, h->d_knn_idx, sizeof(int32_t) * n * k, cudaMemcpyDeviceToHost, h->stream));
    if (dist_out)
        DD_CUDA(h, cudaMemcpyAsync(dist_out, h->d_knn_dist, sizeof(float) * n * k, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

// Launchers for knn_prune.cu (the experimental cluster-ordered kNN lives in its own translation unit; the measured kernels of
// this file keep their instruction sequence): operand tiles, the list-driven kernel (lists of 16), the exact re-ranking.
int dd_knn_launch_prep(dd_handle *h, const float *emb, int64_t n, int64_t n_pad, uint8_t *qa, uint8_t *cb) {
    DD_LAUNCH(h, "knn_prep", tc::k_knn_prep, (unsigned)(n_pad / 8), 112, 0, emb, n, n_pad, reinterpret_cast<uint4 *>(qa),
              reinterpret_cast<uint4 *>(cb));
    return DD_OK;
}

int dd_knn_launch_listed16(dd_handle *h, const uint8_t *qa, const uint8_t *cb, int64_t n, int n_tiles, int n_blocks, int *cand_i,
                           const int *list_off, const int *list_tiles, const int *list_len, const int *block_order,
                           const float *tau_init, float *tau_out, int shard_world, int shard_rank) {
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(tc::k_knn_tc<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    });
    DD_LAUNCH(h, "knn_tc_listed", (tc::k_knn_tc<16, true>), (unsigned)n_blocks, tc::THREADS, tc::SMEM_BYTES, qa, cb, n, n_tiles, 0,
              n_blocks, cand_i, list_off, list_tiles, list_len, block_order, tau_init, tau_out, shard_world, shard_rank);
    return DD_OK;
}

// the same for lists of 40 (k - 1 > 12: PhenoGraph's 30 neighbours keep a margin of 10 filter ranks)
int dd_knn_launch_listed40(dd_handle *h, const uint8_t *qa, const uint8_t *cb, int64_t n, int n_tiles, int n_blocks, int *cand_i,
                           const int *list_off, const int *list_tiles, const int *list_len, const int *block_order,
                           const float *tau_init, float *tau_out, int shard_world, int shard_rank) {
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(tc::k_knn_tc<40, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    });
    DD_LAUNCH(h, "knn_tc_listed", (tc::k_knn_tc<40, true>), (unsigned)n_blocks, tc::THREADS, tc::SMEM_BYTES, qa, cb, n, n_tiles, 0,
              n_blocks, cand_i, list_off, list_tiles, list_len, block_order, tau_init, tau_out, shard_world, shard_rank);
    return DD_OK;
}

// re-ranking of ONE launch's lists (16 or 40 wide) over the permuted rows, for the blocks this rank owns
int dd_knn_launch_refine_lists(dd_handle *h, const float *emb, const int *cand_i, int width, int64_t n, int k, int32_t *idx_out,
                               float *dist_out, int shard_world, int shard_rank) {
    const unsigned grid = (unsigned)((n + 7) / 8);
    if (width <= 32)
        DD_LAUNCH(h, "knn_refine", k_knn_refine_w<1>, grid, 256, 0, emb, cand_i, width, (int64_t)0, n, k, idx_out, dist_out, KnnCert(),
                  (const int32_t *)nullptr, shard_world, shard_rank);
    else
        DD_LAUNCH(h, "knn_refine", k_knn_refine_w<2>, grid, 256, 0, emb, cand_i, width, (int64_t)0, n, k, idx_out, dist_out, KnnCert(),
                  (const int32_t *)nullptr, shard_world, shard_rank);
    return DD_OK;
}

// re-ranking launcher for knn_prune.cu: lists of 40 (one launch's)
int dd_knn_launch_refine40(dd_handle *h, const float *emb, const int *cand_i, int64_t n, int k, int32_t *idx_out, float *dist_out) {
    DD_LAUNCH(h, "knn_refine", k_knn_refine_w<2>, (unsigned)((n + 7) / 8), 256, 0, emb, cand_i, 40, (int64_t)0, n, k, idx_out, dist_out);
    return DD_OK;
}

int dd_knn_launch_refine32(dd_handle *h, const float *emb, const int *cand_i, int64_t n, int k, int32_t *idx_out, float *dist_out) {
    DD_LAUNCH(h, "knn_refine", (k_knn_refine<32, 32>), (unsigned)((n + 7) / 8), 256, 0, emb, cand_i, (int64_t)0, n, k, idx_out,
              dist_out);
    return DD_OK;
}

int dd_knn_launch_refine16(dd_handle *h, const float *emb, const int *cand_i, int64_t n, int k, int32_t *idx_out, float *dist_out) {
    DD_LAUNCH(h, "knn_refine", (k_knn_refine<32, 16>), (unsigned)((n + 7) / 8), 256, 0, emb, cand_i, (int64_t)0, n, k, idx_out,
              dist_out);
    return DD_OK;
}
