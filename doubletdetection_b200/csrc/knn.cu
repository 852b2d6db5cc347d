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

#include <cfloat>

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
__global__ void k_knn_refine(const float *__restrict__ emb, const int *__restrict__ cand_i, int64_t n, int k,
                             int32_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
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

template <int KP, int TL>
int run_knn(dd_handle *h, int k, float *norms, float *cand_d, int *cand_i) {
    const int64_t n = h->emb_rows;
    using S = KnnSmem<KP, TL>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_knn_scan<KP, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S));
        attr_set = true;
    }
    DD_LAUNCH(h, "knn_norms", k_row_norms<KP>, (unsigned)((n + 255) / 256), 256, 0, h->d_emb, n, norms);
    DD_LAUNCH(h, "knn_scan", (k_knn_scan<KP, TL>), (unsigned)((n + BQ - 1) / BQ), 256, sizeof(S), h->d_emb, norms, n,
              cand_d, cand_i);
    DD_LAUNCH(h, "knn_refine", (k_knn_refine<KP, TL>), (unsigned)((n + 7) / 8), 256, 0, h->d_emb, cand_i, n, k,
              h->d_knn_idx, h->d_knn_dist);
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
    const int TL = (k - 1 <= 12) ? 16 : 32;
    const int64_t need = n * k + n + 2 * n * 32;
    if (need > h->cap_knn) {
        if (h->d_knn_idx) cudaFree(h->d_knn_idx);
        if (h->d_knn_dist) cudaFree(h->d_knn_dist);
        h->d_knn_idx = nullptr; h->d_knn_dist = nullptr; h->cap_knn = 0;
        DD_CUDA(h, cudaMalloc(&h->d_knn_idx, sizeof(int32_t) * n * 32));
        DD_CUDA(h, cudaMalloc(&h->d_knn_dist, sizeof(float) * need));
        h->cap_knn = need;
    }
    float *norms = h->d_knn_dist + n * k;
    float *cand_d = norms + n;
    int *cand_i = reinterpret_cast<int *>(cand_d + n * 32);
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
