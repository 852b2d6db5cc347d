// knn_prune.cu -- the exact kNN with cluster-ordered candidate tiles, the fit loop's default from
// 50 000 rows: ordering pre-pass (k-means groups, axis slabs), bounding boxes, tile lists and both list-driven launches on
// the device; 2.3 ms instead of 4.9 ms at 125 k rows with identical output (tests/test_gpu_knn_clustered.py,
// profiles/r2k_knn_clustered_vs_dense.log).  Its own translation unit: knn.cu holds the tcgen05 kernels and their launchers.
#include "dd_internal.h"

#include <cmath>
#include <cstdlib>
#include <vector>

int dd_knn_launch_prep(dd_handle *h, const float *emb, int64_t n, int64_t n_pad, uint8_t *qa, uint8_t *cb);        // knn.cu
int dd_knn_launch_listed16(dd_handle *h, const uint8_t *qa, const uint8_t *cb, int64_t n, int n_tiles, int n_blocks, int *cand_i,
                           const int *list_off, const int *list_tiles, const int *list_len, const int *block_order,
                           const float *tau_init, float *tau_out, int shard_world, int shard_rank);                  // knn.cu
int dd_knn_launch_refine_lists(dd_handle *h, const float *emb, const int *cand_i, int width, int64_t n, int k, int32_t *idx_out,
                               float *dist_out, int shard_world, int shard_rank);                                   // knn.cu
int dd_knn_launch_refine32(dd_handle *h, const float *emb, const int *cand_i, int64_t n, int k, int32_t *idx_out,
                           float *dist_out);                                                                         // knn.cu
int dd_knn_launch_listed40(dd_handle *h, const uint8_t *qa, const uint8_t *cb, int64_t n, int n_tiles, int n_blocks, int *cand_i,
                           const int *list_off, const int *list_tiles, const int *list_len, const int *block_order,
                           const float *tau_init, float *tau_out, int shard_world, int shard_rank);                  // knn.cu
int dd_knn_launch_refine40(dd_handle *h, const float *emb, const int *cand_i, int64_t n, int k, int32_t *idx_out,
                           float *dist_out);                                                                         // knn.cu
int dd_knn_refine_final(dd_handle *h, const float *emb, const int *cand_i, int width, int list_w, int n_lists, int64_t q0,
                        int64_t q1, int64_t n, int k, const int32_t *row_map, int shard_world, int shard_rank);      // knn.cu
int dd_knn_launch_refine16(dd_handle *h, const float *emb, const int *cand_i, int64_t n, int k, int32_t *idx_out,
                           float *dist_out);                                                                         // knn.cu

namespace tc {
constexpr int TILE = 128, QT = 2, TILE_BYTES = TILE * 14 * 16;  // operand tile geometry of knn.cu
}

// The caller supplies a PADDED PERMUTATION of the rows (position -> original
// row or -1; groups of spatially close rows are contiguous and padded to whole 256-row blocks) and the group of every
// block.  Two launches of the list-driven tcgen05 kernel:
//   A  every block against the tiles of its own group        -> an upper bound on each query's k-th distance
//   B  every block against the tiles whose bounding box can still hold something closer than its largest bound
// and the usual float64 re-ranking in the ORIGINAL numbering (so ties break exactly as in the dense kNN).  Valid lower
// bounds only: the result equals dd_knn's.  scripts/knn_listed_experiment.py holds the host reference of every step.
namespace {

constexpr float kPadCoord = 1.0e12f;  // a padding row sits here in dimension 0: its score is ~ -5e23, never selected

__global__ void k_prune_gather(const float *__restrict__ emb, const int32_t *__restrict__ perm, int64_t n_pad,
                               float *__restrict__ emb_p) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (row, 4 columns)
    const int64_t r = t >> 3;
    const int c4 = (int)(t & 7) * 4;
    if (r >= n_pad) return;
    const int src = perm[r];
    float4 v = make_float4(c4 == 0 ? kPadCoord : 0.f, 0.f, 0.f, 0.f);
    if (src >= 0) v = *reinterpret_cast<const float4 *>(emb + (int64_t)src * 32 + c4);
    *reinterpret_cast<float4 *>(emb_p + r * 32 + c4) = v;
}

// one warp per 128-row tile, lane = dimension: bounding box over the real rows of the tile
__global__ void k_prune_boxes(const float *__restrict__ emb_p, const int32_t *__restrict__ perm, int n_tiles,
                              float *__restrict__ lo, float *__restrict__ hi, int32_t *__restrict__ tile_rows,
                              float *__restrict__ lo_t = nullptr, float *__restrict__ hi_t = nullptr) {
    const int lane = threadIdx.x & 31;
    const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= n_tiles) return;
    float mn = 3.0e38f, mx = -3.0e38f;
    int rows = 0;
    for (int r = 0; r < tc::TILE; r++) {
        const int64_t row = (int64_t)t * tc::TILE + r;
        if (perm[row] < 0) continue;  // uniform across the warp
        const float x = emb_p[row * 32 + lane];
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
        rows++;
    }
    lo[(int64_t)t * 32 + lane] = mn;
    hi[(int64_t)t * 32 + lane] = mx;
    if (lo_t) {  // dimension-major copies: one thread per tile reads them coalesced (k_lists_other)
        lo_t[(int64_t)lane * n_tiles + t] = mn;
        hi_t[(int64_t)lane * n_tiles + t] = mx;
    }
    if (lane == 0) tile_rows[t] = rows;
}

// one CTA (256 threads) per 256-row block: the largest k-th distance^2 launch A found for its real rows (inf if a row
// found fewer than k - 1 neighbours in its own group)
__global__ void __launch_bounds__(256) k_prune_threshold(const int32_t *__restrict__ perm, const int32_t *__restrict__ idx_a,
                                                         const float *__restrict__ dist_a, int k, double *__restrict__ thr,
                                                         int shard_world = 1, int shard_rank = 0) {
    __shared__ double s_max[8];
    if (shard_world > 1 && ((int)blockIdx.x % shard_world) != shard_rank) {  // another rank's block (uniform)
        if (threadIdx.x == 0) thr[blockIdx.x] = 0.0;
        return;
    }
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double v = 0.0;
    if (perm[row] >= 0) {
        const double d = (double)dist_a[row * k + k - 1];
        v = idx_a[row * k + k - 1] >= 0 ? d * d : INFINITY;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = s_max[0];
        for (int w = 1; w < 8; w++) m = fmax(m, s_max[w]);
        thr[blockIdx.x] = m;
    }
}

// one CTA (256 threads) per block: the tiles whose box-to-box distance^2 to the block's box is within its threshold.
// First pass (off == nullptr): len[b] = how many; second pass: the tiles themselves, compacted at list[off[b] ...].
__global__ void __launch_bounds__(256) k_prune_lists(const float *__restrict__ lo, const float *__restrict__ hi,
                                                     const int32_t *__restrict__ tile_rows, const double *__restrict__ thr,
                                                     int n_tiles, const int32_t *__restrict__ off,
                                                     int32_t *__restrict__ list, int32_t *__restrict__ len) {
    __shared__ float q_lo[32], q_hi[32];
    __shared__ int s_warp[8], s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;
    if (tid < 32) {
        float mn = 3.0e38f, mx = -3.0e38f;
        for (int h2 = 0; h2 < tc::QT; h2++) {
            const int t = b * tc::QT + h2;
            if (t < n_tiles && tile_rows[t] > 0) {
                mn = fminf(mn, lo[(int64_t)t * 32 + tid]);
                mx = fmaxf(mx, hi[(int64_t)t * 32 + tid]);
            }
        }
        q_lo[tid] = mn;
        q_hi[tid] = mx;
    }
    if (tid == 0) s_base = 0;
    __syncthreads();
    const double limit = thr[b] * (1.0 + 1.0e-5);
    for (int t0 = 0; t0 < n_tiles; t0 += 256) {
        const int t = t0 + tid;
        bool need = false;
        if (t < n_tiles && tile_rows[t] > 0) {
            double lb = 0.0;
            for (int c = 0; c < 32; c++) {
                const float g = fmaxf(0.f, fmaxf(lo[(int64_t)t * 32 + c] - q_hi[c], q_lo[c] - hi[(int64_t)t * 32 + c]));
                lb += (double)g * (double)g;
            }
            need = lb <= limit;
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (lane == 0) s_warp[wl] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < wl; w++) before += s_warp[w];
        if (need && off) list[off[b] + before + __popc(m & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < 8; w++) total += s_warp[w];
            s_base += total;
        }
        __syncthreads();
    }
    if (tid == 0 && !off) len[b] = s_base;
}

__global__ void k_prune_offsets(const int32_t *__restrict__ len, int n_blocks, int32_t *__restrict__ off) {
    if (blockIdx.x || threadIdx.x) return;
    int run = 0;
    for (int b = 0; b < n_blocks; b++) {
        off[b] = run;
        run += len[b];
    }
    off[n_blocks] = run;
}

// candidate lists of launch B (permuted numbering, one row per permuted position) -> original numbering and row order
__global__ void k_prune_translate(const int32_t *__restrict__ perm, const int *__restrict__ cand_p, int64_t n_pad,
                                  int *__restrict__ cand_o, int out_stride = 16, int out_col0 = 0, int list_w = 16,
                                  int shard_world = 1, int shard_rank = 0) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / list_w;
    const int l = (int)(t % list_w);
    if (r >= n_pad) return;
    if (shard_world > 1 && ((r >> 8) % shard_world) != shard_rank) return;  // another rank's block
    const int o = perm[r];
    if (o < 0) return;
    const int c = cand_p[r * list_w + l];
    int out = 0x7fffffff;
    if (c != 0x7fffffff && c >= 0 && c < n_pad) {
        const int oc = perm[c];
        if (oc >= 0) out = oc;
    }
    cand_o[(int64_t)o * out_stride + out_col0 + l] = out;
}

template <typename T>
struct PruneBuf {
    T *p = nullptr;
    ~PruneBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, sizeof(T) * (count ? count : 1)); }
};

}  // namespace

extern "C" int dd_knn_pruned(dd_handle *h, int32_t k, int64_t n_pad, const int32_t *perm, const int32_t *block_group,
                             int32_t *idx_out, float *dist_out, int64_t *stats_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_knn_pruned: null handle");
    if (!perm || !block_group || !idx_out) return dd_fail(h, DD_ERR_ARG, "dd_knn_pruned: null argument");
    if (!h->emb_valid || h->KP != 32) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_knn_pruned: needs an embedding of <= 32 components");
    if (k < 2 || k > 13) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_knn_pruned: k must be in [2, 13]");
    h->knn_last_k = k;
    const int64_t n = h->emb_rows;
    if (n_pad < n || n_pad % 256 != 0 || n_pad >= (1ll << 31) - 1) return dd_fail(h, DD_ERR_ARG, "dd_knn_pruned: bad padded size");
    const int n_blocks = (int)(n_pad / 256), n_tiles = (int)(n_pad / tc::TILE);
    if ((int64_t)n_blocks * n_tiles >= (1ll << 31) - 1) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_knn_pruned: list table too large");
    // the permutation must hit every original row exactly once
    {
        std::vector<uint8_t> seen((size_t)n, 0);
        int64_t real = 0;
        for (int64_t r = 0; r < n_pad; r++) {
            const int o = perm[r];
            if (o < 0) continue;
            if (o >= n || seen[o]) return dd_fail(h, DD_ERR_ARG, "dd_knn_pruned: perm is not a padded permutation of the rows");
            seen[o] = 1;
            real++;
        }
        if (real != n) return dd_fail(h, DD_ERR_ARG, "dd_knn_pruned: perm misses rows");
    }
    // launch A lists on the host: the non-empty tiles of the block's own group (groups are contiguous runs of blocks)
    std::vector<int32_t> off_a((size_t)n_blocks + 1, 0), tiles_a;
    {
        std::vector<uint8_t> tile_real((size_t)n_tiles, 0);
        for (int64_t r = 0; r < n_pad; r++)
            if (perm[r] >= 0) tile_real[r / tc::TILE] = 1;
        int b0 = 0;
        while (b0 < n_blocks) {
            int b1 = b0;
            while (b1 < n_blocks && block_group[b1] == block_group[b0]) b1++;
            for (int b = b0; b < b1; b++) {
                for (int t = b0 * tc::QT; t < b1 * tc::QT; t++)
                    if (tile_real[t]) tiles_a.push_back(t);
                off_a[b + 1] = (int32_t)tiles_a.size();
            }
            b0 = b1;
        }
    }
    DD_CUDA(h, cudaSetDevice(h->device));
    PruneBuf<int32_t> d_perm, d_tile_rows, d_off_a, d_tiles_a, d_off_b, d_list_b, d_len_b, d_idx_a;
    PruneBuf<int> d_cand_p, d_cand_o;
    PruneBuf<float> d_emb_p, d_lo, d_hi, d_dist_a;
    PruneBuf<double> d_thr;
    if (d_perm.alloc(n_pad) || d_tile_rows.alloc(n_tiles) || d_off_a.alloc(n_blocks + 1) || d_tiles_a.alloc(tiles_a.size()) ||
        d_off_b.alloc(n_blocks + 1) || d_list_b.alloc((size_t)n_blocks * n_tiles) || d_len_b.alloc(n_blocks) ||
        d_idx_a.alloc((size_t)n_pad * k) || d_cand_p.alloc((size_t)n_pad * 16) || d_cand_o.alloc((size_t)n * 16) ||
        d_emb_p.alloc((size_t)n_pad * 32) || d_lo.alloc((size_t)n_tiles * 32) || d_hi.alloc((size_t)n_tiles * 32) ||
        d_dist_a.alloc((size_t)n_pad * k) || d_thr.alloc(n_blocks))
        return dd_fail(h, DD_ERR_NOMEM, "dd_knn_pruned: device buffers");
    DD_CUDA(h, cudaMemcpyAsync(d_perm.p, perm, sizeof(int32_t) * n_pad, cudaMemcpyHostToDevice, h->stream));
    DD_CUDA(h, cudaMemcpyAsync(d_off_a.p, off_a.data(), sizeof(int32_t) * (n_blocks + 1), cudaMemcpyHostToDevice, h->stream));
    if (!tiles_a.empty())
        DD_CUDA(h, cudaMemcpyAsync(d_tiles_a.p, tiles_a.data(), sizeof(int32_t) * tiles_a.size(), cudaMemcpyHostToDevice, h->stream));
    // a row whose own group holds fewer than k - 1 other rows leaves list positions unwritten in launch A: -1 = "not found"
    DD_CUDA(h, cudaMemsetAsync(d_idx_a.p, 0xff, sizeof(int32_t) * (size_t)n_pad * k, h->stream));
    DD_TRY(dd_stage_begin(h));
    // operand tiles of the permuted, padded embedding
    const int64_t op_bytes = (int64_t)n_tiles * tc::TILE_BYTES;
    DD_TRY(dd_reserve(h, &h->d_knn_ops, &h->cap_knn_ops, 2 * op_bytes));
    uint8_t *qa = h->d_knn_ops, *cb = h->d_knn_ops + op_bytes;
    DD_LAUNCH(h, "prune_gather", k_prune_gather, (unsigned)((n_pad * 8 + 255) / 256), 256, 0, h->d_emb, d_perm.p, n_pad, d_emb_p.p);
    DD_TRY(dd_knn_launch_prep(h, d_emb_p.p, n_pad, n_pad, qa, cb));
    DD_LAUNCH(h, "prune_boxes", k_prune_boxes, (unsigned)(((int64_t)n_tiles * 32 + 255) / 256), 256, 0, d_emb_p.p, d_perm.p, n_tiles,
              d_lo.p, d_hi.p, d_tile_rows.p);
    // launch A + exact re-ranking in the permuted numbering -> thresholds
    DD_TRY(dd_knn_launch_listed16(h, qa, cb, n_pad, n_tiles, n_blocks, d_cand_p.p, d_off_a.p, d_tiles_a.p, nullptr, nullptr, nullptr,
                                  nullptr, 1, 0));
    DD_TRY(dd_knn_launch_refine16(h, d_emb_p.p, d_cand_p.p, n_pad, (int)k, d_idx_a.p, d_dist_a.p));
    DD_LAUNCH(h, "prune_threshold", k_prune_threshold, (unsigned)n_blocks, 256, 0, d_perm.p, d_idx_a.p, d_dist_a.p, (int)k, d_thr.p);
    DD_LAUNCH(h, "prune_lists", k_prune_lists, (unsigned)n_blocks, 256, 0, d_lo.p, d_hi.p, d_tile_rows.p, d_thr.p, n_tiles,
              (const int32_t *)nullptr, d_list_b.p, d_len_b.p);
    DD_LAUNCH(h, "prune_offsets", k_prune_offsets, 1, 1, 0, d_len_b.p, n_blocks, d_off_b.p);
    DD_LAUNCH(h, "prune_lists", k_prune_lists, (unsigned)n_blocks, 256, 0, d_lo.p, d_hi.p, d_tile_rows.p, d_thr.p, n_tiles,
              (const int32_t *)d_off_b.p, d_list_b.p, d_len_b.p);
    // launch B, back to the original numbering, exact re-ranking there
    DD_TRY(dd_knn_launch_listed16(h, qa, cb, n_pad, n_tiles, n_blocks, d_cand_p.p, d_off_b.p, d_list_b.p, nullptr, nullptr, nullptr,
                                  nullptr, 1, 0));
    DD_LAUNCH(h, "prune_translate", k_prune_translate, (unsigned)((n_pad * 16 + 255) / 256), 256, 0, d_perm.p, d_cand_p.p, n_pad,
              d_cand_o.p, 16, 0, 16);
    // output buffers of the ordinary kNN (sized by an earlier dd_knn call on this embedding, or here)
    if (!h->d_knn_idx || h->cap_knn < n * k + n + ((n + 255) / 256 * 256) * 72)
        return dd_fail(h, DD_ERR_ARG, "dd_knn_pruned: call dd_knn on this embedding first (it sizes the output buffers)");
    DD_TRY(dd_knn_launch_refine16(h, h->d_emb, d_cand_o.p, n, (int)k, h->d_knn_idx, h->d_knn_dist));
    DD_TRY(dd_stage_end(h, "knn"));
    std::vector<int32_t> len_b((size_t)n_blocks);
    DD_CUDA(h, cudaMemcpyAsync(len_b.data(), d_len_b.p, sizeof(int32_t) * n_blocks, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaMemcpyAsync(idx_out, h->d_knn_idx, sizeof(int32_t) * n * k, cudaMemcpyDeviceToHost, h->stream));
    if (dist_out)
        DD_CUDA(h, cudaMemcpyAsync(dist_out, h->d_knn_dist, sizeof(float) * n * k, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (stats_out) {
        int64_t pairs_b = 0;
        for (int32_t v : len_b) pairs_b += v;
        stats_out[0] = (int64_t)tiles_a.size();
        stats_out[1] = pairs_b;
        stats_out[2] = n_blocks;
        stats_out[3] = n_tiles;
    }
    return DD_OK;
}

// ====================================================================================================================
// The cluster-ordered exact kNN as the fit loop runs it (dd_dev_knn -> here for >= kClusteredMinRows rows, k <= 13): the
// ordering itself is found on the device as well, nothing synchronises with the host.
//
//   1. groups: kGroups k-means centroids on the embedding, warm-started from the previous call on this handle (the fit
//      loop's embeddings of consecutive iterations share their leading components); each call is one Lloyd step (five
//      after a cold start) -- k_km_assign / k_km_update / k_grp_axis.  The grouping only decides how tight the bounds are: any
//      grouping gives the exact result.
//   2. order inside a group: along the coordinate axis on which the group varies most (the bounding boxes are axis
//      aligned, so slabs along an axis are what tightens them), in kBins buckets between mean -/+ 2.5 sigma --
//      k_grp_axis / k_bucket_count.  Groups are padded to whole 256-row query blocks -- k_bucket_layout / k_bucket_scatter.
//   3. gather + operand tiles + per-tile boxes of the permuted embedding (kernels above).
//   4. launch A: every block against the tiles of its own group; exact re-ranking there -> each row's k-th distance, an
//      upper bound of its true one; the block's threshold is the largest of its rows.
//   5. launch B: every block against the tiles of the OTHER groups whose box-to-box distance is within the threshold, longest
//      lists first; a row starts from its launch-A filter threshold (tau) so that only improvements are collected.
//   6. both candidate lists back to the original numbering, exact float64 re-ranking of the <= 32 candidates there
//      (ties break by original index, as in the dense kNN).
namespace {

constexpr int kGroups = 128, kBins = 64, kBuckets = kGroups * kBins;

__global__ void k_km_init(const float *__restrict__ emb, int64_t n, float *__restrict__ cent) {
    const int g = blockIdx.x, c = threadIdx.x;  // 32 threads
    const int64_t row = (int64_t)g * n / kGroups + n / (2 * kGroups);
    cent[g * 32 + c] = emb[min(row, n - 1) * 32 + c];
}

// one thread per row: nearest centroid (largest x.c - |c|^2 / 2).  Per-group counts, sums (centroid update) and, with STATS,
// sums of squares (axis choice) are accumulated in shared memory by a persistent grid (each CTA owns a contiguous range of
// rows) and flushed with one global atomic per touched (group, dimension) and CTA.
constexpr int kAssignThreads = 256;
constexpr size_t kAssignSmem = sizeof(float) * (kGroups * 32 * 3 + kGroups) + sizeof(int) * kGroups;
template <bool STATS>
__global__ void __launch_bounds__(kAssignThreads, 2) k_km_assign(const float *__restrict__ emb, int64_t n, const float *__restrict__ cent,
                                                              int32_t *__restrict__ label, float *__restrict__ acc_sum,
                                                              float *__restrict__ acc_sq, int32_t *__restrict__ acc_cnt) {
    extern __shared__ __align__(16) float km_sm[];
    float *s_c = km_sm;                    // kGroups x 32 (every thread reads the same address: broadcast, no padding needed)
    float *s_sum = s_c + kGroups * 32;
    float *s_sq = s_sum + kGroups * 32;
    float *s_h = s_sq + kGroups * 32;      // |c|^2 / 2
    int *s_cnt = reinterpret_cast<int *>(s_h + kGroups);
    for (int e = threadIdx.x; e < kGroups * 32; e += kAssignThreads) {
        s_c[e] = cent[e];
        s_sum[e] = 0.f;
        s_sq[e] = 0.f;
    }
    __syncthreads();
    if (threadIdx.x < kGroups) {
        float h = 0.f;
        for (int c = 0; c < 32; c++) h += s_c[threadIdx.x * 32 + c] * s_c[threadIdx.x * 32 + c];
        s_h[threadIdx.x] = 0.5f * h;
        s_cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(n, r0 + per);
    for (int64_t row = r0 + threadIdx.x; row < r1; row += kAssignThreads) {
        float x[32];
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(emb + row * 32 + c);
            x[c] = v.x; x[c + 1] = v.y; x[c + 2] = v.z; x[c + 3] = v.w;
        }
        int best = 0;
        float best_t = -3.0e38f;
#pragma unroll 2
        for (int g = 0; g < kGroups; g++) {
            const float4 *c4 = reinterpret_cast<const float4 *>(s_c + g * 32);
            float t0 = -s_h[g], t1 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                const float4 a = c4[q], b = c4[q + 1];
                t0 = fmaf(x[4 * q], a.x, t0); t0 = fmaf(x[4 * q + 1], a.y, t0);
                t0 = fmaf(x[4 * q + 2], a.z, t0); t0 = fmaf(x[4 * q + 3], a.w, t0);
                t1 = fmaf(x[4 * q + 4], b.x, t1); t1 = fmaf(x[4 * q + 5], b.y, t1);
                t1 = fmaf(x[4 * q + 6], b.z, t1); t1 = fmaf(x[4 * q + 7], b.w, t1);
            }
            const float t = t0 + t1;
            if (t > best_t) {
                best_t = t;
                best = g;
            }
        }
        label[row] = best;
        atomicAdd(s_cnt + best, 1);
#pragma unroll
        for (int c = 0; c < 32; c++) {
            atomicAdd(s_sum + best * 32 + c, x[c]);
            if (STATS) atomicAdd(s_sq + best * 32 + c, x[c] * x[c]);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < kGroups * 32; e += kAssignThreads) {
        if (s_cnt[e >> 5] == 0) continue;
        atomicAdd(acc_sum + e, s_sum[e]);
        if (STATS) atomicAdd(acc_sq + e, s_sq[e]);
    }
    if (threadIdx.x < kGroups && s_cnt[threadIdx.x] > 0) atomicAdd(acc_cnt + threadIdx.x, s_cnt[threadIdx.x]);
}

// new centroid = mean of its rows (an empty group keeps its old centroid); clears the accumulators
__global__ void k_km_update(float *__restrict__ cent, float *__restrict__ acc_sum, int32_t *__restrict__ acc_cnt) {
    const int g = blockIdx.x, c = threadIdx.x;  // 32 threads
    const int cnt = acc_cnt[g];
    const float s = acc_sum[g * 32 + c];
    __syncwarp();
    if (cnt > 0) cent[g * 32 + c] = s / (float)cnt;
    acc_sum[g * 32 + c] = 0.f;
    if (c == 0) acc_cnt[g] = 0;
}

// per group: the coordinate axis with the largest variance, and the affine map of that coordinate onto the kBins buckets
// (mean -/+ 2.5 sigma; outliers land in the end buckets).  Also moves the centroid to the group's mean for the next call
// and clears the accumulators.
__global__ void k_grp_axis(float *__restrict__ cent, float *__restrict__ acc_sum, float *__restrict__ acc_sq,
                           int32_t *__restrict__ acc_cnt, int32_t *__restrict__ axis, float *__restrict__ bin_lo,
                           float *__restrict__ bin_scale) {
    const int g = blockIdx.x, c = threadIdx.x;  // 32 threads
    const int cnt = acc_cnt[g];
    const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
    const float mean = acc_sum[g * 32 + c] * inv;
    float var = fmaxf(acc_sq[g * 32 + c] * inv - mean * mean, 0.f);
    __syncwarp();
    if (cnt > 0) cent[g * 32 + c] = mean;
    acc_sum[g * 32 + c] = 0.f;
    acc_sq[g * 32 + c] = 0.f;
    float best_var = var;
    int best = c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_var, o);
        const int ob = __shfl_xor_sync(0xffffffffu, best, o);
        if (ov > best_var || (ov == best_var && ob < best)) {
            best_var = ov;
            best = ob;
        }
    }
    const float m_best = __shfl_sync(0xffffffffu, mean, best);
    if (c == 0) {
        const float sd = sqrtf(best_var);
        axis[g] = best;
        bin_lo[g] = m_best - 2.5f * sd;
        bin_scale[g] = sd > 0.f ? (float)kBins / (5.f * sd) : 0.f;
        acc_cnt[g] = 0;
    }
}

__global__ void k_bucket_count(const float *__restrict__ emb, int64_t n, const int32_t *__restrict__ label,
                               const int32_t *__restrict__ axis, const float *__restrict__ bin_lo,
                               const float *__restrict__ bin_scale, int32_t *__restrict__ bucket, int32_t *__restrict__ hist) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int g = label[row];
    const float x = emb[row * 32 + axis[g]];
    const float f = (x - bin_lo[g]) * bin_scale[g];
    const int b = g * kBins + max(0, min(kBins - 1, (int)f));
    bucket[row] = b;
    atomicAdd(hist + b, 1);
}

// one CTA, kGroups threads... groups padded to whole 256-row blocks; first position of every bucket; per-block group (-1:
// block unused) and per-group tile range.  info[0] = padded rows in use.
__global__ void __launch_bounds__(kGroups) k_bucket_layout(const int32_t *__restrict__ hist, int32_t *__restrict__ start,
                                                           int32_t *__restrict__ cursor, int32_t *__restrict__ block_group,
                                                           int n_blocks_max, int32_t *__restrict__ group_tile0,
                                                           int32_t *__restrict__ group_tiles, int32_t *__restrict__ info) {
    __shared__ int s_pad[kGroups], s_start[kGroups + 1];
    const int g = threadIdx.x;
    int size = 0;
    for (int b = 0; b < kBins; b++) size += hist[g * kBins + b];
    s_pad[g] = (size + 255) / 256 * 256;
    __syncthreads();
    if (g == 0) {
        int run = 0;
        for (int i = 0; i < kGroups; i++) {
            s_start[i] = run;
            run += s_pad[i];
        }
        s_start[kGroups] = run;
        info[0] = run;
    }
    __syncthreads();
    int run = s_start[g];
    for (int b = 0; b < kBins; b++) {
        start[g * kBins + b] = run;
        cursor[g * kBins + b] = 0;
        run += hist[g * kBins + b];
    }
    group_tile0[g] = s_start[g] / tc::TILE;
    group_tiles[g] = (size + tc::TILE - 1) / tc::TILE;  // tiles that hold real rows
    for (int b = s_start[g] / 256; b < s_start[g + 1] / 256; b++) block_group[b] = g;
    for (int b = s_start[kGroups] / 256 + g; b < n_blocks_max; b += kGroups) block_group[b] = -1;
}

__global__ void k_bucket_scatter(int64_t n, const int32_t *__restrict__ bucket, const int32_t *__restrict__ start,
                                 int32_t *__restrict__ cursor, int32_t *__restrict__ perm) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const int b = bucket[row];
    perm[start[b] + atomicAdd(cursor + b, 1)] = (int32_t)row;
}

// launch A lists: the tiles of the block's own group that hold real rows, at stride n_tiles_max -- starting with the block's
// own two tiles and moving outwards (the group is ordered in slabs: the nearest slabs come first, so a row's list fills
// with near-final candidates at once and the later tiles rarely pass its threshold)
__global__ void k_lists_own(const int32_t *__restrict__ block_group, const int32_t *__restrict__ group_tile0,
                            const int32_t *__restrict__ group_tiles, int n_tiles_max, int32_t *__restrict__ off,
                            int32_t *__restrict__ list, int32_t *__restrict__ len) {
    const int b = blockIdx.x;
    const int g = block_group[b];
    const int t0 = g >= 0 ? group_tile0[g] : 0, nt = g >= 0 ? group_tiles[g] : 0;
    // position of the block's first tile inside the group's real tiles (the group's last block may hold fewer real tiles)
    const int own = min(max(b * tc::QT - t0, 0), max(nt - 1, 0));
    for (int i = threadIdx.x; i < nt; i += blockDim.x) {
        // i-th entry: own, own + 1, own - 1, own + 2, own - 2, ... folded back into [0, nt)
        const int step = (i + 1) / 2;
        int t = (i & 1) ? own + step : own - step;
        // once one side runs out the other side continues
        const int left = own, right = nt - 1 - own;  // tiles available on either side
        if (step > min(left, right)) {
            const int m = min(left, right);
            const int rest = i - 2 * m;  // entries after the symmetric part (i >= 2 m + 1)
            t = left < right ? own + m + rest : own - m - rest;
        }
        list[(int64_t)b * n_tiles_max + i] = t0 + t;
    }
    if (threadIdx.x == 0) {
        off[b] = b * n_tiles_max;
        len[b] = nt;
    }
}

// launch B lists in ONE pass (fixed stride): the tiles of OTHER groups whose box-to-box distance^2 to the block's box is
// within the block's threshold.  One CTA (256 threads) per block; order inside the list = tile order.
__global__ void __launch_bounds__(256) k_lists_other(const float *__restrict__ lo, const float *__restrict__ hi,
                                                     const float *__restrict__ lo_t, const float *__restrict__ hi_t,
                                                     const int32_t *__restrict__ tile_rows, const double *__restrict__ thr,
                                                     const int32_t *__restrict__ block_group, int n_tiles_max,
                                                     int32_t *__restrict__ list, int32_t *__restrict__ len,
                                                     int shard_world = 1, int shard_rank = 0) {
    __shared__ float q_lo[32], q_hi[32];
    __shared__ int s_warp[8], s_base;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wl = tid >> 5;
    const int g = block_group[b];
    if (g < 0 || (shard_world > 1 && (b % shard_world) != shard_rank)) {  // unused block, or another rank's (uniform)
        if (tid == 0) len[b] = 0;
        return;
    }
    if (tid < 32) {
        float mn = 3.0e38f, mx = -3.0e38f;
        for (int h2 = 0; h2 < tc::QT; h2++) {
            const int t = b * tc::QT + h2;
            if (tile_rows[t] > 0) {
                mn = fminf(mn, lo[(int64_t)t * 32 + tid]);
                mx = fmaxf(mx, hi[(int64_t)t * 32 + tid]);
            }
        }
        q_lo[tid] = mn;
        q_hi[tid] = mx;
    }
    if (tid == 0) s_base = 0;
    __syncthreads();
    const double limit = thr[b] * (1.0 + 1.0e-5);
    int32_t *mine = list + (int64_t)b * n_tiles_max;
    for (int t0 = 0; t0 < n_tiles_max; t0 += 256) {
        const int t = t0 + tid;
        bool need = false;
        if (t < n_tiles_max && tile_rows[t] > 0 && block_group[t / tc::QT] != g) {
            double lb = 0.0;
            for (int c = 0; c < 32; c++) {
                const float gap = fmaxf(0.f, fmaxf(lo_t[(int64_t)c * n_tiles_max + t] - q_hi[c], q_lo[c] - hi_t[(int64_t)c * n_tiles_max + t]));
                lb += (double)gap * (double)gap;
            }
            need = lb <= limit;
        }
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (lane == 0) s_warp[wl] = __popc(m);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < wl; w++) before += s_warp[w];
        if (need) mine[before + __popc(m & ((1u << lane) - 1u))] = t;
        __syncthreads();
        if (tid == 0) {
            int total = 0;
            for (int w = 0; w < 8; w++) total += s_warp[w];
            s_base += total;
        }
        __syncthreads();
    }
    if (tid == 0) len[b] = s_base;
}

// blocks by decreasing list length (rank by counting; ties by block index): order[rank] = block
__global__ void k_block_order(const int32_t *__restrict__ len, int n_blocks, int32_t *__restrict__ order) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const int mine = len[b];
    int rank = 0;
    for (int o = 0; o < n_blocks; o++) {
        const int l = len[o];
        rank += (l > mine) || (l == mine && o < b);
    }
    order[rank] = b;
}

// info[2], info[3] = block-tile pairs of launch A and launch B (read back asynchronously: the fit loop uses it to notice
// embeddings without cluster structure, where the ordering does not pay)
__global__ void k_sum_pairs(const int32_t *__restrict__ len_a, const int32_t *__restrict__ len_b, int n_blocks,
                            int32_t *__restrict__ info) {
    __shared__ long long s_a[256], s_b[256];
    long long a = 0, b = 0;
    for (int i = threadIdx.x; i < n_blocks; i += 256) {
        a += len_a[i];
        b += len_b[i];
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 256; i++) {
            a += s_a[i];
            b += s_b[i];
        }
        info[2] = (int32_t)min(a, 0x7fffffffll);
        info[3] = (int32_t)min(b, 0x7fffffffll);
    }
}

struct ClusteredBuffers {
    int32_t *label, *bucket, *acc_cnt, *axis, *hist, *start, *cursor, *block_group, *group_tile0, *group_tiles, *info, *perm,
        *tile_rows, *off, *list_a, *len_a, *list_b, *len_b, *order, *idx_a;
    int *cand_a, *cand_b, *cand_o;
    float *cent, *acc_sum, *acc_sq, *bin_lo, *bin_scale, *emb_p, *lo, *hi, *lo_t, *hi_t, *tau, *dist_a;
    double *thr;
};

}  // namespace

// Smallest embedding for which the ordering pays (clusters must be larger than tiles; measured: DESIGN.md section 5)
static constexpr int64_t kClusteredMinRows = 50000;

bool dd_knn_clustered_applies(const dd_handle *h, int32_t k) {
    static const bool off = getenv("DD_KNN_DENSE") != nullptr;
    if (off || h->knn_mode == 1 || h->KP != 32 || k < 2 || k > 31) return false;
    if (h->knn_mode == 2) return h->emb_rows >= 512;
    if (h->emb_rows < kClusteredMinRows) return false;
    // cell-block sharding: every rank must take the same branch (the calls below contain collectives), so the choice depends
    // on the size only -- no feedback from the pair counts, which differ from rank to rank
    if (dd_sharded(h)) return true;
    // by size -- unless an earlier call on this problem size reported that the ordering does not pay: on an embedding
    // without cluster structure the bounds exclude nothing and the padded order visits MORE pairs than the all-tiles kernel
    // (uniform points: 1.2 x).  The counts arrive asynchronously (pinned host memory), one or two calls late.
    if (h->h_knn_cl_pairs && h->knn_cl_rows == h->emb_rows) {
        const int64_t visited = (int64_t)h->h_knn_cl_pairs[0] + h->h_knn_cl_pairs[1];
        const int64_t dense = ((h->emb_rows + 255) / 256) * ((h->emb_rows + 127) / 128);
        if (visited > dense * 8 / 10) return false;
    }
    return true;
}

// 0 = choose by size (default), 1 = always the all-tiles kernel, 2 = always the cluster-ordered path (tests, A/B timing)
extern "C" int dd_set_knn_mode(dd_handle *h, int32_t mode) {
    if (!h || mode < 0 || mode > 2) return dd_fail(h, DD_ERR_ARG, "dd_set_knn_mode: mode must be 0, 1 or 2");
    h->knn_mode = mode;
    return DD_OK;
}

// Asynchronous on h->stream; result in h->d_knn_idx / h->d_knn_dist (allocated by dd_dev_knn).
int dd_dev_knn_clustered(dd_handle *h, int32_t k) {
    const int64_t n = h->emb_rows;
    const int64_t P = (n + 255) / 256 * 256 + (int64_t)kGroups * 256;  // padded rows: every group wastes < 256
    const int T = (int)(P / tc::TILE), B = (int)(P / 256);
    if ((int64_t)B * T >= (1ll << 31) - 1) return dd_fail(h, DD_ERR_UNSUPPORTED, "clustered knn: list table too large");
    const int TL = (k - 1 <= 12) ? 16 : 40;  // candidates kept per row and launch (as in dd_dev_knn)
    // Cell-block sharding: every rank holds the whole (all-gathered) embedding and runs the pre-pass; rank 0's ordering is
    // broadcast (the k-means sums and the bucket cursors are atomics: their rounding / order differs between ranks), the
    // 256-row blocks of the permuted order are dealt to the ranks round-robin, every rank re-ranks the rows of its blocks
    // into a zeroed result buffer, and the buffers are summed (one writer per word).
    const int W = dd_sharded(h) ? h->world : 1, R = dd_sharded(h) ? h->rank : 0;
    // ---- one grow-only allocation, carved up
    size_t bytes = 0;
    auto take = [&](size_t b) {
        const size_t at = bytes;
        bytes += (b + 255) / 256 * 256;
        return at;
    };
    const size_t o_label = take(4 * n), o_bucket = take(4 * n), o_acc_cnt = take(4 * kGroups), o_axis = take(4 * kGroups),
                 o_hist = take(4 * kBuckets), o_start = take(4 * kBuckets), o_cursor = take(4 * kBuckets), o_bg = take(4 * B),
                 o_gt0 = take(4 * kGroups), o_gt = take(4 * kGroups), o_info = take(64), o_perm = take(4 * P),
                 o_trows = take(4 * T), o_off = take(4 * B), o_list_a = take(4 * (size_t)B * T), o_len_a = take(4 * B),
                 o_list_b = take(4 * (size_t)B * T), o_len_b = take(4 * B), o_order = take(4 * B), o_idx_a = take(4 * P * k),
                 o_cand_a = take(4 * P * TL), o_cand_b = take(4 * P * TL), o_cand_o = take(4 * n * 2 * TL),
                 o_cent = take(4 * kGroups * 32), o_acc_sum = take(4 * kGroups * 32), o_acc_sq = take(4 * kGroups * 32),
                 o_bin_lo = take(4 * kGroups), o_bin_scale = take(4 * kGroups), o_emb_p = take(4 * P * 32),
                 o_lo = take(4 * (size_t)T * 32), o_hi = take(4 * (size_t)T * 32), o_tau = take(4 * P), o_dist_a = take(4 * P * k),
                 o_thr = take(8 * B), o_lo_t = take(4 * (size_t)T * 32), o_hi_t = take(4 * (size_t)T * 32);
    bool cold = false;
    if ((int64_t)bytes > h->cap_knn_cl || h->knn_cl_rows != n || h->knn_cl_tl != TL) {
        if ((int64_t)bytes > h->cap_knn_cl) {
            if (h->d_knn_cl) cudaFree(h->d_knn_cl);
            h->d_knn_cl = nullptr;
            h->cap_knn_cl = 0;
            DD_CUDA(h, cudaMalloc(&h->d_knn_cl, bytes));
            h->cap_knn_cl = (int64_t)bytes;
        }
        h->knn_cl_rows = n;
        h->knn_cl_tl = TL;  // the carving depends on it
        cold = true;
        if (!h->h_knn_cl_pairs && cudaMallocHost(&h->h_knn_cl_pairs, 2 * sizeof(int32_t)) != cudaSuccess) h->h_knn_cl_pairs = nullptr;
        if (h->h_knn_cl_pairs) h->h_knn_cl_pairs[0] = h->h_knn_cl_pairs[1] = 0;  // nothing known about this problem size yet  // no centroids from an earlier call on this problem size
    }
    uint8_t *base = h->d_knn_cl;
    ClusteredBuffers b;
    b.label = (int32_t *)(base + o_label); b.bucket = (int32_t *)(base + o_bucket); b.acc_cnt = (int32_t *)(base + o_acc_cnt);
    b.axis = (int32_t *)(base + o_axis); b.hist = (int32_t *)(base + o_hist); b.start = (int32_t *)(base + o_start);
    b.cursor = (int32_t *)(base + o_cursor); b.block_group = (int32_t *)(base + o_bg); b.group_tile0 = (int32_t *)(base + o_gt0);
    b.group_tiles = (int32_t *)(base + o_gt); b.info = (int32_t *)(base + o_info); b.perm = (int32_t *)(base + o_perm);
    b.tile_rows = (int32_t *)(base + o_trows); b.off = (int32_t *)(base + o_off); b.list_a = (int32_t *)(base + o_list_a);
    b.len_a = (int32_t *)(base + o_len_a); b.list_b = (int32_t *)(base + o_list_b); b.len_b = (int32_t *)(base + o_len_b);
    b.order = (int32_t *)(base + o_order); b.idx_a = (int32_t *)(base + o_idx_a); b.cand_a = (int *)(base + o_cand_a);
    b.cand_b = (int *)(base + o_cand_b); b.cand_o = (int *)(base + o_cand_o); b.cent = (float *)(base + o_cent);
    b.acc_sum = (float *)(base + o_acc_sum); b.acc_sq = (float *)(base + o_acc_sq); b.bin_lo = (float *)(base + o_bin_lo);
    b.bin_scale = (float *)(base + o_bin_scale); b.emb_p = (float *)(base + o_emb_p); b.lo = (float *)(base + o_lo);
    b.hi = (float *)(base + o_hi); b.tau = (float *)(base + o_tau); b.dist_a = (float *)(base + o_dist_a);
    b.thr = (double *)(base + o_thr);
    b.lo_t = (float *)(base + o_lo_t); b.hi_t = (float *)(base + o_hi_t);
    const int64_t op_bytes = (int64_t)T * tc::TILE_BYTES;
    DD_TRY(dd_reserve(h, &h->d_knn_ops, &h->cap_knn_ops, 2 * op_bytes));
    uint8_t *qa = h->d_knn_ops, *cb = h->d_knn_ops + op_bytes;
    const unsigned row_ctas = (unsigned)((n + 255) / 256);
    const unsigned assign_grid = (unsigned)std::min<int64_t>(2 * h->num_sms, (n + kAssignThreads - 1) / kAssignThreads);
    static dd_once_per_device assign_attr;  // function attributes are per device
    assign_attr.run(h->device, [&] {
        cudaFuncSetAttribute(k_km_assign<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAssignSmem);
        cudaFuncSetAttribute(k_km_assign<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAssignSmem);
    });

    // ---- 1. groups
    if (cold) {
        DD_LAUNCH(h, "kcl_init", k_km_init, kGroups, 32, 0, h->d_emb, n, b.cent);
        DD_CUDA(h, cudaMemsetAsync(b.acc_cnt, 0, 4 * kGroups, h->stream));
        DD_CUDA(h, cudaMemsetAsync(b.acc_sum, 0, 4 * kGroups * 32, h->stream));
        DD_CUDA(h, cudaMemsetAsync(b.acc_sq, 0, 4 * kGroups * 32, h->stream));
    }
    // warm: the final assignment below is itself one Lloyd step (k_grp_axis moves every centroid to its group's mean)
    for (int pass = 0; pass < (cold ? 4 : 0); pass++) {
        DD_LAUNCH(h, "kcl_assign", k_km_assign<false>, assign_grid, kAssignThreads, kAssignSmem, h->d_emb, n, b.cent, b.label,
                  b.acc_sum, b.acc_sq, b.acc_cnt);
        DD_LAUNCH(h, "kcl_update", k_km_update, kGroups, 32, 0, b.cent, b.acc_sum, b.acc_cnt);
    }
    DD_LAUNCH(h, "kcl_assign", k_km_assign<true>, assign_grid, kAssignThreads, kAssignSmem, h->d_emb, n, b.cent, b.label, b.acc_sum,
              b.acc_sq, b.acc_cnt);
    // ---- 2. order inside the groups, padded layout
    DD_LAUNCH(h, "kcl_axis", k_grp_axis, kGroups, 32, 0, b.cent, b.acc_sum, b.acc_sq, b.acc_cnt, b.axis, b.bin_lo, b.bin_scale);
    DD_CUDA(h, cudaMemsetAsync(b.hist, 0, 4 * kBuckets, h->stream));
    DD_LAUNCH(h, "kcl_bucket", k_bucket_count, row_ctas, 256, 0, h->d_emb, n, b.label, b.axis, b.bin_lo, b.bin_scale, b.bucket, b.hist);
    DD_LAUNCH(h, "kcl_layout", k_bucket_layout, 1, kGroups, 0, b.hist, b.start, b.cursor, b.block_group, B, b.group_tile0,
              b.group_tiles, b.info);
    DD_CUDA(h, cudaMemsetAsync(b.perm, 0xff, 4 * P, h->stream));
    DD_LAUNCH(h, "kcl_scatter", k_bucket_scatter, row_ctas, 256, 0, n, b.bucket, b.start, b.cursor, b.perm);
    if (W > 1) {
        DD_TRY(dd_comm_bcast(h, b.perm, 4 * P, 0));
        DD_TRY(dd_comm_bcast(h, b.block_group, 4 * (int64_t)B, 0));
        DD_TRY(dd_comm_bcast(h, b.group_tile0, 4 * kGroups, 0));
        DD_TRY(dd_comm_bcast(h, b.group_tiles, 4 * kGroups, 0));
    }
    // ---- 3. permuted embedding, operand tiles, boxes
    DD_LAUNCH(h, "prune_gather", k_prune_gather, (unsigned)((P * 8 + 255) / 256), 256, 0, h->d_emb, b.perm, P, b.emb_p);
    DD_TRY(dd_knn_launch_prep(h, b.emb_p, P, P, qa, cb));
    DD_LAUNCH(h, "prune_boxes", k_prune_boxes, (unsigned)(((int64_t)T * 32 + 255) / 256), 256, 0, b.emb_p, b.perm, T, b.lo, b.hi,
              b.tile_rows, b.lo_t, b.hi_t);
    // ---- 4. launch A (own group) -> thresholds
    DD_LAUNCH(h, "kcl_lists_own", k_lists_own, (unsigned)B, 128, 0, b.block_group, b.group_tile0, b.group_tiles, T, b.off, b.list_a,
              b.len_a);
    DD_LAUNCH(h, "kcl_order", k_block_order, (unsigned)((B + 255) / 256), 256, 0, b.len_a, B, b.order);
    if (TL == 16)
        DD_TRY(dd_knn_launch_listed16(h, qa, cb, P, T, B, b.cand_a, b.off, b.list_a, b.len_a, b.order, nullptr, b.tau, W, R));
    else
        DD_TRY(dd_knn_launch_listed40(h, qa, cb, P, T, B, b.cand_a, b.off, b.list_a, b.len_a, b.order, nullptr, b.tau, W, R));
    DD_CUDA(h, cudaMemsetAsync(b.idx_a, 0xff, sizeof(int32_t) * (size_t)P * k, h->stream));  // -1 = "not found"
    DD_TRY(dd_knn_launch_refine_lists(h, b.emb_p, b.cand_a, TL, P, (int)k, b.idx_a, b.dist_a, W, R));
    DD_LAUNCH(h, "prune_threshold", k_prune_threshold, (unsigned)B, 256, 0, b.perm, b.idx_a, b.dist_a, (int)k, b.thr, W, R);
    // ---- 5. launch B (other groups within the bound), longest lists first, starting from launch A's filter thresholds
    DD_LAUNCH(h, "kcl_lists_other", k_lists_other, (unsigned)B, 256, 0, b.lo, b.hi, b.lo_t, b.hi_t, b.tile_rows, b.thr, b.block_group,
              T, b.list_b, b.len_b, W, R);
    DD_LAUNCH(h, "kcl_order", k_block_order, (unsigned)((B + 255) / 256), 256, 0, b.len_b, B, b.order);
    if (TL == 16)
        DD_TRY(dd_knn_launch_listed16(h, qa, cb, P, T, B, b.cand_b, b.off, b.list_b, b.len_b, b.order, b.tau, nullptr, W, R));
    else
        DD_TRY(dd_knn_launch_listed40(h, qa, cb, P, T, B, b.cand_b, b.off, b.list_b, b.len_b, b.order, b.tau, nullptr, W, R));
    DD_LAUNCH(h, "kcl_pairs", k_sum_pairs, 1, 256, 0, b.len_a, b.len_b, B, b.info);
    if (h->h_knn_cl_pairs)
        DD_CUDA(h, cudaMemcpyAsync(h->h_knn_cl_pairs, b.info + 2, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    // ---- 6. back to the original numbering, exact re-ranking of both lists together
    DD_LAUNCH(h, "prune_translate", k_prune_translate, (unsigned)((P * TL + 255) / 256), 256, 0, b.perm, b.cand_a, P, b.cand_o, 2 * TL, 0,
              TL, W, R);
    DD_LAUNCH(h, "prune_translate", k_prune_translate, (unsigned)((P * TL + 255) / 256), 256, 0, b.perm, b.cand_b, P, b.cand_o, 2 * TL, TL,
              TL, W, R);
    // (with the filter's certificate over both lists, and the float64 fix-up of the rows it cannot clear)
    if (W == 1) {
        DD_TRY(dd_knn_refine_final(h, h->d_emb, b.cand_o, 2 * TL, TL, 2, 0, n, n, (int)k, nullptr, 1, 0));
    } else {
        // this rank's rows = the real rows of its blocks of the permuted order: iterate over permuted positions, perm maps them
        // to original rows; everything else stays zero and the ranks' buffers are summed
        DD_CUDA(h, cudaMemsetAsync(h->d_knn_idx, 0, sizeof(int32_t) * (size_t)n * k, h->stream));
        DD_CUDA(h, cudaMemsetAsync(h->d_knn_dist, 0, sizeof(float) * (size_t)n * k, h->stream));
        DD_TRY(dd_knn_refine_final(h, h->d_emb, b.cand_o, 2 * TL, TL, 2, 0, P, n, (int)k, b.perm, W, R));
        DD_TRY(dd_comm_allreduce_i32(h, h->d_knn_idx, n * k));
        DD_TRY(dd_comm_allreduce_i32(h, reinterpret_cast<int32_t *>(h->d_knn_dist), n * k));
    }
    return DD_OK;
}

// Inspection: how much of the dense work the last clustered kNN on this handle did.  stats_out[0..3] = block-tile pairs of
// launch A, of launch B, blocks in use, tiles in use.
extern "C" int dd_knn_clustered_stats(dd_handle *h, int64_t *stats_out) {
    if (!h || !stats_out || !h->d_knn_cl || h->knn_cl_rows <= 0) return dd_fail(h, DD_ERR_ARG, "dd_knn_clustered_stats: no clustered kNN ran");
    DD_CUDA(h, cudaSetDevice(h->device));
    const int64_t n = h->knn_cl_rows;
    const int64_t P = (n + 255) / 256 * 256 + (int64_t)kGroups * 256;
    const int B = (int)(P / 256);
    // same carving as dd_dev_knn_clustered (offsets recomputed)
    size_t bytes = 0;
    auto take = [&](size_t b) { const size_t at = bytes; bytes += (b + 255) / 256 * 256; return at; };
    const int T = (int)(P / tc::TILE);
    take(4 * n); take(4 * n); take(4 * kGroups); take(4 * kGroups); take(4 * kBuckets); take(4 * kBuckets); take(4 * kBuckets);
    take(4 * B); take(4 * kGroups); take(4 * kGroups);
    const size_t o_info = take(64);
    take(4 * P); take(4 * T); take(4 * B); take(4 * (size_t)B * T);
    const size_t o_len_a = take(4 * B);
    take(4 * (size_t)B * T);
    const size_t o_len_b = take(4 * B);
    std::vector<int32_t> la(B), lb(B);
    int32_t info[16];
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    DD_CUDA(h, cudaMemcpy(la.data(), h->d_knn_cl + o_len_a, 4 * B, cudaMemcpyDeviceToHost));
    DD_CUDA(h, cudaMemcpy(lb.data(), h->d_knn_cl + o_len_b, 4 * B, cudaMemcpyDeviceToHost));
    DD_CUDA(h, cudaMemcpy(info, h->d_knn_cl + o_info, 64, cudaMemcpyDeviceToHost));
    int64_t pa = 0, pb = 0;
    for (int i = 0; i < B; i++) { pa += la[i]; pb += lb[i]; }
    stats_out[0] = pa; stats_out[1] = pb; stats_out[2] = info[0] / 256; stats_out[3] = info[0] / tc::TILE;
    return DD_OK;
}
