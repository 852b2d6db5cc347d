// knn_prune.cu -- EXPERIMENTAL (never run on hardware; DESIGN.md section 5 "next lever"): the exact kNN with cluster-ordered
// candidate tiles, everything after the ordering on the device.  Its own translation unit: knn.cu only gains three launchers,
// its measured kernels keep their instruction sequence.
#include "dd_internal.h"

#include <cmath>
#include <vector>

int dd_knn_launch_prep(dd_handle *h, const float *emb, int64_t n, int64_t n_pad, uint8_t *qa, uint8_t *cb);        // knn.cu
int dd_knn_launch_listed16(dd_handle *h, const uint8_t *qa, const uint8_t *cb, int64_t n, int n_tiles, int n_blocks, int *cand_i,
                           const int *list_off, const int *list_tiles);                                              // knn.cu
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
                              float *__restrict__ lo, float *__restrict__ hi, int32_t *__restrict__ tile_rows) {
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
    if (lane == 0) tile_rows[t] = rows;
}

// one CTA (256 threads) per 256-row block: the largest k-th distance^2 launch A found for its real rows (inf if a row
// found fewer than k - 1 neighbours in its own group)
__global__ void __launch_bounds__(256) k_prune_threshold(const int32_t *__restrict__ perm, const int32_t *__restrict__ idx_a,
                                                         const float *__restrict__ dist_a, int k, double *__restrict__ thr) {
    __shared__ double s_max[8];
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
                                  int *__restrict__ cand_o) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t >> 4;
    const int l = (int)(t & 15);
    if (r >= n_pad) return;
    const int o = perm[r];
    if (o < 0) return;
    const int c = cand_p[r * 16 + l];
    int out = 0x7fffffff;
    if (c != 0x7fffffff && c >= 0 && c < n_pad) {
        const int oc = perm[c];
        if (oc >= 0) out = oc;
    }
    cand_o[(int64_t)o * 16 + l] = out;
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
    DD_TRY(dd_knn_launch_listed16(h, qa, cb, n_pad, n_tiles, n_blocks, d_cand_p.p, d_off_a.p, d_tiles_a.p));
    DD_TRY(dd_knn_launch_refine16(h, d_emb_p.p, d_cand_p.p, n_pad, (int)k, d_idx_a.p, d_dist_a.p));
    DD_LAUNCH(h, "prune_threshold", k_prune_threshold, (unsigned)n_blocks, 256, 0, d_perm.p, d_idx_a.p, d_dist_a.p, (int)k, d_thr.p);
    DD_LAUNCH(h, "prune_lists", k_prune_lists, (unsigned)n_blocks, 256, 0, d_lo.p, d_hi.p, d_tile_rows.p, d_thr.p, n_tiles,
              (const int32_t *)nullptr, d_list_b.p, d_len_b.p);
    DD_LAUNCH(h, "prune_offsets", k_prune_offsets, 1, 1, 0, d_len_b.p, n_blocks, d_off_b.p);
    DD_LAUNCH(h, "prune_lists", k_prune_lists, (unsigned)n_blocks, 256, 0, d_lo.p, d_hi.p, d_tile_rows.p, d_thr.p, n_tiles,
              (const int32_t *)d_off_b.p, d_list_b.p, d_len_b.p);
    // launch B, back to the original numbering, exact re-ranking there
    DD_TRY(dd_knn_launch_listed16(h, qa, cb, n_pad, n_tiles, n_blocks, d_cand_p.p, d_off_b.p, d_list_b.p));
    DD_LAUNCH(h, "prune_translate", k_prune_translate, (unsigned)((n_pad * 16 + 255) / 256), 256, 0, d_perm.p, d_cand_p.p, n_pad,
              d_cand_o.p);
    // output buffers of the ordinary kNN (sized by an earlier dd_knn call on this embedding, or here)
    if (!h->d_knn_idx || h->cap_knn < n * k + n + 2 * ((n + 255) / 256 * 256) * 32)
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
