// pca_exact.cu -- the device side of sklearn's EXACT PCA branches (svd_solver "covariance_eigh" and "full", which
// PCA(svd_solver="auto") picks for <= 1000 genes with >= 10x as many augmented cells, for tiny matrices, and when
// n_components >= 0.8 min(shape): sklearn/decomposition/_pca.py:524-536, 560-640; reached from doubletdetection.py:309-314).
// Both are the top principal components of the centred matrix, so the path is
//     Gram matrix of the centred matrix on its SMALLER side, float64 (here)  ->  symmetric eigendecomposition of that
//     <= 1000 x 1000 matrix (host LAPACK, called by the Python shim: O(G^3), independent of the number of cells)  ->
//     projection X_pca = (D - mean) V of all augmented cells (here).
// Everything that scales with the number of cells runs on the device; products are exact in float64 (float32 inputs), so
// the result is the float64 truth rounded once -- sklearn's own float32 run is ~1e-5 away from it.
#include <vector>

#include "dd_internal.h"

namespace {

constexpr int GT = 64, GK = 32;  // output tile (GT x GT), reduction chunk

// out[o1][o2] += sum_k v(k, o1) v(k, o2),  v = D - mean[gene];  TR == false: k = row, o = gene;  TR == true: k = gene, o = row.
// One CTA = one GT x GT tile of the upper triangle x one slice of the reduction range; 16 x 16 threads, 4 x 4 outputs each.
template <bool TR>
__global__ void __launch_bounds__(256) k_gram64(const float *__restrict__ D, int64_t ld, int64_t n_rows, int n_genes,
                                                const double *__restrict__ mean, double *__restrict__ out, int n_out,
                                                int64_t red_len, int64_t red_per_split) {
    if (blockIdx.x > blockIdx.y) return;  // symmetric: the host mirrors the upper triangle
    __shared__ double sa[GK][GT], sb[GK][GT];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t o1 = (int64_t)blockIdx.x * GT, o2 = (int64_t)blockIdx.y * GT;
    const int64_t k0 = (int64_t)blockIdx.z * red_per_split, k1 = min(red_len, k0 + red_per_split);
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
    auto value = [&](int64_t k, int64_t o) -> double {
        const int64_t row = TR ? o : k, gene = TR ? k : o;
        if (row >= n_rows || gene >= n_genes) return 0.0;
        return (double)D[row * ld + gene] - mean[gene];
    };
    for (int64_t kc = k0; kc < k1; kc += GK) {
        for (int e = threadIdx.x; e < GK * GT; e += 256) {
            // consecutive threads walk the contiguous direction of D: genes
            const int kk = TR ? (e & (GK - 1)) : (e / GT), oo = TR ? (e / GK) : (e & (GT - 1));
            const bool in = kc + kk < k1;
            sa[kk][oo] = in ? value(kc + kk, o1 + oo) : 0.0;
            sb[kk][oo] = in ? value(kc + kk, o2 + oo) : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < GK; kk++) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) a[i] = sa[kk][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < 4; j++) b[j] = sb[kk][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t r = o1 + ty + 16 * i, c = o2 + tx + 16 * j;
            if (r < n_out && c < n_out) atomicAdd(out + r * n_out + c, acc[i][j]);
        }
}

__global__ void k_mean64(const double *__restrict__ colsum, int n_genes, double inv_n, double *__restrict__ mean) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_genes) mean[g] = colsum[g] * inv_n;
}

// emb[r][c] = sum_g (D[r][g] - mean[g]) V[g][c]  (float64 accumulate, float32 out, zero beyond n_comp): 8 rows x 32
// components per CTA, the components streamed through shared memory in chunks of 128 genes
constexpr int PR = 8, PG = 128;
__global__ void __launch_bounds__(256) k_project(const float *__restrict__ D, int64_t ld, int64_t n_rows, int n_genes,
                                                 const double *__restrict__ mean, const double *__restrict__ V, int n_comp,
                                                 float *__restrict__ emb, int KP) {
    __shared__ double sv[PG][32];
    __shared__ double sd[PR][PG];
    const int c = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * PR + rl;
    for (int c0 = 0; c0 < KP; c0 += 32) {
        double acc = 0.0;
        for (int g0 = 0; g0 < n_genes; g0 += PG) {
            for (int e = threadIdx.x; e < PG * 32; e += 256) {
                const int g = g0 + (e >> 5), cc = c0 + (e & 31);
                sv[e >> 5][e & 31] = (g < n_genes && cc < n_comp) ? V[(int64_t)g * n_comp + cc] : 0.0;
            }
            for (int e = threadIdx.x; e < PR * PG; e += 256) {
                const int64_t r = (int64_t)blockIdx.x * PR + e / PG;
                const int g = g0 + e % PG;
                sd[e / PG][e % PG] = (r < n_rows && g < n_genes) ? (double)D[r * ld + g] - mean[g] : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int g = 0; g < PG; g++) acc = fma(sd[rl][g], sv[g][c], acc);
            __syncthreads();
        }
        if (row < n_rows) emb[row * KP + c0 + c] = (float)acc;
    }
}

}  // namespace

// Float64 Gram matrix of the centred dense matrix on one side: transposed == 0 -> G x G (sum over cells: (A - 1) times the
// covariance matrix sklearn's covariance_eigh factorises), transposed != 0 -> A x A (sum over genes).  out: row-major,
// symmetric, n x n with n = G resp. A.
extern "C" int dd_centered_gram(dd_handle *h, int32_t transposed, double *out) {
    if (!h || !out) return dd_fail(h, DD_ERR_ARG, "dd_centered_gram: null argument");
    if (!h->dense_valid) return dd_fail(h, DD_ERR_ARG, "dd_centered_gram: no dense matrix (call dd_normalise_log first)");
    if (dd_sharded(h)) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_centered_gram: not available on a cell-sharded handle");
    DD_CUDA(h, cudaSetDevice(h->device));
    const int64_t A = h->A, G = h->G;
    const int64_t n = transposed ? A : G, red = transposed ? G : A;
    if (n > 16384) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_centered_gram: more than 16384 rows / columns on the Gram side");
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_colstats(h, false));
    double *d_mean = nullptr, *d_out = nullptr;
    DD_CUDA(h, cudaMalloc(&d_mean, sizeof(double) * G));
    if (cudaMalloc(&d_out, sizeof(double) * n * n) != cudaSuccess) {
        cudaFree(d_mean);
        return dd_fail(h, DD_ERR_NOMEM, "dd_centered_gram: device buffers");
    }
    cudaMemsetAsync(d_out, 0, sizeof(double) * n * n, h->stream);
    int rc = DD_OK;
    do {
        dd_launch_begin(h);
        k_mean64<<<(unsigned)((G + 255) / 256), 256, 0, h->stream>>>(h->d_colsum, (int)G, 1.0 / (double)A, d_mean);
        if ((rc = dd_launch_end(h, "mean64")) != DD_OK) break;
        const int tiles = (int)((n + GT - 1) / GT);
        const int64_t upper = (int64_t)tiles * (tiles + 1) / 2;
        int splits = (int)std::max<int64_t>(1, std::min<int64_t>((4 * h->num_sms + upper - 1) / upper, (red + GK - 1) / GK));
        int64_t per = ((red + splits - 1) / splits + GK - 1) / GK * GK;
        splits = (int)((red + per - 1) / per);
        dd_launch_begin(h);
        if (transposed)
            k_gram64<true><<<dim3(tiles, tiles, splits), 256, 0, h->stream>>>(h->d_dense, h->ld, A, (int)G, d_mean, d_out, (int)n, red, per);
        else
            k_gram64<false><<<dim3(tiles, tiles, splits), 256, 0, h->stream>>>(h->d_dense, h->ld, A, (int)G, d_mean, d_out, (int)n, red, per);
        if ((rc = dd_launch_end(h, "gram64")) != DD_OK) break;
        if ((rc = dd_stage_end(h, "gram")) != DD_OK) break;
        if (cudaMemcpyAsync(out, d_out, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
            cudaStreamSynchronize(h->stream) != cudaSuccess)
            rc = dd_fail(h, DD_ERR_CUDA, "dd_centered_gram: copy back failed");
    } while (false);
    cudaFree(d_mean);
    cudaFree(d_out);
    if (rc != DD_OK) return rc;
    for (int64_t r = 0; r < n; r++)  // mirror the upper triangle
        for (int64_t c = 0; c < r; c++) out[r * n + c] = out[c * n + r];
    return DD_OK;
}

// X_pca = (D - mean) V for components V (float64, G x n_comp row-major, already sign-fixed): leaves the float32 A x n_comp
// embedding on the device for dd_knn (like dd_pca); emb_out (A x n_comp float32) may be NULL.
extern "C" int dd_project(dd_handle *h, int32_t n_comp, const double *components, float *emb_out) {
    if (!h || !components) return dd_fail(h, DD_ERR_ARG, "dd_project: null argument");
    if (!h->dense_valid) return dd_fail(h, DD_ERR_ARG, "dd_project: no dense matrix (call dd_normalise_log first)");
    if (dd_sharded(h)) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_project: not available on a cell-sharded handle");
    if (n_comp < 1 || n_comp > 64) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_project: n_components must be in [1, 64]");
    DD_CUDA(h, cudaSetDevice(h->device));
    const int64_t A = h->A, G = h->G;
    const int KP = n_comp <= 32 ? 32 : 64;
    DD_TRY(dd_emb_reserve(h, A, KP));
    h->C = n_comp;
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_colstats(h, false));
    double *d_mean = nullptr, *d_v = nullptr;
    DD_CUDA(h, cudaMalloc(&d_mean, sizeof(double) * G));
    if (cudaMalloc(&d_v, sizeof(double) * G * n_comp) != cudaSuccess) {
        cudaFree(d_mean);
        return dd_fail(h, DD_ERR_NOMEM, "dd_project: device buffers");
    }
    int rc = DD_OK;
    do {
        if (cudaMemcpyAsync(d_v, components, sizeof(double) * G * n_comp, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
            rc = dd_fail(h, DD_ERR_CUDA, "dd_project: upload failed");
            break;
        }
        dd_launch_begin(h);
        k_mean64<<<(unsigned)((G + 255) / 256), 256, 0, h->stream>>>(h->d_colsum, (int)G, 1.0 / (double)A, d_mean);
        if ((rc = dd_launch_end(h, "mean64")) != DD_OK) break;
        dd_launch_begin(h);
        k_project<<<(unsigned)((A + PR - 1) / PR), 256, 0, h->stream>>>(h->d_dense, h->ld, A, (int)G, d_mean, d_v, n_comp, h->d_emb, KP);
        if ((rc = dd_launch_end(h, "project")) != DD_OK) break;
        if ((rc = dd_stage_end(h, "pca")) != DD_OK) break;
        if (emb_out &&
            cudaMemcpy2DAsync(emb_out, sizeof(float) * n_comp, h->d_emb, sizeof(float) * KP, sizeof(float) * n_comp, A,
                              cudaMemcpyDeviceToHost, h->stream) != cudaSuccess)
            rc = dd_fail(h, DD_ERR_CUDA, "dd_project: copy back failed");
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) rc = dd_fail(h, DD_ERR_CUDA, "dd_project: device error");
    } while (false);
    cudaFree(d_mean);
    cudaFree(d_v);
    if (rc != DD_OK) return rc;
    h->emb_rows = A;
    h->emb_valid = true;
    return DD_OK;
}
