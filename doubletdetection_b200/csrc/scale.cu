// scale.cu -- per-gene statistics of the dense augmented matrix and the optional
// sc.pp.scale(max_value=15) step (doubletdetection.py:302-303; upstream semantics restated in
// oracle/upstream.py:pp_scale -- float64 mean / variance (ddof=1), centre and divide rounded to float32
// after each step, clip to [-max, max]).
#include "dd_internal.h"

namespace {

constexpr int kColsPerCta = 256;  // 64 threads x float4
constexpr int kRowLanes = 4;
constexpr int kRowsPerCta = 256;

// column sums (and sums of float32-rounded squares) accumulated in double
template <bool WITH_SQ>
__global__ void k_colstats(const float *__restrict__ dense, int64_t n_rows, int ld, double *__restrict__ colsum,
                           double *__restrict__ colsumsq) {
    __shared__ double red[kRowLanes][kColsPerCta];
    __shared__ double redsq[WITH_SQ ? kRowLanes : 1][WITH_SQ ? kColsPerCta : 1];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int col = blockIdx.x * kColsPerCta + 4 * tx;
    const int64_t r0 = (int64_t)blockIdx.y * kRowsPerCta;
    const int64_t r1 = min(r0 + kRowsPerCta, n_rows);
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (col < ld) {
        for (int64_t r = r0 + ty; r < r1; r += kRowLanes) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(dense + r * ld + col));
            const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                s[i] += (double)x[i];
                if (WITH_SQ) q[i] += (double)__fmul_rn(x[i], x[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        red[ty][4 * tx + i] = s[i];
        if (WITH_SQ) redsq[ty][4 * tx + i] = q[i];
    }
    __syncthreads();
    const int t = ty * blockDim.x + tx;  // 256 threads, one column each
    const int c = blockIdx.x * kColsPerCta + t;
    if (c < ld) {
        double a = 0, b = 0;
#pragma unroll
        for (int y = 0; y < kRowLanes; y++) {
            a += red[y][t];
            if (WITH_SQ) b += redsq[y][t];
        }
        atomicAdd(colsum + c, a);
        if (WITH_SQ) atomicAdd(colsumsq + c, b);
    }
}

__global__ void k_scale_apply(float *__restrict__ dense, int64_t n_rows, int64_t n_total, int n_genes, int ld,
                              const double *__restrict__ colsum, const double *__restrict__ colsumsq,
                              float max_value) {
    const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (col >= ld) return;
    double mean[4], sd[4];
    const double n = (double)n_total;  // rows over all ranks (the statistics are all-reduced), n_rows are local
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int c = col + i;
        if (c < n_genes) {
            const double m = colsum[c] / n;
            const double msq = colsumsq[c] / n;
            double var = (msq - m * m) * (n / (n - 1.0));
            double s = sqrt(var);
            if (s == 0.0) s = 1.0;
            mean[i] = m;
            sd[i] = s;
        } else {
            mean[i] = 0.0;
            sd[i] = 1.0;
        }
    }
    const int64_t r0 = (int64_t)blockIdx.y * kRowsPerCta;
    const int64_t r1 = min(r0 + kRowsPerCta, n_rows);
    for (int64_t r = r0; r < r1; r++) {
        float4 *p = reinterpret_cast<float4 *>(dense + r * ld + col);
        float4 v = *p;
        float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (col + i < n_genes) {
                float y = (float)((double)x[i] - mean[i]);
                y = (float)((double)y / sd[i]);
                if (max_value > 0.f) y = fminf(fmaxf(y, -max_value), max_value);
                x[i] = y;
            }
        }
        *p = make_float4(x[0], x[1], x[2], x[3]);
    }
}

}  // namespace

int dd_dev_colstats(dd_handle *h, bool with_sq) {
    if (!h->dense_valid) return dd_fail(h, DD_ERR_ARG, "column statistics: no dense matrix");
    if (h->ld > h->cap_cols) {
        if (h->d_colsum) cudaFree(h->d_colsum);
        if (h->d_colsumsq) cudaFree(h->d_colsumsq);
        h->d_colsum = h->d_colsumsq = nullptr;
        h->cap_cols = 0;
        DD_CUDA(h, cudaMalloc(&h->d_colsum, sizeof(double) * h->ld));
        DD_CUDA(h, cudaMalloc(&h->d_colsumsq, sizeof(double) * h->ld));
        h->cap_cols = h->ld;
    }
    DD_CUDA(h, cudaMemsetAsync(h->d_colsum, 0, sizeof(double) * h->ld, h->stream));
    if (with_sq) DD_CUDA(h, cudaMemsetAsync(h->d_colsumsq, 0, sizeof(double) * h->ld, h->stream));
    dim3 grid((unsigned)((h->ld + kColsPerCta - 1) / kColsPerCta), (unsigned)((h->A + kRowsPerCta - 1) / kRowsPerCta));
    dim3 block(kColsPerCta / 4, kRowLanes);
    if (with_sq)
        DD_LAUNCH(h, "colstats_sq", k_colstats<true>, grid, block, 0, h->d_dense, h->A, (int)h->ld, h->d_colsum, h->d_colsumsq);
    else
        DD_LAUNCH(h, "colstats", k_colstats<false>, grid, block, 0, h->d_dense, h->A, (int)h->ld, h->d_colsum, h->d_colsumsq);
    // cell-block sharding: sums over the cells of all ranks
    DD_TRY(dd_comm_allreduce_f64(h, h->d_colsum, h->ld));
    if (with_sq) DD_TRY(dd_comm_allreduce_f64(h, h->d_colsumsq, h->ld));
    return DD_OK;
}

int dd_dev_standard_scale(dd_handle *h, float max_value) {
    if (h->A_glob < 2) return dd_fail(h, DD_ERR_ARG, "standard scaling needs at least two rows");
    DD_TRY(dd_dev_colstats(h, true));
    dim3 grid((unsigned)((h->ld / 4 + 63) / 64), (unsigned)((h->A + kRowsPerCta - 1) / kRowsPerCta));
    DD_LAUNCH(h, "scale_apply", k_scale_apply, grid, 64, 0, h->d_dense, h->A, h->A_glob, (int)h->G, (int)h->ld, h->d_colsum,
              h->d_colsumsq, max_value);
    h->emb_valid = false;
    return DD_OK;
}

extern "C" int dd_standard_scale(dd_handle *h, float max_value) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_standard_scale: null handle");
    if (!h->dense_valid) return dd_fail(h, DD_ERR_ARG, "dd_standard_scale: no dense matrix");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_standard_scale(h, max_value));
    DD_TRY(dd_stage_end(h, "scale"));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}
