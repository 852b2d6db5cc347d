// hvg.cu -- the highly-variable-gene selection of fit()'s prologue (doubletdetection.py:165-176) on the device:
//
//   gene_variances = raw.power(2).mean(axis=0) - raw.mean(axis=0) ** 2        (float32, scipy)
//   top_var_genes_ = np.argsort(gene_variances)[-n_top:]                      (host: numpy's own sort decides ties)
//   raw = raw.tocsc()[:, top_var_genes_].tocsr()                              (columns in ascending-variance order, Q1)
//
// scipy computes both means as  ones(1, N) @ (raw * float32(1 / N))  with csc_matvecs: ONE float32 accumulator per gene,
// the entries of a gene added in ROW ORDER, each multiplied by float32(1 / N) first.  To reproduce `top_var_genes_` bit for
// bit the device does exactly that: the values are brought into column-major order by a STABLE counting sort (rows stay in
// order inside a gene), then one thread per gene adds them sequentially with explicit round-to-nearest float32 operations.
// The argsort stays on the host (G floats) so that ties break exactly as in the reference; the column subset that follows
// (re-numbered to the position in top_var_genes_, rows re-sorted by the new column id) is done here again.
//
//   k_hvg_hist      per (row block, gene) entry counts                 HBM, nnz * 4 B read + atomics into a B x G table
//   k_hvg_bases     per gene: exclusive scan over the row blocks       B x G * 4 B
//   k_hvg_scatter   one CTA per row block, its rows one after the other: values to column-major positions
//   k_hvg_moments   one thread per gene: sequential float32 sums       nnz * 4 B read
//   k_sel_count / k_sel_fill   one warp per row: selected entries, ordered by the new column id through a bitmap
#include "dd_internal.h"

#include <algorithm>
#include <vector>

namespace {

constexpr int kRowsPerBlock = 128;

__global__ void k_hvg_hist(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, int64_t n_rows, int n_genes,
                           int32_t *__restrict__ table) {
    const int b = blockIdx.x;
    const int64_t r0 = (int64_t)b * kRowsPerBlock, r1 = min(r0 + kRowsPerBlock, n_rows);
    const int s = indptr[r0], e = indptr[r1];
    int32_t *row = table + (int64_t)b * n_genes;
    for (int p = s + threadIdx.x; p < e; p += blockDim.x) atomicAdd(row + indices[p], 1);
}

// table[b][g]: count -> first column-major position of block b's entries of gene g, RELATIVE to the gene's start;
// col_count[g] = entries of the gene
__global__ void k_hvg_bases(int32_t *__restrict__ table, int n_blocks, int n_genes, int32_t *__restrict__ col_count) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_genes) return;
    int run = 0;
    for (int b = 0; b < n_blocks; b++) {
        const int64_t at = (int64_t)b * n_genes + g;
        const int c = table[at];
        table[at] = run;
        run += c;
    }
    col_count[g] = run;
}

// rows of the block strictly one after the other (a row's entries have distinct genes: no conflict inside a row)
__global__ void k_hvg_scatter(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                              const float *__restrict__ data, int64_t n_rows, int n_genes, int32_t *table,
                              const int32_t *__restrict__ col_start, float *__restrict__ col_val) {
    const int b = blockIdx.x;
    const int64_t r0 = (int64_t)b * kRowsPerBlock, r1 = min(r0 + kRowsPerBlock, n_rows);
    int32_t *row = table + (int64_t)b * n_genes;
    for (int64_t r = r0; r < r1; r++) {
        const int s = indptr[r], e = indptr[r + 1];
        for (int p = s + threadIdx.x; p < e; p += blockDim.x) {
            const int g = indices[p];
            const int pos = row[g];
            row[g] = pos + 1;
            col_val[(int64_t)col_start[g] + pos] = data[p];
        }
        __syncthreads();  // the next row must see this row's cursor updates (same CTA, global memory)
    }
}

// scipy: (raw * float32(1/N)) summed in row order; power(2) first for the second moment; var = m2 - m1 * m1 (float32)
__global__ void k_hvg_moments(const int32_t *__restrict__ col_start, const float *__restrict__ col_val, int n_genes, float inv_n,
                              float *__restrict__ var_out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_genes) return;
    const int s = col_start[g], e = col_start[g + 1];
    float m1 = 0.f, m2 = 0.f;
    int p = s;
    for (; p + 4 <= e; p += 4) {  // four loads in flight, the additions strictly in order
        const float v0 = col_val[p], v1 = col_val[p + 1], v2 = col_val[p + 2], v3 = col_val[p + 3];
        m1 = __fadd_rn(m1, __fmul_rn(v0, inv_n));
        m2 = __fadd_rn(m2, __fmul_rn(__fmul_rn(v0, v0), inv_n));
        m1 = __fadd_rn(m1, __fmul_rn(v1, inv_n));
        m2 = __fadd_rn(m2, __fmul_rn(__fmul_rn(v1, v1), inv_n));
        m1 = __fadd_rn(m1, __fmul_rn(v2, inv_n));
        m2 = __fadd_rn(m2, __fmul_rn(__fmul_rn(v2, v2), inv_n));
        m1 = __fadd_rn(m1, __fmul_rn(v3, inv_n));
        m2 = __fadd_rn(m2, __fmul_rn(__fmul_rn(v3, v3), inv_n));
    }
    for (; p < e; p++) {
        const float v = col_val[p];
        m1 = __fadd_rn(m1, __fmul_rn(v, inv_n));
        m2 = __fadd_rn(m2, __fmul_rn(__fmul_rn(v, v), inv_n));
    }
    var_out[g] = __fsub_rn(m2, __fmul_rn(m1, m1));
}

// ---- column subset -------------------------------------------------------------------------------------------------
__global__ void k_sel_count(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                            const int32_t *__restrict__ new_id, int64_t n_rows, int32_t *__restrict__ count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        int c = 0;
        for (int p = indptr[r] + lane; p < indptr[r + 1]; p += 32) c += new_id[indices[p]] >= 0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) count[r] = c;
    }
}

// one warp per row: bit j of the warp's bitmap = "new column j occurs in this row"; an entry's place in the output row is
// the number of set bits below its own (rows of the result have sorted, duplicate-free column ids, like .tocsr())
__global__ void k_sel_fill(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                           const int32_t *__restrict__ new_id, int64_t n_rows, int n_words,
                           const int32_t *__restrict__ out_indptr, int32_t *__restrict__ out_indices,
                           float *__restrict__ out_data) {
    extern __shared__ uint32_t s_bits[];  // per warp: n_words bitmap words + n_words prefix counts
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    uint32_t *bits = s_bits + (size_t)wl * 2 * n_words;
    uint32_t *pre = bits + n_words;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wl;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        const int s = indptr[r], e = indptr[r + 1];
        const int o = out_indptr[r];
        if (out_indptr[r + 1] == o) continue;  // whole warp
        for (int w = lane; w < n_words; w += 32) bits[w] = 0u;
        __syncwarp();
        for (int p = s + lane; p < e; p += 32) {
            const int j = new_id[indices[p]];
            if (j >= 0) atomicOr(bits + (j >> 5), 1u << (j & 31));
        }
        __syncwarp();
        int carry = 0;
        for (int w0 = 0; w0 < n_words; w0 += 32) {  // exclusive prefix of the word popcounts
            const int w = w0 + lane;
            const int c = w < n_words ? __popc(bits[w]) : 0;
            int incl = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += u;
            }
            if (w < n_words) pre[w] = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        for (int p = s + lane; p < e; p += 32) {
            const int j = new_id[indices[p]];
            if (j < 0) continue;
            const int at = o + (int)pre[j >> 5] + __popc(bits[j >> 5] & ((1u << (j & 31)) - 1u));
            out_indices[at] = j;
            out_data[at] = data[p];
        }
        __syncwarp();
    }
}

__global__ void k_scan_i32(const int32_t *__restrict__ count, int64_t n, int32_t *__restrict__ out) {
    // exclusive scan by one CTA of 1024 threads (n <= a few 10^6): thread t owns a contiguous slice
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int64_t per = (n + 1023) / 1024;
    const int64_t b = min((int64_t)t * per, n), e = min(b + per, n);
    long long s = 0;
    for (int64_t i = b; i < e; i++) s += count[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; i++) {
            const long long v = part[i];
            part[i] = run;
            run += v;
        }
        out[n] = run > 0x7fffffffll ? -1 : (int32_t)run;
    }
    __syncthreads();
    long long run = part[t];
    for (int64_t i = b; i < e; i++) {
        out[i] = (int32_t)run;
        run += count[i];
    }
}

template <typename T>
struct Tmp {
    T *p = nullptr;
    ~Tmp() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t count) { return cudaMalloc((void **)&p, sizeof(T) * (count ? count : 1)); }
};

}  // namespace

int dd_finish_upload(dd_handle *h);  // csr.cu: library sizes of the CSR the handle holds

// gene_variances (:166-169) of the uploaded counts, float32, bit for bit what scipy returns.
extern "C" int dd_hvg_variances(dd_handle *h, float *var_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_hvg_variances: null handle");
    if (!var_out || !h->d_indptr) return dd_fail(h, DD_ERR_ARG, "dd_hvg_variances: upload the counts first");
    DD_CUDA(h, cudaSetDevice(h->device));
    const int64_t N = h->N, G = h->G, nnz = h->nnz;
    const int n_blocks = (int)((N + kRowsPerBlock - 1) / kRowsPerBlock);
    if ((int64_t)n_blocks * G >= (1ll << 33)) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_hvg_variances: matrix too large for the block table");
    Tmp<int32_t> table, col_start, col_count;
    Tmp<float> col_val, var;
    if (table.alloc((size_t)n_blocks * G) || col_start.alloc(G + 1) || col_count.alloc(G) || col_val.alloc(nnz) || var.alloc(G))
        return dd_fail(h, DD_ERR_NOMEM, "dd_hvg_variances: device buffers");
    DD_TRY(dd_stage_begin(h));
    DD_CUDA(h, cudaMemsetAsync(table.p, 0, sizeof(int32_t) * (size_t)n_blocks * G, h->stream));
    const int g_int = (int)G;
    DD_LAUNCH(h, "hvg_hist", k_hvg_hist, (unsigned)n_blocks, 256, 0, h->d_indptr, h->d_indices, N, g_int, table.p);
    DD_LAUNCH(h, "hvg_bases", k_hvg_bases, (unsigned)((G + 127) / 128), 128, 0, table.p, n_blocks, g_int, col_count.p);
    DD_LAUNCH(h, "hvg_scan", k_scan_i32, 1, 1024, 0, col_count.p, G, col_start.p);
    DD_LAUNCH(h, "hvg_scatter", k_hvg_scatter, (unsigned)n_blocks, 128, 0, h->d_indptr, h->d_indices, h->d_data, N, g_int, table.p,
              col_start.p, col_val.p);
    // scipy multiplies by the Python float 1.0 / N, which NumPy casts to the array's float32 first
    const float inv_n = (float)(1.0 / (double)N);
    DD_LAUNCH(h, "hvg_moments", k_hvg_moments, (unsigned)((G + 63) / 64), 64, 0, col_start.p, col_val.p, g_int, inv_n, var.p);
    DD_TRY(dd_stage_end(h, "hvg"));
    DD_CUDA(h, cudaMemcpyAsync(var_out, var.p, sizeof(float) * G, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

// raw = raw.tocsc()[:, genes].tocsr() (:172-176): column j of the result is gene genes[j] (top_var_genes_ keeps the
// ascending-variance order); rows come out with sorted column ids.  Replaces the handle's matrix and recomputes
// the library sizes (:182).
extern "C" int dd_select_genes(dd_handle *h, int64_t n_sel, const int64_t *genes) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_select_genes: null handle");
    if (!h->d_indptr) return dd_fail(h, DD_ERR_ARG, "dd_select_genes: upload the counts first");
    if (n_sel < 1 || !genes) return dd_fail(h, DD_ERR_ARG, "dd_select_genes: empty selection");
    if (h->counts_borrowed) return dd_fail(h, DD_ERR_ARG, "dd_select_genes: the count matrix belongs to another handle");
    const int64_t N = h->N, G = h->G;
    std::vector<int32_t> new_id((size_t)G, -1);
    for (int64_t j = 0; j < n_sel; j++) {
        if (genes[j] < 0 || genes[j] >= G) return dd_fail(h, DD_ERR_ARG, "dd_select_genes: gene index out of range");
        if (new_id[genes[j]] >= 0) return dd_fail(h, DD_ERR_ARG, "dd_select_genes: duplicate gene");
        new_id[genes[j]] = (int32_t)j;
    }
    DD_CUDA(h, cudaSetDevice(h->device));
    Tmp<int32_t> d_new_id, count, out_indptr;
    if (d_new_id.alloc(G) || count.alloc(N) || out_indptr.alloc(N + 1)) return dd_fail(h, DD_ERR_NOMEM, "dd_select_genes: device buffers");
    DD_CUDA(h, cudaMemcpyAsync(d_new_id.p, new_id.data(), sizeof(int32_t) * G, cudaMemcpyHostToDevice, h->stream));
    DD_TRY(dd_stage_begin(h));
    const int grid = h->num_sms * 8;
    DD_LAUNCH(h, "sel_count", k_sel_count, grid, 256, 0, h->d_indptr, h->d_indices, d_new_id.p, N, count.p);
    DD_LAUNCH(h, "sel_scan", k_scan_i32, 1, 1024, 0, count.p, N, out_indptr.p);
    int32_t total = 0;
    DD_CUDA(h, cudaMemcpyAsync(&total, out_indptr.p + N, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (total < 0) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_select_genes: nnz exceeds int32 indexing");
    int32_t *out_indices = nullptr;
    float *out_data = nullptr;
    // + 4 entries: the bulk-copy staging of a row reads whole 16-byte groups around it (as in dd_upload_counts)
    DD_CUDA(h, cudaMalloc(&out_indices, sizeof(int32_t) * (std::max<int64_t>(total, 1) + 4)));
    if (cudaMalloc(&out_data, sizeof(float) * (std::max<int64_t>(total, 1) + 4)) != cudaSuccess) {
        cudaFree(out_indices);
        return dd_fail(h, DD_ERR_NOMEM, "dd_select_genes: device buffers");
    }
    const int n_words = (int)((n_sel + 31) / 32);
    int warps = 8;
    while (warps > 1 && (size_t)warps * 2 * n_words * 4 > 96 * 1024) warps >>= 1;
    const size_t smem = (size_t)warps * 2 * n_words * 4;
    int rc = DD_OK;
    if (smem > 200 * 1024) rc = dd_fail(h, DD_ERR_UNSUPPORTED, "dd_select_genes: selection too wide for the per-row bitmap");
    if (rc == DD_OK && smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_sel_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        rc = dd_fail(h, DD_ERR_CUDA, "dd_select_genes: shared memory attribute");
    if (rc == DD_OK && total > 0) {
        dd_launch_begin(h);
        k_sel_fill<<<grid, warps * 32, smem, h->stream>>>(h->d_indptr, h->d_indices, h->d_data, d_new_id.p, N, n_words, out_indptr.p,
                                                          out_indices, out_data);
        rc = dd_launch_end(h, "sel_fill");
    }
    if (rc == DD_OK && cudaMemcpyAsync(h->d_indptr, out_indptr.p, sizeof(int32_t) * (N + 1), cudaMemcpyDeviceToDevice, h->stream) != cudaSuccess)
        rc = dd_fail(h, DD_ERR_CUDA, "dd_select_genes: indptr copy");
    if (rc == DD_OK && cudaStreamSynchronize(h->stream) != cudaSuccess) rc = dd_fail(h, DD_ERR_CUDA, "dd_select_genes: device error");
    if (rc != DD_OK) {
        cudaFree(out_indices);
        cudaFree(out_data);
        return rc;
    }
    cudaFree(h->d_indices);
    cudaFree(h->d_data);
    h->d_indices = out_indices;
    h->d_data = out_data;
    h->cap_nnz = std::max<int64_t>(total, 1);
    h->nnz = total;
    h->G = n_sel;
    h->ld = dd_round_up(n_sel, 32);
    h->synth_csr_valid = false; h->dense_valid = false; h->emb_valid = false; h->M = 0; h->A = 0;
    DD_TRY(dd_finish_upload(h));
    DD_TRY(dd_stage_end(h, "select_genes"));
    return DD_OK;
}

extern "C" int dd_counts_nnz(dd_handle *h, int64_t *nnz_out) {
    if (!h || !nnz_out || !h->d_indptr) return dd_fail(h, DD_ERR_ARG, "dd_counts_nnz: upload the counts first");
    *nnz_out = h->nnz;
    return DD_OK;
}

extern "C" int dd_download_counts(dd_handle *h, int32_t *indptr_out, int32_t *indices_out, float *data_out) {
    if (!h || !h->d_indptr || !indptr_out) return dd_fail(h, DD_ERR_ARG, "dd_download_counts: upload the counts first");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_CUDA(h, cudaMemcpyAsync(indptr_out, h->d_indptr, sizeof(int32_t) * (h->N + 1), cudaMemcpyDeviceToHost, h->stream));
    if (h->nnz > 0) {
        if (!indices_out || !data_out) return dd_fail(h, DD_ERR_ARG, "dd_download_counts: null output");
        DD_CUDA(h, cudaMemcpyAsync(indices_out, h->d_indices, sizeof(int32_t) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
        DD_CUDA(h, cudaMemcpyAsync(data_out, h->d_data, sizeof(float) * h->nnz, cudaMemcpyDeviceToHost, h->stream));
    }
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}
