// csr.cu -- the sparse half of _one_fit: library sizes, _createDoublets (CSR row-pair add) and the
// fused "pair-add + L1 normalise + x median + log(x + pseudocount) -> dense" build.
//
// Reference lines (doubletdetection/doubletdetection.py):
//   :182-184  _lib_size, memoised L1-normalised originals      -> k_row_sums
//   :397-399  raw[parents[:,0]] + raw[parents[:,1]]             -> k_synth_count / k_synth_fill
//   :288-295  synth lib sizes, L1 normalise, vstack, x median, log(. + pc) dense
//                                                                -> k_dense_rows (originals and synthetics)
// All kernels are HBM-bound byte movers: one warp owns one row, stages it in a per-warp shared-memory
// row buffer (scatter in, 128-bit coalesced streaming stores out) and the grid is a multiple of the
// SM count with warps striding over rows.
#include "dd_internal.h"

#include <algorithm>
#include <cstdlib>

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMaxChunk = 8192;  // columns staged per pass (upper bound): 32 KB per warp

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// _lib_size (float32 row sums, :182) and the double-precision sum of |x| that sklearn's
// inplace_csr_row_normalize_l1 divides by (:184).  Integer-valued counts make both exact in any
// summation order.
__global__ void k_row_sums(const int32_t *__restrict__ indptr, const float *__restrict__ data, int64_t n_rows,
                           float *__restrict__ lib, double *__restrict__ l1, int *__restrict__ any_negative) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        const int s = indptr[row], e = indptr[row + 1];
        float acc = 0.f;
        double acc1 = 0.0;
        for (int p = s + lane; p < e; p += 32) {
            const float v = __ldg(data + p);
            acc += v;
            acc1 += fabs((double)v);
            if (v < 0.f) any_negative[0] = 1;
            if (!isfinite(v)) any_negative[1] = 1;  // NaN / inf: what check_array(ensure_all_finite=True) looks for (:149-155)
        }
        acc = warp_sum_f(acc);
        acc1 = warp_sum_d(acc1);
        if (lane == 0) {
            lib[row] = acc;
            l1[row] = acc1;
        }
    }
}

// Stage columns [c0, c0+cw) of (row a + row b) into buf.  buf is per-warp shared memory.
__device__ __forceinline__ void stage_pair_sum(float *buf, int c0, int cw, const int32_t *__restrict__ indices,
                                               const float *__restrict__ data, int sa, int ea, int sb, int eb,
                                               int lane) {
    for (int j = lane; j < cw; j += 32) buf[j] = 0.f;
    __syncwarp();
    for (int p = sa + lane; p < ea; p += 32) {
        const int c = __ldg(indices + p) - c0;
        if ((unsigned)c < (unsigned)cw) buf[c] = __ldg(data + p);
    }
    __syncwarp();
    for (int p = sb + lane; p < eb; p += 32) {
        const int c = __ldg(indices + p) - c0;
        if ((unsigned)c < (unsigned)cw) buf[c] += __ldg(data + p);  // columns are unique within a row
    }
    __syncwarp();
}

// Pass 1 of _createDoublets: nnz of every synthetic row (entries whose sum is non-zero, exactly the
// entries scipy's canonical csr_plus_csr keeps) and its float32 library size (:288).
__global__ void k_synth_count(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                              const float *__restrict__ data, const int64_t *__restrict__ parents,
                              int64_t n_synth, int n_genes, int chunk, int32_t *__restrict__ count,
                              float *__restrict__ slib) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *buf = smem + (size_t)w * chunk;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + w;
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;
    for (int64_t r = warp; r < n_synth; r += n_warps) {
        const int64_t pa = parents[2 * r], pb = parents[2 * r + 1];
        const int sa = indptr[pa], ea = indptr[pa + 1], sb = indptr[pb], eb = indptr[pb + 1];
        int cnt = 0;
        float sum = 0.f;
        for (int c0 = 0; c0 < n_genes; c0 += chunk) {
            const int cw = min(chunk, n_genes - c0);
            stage_pair_sum(buf, c0, cw, indices, data, sa, ea, sb, eb, lane);
            for (int j = lane; j < cw; j += 32) {
                const float v = buf[j];
                cnt += (v != 0.f);
                sum += v;
            }
            __syncwarp();
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        sum = warp_sum_f(sum);
        if (lane == 0) {
            count[r] = cnt;
            slib[r] = sum;
        }
    }
}

// Exclusive scan of count[0..n) into indptr[0..n]; single CTA (n is a few 10^5 at most).
__global__ void k_exclusive_scan(const int32_t *__restrict__ count, int64_t n, int32_t *__restrict__ indptr) {
    __shared__ long long part[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int64_t per = (n + nt - 1) / nt;
    const int64_t b = min((int64_t)t * per, n), e = min(b + per, n);
    long long s = 0;
    for (int64_t i = b; i < e; i++) s += count[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int i = 0; i < nt; i++) {
            const long long v = part[i];
            part[i] = run;
            run += v;
        }
        indptr[n] = (int32_t)run;
    }
    __syncthreads();
    long long run = part[t];
    for (int64_t i = b; i < e; i++) {
        indptr[i] = (int32_t)run;
        run += count[i];
    }
}

// Pass 2 of _createDoublets: write the merged rows in column order (canonical CSR).
__global__ void k_synth_fill(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                             const float *__restrict__ data, const int64_t *__restrict__ parents,
                             int64_t n_synth, int n_genes, int chunk, const int32_t *__restrict__ sindptr,
                             int32_t *__restrict__ sindices, float *__restrict__ sdata) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *buf = smem + (size_t)w * chunk;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + w;
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int64_t r = warp; r < n_synth; r += n_warps) {
        const int64_t pa = parents[2 * r], pb = parents[2 * r + 1];
        const int sa = indptr[pa], ea = indptr[pa + 1], sb = indptr[pb], eb = indptr[pb + 1];
        int out = sindptr[r];
        for (int c0 = 0; c0 < n_genes; c0 += chunk) {
            const int cw = min(chunk, n_genes - c0);
            stage_pair_sum(buf, c0, cw, indices, data, sa, ea, sb, eb, lane);
            for (int base = 0; base < cw; base += 32) {
                const int j = base + lane;
                const float v = j < cw ? buf[j] : 0.f;
                const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
                if (v != 0.f) {
                    const int pos = out + __popc(m & lt_mask);
                    sindices[pos] = c0 + j;
                    sdata[pos] = v;
                }
                out += __popc(m);
            }
            __syncwarp();
        }
    }
}

// log( (x / l1) * median + pc ) with the reference's float32 rounding after every step:
//   sklearn divides in double and stores float32 (sparsefuncs_fast.pyx:516-545), scipy multiplies the
//   float32 data by the float32 median (:293), numpy adds the pseudocount and takes the float32 log (:295).
__device__ __forceinline__ float norm_log(float x, double l1, float median, float pc) {
    const float normed = l1 != 0.0 ? (float)((double)x / l1) : x;
    const float scaled = __fmul_rn(normed, median);
    // pseudocount == 1 is the reference's sparse branch: np.log1p on the stored entries (:296-297), log(1) = 0 elsewhere
    return pc == 1.0f ? log1pf(scaled) : logf(__fadd_rn(scaled, pc));
}

// The same with the row's reciprocal 1 / l1 hoisted out of the loop: q = x * inv corrected by one Newton step on the residual
// (r = x - q l1 exactly, by FMA; q + r inv) is the correctly rounded double quotient x / l1 for the integer-valued counts
// and row sums of this path (and within 1e-16 relative of it in general -- invisible after the rounding to float32), at
// three float64 instructions instead of the ~35 of a float64 division.  The dense build is instruction-bound on it.
__device__ __forceinline__ float norm_log_inv(float x, double l1, double inv, float median, float pc) {
    float normed = x;
    if (l1 != 0.0) {
        const double xd = (double)x;
        const double q = xd * inv;
        normed = (float)fma(fma(-q, l1, xd), inv, q);
    }
    const float scaled = __fmul_rn(normed, median);
    return pc == 1.0f ? log1pf(scaled) : logf(__fadd_rn(scaled, pc));
}

// lower_bound of `col` in the sorted index range [s, e); returns e if absent / position of first >= col
__device__ __forceinline__ int lower_bound_idx(const int32_t *__restrict__ indices, int s, int e, int col) {
    while (s < e) {
        const int m = (s + e) >> 1;
        if (__ldg(indices + m) < col) s = m + 1; else e = m;
    }
    return s;
}

// Dense rows of the augmented matrix, originals (row < n_cells) and synthetics in one launch.
// One warp per row.  The per-warp shared-memory row buffer always holds log(pc) (0 in the pad columns); the
// warp scatters the TRANSFORMED non-zeros into it (one division + log per stored entry, all lanes busy),
// streams the buffer out with 128-bit stores and restores the constant on the way.  A synthetic row is
// the sorted merge of its two parent rows: every entry looks its column up in the other parent by binary
// search (the rows sit in L1/L2), so no raw staging pass is needed.
__global__ void k_dense_rows(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                             const float *__restrict__ data, const double *__restrict__ l1_rows,
                             const int64_t *__restrict__ parents, int64_t n_cells, int64_t n_synth, int n_genes,
                             int ld, int chunk, float median, float pc, int l1_additive,
                             float *__restrict__ dense) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *buf = smem + (size_t)w * chunk;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerCta + w;
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerCta;
    const int64_t n_rows = n_cells + n_synth;
    const float logpc = logf(pc);
    const int n_chunks = (ld + chunk - 1) / chunk;
    if (n_chunks == 1) {  // the buffer keeps its constant fill across rows
        for (int j = lane; j < chunk; j += 32) buf[j] = j < n_genes ? logpc : 0.f;
        __syncwarp();
    }
    for (int64_t row = warp; row < n_rows; row += n_warps) {
        const bool synth = row >= n_cells;
        int sa, ea, sb = 0, eb = 0;
        double l1;
        if (!synth) {
            sa = indptr[row];
            ea = indptr[row + 1];
            l1 = l1_rows[row];
        } else {
            const int64_t r = row - n_cells;
            const int64_t pa = parents[2 * r], pb = parents[2 * r + 1];
            sa = indptr[pa];
            ea = indptr[pa + 1];
            sb = indptr[pb];
            eb = indptr[pb + 1];
            if (l1_additive) {
                // non-negative counts: |a + b| = |a| + |b|, exact in double for integer-valued data
                l1 = l1_rows[pa] + l1_rows[pb];
            } else {
                double acc = 0.0;
                for (int p = sa + lane; p < ea; p += 32) {
                    const int c = __ldg(indices + p);
                    const int q = lower_bound_idx(indices, sb, eb, c);
                    float v = __ldg(data + p);
                    if (q < eb && __ldg(indices + q) == c) v += __ldg(data + q);
                    acc += fabs((double)v);
                }
                for (int p = sb + lane; p < eb; p += 32) {
                    const int c = __ldg(indices + p);
                    const int q = lower_bound_idx(indices, sa, ea, c);
                    if (!(q < ea && __ldg(indices + q) == c)) acc += fabs((double)__ldg(data + p));
                }
                l1 = warp_sum_d(acc);
            }
        }
        float *out_row = dense + row * (int64_t)ld;
        for (int c0 = 0; c0 < ld; c0 += chunk) {
            const int cw = min(chunk, ld - c0);            // multiple of 32
            const int cg = max(0, min(cw, n_genes - c0));  // real gene columns in this chunk
            if (n_chunks > 1) {
                for (int j = lane; j < cw; j += 32) buf[j] = j < cg ? logpc : 0.f;
                __syncwarp();
            }
            // scatter the transformed non-zeros of this column range
            for (int p = sa + lane; p < ea; p += 32) {
                const int col = __ldg(indices + p);
                const int c = col - c0;
                if ((unsigned)c < (unsigned)cg) {
                    float v = __ldg(data + p);
                    if (synth) {
                        const int q = lower_bound_idx(indices, sb, eb, col);
                        if (q < eb && __ldg(indices + q) == col) v += __ldg(data + q);
                    }
                    if (v != 0.f) buf[c] = norm_log(v, l1, median, pc);
                }
            }
            if (synth) {
                for (int p = sb + lane; p < eb; p += 32) {
                    const int col = __ldg(indices + p);
                    const int c = col - c0;
                    if ((unsigned)c < (unsigned)cg) {
                        const int q = lower_bound_idx(indices, sa, ea, col);
                        if (!(q < ea && __ldg(indices + q) == col)) {  // columns of both parents were done above
                            const float v = __ldg(data + p);
                            if (v != 0.f) buf[c] = norm_log(v, l1, median, pc);
                        }
                    }
                }
            }
            __syncwarp();
            // stream out; with a single chunk the buffer is reset to its constant fill for the next row
            for (int j = 4 * lane; j < cw; j += 128) {
                const float4 v = *reinterpret_cast<const float4 *>(buf + j);
                __stcs(reinterpret_cast<float4 *>(out_row + c0 + j), v);
                if (n_chunks == 1)
                    *reinterpret_cast<float4 *>(buf + j) =
                        make_float4(j < cg ? logpc : 0.f, j + 1 < cg ? logpc : 0.f, j + 2 < cg ? logpc : 0.f,
                                    j + 3 < cg ? logpc : 0.f);
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Dense rows without shared-memory staging (the default).  The matrix is ~85-90 % the constant log(pc), so a
// row is (1) a coalesced 128-bit fill of the constant, (2) a scatter of the transformed non-zeros on top of it.
// Both land in L2 (the row was written a few hundred cycles earlier), which merges them before the lines are
// evicted to HBM once: DRAM traffic stays the algorithmic 8 B per stored entry + 4 B per dense element, there is
// no per-warp row buffer, hence no column chunking and full occupancy.
// A synthetic row merges its two parents THROUGH the row itself instead of searching:
//   phase 1  parent a's entries leave a tag (a NaN pattern carrying the entry's offset) at their columns;
//   phase 2  parent b's entries read their column back: tag -> both parents have it, write T(va + vb);
//            constant -> only b has it, write T(vb);
//   phase 3  parent a's entries whose tag is still there are a-only: write T(va).
// T(v) = log(v / l1 * median + pc) can be finite, +-inf or the canonical NaN, never a tag (exponent all ones,
// quiet bit clear, non-zero payload), and the fill log(pc) cannot be one either.  __syncwarp() orders the
// phases (the row belongs to one warp); the read-backs bypass L1 (ld.global.cg).
// Rows are dealt synthetic-first: they are the long jobs.
// Cell-block sharding: the launch covers originals [n0, n0 + n_loc) and synthetics [m0, m0 + m_loc) and writes
// them as local rows [0, n_loc + m_loc).
constexpr uint32_t kTagBits = 0x7f800000u, kTagMask = 0xffc00000u, kTagPayload = 0x003fffffu;
__device__ __forceinline__ bool is_tag(uint32_t x) { return (x & kTagMask) == kTagBits && (x & kTagPayload) != 0u; }

__global__ void __launch_bounds__(256) k_dense_rows_l2(const int32_t *__restrict__ indptr,
                                                       const int32_t *__restrict__ indices,
                                                       const float *__restrict__ data,
                                                       const double *__restrict__ l1_rows,
                                                       const int64_t *__restrict__ parents, int64_t n0, int64_t n_loc,
                                                       int64_t m0, int64_t m_loc, int n_genes, int ld, float median,
                                                       float pc, float *__restrict__ dense) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t n_rows = n_loc + m_loc;
    const float logpc = logf(pc);
    for (int64_t w = warp; w < n_rows; w += n_warps) {
        const bool synth = w < m_loc;
        const int64_t row = synth ? n_loc + w : w - m_loc;  // local output row
        float *out_row = dense + row * (int64_t)ld;
        int sa, ea, sb = 0, eb = 0;
        double l1;
        if (!synth) {
            const int64_t src = n0 + row;
            sa = __ldg(indptr + src);
            ea = __ldg(indptr + src + 1);
            l1 = __ldg(l1_rows + src);
        } else {
            const int64_t r = m0 + w;
            const int64_t pa = __ldg(parents + 2 * r), pb = __ldg(parents + 2 * r + 1);
            sa = __ldg(indptr + pa);
            ea = __ldg(indptr + pa + 1);
            sb = __ldg(indptr + pb);
            eb = __ldg(indptr + pb + 1);
            l1 = __ldg(l1_rows + pa) + __ldg(l1_rows + pb);  // non-negative counts: |a + b| = |a| + |b|, exact
        }
        // (1) constant fill, pad columns zero
        for (int j = 4 * lane; j < ld; j += 128) {
            float4 v;
            v.x = j < n_genes ? logpc : 0.f;
            v.y = j + 1 < n_genes ? logpc : 0.f;
            v.z = j + 2 < n_genes ? logpc : 0.f;
            v.w = j + 3 < n_genes ? logpc : 0.f;
            *reinterpret_cast<float4 *>(out_row + j) = v;
        }
        __syncwarp();
        if (!synth) {
#pragma unroll 4
            for (int p = sa + lane; p < ea; p += 32) {
                const int col = __ldg(indices + p);
                const float v = __ldg(data + p);
                if (v != 0.f) out_row[col] = norm_log(v, l1, median, pc);
            }
        } else {
            uint32_t *out_bits = reinterpret_cast<uint32_t *>(out_row);
#pragma unroll 4
            for (int p = sa + lane; p < ea; p += 32) out_bits[__ldg(indices + p)] = kTagBits | (uint32_t)(p - sa + 1);
            __syncwarp();
#pragma unroll 4
            for (int p = sb + lane; p < eb; p += 32) {
                const int col = __ldg(indices + p);
                float v = __ldg(data + p);
                const uint32_t x = __ldcg(out_bits + col);
                if (is_tag(x)) v += __ldg(data + sa + (int)(x & kTagPayload) - 1);
                out_row[col] = v != 0.f ? norm_log(v, l1, median, pc) : logpc;
            }
            __syncwarp();
#pragma unroll 4
            for (int p = sa + lane; p < ea; p += 32) {
                const int col = __ldg(indices + p);
                const float v = __ldg(data + p);
                if (is_tag(__ldcg(out_bits + col))) out_row[col] = v != 0.f ? norm_log(v, l1, median, pc) : logpc;
            }
        }
        // the next row of this warp touches other addresses: no barrier needed here
    }
}

// ------------------------------------------------------------------------------------------------
// Dense rows with the CSR segments STAGED BY THE TMA ENGINE (DD_DENSE_V=2).  Same row algorithm as k_dense_rows_l2
// (constant fill + scatter through L2, synthetic rows merged through the row), but the gathers that bound that
// kernel are taken off the warps: every warp owns a ring of kTmaDepth row slots in shared memory, lane 0 issues
// cp.async.bulk copies of the next rows' index / value segments (whole 16-byte groups around the segment) completing on
// one mbarrier per slot, and the warp only ever waits for data that was requested kTmaDepth rows earlier.  Few warps
// (8 per SM) keep ~60 KB of CSR reads in flight per SM; rows whose segments exceed a slot fall back to direct loads.
constexpr int kTmaWarps = 8, kTmaDepth = 3, kTmaSeg = 520;                  // entries per staged segment (16-byte multiple)
constexpr int kTmaSlotBytes = 4 * kTmaSeg * 4;                               // idx a | val a | idx b | val b
constexpr int kTmaWarpBytes = kTmaDepth * kTmaSlotBytes;
struct TmaRowDesc {
    double l1;
    long long row;        // local output row
    int sa, na, sb, nb;   // global start / length of the (one or two) source segments
    int oa, ob;           // offset of the first real entry inside the staged groups
    int staged, synth;
};
constexpr int kTmaTable = 64;  // row descriptors per warp: two halves of 32, refilled by all lanes together
constexpr size_t kTmaSmemBytes =
    (size_t)kTmaWarps * kTmaWarpBytes + (size_t)kTmaWarps * kTmaTable * sizeof(TmaRowDesc) + kTmaWarps * kTmaDepth * 8 + 64;

__device__ __forceinline__ uint32_t csr_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void csr_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(csr_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void csr_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(csr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void csr_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = csr_smem_u32(bar);
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
__device__ __forceinline__ void csr_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     csr_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(csr_smem_u32(bar))
                 : "memory");
}

// one row from (index, value) lists that live in shared memory (staged) or in global memory (fallback)
__device__ __forceinline__ void dense_row_from_lists(const int32_t *ia, const float *da, int na, const int32_t *ib,
                                                     const float *db, int nb, bool synth, double l1, float median, float pc,
                                                     float logpc, int n_genes, int ld, float *out_row, int lane, int dbg = 0) {
    if (!(dbg & 1))  // timing experiments only (DD_DENSE_DBG): 1 = no fill, 2 = no scatter
    for (int j = 4 * lane; j < ld; j += 128) {
        float4 v;
        v.x = j < n_genes ? logpc : 0.f;
        v.y = j + 1 < n_genes ? logpc : 0.f;
        v.z = j + 2 < n_genes ? logpc : 0.f;
        v.w = j + 3 < n_genes ? logpc : 0.f;
        *reinterpret_cast<float4 *>(out_row + j) = v;
    }
    __syncwarp();
    if (dbg & 2) return;
    if (!synth) {
#pragma unroll 4
        for (int p = lane; p < na; p += 32) {
            const float v = da[p];
            if (v != 0.f) out_row[ia[p]] = norm_log(v, l1, median, pc);
        }
        return;
    }
    uint32_t *out_bits = reinterpret_cast<uint32_t *>(out_row);
#pragma unroll 4
    for (int p = lane; p < na; p += 32) out_bits[ia[p]] = kTagBits | (uint32_t)(p + 1);
    __syncwarp();
#pragma unroll 4
    for (int p = lane; p < nb; p += 32) {
        const int col = ib[p];
        float v = db[p];
        const uint32_t x = __ldcg(out_bits + col);
        if (is_tag(x)) v += da[(int)(x & kTagPayload) - 1];
        out_row[col] = v != 0.f ? norm_log(v, l1, median, pc) : logpc;
    }
    __syncwarp();
#pragma unroll 4
    for (int p = lane; p < na; p += 32) {
        const int col = ia[p];
        const float v = da[p];
        if (is_tag(__ldcg(out_bits + col))) out_row[col] = v != 0.f ? norm_log(v, l1, median, pc) : logpc;
    }
}

__global__ void __launch_bounds__(kTmaWarps * 32, 1)
    k_dense_rows_tma(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                     const double *__restrict__ l1_rows, const int64_t *__restrict__ parents, int64_t n0, int64_t n_loc,
                     int64_t m0, int64_t m_loc, int n_genes, int ld, float median, float pc, float *__restrict__ dense, int dbg) {
    extern __shared__ __align__(128) uint8_t tma_smem[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    uint8_t *ring = tma_smem + (size_t)wl * kTmaWarpBytes;
    TmaRowDesc *table = reinterpret_cast<TmaRowDesc *>(tma_smem + (size_t)kTmaWarps * kTmaWarpBytes) + wl * kTmaTable;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tma_smem + (size_t)kTmaWarps * kTmaWarpBytes +
                                                  (size_t)kTmaWarps * kTmaTable * sizeof(TmaRowDesc)) + wl * kTmaDepth;
    if (lane == 0) {
        for (int d = 0; d < kTmaDepth; d++) csr_mbar_init(bars + d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t gw = (int64_t)blockIdx.x * kTmaWarps + wl, n_warps = (int64_t)gridDim.x * kTmaWarps;
    const int64_t n_rows = n_loc + m_loc;
    const int64_t my_rows = gw < n_rows ? (n_rows - gw + n_warps - 1) / n_warps : 0;
    const float logpc = logf(pc);

    // all lanes: describe 32 of this warp's rows at once (the dependent parents -> indptr loads run in parallel)
    auto describe = [&](int64_t k_first) {
        const int64_t k = k_first + lane;
        if (k < my_rows) {
            const int64_t w = gw + k * n_warps;
            TmaRowDesc d;
            d.synth = w < m_loc;
            d.row = d.synth ? n_loc + w : w - m_loc;
            d.sb = 0; d.nb = 0;
            if (!d.synth) {
                const int64_t src = n0 + d.row;
                d.sa = __ldg(indptr + src);
                d.na = __ldg(indptr + src + 1) - d.sa;
                d.l1 = __ldg(l1_rows + src);
            } else {
                const int64_t r = m0 + w;
                const int64_t pa = __ldg(parents + 2 * r), pb = __ldg(parents + 2 * r + 1);
                d.sa = __ldg(indptr + pa);
                d.na = __ldg(indptr + pa + 1) - d.sa;
                d.sb = __ldg(indptr + pb);
                d.nb = __ldg(indptr + pb + 1) - d.sb;
                d.l1 = __ldg(l1_rows + pa) + __ldg(l1_rows + pb);
            }
            const int a0 = d.sa & ~3, a1 = (d.sa + d.na + 3) & ~3, b0 = d.sb & ~3, b1 = (d.sb + d.nb + 3) & ~3;
            d.oa = d.sa - a0;
            d.ob = d.sb - b0;
            const int ga = d.na > 0 ? a1 - a0 : 0, gb = d.nb > 0 ? b1 - b0 : 0;
            d.staged = (ga <= kTmaSeg && gb <= kTmaSeg && ga + gb > 0) ? 1 : 0;
            table[k % kTmaTable] = d;
        }
    };
    // lane 0: request the segments of row k (its descriptor is in the table)
    auto issue = [&](int64_t k) {
        const TmaRowDesc &d = table[k % kTmaTable];
        if (!d.staged) return;
        const int slot = (int)(k % kTmaDepth);
        const int a0 = d.sa & ~3, a1 = (d.sa + d.na + 3) & ~3, b0 = d.sb & ~3, b1 = (d.sb + d.nb + 3) & ~3;
        const int ga = d.na > 0 ? a1 - a0 : 0, gb = d.nb > 0 ? b1 - b0 : 0;
        uint8_t *sl = ring + (size_t)slot * kTmaSlotBytes;
        // the slot was read with ordinary loads by the whole warp (__syncwarp before this call): order those before the
        // async-proxy writes of the new copies
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        csr_mbar_expect_tx(bars + slot, (uint32_t)(2 * (ga + gb) * 4));
        if (ga > 0) {
            csr_bulk_g2s(sl, indices + a0, (uint32_t)ga * 4, bars + slot);
            csr_bulk_g2s(sl + kTmaSeg * 4, data + a0, (uint32_t)ga * 4, bars + slot);
        }
        if (gb > 0) {
            csr_bulk_g2s(sl + 2 * kTmaSeg * 4, indices + b0, (uint32_t)gb * 4, bars + slot);
            csr_bulk_g2s(sl + 3 * kTmaSeg * 4, data + b0, (uint32_t)gb * 4, bars + slot);
        }
    };

    describe(0);
    describe(32);
    __syncwarp();
    if (lane == 0)
        for (int64_t k = 0; k < my_rows && k < kTmaDepth; k++) issue(k);
    uint32_t phase_bits = 0;  // per-slot parity of the next completion to wait for
    for (int64_t k = 0; k < my_rows; k++) {
        const int slot = (int)(k % kTmaDepth);
        if (k > 0 && (k & 31) == 0) {  // rows [k, k + 32) are described; describe [k + 32, k + 64) into the other half
            describe(k + 32);
            __syncwarp();
        }
        const TmaRowDesc d = table[k % kTmaTable];
        float *out_row = dense + d.row * (int64_t)ld;
        if (d.staged) {
            csr_mbar_wait(bars + slot, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            const uint8_t *sl = ring + (size_t)slot * kTmaSlotBytes;
            const int32_t *ia = reinterpret_cast<const int32_t *>(sl) + d.oa;
            const float *da = reinterpret_cast<const float *>(sl + kTmaSeg * 4) + d.oa;
            const int32_t *ib = reinterpret_cast<const int32_t *>(sl + 2 * kTmaSeg * 4) + d.ob;
            const float *db = reinterpret_cast<const float *>(sl + 3 * kTmaSeg * 4) + d.ob;
            dense_row_from_lists(ia, da, d.na, ib, db, d.nb, d.synth != 0, d.l1, median, pc, logpc, n_genes, ld, out_row, lane, dbg);
        } else {
            dense_row_from_lists(indices + d.sa, data + d.sa, d.na, indices + d.sb, data + d.sb, d.nb, d.synth != 0, d.l1, median,
                                 pc, logpc, n_genes, ld, out_row, lane, dbg);
        }
        __syncwarp();  // every lane is done with the slot
        if (lane == 0 && k + kTmaDepth < my_rows) issue(k + kTmaDepth);
    }
}

// the same row algorithm on a shared-memory row buffer (no L2 round trips for the tag merge)
__device__ __forceinline__ void dense_row_in_smem(const int32_t *ia, const float *da, int na, const int32_t *ib, const float *db,
                                                  int nb, bool synth, double l1, float median, float pc, float logpc,
                                                  int n_genes, int ld, float *buf, int lane) {
    for (int j = 4 * lane; j < ld; j += 128) {
        float4 v;
        v.x = j < n_genes ? logpc : 0.f;
        v.y = j + 1 < n_genes ? logpc : 0.f;
        v.z = j + 2 < n_genes ? logpc : 0.f;
        v.w = j + 3 < n_genes ? logpc : 0.f;
        *reinterpret_cast<float4 *>(buf + j) = v;
    }
    __syncwarp();
    if (!synth) {
#pragma unroll 4
        for (int p = lane; p < na; p += 32) {
            const float v = da[p];
            if (v != 0.f) buf[ia[p]] = norm_log(v, l1, median, pc);
        }
        return;
    }
    uint32_t *bits = reinterpret_cast<uint32_t *>(buf);
#pragma unroll 4
    for (int p = lane; p < na; p += 32) bits[ia[p]] = kTagBits | (uint32_t)(p + 1);
    __syncwarp();
#pragma unroll 4
    for (int p = lane; p < nb; p += 32) {
        const int col = ib[p];
        float v = db[p];
        const uint32_t x = bits[col];
        if (is_tag(x)) v += da[(int)(x & kTagPayload) - 1];
        buf[col] = v != 0.f ? norm_log(v, l1, median, pc) : logpc;
    }
    __syncwarp();
#pragma unroll 4
    for (int p = lane; p < na; p += 32) {
        const int col = ia[p];
        const float v = da[p];
        if (is_tag(bits[col])) buf[col] = v != 0.f ? norm_log(v, l1, median, pc) : logpc;
    }
}

// DD_DENSE_V=3: as k_dense_rows_tma, but the row is assembled in a SHARED-MEMORY row buffer (constant fill, scatter and the
// tag merge of synthetic rows all stay on chip) and leaves as ONE bulk store (cp.async.bulk shared -> global): no 4-byte
// scatter stores reach L2 (they cost 0.44 ms of the 0.78 ms at c3), HBM sees the matrix exactly once, as full rows.  Two row
// buffers per warp overlap a row's store with the next row's assembly; the number of warps per CTA follows from the row size.
__global__ void __launch_bounds__(256, 1)
    k_dense_rows_tma_smem(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                     const double *__restrict__ l1_rows, const int64_t *__restrict__ parents, int64_t n0, int64_t n_loc,
                     int64_t m0, int64_t m_loc, int n_genes, int ld, float median, float pc, float *__restrict__ dense, int dbg) {
    extern __shared__ __align__(128) uint8_t tma_smem[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const size_t row_bytes = (size_t)ld * 4;  // multiple of 128
    const size_t warp_bytes = kTmaWarpBytes + 2 * row_bytes + kTmaTable * sizeof(TmaRowDesc) + kTmaDepth * 8 + 40;  // 16-byte multiple
    uint8_t *wbase = tma_smem + (size_t)wl * warp_bytes;
    uint8_t *ring = wbase;
    float *rowbuf = reinterpret_cast<float *>(wbase + kTmaWarpBytes);
    TmaRowDesc *table = reinterpret_cast<TmaRowDesc *>(wbase + kTmaWarpBytes + 2 * row_bytes);
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + kTmaWarpBytes + 2 * row_bytes + kTmaTable * sizeof(TmaRowDesc));
    if (lane == 0) {
        for (int d = 0; d < kTmaDepth; d++) csr_mbar_init(bars + d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int64_t gw = (int64_t)blockIdx.x * nw + wl, n_warps = (int64_t)gridDim.x * nw;
    const int64_t n_rows = n_loc + m_loc;
    const int64_t my_rows = gw < n_rows ? (n_rows - gw + n_warps - 1) / n_warps : 0;
    const float logpc = logf(pc);

    // all lanes: describe 32 of this warp's rows at once (the dependent parents -> indptr loads run in parallel)
    auto describe = [&](int64_t k_first) {
        const int64_t k = k_first + lane;
        if (k < my_rows) {
            const int64_t w = gw + k * n_warps;
            TmaRowDesc d;
            d.synth = w < m_loc;
            d.row = d.synth ? n_loc + w : w - m_loc;
            d.sb = 0; d.nb = 0;
            if (!d.synth) {
                const int64_t src = n0 + d.row;
                d.sa = __ldg(indptr + src);
                d.na = __ldg(indptr + src + 1) - d.sa;
                d.l1 = __ldg(l1_rows + src);
            } else {
                const int64_t r = m0 + w;
                const int64_t pa = __ldg(parents + 2 * r), pb = __ldg(parents + 2 * r + 1);
                d.sa = __ldg(indptr + pa);
                d.na = __ldg(indptr + pa + 1) - d.sa;
                d.sb = __ldg(indptr + pb);
                d.nb = __ldg(indptr + pb + 1) - d.sb;
                d.l1 = __ldg(l1_rows + pa) + __ldg(l1_rows + pb);
            }
            const int a0 = d.sa & ~3, a1 = (d.sa + d.na + 3) & ~3, b0 = d.sb & ~3, b1 = (d.sb + d.nb + 3) & ~3;
            d.oa = d.sa - a0;
            d.ob = d.sb - b0;
            const int ga = d.na > 0 ? a1 - a0 : 0, gb = d.nb > 0 ? b1 - b0 : 0;
            d.staged = (ga <= kTmaSeg && gb <= kTmaSeg && ga + gb > 0) ? 1 : 0;
            table[k % kTmaTable] = d;
        }
    };
    // lane 0: request the segments of row k (its descriptor is in the table)
    auto issue = [&](int64_t k) {
        const TmaRowDesc &d = table[k % kTmaTable];
        if (!d.staged) return;
        const int slot = (int)(k % kTmaDepth);
        const int a0 = d.sa & ~3, a1 = (d.sa + d.na + 3) & ~3, b0 = d.sb & ~3, b1 = (d.sb + d.nb + 3) & ~3;
        const int ga = d.na > 0 ? a1 - a0 : 0, gb = d.nb > 0 ? b1 - b0 : 0;
        uint8_t *sl = ring + (size_t)slot * kTmaSlotBytes;
        // the slot was read with ordinary loads by the whole warp (__syncwarp before this call): order those before the
        // async-proxy writes of the new copies
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        csr_mbar_expect_tx(bars + slot, (uint32_t)(2 * (ga + gb) * 4));
        if (ga > 0) {
            csr_bulk_g2s(sl, indices + a0, (uint32_t)ga * 4, bars + slot);
            csr_bulk_g2s(sl + kTmaSeg * 4, data + a0, (uint32_t)ga * 4, bars + slot);
        }
        if (gb > 0) {
            csr_bulk_g2s(sl + 2 * kTmaSeg * 4, indices + b0, (uint32_t)gb * 4, bars + slot);
            csr_bulk_g2s(sl + 3 * kTmaSeg * 4, data + b0, (uint32_t)gb * 4, bars + slot);
        }
    };

    describe(0);
    describe(32);
    __syncwarp();
    if (lane == 0)
        for (int64_t k = 0; k < my_rows && k < kTmaDepth; k++) issue(k);
    uint32_t phase_bits = 0;  // per-slot parity of the next completion to wait for
    for (int64_t k = 0; k < my_rows; k++) {
        const int slot = (int)(k % kTmaDepth);
        if (k > 0 && (k & 31) == 0) {  // rows [k, k + 32) are described; describe [k + 32, k + 64) into the other half
            describe(k + 32);
            __syncwarp();
        }
        const TmaRowDesc d = table[k % kTmaTable];
        float *buf = rowbuf + (size_t)(k & 1) * ld;
        // the bulk store that last used this buffer (row k - 2) must have finished READING it: at most one younger store
        // (row k - 1) may still be in flight
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        if (d.staged) {
            csr_mbar_wait(bars + slot, (phase_bits >> slot) & 1u);
            phase_bits ^= 1u << slot;
            const uint8_t *sl = ring + (size_t)slot * kTmaSlotBytes;
            const int32_t *ia = reinterpret_cast<const int32_t *>(sl) + d.oa;
            const float *da = reinterpret_cast<const float *>(sl + kTmaSeg * 4) + d.oa;
            const int32_t *ib = reinterpret_cast<const int32_t *>(sl + 2 * kTmaSeg * 4) + d.ob;
            const float *db = reinterpret_cast<const float *>(sl + 3 * kTmaSeg * 4) + d.ob;
            dense_row_in_smem(ia, da, d.na, ib, db, d.nb, d.synth != 0, d.l1, median, pc, logpc, n_genes, ld, buf, lane);
        } else {
            dense_row_in_smem(indices + d.sa, data + d.sa, d.na, indices + d.sb, data + d.sb, d.nb, d.synth != 0, d.l1, median, pc,
                              logpc, n_genes, ld, buf, lane);
        }
        __syncwarp();  // the row is complete in shared memory
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy read
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dense + d.row * (int64_t)ld),
                         "r"(csr_smem_u32(buf)), "r"((uint32_t)row_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();  // every lane is done with the slot
        if (lane == 0 && k + kTmaDepth < my_rows) issue(k + kTmaDepth);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all rows have landed before the CTA retires
}

// DD_DENSE_V=4 (experimental, written after the round's GPU budget was spent: NOT yet run on hardware): the occupancy fix
// for variant 3.  The row is still assembled in shared memory and leaves as ONE bulk store, but a warp owns a single row
// buffer and no staging ring (the ring cost 25 KB per warp, twice the row buffers), so 16-18 warps per SM are resident
// instead of 4 and cover each other's per-row latency chain (descriptor loads -> list gathers -> merge -> store drain).
// The lists are read straight from global memory (coalesced, through L1); the descriptor of the next row is loaded one row
// ahead.  Per SM: rows in flight x 12 KB leave through the bulk-copy engine, nothing is scattered into L2.
struct V4Desc {
    double l1;
    int64_t row;
    int sa, na, sb, nb;
    bool synth;
};
constexpr int kV4MaxWarps = 18;

__global__ void __launch_bounds__(kV4MaxWarps * 32, 1)
    k_dense_rows_v4(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                    const double *__restrict__ l1_rows, const int64_t *__restrict__ parents, int64_t n0, int64_t n_loc,
                    int64_t m0, int64_t m_loc, int n_genes, int ld, float median, float pc, float *__restrict__ dense) {
    extern __shared__ __align__(128) uint8_t tma_smem[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)ld * 4u;  // multiple of 128
    float *buf = reinterpret_cast<float *>(tma_smem + (size_t)wl * row_bytes);
    const int64_t gw = (int64_t)blockIdx.x * nw + wl, n_warps = (int64_t)gridDim.x * nw;
    const int64_t n_rows = n_loc + m_loc;
    const int64_t my_rows = gw < n_rows ? (n_rows - gw + n_warps - 1) / n_warps : 0;
    const float logpc = logf(pc);

    // every lane loads the same words (one broadcast transaction each)
    auto load_desc = [&](int64_t k) {
        V4Desc d;
        const int64_t w = gw + k * n_warps;
        d.synth = w < m_loc;
        d.row = d.synth ? n_loc + w : w - m_loc;
        d.sb = 0;
        d.nb = 0;
        if (!d.synth) {
            const int64_t src = n0 + d.row;
            d.sa = __ldg(indptr + src);
            d.na = __ldg(indptr + src + 1) - d.sa;
            d.l1 = __ldg(l1_rows + src);
        } else {
            const int64_t r = m0 + w;
            const int64_t pa = __ldg(parents + 2 * r), pb = __ldg(parents + 2 * r + 1);
            d.sa = __ldg(indptr + pa);
            d.na = __ldg(indptr + pa + 1) - d.sa;
            d.sb = __ldg(indptr + pb);
            d.nb = __ldg(indptr + pb + 1) - d.sb;
            d.l1 = __ldg(l1_rows + pa) + __ldg(l1_rows + pb);
        }
        return d;
    };

    V4Desc cur = my_rows > 0 ? load_desc(0) : V4Desc{};
    for (int64_t k = 0; k < my_rows; k++) {
        V4Desc nxt = cur;
        if (k + 1 < my_rows) nxt = load_desc(k + 1);  // consumed one row later: the loads overlap this row's work
        // the bulk store of the previous row must have finished READING the buffer before it is refilled
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        dense_row_in_smem(indices + cur.sa, data + cur.sa, cur.na, indices + cur.sb, data + cur.sb, cur.nb, cur.synth, cur.l1,
                          median, pc, logpc, n_genes, ld, buf, lane);
        __syncwarp();  // the row is complete in shared memory
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy read
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dense + cur.row * (int64_t)ld),
                         "r"(csr_smem_u32(buf)), "r"(row_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
        cur = nxt;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all rows have landed before the CTA retires
}

// DD_DENSE_V=5: variant 4 with the gathers taken off the row's critical path.  Variant 4 loads a row's index / value lists
// inside the scatter loops, four at a time: with 16-18 warps per SM that is ~5 KB of CSR reads in flight per SM, a
// fraction of what HBM's latency x bandwidth product asks for (~70 KB per SM) -- the kernel waits for its own gathers
// (0.85 ms).  Here a warp issues ALL loads of its next row (up to kV5PF entries per lane and parent, i.e. 384 per list) into
// registers before it does anything else with the current one, so 18 warps keep ~60 KB in flight per SM; rows with longer lists
// finish from global memory.  Same row algorithm (constant fill, scatter, tag merge of the two parents in the shared-memory
// row buffer, one bulk store per row).
template <int kV5PF>
struct V5Regs {
    int ia[kV5PF], ib[kV5PF];
    float va[kV5PF], vb[kV5PF];
};

// <12, 16>: lists up to 384 entries from registers, 16 warps (121 registers); <10, 18>: 320 entries, 18 warps
template <int kV5PF, int kV5MaxWarps>
__global__ void __launch_bounds__(kV5MaxWarps * 32, 1)
    k_dense_rows_v5(const int32_t *__restrict__ indptr, const int32_t *__restrict__ indices, const float *__restrict__ data,
                    const double *__restrict__ l1_rows, const int64_t *__restrict__ parents, int64_t n0, int64_t n_loc,
                    int64_t m0, int64_t m_loc, int n_genes, int ld, float median, float pc, float *__restrict__ dense) {
    extern __shared__ __align__(128) uint8_t tma_smem[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t row_bytes = (uint32_t)ld * 4u;  // multiple of 128
    float *buf = reinterpret_cast<float *>(tma_smem + (size_t)wl * row_bytes);
    uint32_t *bits = reinterpret_cast<uint32_t *>(buf);
    const int64_t gw = (int64_t)blockIdx.x * nw + wl, n_warps = (int64_t)gridDim.x * nw;
    const int64_t n_rows = n_loc + m_loc;
    const int64_t my_rows = gw < n_rows ? (n_rows - gw + n_warps - 1) / n_warps : 0;
    const float logpc = logf(pc);

    auto load_desc = [&](int64_t k) {
        V4Desc d;
        const int64_t w = gw + k * n_warps;
        d.synth = w < m_loc;
        d.row = d.synth ? n_loc + w : w - m_loc;
        d.sb = 0;
        d.nb = 0;
        if (!d.synth) {
            const int64_t src = n0 + d.row;
            d.sa = __ldg(indptr + src);
            d.na = __ldg(indptr + src + 1) - d.sa;
            d.l1 = __ldg(l1_rows + src);
        } else {
            const int64_t r = m0 + w;
            const int64_t pa = __ldg(parents + 2 * r), pb = __ldg(parents + 2 * r + 1);
            d.sa = __ldg(indptr + pa);
            d.na = __ldg(indptr + pa + 1) - d.sa;
            d.sb = __ldg(indptr + pb);
            d.nb = __ldg(indptr + pb + 1) - d.sb;
            d.l1 = __ldg(l1_rows + pa) + __ldg(l1_rows + pb);
        }
        return d;
    };
    auto load_lists = [&](const V4Desc &d, V5Regs<kV5PF> &r) {
#pragma unroll
        for (int j = 0; j < kV5PF; j++) {
            const int p = lane + 32 * j;
            const bool ina = p < d.na, inb = p < d.nb;
            r.ia[j] = ina ? __ldg(indices + d.sa + p) : -1;
            r.va[j] = ina ? __ldg(data + d.sa + p) : 0.f;
            r.ib[j] = inb ? __ldg(indices + d.sb + p) : -1;
            r.vb[j] = inb ? __ldg(data + d.sb + p) : 0.f;
        }
    };

    V4Desc cur = my_rows > 0 ? load_desc(0) : V4Desc{};
    V5Regs<kV5PF> r;
    if (my_rows > 0) load_lists(cur, r);
    for (int64_t k = 0; k < my_rows; k++) {
        V4Desc nxt = cur;
        if (k + 1 < my_rows) nxt = load_desc(k + 1);  // consumed at the end of this iteration
        // the bulk store of the previous row must have finished READING the buffer before it is refilled
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int j = 4 * lane; j < ld; j += 128) {
            float4 v = make_float4(logpc, logpc, logpc, logpc);
            if (j + 3 >= n_genes) {  // the last groups: pad columns are zero
                v.x = j < n_genes ? logpc : 0.f;
                v.y = j + 1 < n_genes ? logpc : 0.f;
                v.z = j + 2 < n_genes ? logpc : 0.f;
                v.w = 0.f;
            }
            *reinterpret_cast<float4 *>(buf + j) = v;
        }
        __syncwarp();
        const double l1 = cur.l1;
        const double inv = l1 != 0.0 ? 1.0 / l1 : 0.0;
        const int32_t *ia = indices + cur.sa, *ib = indices + cur.sb;
        const float *da = data + cur.sa, *db = data + cur.sb;
        if (!cur.synth) {
#pragma unroll
            for (int j = 0; j < kV5PF; j++)
                if (r.ia[j] >= 0 && r.va[j] != 0.f) buf[r.ia[j]] = norm_log_inv(r.va[j], l1, inv, median, pc);
            for (int p = lane + 32 * kV5PF; p < cur.na; p += 32) {
                const float v = da[p];
                if (v != 0.f) buf[ia[p]] = norm_log_inv(v, l1, inv, median, pc);
            }
        } else {
#pragma unroll
            for (int j = 0; j < kV5PF; j++)
                if (r.ia[j] >= 0) bits[r.ia[j]] = kTagBits | (uint32_t)(lane + 32 * j + 1);
            for (int p = lane + 32 * kV5PF; p < cur.na; p += 32) bits[ia[p]] = kTagBits | (uint32_t)(p + 1);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kV5PF; j++)
                if (r.ib[j] >= 0) {
                    float v = r.vb[j];
                    const uint32_t x = bits[r.ib[j]];
                    if (is_tag(x)) v += da[(int)(x & kTagPayload) - 1];
                    buf[r.ib[j]] = v != 0.f ? norm_log_inv(v, l1, inv, median, pc) : logpc;
                }
            for (int p = lane + 32 * kV5PF; p < cur.nb; p += 32) {
                const int col = ib[p];
                float v = db[p];
                const uint32_t x = bits[col];
                if (is_tag(x)) v += da[(int)(x & kTagPayload) - 1];
                buf[col] = v != 0.f ? norm_log_inv(v, l1, inv, median, pc) : logpc;
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < kV5PF; j++)
                if (r.ia[j] >= 0 && is_tag(bits[r.ia[j]])) buf[r.ia[j]] = r.va[j] != 0.f ? norm_log_inv(r.va[j], l1, inv, median, pc) : logpc;
            for (int p = lane + 32 * kV5PF; p < cur.na; p += 32) {
                const int col = ia[p];
                const float v = da[p];
                if (is_tag(bits[col])) buf[col] = v != 0.f ? norm_log_inv(v, l1, inv, median, pc) : logpc;
            }
        }
        __syncwarp();  // the row is complete in shared memory
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async-proxy read
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dense + cur.row * (int64_t)ld),
                         "r"(csr_smem_u32(buf)), "r"(row_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
        cur = nxt;
        if (k + 1 < my_rows) load_lists(cur, r);  // every gather of the next row goes out now
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all rows have landed before the CTA retires
}

int pick_chunk(int64_t ld) { return (int)std::min<int64_t>(ld, kMaxChunk); }

// The dense build is latency-bound with one 12 KB row buffer per warp (16 warps / SM); staging 1024 columns
// at a time (4 KB per warp) lets 48 warps / SM overlap their gather / scatter / store phases.
int pick_dense_chunk(int64_t ld) {
    static const int env = getenv("DD_DENSE_CHUNK") ? atoi(getenv("DD_DENSE_CHUNK")) : 1024;
    return (int)std::min<int64_t>(ld, std::max(32, env / 32 * 32));
}

int grid_for(dd_handle *h, size_t smem_bytes) {
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (200 * 1024) / std::max<size_t>(smem_bytes + 1024, 1)));
    return h->num_sms * per_sm;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int dd_finish_upload(dd_handle *h);

// a handle that borrowed another handle's count matrix (dd_share_counts) forgets the pointers instead of freeing them
void dd_drop_borrowed_counts(dd_handle *h) {
    if (!h->counts_borrowed) return;
    h->d_indptr = nullptr; h->d_indices = nullptr; h->d_data = nullptr; h->d_lib = nullptr; h->d_l1 = nullptr;
    h->cap_rows = 0; h->cap_nnz = 0;
    h->counts_borrowed = false;
}

// Second (third, ...) pipeline on the same GPU: `dst` works on `src`'s resident count matrix and library sizes (read-only
// for the fit loop) instead of uploading its own copy.  `src` must stay alive and keep its matrix while `dst` uses it.
extern "C" int dd_share_counts(dd_handle *dst, const dd_handle *src) {
    if (!dst || !src || dst == src) return dd_fail(dst, DD_ERR_ARG, "dd_share_counts: two different handles are needed");
    if (!src->d_indptr) return dd_fail(dst, DD_ERR_ARG, "dd_share_counts: the source handle holds no counts");
    if (dst->device != src->device) return dd_fail(dst, DD_ERR_ARG, "dd_share_counts: the handles live on different devices");
    DD_CUDA(dst, cudaSetDevice(dst->device));
    if (!dst->counts_borrowed) {
        for (void *p : {(void *)dst->d_indptr, (void *)dst->d_indices, (void *)dst->d_data, (void *)dst->d_lib, (void *)dst->d_l1})
            if (p) cudaFree(p);
    }
    dst->d_indptr = src->d_indptr; dst->d_indices = src->d_indices; dst->d_data = src->d_data;
    dst->d_lib = src->d_lib; dst->d_l1 = src->d_l1;
    dst->cap_rows = 0; dst->cap_nnz = 0;
    dst->counts_borrowed = true;
    dst->N = src->N; dst->G = src->G; dst->nnz = src->nnz; dst->ld = src->ld;
    dst->h_lib = src->h_lib;
    dst->nonneg = src->nonneg;
    dst->all_finite = src->all_finite;
    dst->synth_csr_valid = false; dst->dense_valid = false; dst->emb_valid = false; dst->M = 0; dst->A = 0;
    return DD_OK;
}

extern "C" int dd_upload_counts(dd_handle *h, int64_t n_cells, int64_t n_genes, const int32_t *indptr,
                                const int32_t *indices, const float *data) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_upload_counts: null handle");
    if (n_cells <= 0 || n_genes <= 0 || !indptr) return dd_fail(h, DD_ERR_ARG, "dd_upload_counts: empty matrix");
    if (n_genes >= (1ll << 31) - 64 || n_cells >= (1ll << 31) - 64)
        return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_upload_counts: dimensions exceed int32 indexing");
    const int64_t nnz = indptr[n_cells];
    if (indptr[0] != 0 || nnz < 0) return dd_fail(h, DD_ERR_ARG, "dd_upload_counts: bad indptr");
    if (nnz > 0 && (!indices || !data)) return dd_fail(h, DD_ERR_ARG, "dd_upload_counts: null indices/data");
    DD_CUDA(h, cudaSetDevice(h->device));
    dd_drop_borrowed_counts(h);
    h->N = n_cells; h->G = n_genes; h->nnz = nnz;
    h->ld = dd_round_up(n_genes, 32);
    h->synth_csr_valid = false; h->dense_valid = false; h->emb_valid = false; h->M = 0; h->A = 0;
    // grow-only buffers: a second fit() on the same classifier re-uses the allocations
    if (n_cells + 1 > h->cap_rows) {
        for (void *p : {(void *)h->d_indptr, (void *)h->d_lib, (void *)h->d_l1})
            if (p) cudaFree(p);
        h->d_indptr = nullptr; h->d_lib = nullptr; h->d_l1 = nullptr; h->cap_rows = 0;
        DD_CUDA(h, cudaMalloc(&h->d_indptr, sizeof(int32_t) * (n_cells + 1)));
        DD_CUDA(h, cudaMalloc(&h->d_lib, sizeof(float) * n_cells));
        DD_CUDA(h, cudaMalloc(&h->d_l1, sizeof(double) * (n_cells + 1)));
        h->cap_rows = n_cells + 1;
    }
    if (std::max<int64_t>(nnz, 1) > h->cap_nnz) {
        if (h->d_indices) cudaFree(h->d_indices);
        if (h->d_data) cudaFree(h->d_data);
        h->d_indices = nullptr; h->d_data = nullptr; h->cap_nnz = 0;
        // + 4 entries: the bulk-copy staging of a row reads whole 16-byte groups around it
        DD_CUDA(h, cudaMalloc(&h->d_indices, sizeof(int32_t) * (std::max<int64_t>(nnz, 1) + 4)));
        DD_CUDA(h, cudaMalloc(&h->d_data, sizeof(float) * (std::max<int64_t>(nnz, 1) + 4)));
        h->cap_nnz = std::max<int64_t>(nnz, 1);
    }
    DD_CUDA(h, cudaMemcpyAsync(h->d_indptr, indptr, sizeof(int32_t) * (n_cells + 1), cudaMemcpyHostToDevice, h->stream));
    if (nnz > 0) {
        DD_CUDA(h, cudaMemcpyAsync(h->d_indices, indices, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, h->stream));
        DD_CUDA(h, cudaMemcpyAsync(h->d_data, data, sizeof(float) * nnz, cudaMemcpyHostToDevice, h->stream));
    }
    return dd_finish_upload(h);
}

// _lib_size (:182) + L1 sums + sign check of the CSR the handle holds (uploaded, or subset on the device by dd_select_genes)
int dd_finish_upload(dd_handle *h) {
    const int64_t n_cells = h->N;
    const int grid = h->num_sms * 8;
    int *d_neg = reinterpret_cast<int *>(h->d_l1 + n_cells);  // one spare slot behind the L1 sums
    DD_CUDA(h, cudaMemsetAsync(d_neg, 0, 2 * sizeof(int), h->stream));
    DD_LAUNCH(h, "row_sums", k_row_sums, grid, 256, 0, h->d_indptr, h->d_data, n_cells, h->d_lib, h->d_l1, d_neg);
    int neg[2] = {0, 0};
    DD_CUDA(h, cudaMemcpyAsync(neg, d_neg, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    h->h_lib.resize(n_cells);
    DD_CUDA(h, cudaMemcpyAsync(h->h_lib.data(), h->d_lib, sizeof(float) * n_cells, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->nonneg = neg[0] == 0;
    h->all_finite = neg[1] == 0;
    return DD_OK;
}

// 1 if the uploaded matrix holds only finite values (the device looked at every entry while summing the rows): lets the
// caller skip the host-side finiteness scan of check_array (:149-155) and run it only to produce sklearn's error message.
extern "C" int dd_counts_all_finite(dd_handle *h, int32_t *out) {
    if (!h || !out || !h->d_indptr) return dd_fail(h, DD_ERR_ARG, "dd_counts_all_finite: upload the counts first");
    *out = h->all_finite ? 1 : 0;
    return DD_OK;
}

extern "C" int dd_get_lib_size(dd_handle *h, float *out) {
    if (!h || !out || !h->d_indptr) return dd_fail(h, DD_ERR_ARG, "dd_get_lib_size: no counts uploaded");
    std::copy(h->h_lib.begin(), h->h_lib.end(), out);
    return DD_OK;
}

// Upload `choices` (:394) for one iteration; validates the indices on the host.
int dd_set_parents(dd_handle *h, int64_t n_synth, const int64_t *parents) {
    if (!h || !h->d_indptr) return dd_fail(h, DD_ERR_ARG, "parents: no counts uploaded");
    if (n_synth < 0 || (n_synth > 0 && !parents)) return dd_fail(h, DD_ERR_ARG, "parents: bad arguments");
    for (int64_t i = 0; i < 2 * n_synth; i++)
        if (parents[i] < 0 || parents[i] >= h->N) return dd_fail(h, DD_ERR_ARG, "parents: index out of range");
    DD_CUDA(h, cudaSetDevice(h->device));
    if (n_synth > h->cap_M) {
        for (void *p : {(void *)h->d_parents, (void *)h->d_sindptr, (void *)h->d_scount, (void *)h->d_slib})
            if (p) cudaFree(p);
        h->d_parents = nullptr; h->d_sindptr = nullptr; h->d_scount = nullptr; h->d_slib = nullptr;
        h->cap_M = 0;
        DD_CUDA(h, cudaMalloc(&h->d_parents, sizeof(int64_t) * 2 * n_synth));
        DD_CUDA(h, cudaMalloc(&h->d_sindptr, sizeof(int32_t) * (n_synth + 1)));
        DD_CUDA(h, cudaMalloc(&h->d_scount, sizeof(int32_t) * n_synth));
        DD_CUDA(h, cudaMalloc(&h->d_slib, sizeof(float) * n_synth));
        h->cap_M = n_synth;
    }
    h->M = n_synth;
    dd_set_block(h);  // A = rows of this rank's block (everything unless the handle shards cells)
    h->synth_csr_valid = false; h->dense_valid = false; h->emb_valid = false;
    if (n_synth > 0)
        DD_CUDA(h, cudaMemcpyAsync(h->d_parents, parents, sizeof(int64_t) * 2 * n_synth, cudaMemcpyHostToDevice,
                                   h->stream));
    return DD_OK;
}

int dd_dev_create_doublets_csr(dd_handle *h) {
    const int chunk = pick_chunk(h->ld);
    const size_t smem = sizeof(float) * chunk * kWarpsPerCta;
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(k_synth_count, cudaFuncAttributeMaxDynamicSharedMemorySize, sizeof(float) * kMaxChunk * kWarpsPerCta);
        cudaFuncSetAttribute(k_synth_fill, cudaFuncAttributeMaxDynamicSharedMemorySize, sizeof(float) * kMaxChunk * kWarpsPerCta);
    });
    const int grid = grid_for(h, smem);
    if (h->M > 0) {
        DD_LAUNCH(h, "synth_count", k_synth_count, grid, kThreads, smem, h->d_indptr, h->d_indices, h->d_data,
                  h->d_parents, h->M, (int)h->G, chunk, h->d_scount, h->d_slib);
    }
    DD_LAUNCH(h, "exclusive_scan", k_exclusive_scan, 1, 1024, 0, h->d_scount, h->M, h->d_sindptr);
    int32_t total = 0;
    DD_CUDA(h, cudaMemcpyAsync(&total, h->d_sindptr + h->M, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (total < 0) return dd_fail(h, DD_ERR_UNSUPPORTED, "synthetic nnz exceeds int32 indexing");
    if (total > h->cap_snnz) {
        if (h->d_sindices) cudaFree(h->d_sindices);
        if (h->d_sdata) cudaFree(h->d_sdata);
        h->d_sindices = nullptr; h->d_sdata = nullptr; h->cap_snnz = 0;
        DD_CUDA(h, cudaMalloc(&h->d_sindices, sizeof(int32_t) * std::max<int64_t>(total, 1)));
        DD_CUDA(h, cudaMalloc(&h->d_sdata, sizeof(float) * std::max<int64_t>(total, 1)));
        h->cap_snnz = total;
    }
    h->snnz = total;
    if (h->M > 0 && total > 0) {
        DD_LAUNCH(h, "synth_fill", k_synth_fill, grid, kThreads, smem, h->d_indptr, h->d_indices, h->d_data,
                  h->d_parents, h->M, (int)h->G, chunk, h->d_sindptr, h->d_sindices, h->d_sdata);
    }
    h->synth_csr_valid = true;
    return DD_OK;
}

extern "C" int dd_create_doublets(dd_handle *h, int64_t n_synth, const int64_t *parents) {
    DD_TRY(dd_set_parents(h, n_synth, parents));
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_create_doublets_csr(h));
    DD_TRY(dd_stage_end(h, "doublets"));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

extern "C" int dd_synth_nnz(dd_handle *h, int64_t *nnz_out) {
    if (!h || !nnz_out || !h->synth_csr_valid) return dd_fail(h, DD_ERR_ARG, "dd_synth_nnz: call dd_create_doublets first");
    *nnz_out = h->snnz;
    return DD_OK;
}

extern "C" int dd_download_synthetics(dd_handle *h, int32_t *indptr_out, int32_t *indices_out, float *data_out) {
    if (!h || !h->synth_csr_valid) return dd_fail(h, DD_ERR_ARG, "dd_download_synthetics: call dd_create_doublets first");
    if (!indptr_out) return dd_fail(h, DD_ERR_ARG, "dd_download_synthetics: null output");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_CUDA(h, cudaMemcpyAsync(indptr_out, h->d_sindptr, sizeof(int32_t) * (h->M + 1), cudaMemcpyDeviceToHost, h->stream));
    if (h->snnz > 0) {
        if (!indices_out || !data_out) return dd_fail(h, DD_ERR_ARG, "dd_download_synthetics: null output");
        DD_CUDA(h, cudaMemcpyAsync(indices_out, h->d_sindices, sizeof(int32_t) * h->snnz, cudaMemcpyDeviceToHost, h->stream));
        DD_CUDA(h, cudaMemcpyAsync(data_out, h->d_sdata, sizeof(float) * h->snnz, cudaMemcpyDeviceToHost, h->stream));
    }
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

extern "C" int dd_get_synth_lib_size(dd_handle *h, float *out) {
    if (!h || !out || !h->synth_csr_valid) return dd_fail(h, DD_ERR_ARG, "dd_get_synth_lib_size: call dd_create_doublets first");
    DD_CUDA(h, cudaSetDevice(h->device));
    if (h->M > 0) DD_CUDA(h, cudaMemcpyAsync(out, h->d_slib, sizeof(float) * h->M, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

// np.median(aug_lib_size) (:289, :293): float32, mean of the two middle values for even length.
float dd_host_median(std::vector<float> &v) {
    const size_t n = v.size();
    if (n == 0) return nanf("");
    const size_t mid = n / 2;
    std::nth_element(v.begin(), v.begin() + mid, v.end());
    const float hi = v[mid];
    if (n & 1) return hi;
    const float lo = *std::max_element(v.begin(), v.begin() + mid);
    return (lo + hi) / 2.0f;
}

extern "C" int dd_median_lib_size(dd_handle *h, float *median_out) {
    if (!h || !median_out || !h->synth_csr_valid) return dd_fail(h, DD_ERR_ARG, "dd_median_lib_size: call dd_create_doublets first");
    std::vector<float> aug(h->N + h->M);
    std::copy(h->h_lib.begin(), h->h_lib.end(), aug.begin());
    DD_TRY(dd_get_synth_lib_size(h, aug.data() + h->N));
    *median_out = dd_host_median(aug);
    return DD_OK;
}

int dd_dev_build_dense(dd_handle *h, float median, float pseudocount) {
    // M == 0 (int(boost_rate * n_cells) == 0, tolerated by the reference) never allocates the parents
    if (!h->d_indptr || (h->M > 0 && !h->d_parents) || h->A == 0)
        return dd_fail(h, DD_ERR_ARG, "normalise: upload counts and parents first");
    const int64_t need = h->A * h->ld;
    DD_TRY(dd_reserve(h, &h->d_dense, &h->cap_dense, need));
    // DD_DENSE_V=0 keeps the shared-memory row-buffer kernel (A/B comparison; also used when the merge-through-
    // the-row trick does not apply: negative values, or rows too long for the tag payload)
    // default 5: rows assembled in shared memory from register-prefetched lists (0.49 ms at c3 = 59 % of the HBM peak); rows too
    // long for >= 8 row buffers per SM (more than ~7000 genes) fall through to variant 1 (fill + scatter through L2, 0.76 ms)
    static const int variant = getenv("DD_DENSE_V") ? atoi(getenv("DD_DENSE_V")) : 5;
    const bool sharded = dd_sharded(h);
    if ((variant != 0 && h->nonneg && h->G < (int64_t)kTagPayload) || sharded) {
        if (!h->nonneg || h->G >= (int64_t)kTagPayload)
            return dd_fail(h, DD_ERR_UNSUPPORTED, "normalise: cell-block sharding needs non-negative counts");
        // 48 registers x 256 threads: five CTAs fit an SM; four (32 warps) keep the grid a single full wave and the rows in
        // flight (32 warps x 148 SMs x 12 KB = 57 MB at 3k genes) inside the 126 MB L2
        static const int warps_per_sm = getenv("DD_DENSE_WARPS") ? atoi(getenv("DD_DENSE_WARPS")) : 32;
        if (variant == 3) {  // TMA-staged CSR segments, rows assembled in shared memory, bulk-stored
            const size_t per_warp = kTmaWarpBytes + 2 * (size_t)h->ld * 4 + kTmaTable * sizeof(TmaRowDesc) + kTmaDepth * 8 + 40;
            const int nw = (int)std::min<size_t>(8, (220 * 1024) / per_warp);
            if (nw >= 1) {
                static dd_once_per_device attr3;  // function attributes are per device
                attr3.run(h->device, [&] {
                    cudaFuncSetAttribute(k_dense_rows_tma_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                });
                DD_LAUNCH(h, "dense_rows", k_dense_rows_tma_smem, h->num_sms, nw * 32, per_warp * nw, h->d_indptr, h->d_indices,
                          h->d_data, h->d_l1, h->d_parents, h->blk_n0, h->blk_n, h->blk_m0, h->blk_m, (int)h->G, (int)h->ld, median,
                          pseudocount, h->d_dense, 0);
                h->dense_valid = true;
                h->emb_valid = false;
                return DD_OK;
            }
        }
        if (variant == 4) {  // rows assembled in shared memory (one buffer per warp, no staging ring), bulk-stored
            static const int v4_warps = getenv("DD_DENSE_WARPS") ? atoi(getenv("DD_DENSE_WARPS")) : 16;
            const size_t row_bytes = (size_t)h->ld * 4;
            const int nw = (int)std::min<size_t>(std::min(std::max(v4_warps, 1), kV4MaxWarps), (224 * 1024) / row_bytes);
            if (nw >= 1) {
                static dd_once_per_device attr4;  // function attributes are per device
                attr4.run(h->device, [&] {
                    cudaFuncSetAttribute(k_dense_rows_v4, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                });
                DD_LAUNCH(h, "dense_rows", k_dense_rows_v4, h->num_sms, nw * 32, row_bytes * nw, h->d_indptr, h->d_indices,
                          h->d_data, h->d_l1, h->d_parents, h->blk_n0, h->blk_n, h->blk_m0, h->blk_m, (int)h->G, (int)h->ld, median,
                          pseudocount, h->d_dense);
                h->dense_valid = true;
                h->emb_valid = false;
                return DD_OK;
            }
        }
        if (variant == 5) {  // variant 4 + the next row's gathers issued into registers a row ahead (the default, see below)
            static const int v5_warps = getenv("DD_DENSE_WARPS") ? atoi(getenv("DD_DENSE_WARPS")) : 16;
            const size_t row_bytes = (size_t)h->ld * 4;
            const int nw = (int)std::min<size_t>(std::min(std::max(v5_warps, 1), 18), (224 * 1024) / row_bytes);
            if (nw >= 8) {
                static dd_once_per_device attr5;  // function attributes are per device
                attr5.run(h->device, [&] {
                    cudaFuncSetAttribute(k_dense_rows_v5<12, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                    cudaFuncSetAttribute(k_dense_rows_v5<10, 18>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                });
                if (nw > 16)
                    DD_LAUNCH(h, "dense_rows", (k_dense_rows_v5<10, 18>), h->num_sms, nw * 32, row_bytes * nw, h->d_indptr, h->d_indices,
                              h->d_data, h->d_l1, h->d_parents, h->blk_n0, h->blk_n, h->blk_m0, h->blk_m, (int)h->G, (int)h->ld,
                              median, pseudocount, h->d_dense);
                else
                    DD_LAUNCH(h, "dense_rows", (k_dense_rows_v5<12, 16>), h->num_sms, nw * 32, row_bytes * nw, h->d_indptr, h->d_indices,
                              h->d_data, h->d_l1, h->d_parents, h->blk_n0, h->blk_n, h->blk_m0, h->blk_m, (int)h->G, (int)h->ld,
                              median, pseudocount, h->d_dense);
                h->dense_valid = true;
                h->emb_valid = false;
                return DD_OK;
            }
        }
        if (variant == 2) {  // TMA-staged CSR segments, one 8-warp CTA per SM
            static const int dbg = getenv("DD_DENSE_DBG") ? atoi(getenv("DD_DENSE_DBG")) : 0;
            static dd_once_per_device tma_attr;  // function attributes are per device
            tma_attr.run(h->device, [&] {
                cudaFuncSetAttribute(k_dense_rows_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes);
            });
            DD_LAUNCH(h, "dense_rows", k_dense_rows_tma, h->num_sms, kTmaWarps * 32, kTmaSmemBytes, h->d_indptr, h->d_indices,
                      h->d_data, h->d_l1, h->d_parents, h->blk_n0, h->blk_n, h->blk_m0, h->blk_m, (int)h->G, (int)h->ld, median,
                      pseudocount, h->d_dense, dbg);
            h->dense_valid = true;
            h->emb_valid = false;
            return DD_OK;
        }
        const int grid = h->num_sms * std::max(1, warps_per_sm / 8);
        DD_LAUNCH(h, "dense_rows", k_dense_rows_l2, grid, 256, 0, h->d_indptr, h->d_indices, h->d_data, h->d_l1,
                  h->d_parents, h->blk_n0, h->blk_n, h->blk_m0, h->blk_m, (int)h->G, (int)h->ld, median, pseudocount,
                  h->d_dense);
    } else {
        const int chunk = pick_dense_chunk(h->ld);
        const size_t smem = sizeof(float) * chunk * kWarpsPerCta;
        static dd_once_per_device attr_set;  // function attributes are per device
        attr_set.run(h->device, [&] {
            cudaFuncSetAttribute(k_dense_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, sizeof(float) * kMaxChunk * kWarpsPerCta);
        });
        const int grid = grid_for(h, smem);
        DD_LAUNCH(h, "dense_rows", k_dense_rows, grid, kThreads, smem, h->d_indptr, h->d_indices, h->d_data, h->d_l1,
                  h->d_parents, h->N, h->M, (int)h->G, (int)h->ld, chunk, median, pseudocount, h->nonneg ? 1 : 0, h->d_dense);
    }
    h->dense_valid = true;
    h->emb_valid = false;
    return DD_OK;
}

extern "C" int dd_normalise_log(dd_handle *h, float median, float pseudocount) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_normalise_log: null handle");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_build_dense(h, median, pseudocount));
    DD_TRY(dd_stage_end(h, "normalise"));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

extern "C" int dd_download_dense(dd_handle *h, int64_t row0, int64_t n_rows, float *out) {
    if (!h || !h->dense_valid) return dd_fail(h, DD_ERR_ARG, "dd_download_dense: no dense matrix");
    if (row0 < 0 || n_rows < 0 || row0 + n_rows > h->A || !out) return dd_fail(h, DD_ERR_ARG, "dd_download_dense: bad range");
    DD_CUDA(h, cudaSetDevice(h->device));
    if (n_rows > 0)
        DD_CUDA(h, cudaMemcpy2DAsync(out, sizeof(float) * h->G, h->d_dense + row0 * h->ld, sizeof(float) * h->ld,
                                     sizeof(float) * h->G, n_rows, cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

extern "C" int dd_upload_dense(dd_handle *h, int64_t n_rows, int64_t n_genes, const float *dense) {
    if (!h || !dense || n_rows <= 0 || n_genes <= 0) return dd_fail(h, DD_ERR_ARG, "dd_upload_dense: bad arguments");
    if (dd_sharded(h)) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_upload_dense: not available on a cell-sharded handle");
    DD_CUDA(h, cudaSetDevice(h->device));
    if (!h->d_indptr) {  // stand-alone use (tests): no counts uploaded
        h->G = n_genes;
        h->ld = dd_round_up(n_genes, 32);
        h->N = n_rows;
        h->M = 0;
    } else if (n_genes != h->G) {
        return dd_fail(h, DD_ERR_ARG, "dd_upload_dense: gene count differs from the uploaded counts");
    }
    h->A = h->A_glob = n_rows;
    DD_TRY(dd_reserve(h, &h->d_dense, &h->cap_dense, h->A * h->ld));
    DD_CUDA(h, cudaMemsetAsync(h->d_dense, 0, sizeof(float) * h->A * h->ld, h->stream));
    DD_CUDA(h, cudaMemcpy2DAsync(h->d_dense, sizeof(float) * h->ld, dense, sizeof(float) * h->G, sizeof(float) * h->G,
                                 n_rows, cudaMemcpyHostToDevice, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->dense_valid = true;
    h->emb_valid = false;
    return DD_OK;
}
