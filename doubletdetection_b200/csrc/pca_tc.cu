// pca_tc.cu -- the two tall-skinny products of the randomized PCA on tcgen05 / TMEM.
//
//   GEMM1   Y (A x 40)  = D (A x G) Q            one CTA = 128 rows of D, all of K = G
//   GEMM2   Z (G x 40) += D^T (G x A) Y'         one CTA = 128 genes x a range of rows (split-K, fp64 atomics)
//
// Both stream the dense matrix D from HBM exactly once per pass (HBM-bound: 20 flop/byte) and need
// float32-class accuracy, so they run as "3xTF32": acc += D_hi B_hi + D_hi B_lo + D_lo B_hi with TF32-exact
// high parts and float32 remainders.  D is the A operand and is split ON THE FLY:
//   * a TMA tensor-map load drops a 16 KB tile of D into shared memory (GEMM1: 128 rows x 32 genes with
//     the 128-byte swizzle; GEMM2: 32 rows x 128 genes, so that TMEM lane = gene gives the transpose for free);
//   * four transform warps (thread = TMEM lane = row of the MMA) read their 32 values, split them into
//     hi / lo and tcgen05.st them into a double-buffered A operand IN TENSOR MEMORY -- the MMA then reads A
//     from TMEM, so the big operand crosses shared memory once instead of three times;
//   * the small operand (Q resp. Y', 48 x 32 per K chunk, hi and lo) is produced by the previous kernel
//     already in the canonical K-major UMMA layout and arrives with one 12 KB bulk copy per chunk;
//   * one thread issues 4 K-steps x 3 tcgen05.mma (kind::tf32, M=128, N=48, A from TMEM) per chunk and
//     commits to the mbarriers that recycle the stage and the TMEM operand buffer.
// 192 threads: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = transform + epilogue.
// TMEM: 256 columns per CTA (2 x 48 accumulator + 2 x 64 operand), two CTAs per SM.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dd_internal.h"
#include "pca_tc.h"

namespace tcg {

constexpr int BK = 32, NB = 48, NS = 3;
constexpr int D_TILE_BYTES = 128 * BK * 4;   // 16384
constexpr int B_PART_BYTES = NB * BK * 4;    // 6144
constexpr int B_TILE_BYTES = 2 * B_PART_BYTES;
constexpr int STAGE_BYTES = D_TILE_BYTES + B_TILE_BYTES;  // 28672 = 28 * 1024
constexpr size_t SMEM_BYTES = (size_t)NS * STAGE_BYTES + 256 + 1024;  // + barriers + alignment slack
constexpr int TM_BIG = 0, TM_SMALL = 48, TM_A = 96, TMEM_COLS = 256;
// The tensor core adds every K=8 partial sum into the fp32 accumulator with truncation, so a long chain of
// accumulations drifts (measured: 1e-4 relative after ~1100 of them).  The dominant hi*hi products therefore
// accumulate for at most FLUSH chunks (4 MMAs each) before the transform warps drain the partial sum into
// registers (round-to-nearest adds); the 2^-11 smaller cross terms have their own accumulator.
constexpr int FLUSH = 8;
constexpr int B_LBO = 128, B_SBO = 1024;
// kind::tf32, fp32 accumulate, K-major operands, M = 128, N = 48
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((48u >> 3) << 17) | ((128u >> 4) << 24);

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
__device__ __forceinline__ void tma_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
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
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no swizzle shared-memory descriptor of one 48 x 32 operand part
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(B_LBO >> 4) << 16;
    d |= (uint64_t)(B_SBO >> 4) << 32;
    d |= 1ull << 46;
    return d;
}

// TR == false: GEMM1 (tile = 128 rows, K runs over genes); TR == true: GEMM2 (tile = 128 genes, K over rows)
template <bool TR, int LP>
__global__ void __launch_bounds__(192, 2) k_tc_gemm(const __grid_constant__ CUtensorMap tmap,
                                                    const uint8_t *__restrict__ bt, const float *__restrict__ mu,
                                                    float *__restrict__ Y, uint8_t *__restrict__ ytiles,
                                                    double *__restrict__ Zacc, int64_t n_rows, int ld,
                                                    int chunks_per_split, int tile_rows) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)NS * STAGE_BYTES);
    uint64_t *full = bars;            // NS
    uint64_t *empty = full + NS;      // NS
    uint64_t *a_ready = empty + NS;   // 2
    uint64_t *a_free = a_ready + 2;   // 2
    uint64_t *part_full = a_free + 2;     // 1: a partial (or the final) accumulator is complete
    uint64_t *part_free = part_full + 1;  // 1: the partial has been drained, the accumulator may restart
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(part_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int chunk0, nk;
    if (!TR) {
        chunk0 = 0;
        nk = ld / BK;
    } else {
        const int total = (int)((n_rows + BK - 1) / BK);
        chunk0 = blockIdx.y * chunks_per_split;
        nk = min(chunks_per_split, total - chunk0);
    }
    // first row (GEMM1: tile_rows <= 128 rows per CTA, sized so that the grid fills whole waves) or first gene
    // (GEMM2: 128 genes) of this CTA
    const int tile0 = blockIdx.x * (TR ? 128 : tile_rows);
    const uint32_t stage_tx = (TR ? D_TILE_BYTES : tile_rows * BK * 4) + B_TILE_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(a_ready + b, 128);
            mbar_init(a_free + b, 1);
        }
        mbar_init(part_full, 1);
        mbar_init(part_free, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (nk > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int c = 0; c < nk; c++) {
                    const int s = c % NS;
                    const uint32_t ph = (c / NS) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    mbar_expect_tx(full + s, stage_tx);
                    uint8_t *sD = smem + (size_t)s * STAGE_BYTES;
                    if (!TR)
                        tma_2d(sD, &tmap, (chunk0 + c) * BK, tile0, full + s);
                    else
                        tma_2d(sD, &tmap, tile0, (chunk0 + c) * BK, full + s);
                    bulk_g2s(sD + D_TILE_BYTES, bt + (size_t)(chunk0 + c) * B_TILE_BYTES, B_TILE_BYTES, full + s);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t big = tmem_base + TM_BIG, small = tmem_base + TM_SMALL;
                for (int c = 0; c < nk; c++) {
                    const int s = c % NS;
                    const uint32_t ph = (c / NS) & 1;
                    const int ab = c & 1;
                    const uint32_t aph = (c >> 1) & 1;
                    const bool first = (c % FLUSH) == 0;
                    mbar_wait(full + s, ph);
                    mbar_wait(a_ready + ab, aph);
                    if (first && c > 0) mbar_wait(part_free, ((c / FLUSH) - 1) & 1);
                    fence_after();
                    const uint32_t sB = smem_u32(smem + (size_t)s * STAGE_BYTES + D_TILE_BYTES);
                    const uint64_t b_hi = make_b_desc(sB), b_lo = make_b_desc(sB + B_PART_BYTES);
                    const uint32_t a_hi = tmem_base + TM_A + ab * 64, a_lo = a_hi + 32;
#pragma unroll
                    for (int ks = 0; ks < BK / 8; ks++) {
                        const uint64_t koff = (uint64_t)(ks * 2 * B_LBO / 16);
                        mma_ts_tf32(big, a_hi + ks * 8, b_hi + koff, kIdesc, (first && ks == 0) ? 0u : 1u);
                        mma_ts_tf32(small, a_hi + ks * 8, b_lo + koff, kIdesc, (c == 0 && ks == 0) ? 0u : 1u);
                        mma_ts_tf32(small, a_lo + ks * 8, b_hi + koff, kIdesc, 1u);
                    }
                    mma_commit(empty + s);
                    mma_commit(a_free + ab);
                    if ((c % FLUSH) == FLUSH - 1 || c == nk - 1) mma_commit(part_full);
                }
            }
        } else {
            const int quad = warp & 3;
            const int r = quad * 32 + lane;  // TMEM lane == row of the MMA tile
            const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
            float sum[48];
#pragma unroll
            for (int i = 0; i < 48; i++) sum[i] = 0.f;
            auto drain = [&](uint32_t col) {
                uint32_t t0[16], t1[16], t2[16];
                tmem_ld16(lane_addr + col, t0);
                tmem_ld16(lane_addr + col + 16, t1);
                tmem_ld16(lane_addr + col + 32, t2);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    sum[i] += __uint_as_float(t0[i]);
                    sum[16 + i] += __uint_as_float(t1[i]);
                    sum[32 + i] += __uint_as_float(t2[i]);
                }
            };
            // column centring on the fly (sklearn: X -= mean_, in float32): GEMM2's thread owns one gene for the
            // whole kernel, GEMM1's thread needs the 32 means of the current gene chunk
            const float mu_g = (TR && tile0 + r < ld) ? __ldg(mu + tile0 + r) : 0.f;
            for (int c = 0; c < nk; c++) {
                const int s = c % NS;
                const uint32_t ph = (c / NS) & 1;
                const int ab = c & 1;
                const uint32_t aph = (c >> 1) & 1;
                float4 m4[8];
                if (!TR) {
#pragma unroll
                    for (int j = 0; j < 8; j++) m4[j] = __ldg(reinterpret_cast<const float4 *>(mu + (size_t)(chunk0 + c) * BK) + j);
                }
                mbar_wait(full + s, ph);
                const uint8_t *sD = smem + (size_t)s * STAGE_BYTES;
                uint32_t hi[32], lo[32];
                if (!TR) {
                    // 128-byte swizzle: 16-byte chunk j of row r sits at chunk position j ^ (r & 7)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const float4 x = *reinterpret_cast<const float4 *>(sD + r * 128 + ((j ^ (r & 7)) << 4));
                        const float xs[4] = {x.x - m4[j].x, x.y - m4[j].y, x.z - m4[j].z, x.w - m4[j].w};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const uint32_t h = __float_as_uint(xs[e]) & 0xffffe000u;
                            hi[4 * j + e] = h;
                            lo[4 * j + e] = __float_as_uint(xs[e] - __uint_as_float(h));
                        }
                    }
                } else {
                    const float *sDf = reinterpret_cast<const float *>(sD);
#pragma unroll
                    for (int kk = 0; kk < 32; kk++) {
                        const float x = sDf[kk * 128 + r] - mu_g;
                        const uint32_t h = __float_as_uint(x) & 0xffffe000u;
                        hi[kk] = h;
                        lo[kk] = __float_as_uint(x - __uint_as_float(h));
                    }
                }
                mbar_wait(a_free + ab, aph ^ 1);
                fence_after();
                tmem_st32(lane_addr + TM_A + ab * 64, hi);
                tmem_st32(lane_addr + TM_A + ab * 64 + 32, lo);
                tmem_st_wait();
                fence_before();
                mbar_arrive(a_ready + ab);
                if ((c % FLUSH) == 0 && c > 0) {  // drain the partial sum of the previous FLUSH chunks
                    mbar_wait(part_full, ((c / FLUSH) - 1) & 1);
                    fence_after();
                    drain(TM_BIG);
                    fence_before();
                    mbar_arrive(part_free);
                }
            }
            // ---- epilogue: last partial + the cross terms
            mbar_wait(part_full, ((nk - 1) / FLUSH) & 1);
            fence_after();
            drain(TM_BIG);
            drain(TM_SMALL);
            uint32_t v[48];
#pragma unroll
            for (int i = 0; i < 48; i++) v[i] = __float_as_uint(sum[i]);
            if (!TR) {
                const int64_t row = (int64_t)tile0 + r;
                if (r < tile_rows && row < n_rows) {
                    if (Y != nullptr) {
#pragma unroll
                        for (int j = 0; j < LP; j += 4)
                            *reinterpret_cast<float4 *>(Y + row * LP + j) =
                                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                            __uint_as_float(v[j + 3]));
                    }
                    if (ytiles != nullptr) {  // operand of the following D^T Y product, TF32 hi / lo parts
#pragma unroll
                        for (int j = 0; j < LP; j++) {
                            const float y = __uint_as_float(v[j]);
                            const float h = __uint_as_float(v[j] & 0xffffe000u);
                            *reinterpret_cast<float *>(ytiles + dd_tc_b_offset(row, j, 0)) = h;
                            *reinterpret_cast<float *>(ytiles + dd_tc_b_offset(row, j, 1)) = y - h;
                        }
                    }
                }
            } else {
                const int gene = tile0 + r;
                if (gene < ld) {
#pragma unroll
                    for (int j = 0; j < LP; j++) atomicAdd(Zacc + (int64_t)gene * LP + j, (double)__uint_as_float(v[j]));
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace tcg

struct dd_tc_state {
    CUtensorMap map_rows;   // box 32 genes x 128 rows, 128B swizzle   (GEMM1)
    CUtensorMap map_genes;  // box 128 genes x 32 rows, no swizzle    (GEMM2)
    const float *dense = nullptr;
    int64_t rows = 0, ld = 0;
    int tile_rows = 128;  // rows per CTA of GEMM1
};

bool dd_tc_pca_enabled() {
    static const bool off = getenv("DD_PCA_FFMA") != nullptr;
    return !off;
}

int dd_tc_prepare(dd_handle *h) {
    if (!h->tc) h->tc = new dd_tc_state();
    dd_tc_state *st = h->tc;
    if (st->dense == h->d_dense && st->rows == h->A && st->ld == h->ld) return DD_OK;
    tcg::EncodeTiledFn enc = tcg::get_encode();
    if (!enc) return dd_fail(h, DD_ERR_CUDA, "pca: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)h->ld, (cuuint64_t)h->A};
    const cuuint64_t strides[1] = {(cuuint64_t)h->ld * sizeof(float)};
    const cuuint32_t estr[2] = {1, 1};
    // GEMM1 grid balance: ceil(A / 128) CTAs rarely fill whole waves of 2 CTAs per SM; shrink the row tile so
    // that the same number of waves is full (the MMA still runs M = 128, the surplus lanes are ignored)
    const int64_t slots = (int64_t)h->num_sms * 2;
    const int64_t waves = std::max<int64_t>(1, ((h->A + 127) / 128 + slots - 1) / slots);
    int tile_rows = (int)((h->A + slots * waves - 1) / (slots * waves));
    tile_rows = std::min(128, std::max(8, (tile_rows + 7) / 8 * 8));
    st->tile_rows = tile_rows;
    const cuuint32_t box_rows[2] = {32, (cuuint32_t)tile_rows};
    const cuuint32_t box_genes[2] = {128, 32};
    CUresult r1 = enc(&st->map_rows, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->d_dense, dims, strides, box_rows, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&st->map_genes, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, h->d_dense, dims, strides, box_genes, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS)
        return dd_fail(h, DD_ERR_CUDA, "pca: cuTensorMapEncodeTiled failed (" + std::to_string((int)r1) + ", " +
                                           std::to_string((int)r2) + ")");
    st->dense = h->d_dense;
    st->rows = h->A;
    st->ld = h->ld;
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(tcg::k_tc_gemm<false, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg::SMEM_BYTES);
        cudaFuncSetAttribute(tcg::k_tc_gemm<true, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcg::SMEM_BYTES);
    });
    return DD_OK;
}

void dd_tc_free(dd_handle *h) {
    delete h->tc;
    h->tc = nullptr;
}

// Y = Dc Q   (Q as canonical hi/lo tiles in h->d_qb, Dc = D - mu centred on the fly).  write_y: store Y row-major
// (h->d_Y); write_tiles: store Y as hi/lo operand tiles (h->d_yb) for the following Dc^T Y.
int dd_tc_gemm_dq(dd_handle *h, bool write_y, bool write_tiles) {
    const int tr = h->tc->tile_rows;
    const unsigned grid = (unsigned)((h->A + tr - 1) / tr);
    DD_LAUNCH(h, "tc_gemm_dq", (tcg::k_tc_gemm<false, 40>), grid, 192, tcg::SMEM_BYTES, h->tc->map_rows, h->d_qb, h->d_mu,
              write_y ? h->d_Y : (float *)nullptr, write_tiles ? h->d_yb : (uint8_t *)nullptr, (double *)nullptr, h->A,
              (int)h->ld, 0, tr);
    return DD_OK;
}

// Zacc += D^T Y'   (Y' as canonical hi/lo tiles in h->d_yb)
int dd_tc_gemm_dty(dd_handle *h) {
    const int gblocks = (int)((h->ld + 127) / 128);
    const int total_chunks = (int)((h->A + tcg::BK - 1) / tcg::BK);
    // two CTAs per SM are resident: aim for two full waves (never a nearly empty third one)
    int splits = std::max(1, (h->num_sms * 4) / gblocks);
    int cps = std::max(1, (total_chunks + splits - 1) / splits);
    splits = (total_chunks + cps - 1) / cps;
    DD_LAUNCH(h, "tc_gemm_dty", (tcg::k_tc_gemm<true, 40>), dim3(gblocks, splits), 192, tcg::SMEM_BYTES, h->tc->map_genes,
              h->d_yb, h->d_mu, (float *)nullptr, (uint8_t *)nullptr, h->d_Zacc, h->A, (int)h->ld, cps, 128);
    return DD_OK;
}

// Host-side packing of Omega (G x L, row-major) into the canonical operand tiles.
void dd_tc_pack_omega(const float *omega, int64_t n_genes, int n_random, int64_t ld, std::vector<uint8_t> &out) {
    const int64_t chunks = ld / tcg::BK;
    out.assign((size_t)chunks * tcg::B_TILE_BYTES, 0);
    for (int64_t g = 0; g < n_genes; g++)
        for (int j = 0; j < n_random; j++) {
            const float x = omega[g * n_random + j];
            uint32_t bits;
            memcpy(&bits, &x, 4);
            bits &= 0xffffe000u;
            float hi;
            memcpy(&hi, &bits, 4);
            const float lo = x - hi;
            memcpy(out.data() + dd_tc_b_offset(g, j, 0), &hi, 4);
            memcpy(out.data() + dd_tc_b_offset(g, j, 1), &lo, 4);
        }
}
