// pca.cu -- sc.tl.pca(svd_solver="auto") on the dense augmented matrix == sklearn's randomized
// truncated SVD (doubletdetection.py:309-314 -> sklearn/decomposition/_pca.py:726-758,
// sklearn/utils/extmath.py:313-385, 560-633, 970-981):
//     Q = Omega;  repeat n_iter: Q = normalise(Dc Q); Q = normalise(Dc^T Q);
//     Q = qr(Dc Q);  B = Q^T Dc;  svd(B);  U = Q Uhat;  flip signs by Vt;  X_pca = U[:, :C] * s[:C]
// where Dc is the column-centred dense matrix D.
//
// B200 design (no centred copy, no host round trips, everything on one stream):
//   * the two tall-skinny products  Y = D Q  (A x L)  and  Z = D^T Y  (G x L)  stream D from HBM once
//     each (HBM-bound: 2L/4 = 20 flop per byte).  This file holds the fp32 CUDA-core version (cp.async
//     3-stage pipelines, conflict-free 128-bit shared-memory reads).
//   * centring is implicit:  Dc Q = D Q - 1 (mean(D Q)),  and  Dc^T Y' = D^T Y' - mu (1^T Y')  where the
//     second term is applied to the G x L result from the exactly accumulated column sums.
//   * sklearn's LU normaliser only fixes the span and the conditioning; the span-equivalent
//     CholeskyQR (Gram in float64 -> Cholesky -> triangular inverse) is used on both the tall and the
//     small side.  svd(B) is taken from the float64 Jacobi eigen-decomposition of B B^T (L x L).
#include "dd_internal.h"
#include "pca_tc.h"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// small-workspace layout (doubles), sized for LP = 64
constexpr int kMaxLP = 64;
constexpr int OFF_GRAM = 0;
constexpr int OFF_RINV = OFF_GRAM + kMaxLP * kMaxLP;
constexpr int OFF_CSUM = OFF_RINV + kMaxLP * kMaxLP;   // column sums of Y
constexpr int OFF_SSUM = OFF_CSUM + kMaxLP;            // column sums of the orthonormalised Y'
constexpr int OFF_EVEC = OFF_SSUM + kMaxLP;
constexpr int OFF_EVAL = OFF_EVEC + kMaxLP * kMaxLP;
constexpr int OFF_T = OFF_EVAL + kMaxLP;               // L x KP embedding transform
constexpr int OFF_FLAG = OFF_T + kMaxLP * kMaxLP;      // != 0: a factorisation broke down
constexpr int OFF_KEYS = OFF_FLAG + 8;                 // per component: packed (|w|, gene, sign) arg-max keys
constexpr int SMALL_DOUBLES = OFF_KEYS + kMaxLP;

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

// ------------------------------------------------------------------------------------------------
// Y (A x LP) = D (A x ld) * Qt^T,  Qt is LP x ld (Q transposed; pad rows / columns are zero)
constexpr int G1_BM = 128, G1_BK = 32, G1_ST = 36, G1_NS = 3;
template <int LP>
constexpr size_t gemm1_smem() { return sizeof(float) * G1_NS * (G1_BM + LP) * G1_ST; }

template <int LP>
__global__ void __launch_bounds__(128) k_gemm_dq(const float *__restrict__ D, const float *__restrict__ Qt,
                                                 float *__restrict__ Y, int64_t n_rows, int ld) {
    constexpr int CPT = LP / 8;
    extern __shared__ __align__(16) float sm[];
    float *As = sm;
    float *Bs = sm + G1_NS * G1_BM * G1_ST;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, rg = lane >> 3, cg = lane & 7;
    const int64_t row0 = (int64_t)blockIdx.x * G1_BM;
    const int nk = ld / G1_BK;

    auto load_stage = [&](int s, int kt) {
        const int k0 = kt * G1_BK;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int c = tid + 128 * i, r = c >> 3, kc = c & 7;
            const int64_t grow = row0 + r;
            const bool ok = grow < n_rows;
            cp_async16(As + (s * G1_BM + r) * G1_ST + kc * 4, D + (ok ? grow : 0) * ld + k0 + kc * 4, ok);
        }
        for (int c = tid; c < LP * 8; c += 128) {
            const int r = c >> 3, kc = c & 7;
            cp_async16(Bs + (s * LP + r) * G1_ST + kc * 4, Qt + (int64_t)r * ld + k0 + kc * 4, true);
        }
    };

    float acc[8][CPT];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < CPT; j++) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < G1_NS - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<G1_NS - 2>();
        __syncthreads();
        const int nxt = kt + G1_NS - 1;
        if (nxt < nk) load_stage(nxt % G1_NS, nxt);
        cp_async_commit();
        const int s = kt % G1_NS;
        const float *a_base = As + (s * G1_BM + warp * 32 + rg) * G1_ST;
        const float *b_base = Bs + (s * LP + cg) * G1_ST;
#pragma unroll
        for (int kk = 0; kk < G1_BK; kk += 4) {
            float4 a[8], b[CPT];
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = *reinterpret_cast<const float4 *>(a_base + 4 * i * G1_ST + kk);
#pragma unroll
            for (int j = 0; j < CPT; j++) b[j] = *reinterpret_cast<const float4 *>(b_base + 8 * j * G1_ST + kk);
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < CPT; j++) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t row = row0 + warp * 32 + rg + 4 * i;
        if (row < n_rows) {
#pragma unroll
            for (int j = 0; j < CPT; j++) Y[row * LP + cg + 8 * j] = acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Zacc (ld x LP, double) += D[r_begin:r_end, :]^T * Y[r_begin:r_end, :]   (split over row ranges)
constexpr int G2_BG = 256, G2_BK = 32, G2_NS = 3;
template <int LP>
constexpr size_t gemm2_smem() { return sizeof(float) * G2_NS * G2_BK * (G2_BG + LP); }

template <int LP>
__global__ void __launch_bounds__(128) k_gemm_dty(const float *__restrict__ D, const float *__restrict__ Y,
                                                  double *__restrict__ Zacc, int64_t n_rows, int ld,
                                                  int rows_per_split) {
    constexpr int CPT = LP / 4;
    constexpr int YCH = LP / 4;  // 16-byte chunks per row of Y
    extern __shared__ __align__(16) float sm[];
    float *As = sm;                               // [NS][BK][BG]
    float *Bs = sm + G2_NS * G2_BK * G2_BG;       // [NS][BK][LP]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane & 7, cq = lane >> 3;
    const int g0 = blockIdx.x * G2_BG;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_split;
    const int64_t r_end = min(n_rows, r_begin + rows_per_split);
    const int nk = (int)((r_end - r_begin + G2_BK - 1) / G2_BK);

    auto load_stage = [&](int s, int kt) {
        const int64_t rbase = r_begin + (int64_t)kt * G2_BK;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int c = tid + 128 * i, r = c >> 6, gc = c & 63;
            const int64_t grow = rbase + r;
            const int gcol = g0 + gc * 4;
            const bool ok = grow < r_end && gcol < ld;
            cp_async16(As + (s * G2_BK + r) * G2_BG + gc * 4, D + (ok ? grow * ld + gcol : 0), ok);
        }
        for (int c = tid; c < G2_BK * YCH; c += 128) {
            const int r = c / YCH, cc = c % YCH;
            const int64_t grow = rbase + r;
            const bool ok = grow < r_end;
            cp_async16(Bs + (s * G2_BK + r) * LP + cc * 4, Y + (ok ? grow * LP + cc * 4 : 0), ok);
        }
    };

    float acc[8][CPT];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < CPT; j++) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < G2_NS - 1; s++) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<G2_NS - 2>();
        __syncthreads();
        const int nxt = kt + G2_NS - 1;
        if (nxt < nk) load_stage(nxt % G2_NS, nxt);
        cp_async_commit();
        const int s = kt % G2_NS;
        const float *a_ptr = As + s * G2_BK * G2_BG + warp * 64 + gq * 4;
        const float *b_ptr = Bs + s * G2_BK * LP + cq * CPT;
#pragma unroll 4
        for (int k = 0; k < G2_BK; k++) {
            const float4 a0 = *reinterpret_cast<const float4 *>(a_ptr + k * G2_BG);
            const float4 a1 = *reinterpret_cast<const float4 *>(a_ptr + k * G2_BG + 32);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j += 2) {
                const float2 t = *reinterpret_cast<const float2 *>(b_ptr + k * LP + j);
                b[j] = t.x;
                b[j + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < CPT; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int gene = g0 + warp * 64 + gq * 4 + (i & 3) + (i >> 2) * 32;
        if (gene < ld) {
#pragma unroll
            for (int j = 0; j < CPT; j++) atomicAdd(Zacc + (int64_t)gene * LP + cq * CPT + j, (double)acc[i][j]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Gram matrix (and column sums) of a tall matrix in float64.
//   MODE 0: X = Y (float, n x LP).
//   MODE 1: X = Zacc - mu * ssum^T (double, n = ld rows = genes), written back in place.
constexpr int GR_ROWS = 64;
template <int LP, int MODE>
__global__ void __launch_bounds__(256) k_gram(const float *__restrict__ Yf, double *__restrict__ Zd, int64_t n_rows,
                                              const double *__restrict__ colsum_mu, double inv_n_mu,
                                              const double *__restrict__ ssum, double *__restrict__ gram,
                                              double *__restrict__ csum) {
    constexpr int BPT = LP / 8;
    __shared__ double tile[GR_ROWS][LP + 1];
    const int tid = threadIdx.x, grp = tid >> 6, ti = (tid & 63) >> 3, tj = tid & 7;
    double acc[BPT][BPT];
#pragma unroll
    for (int a = 0; a < BPT; a++)
#pragma unroll
        for (int b = 0; b < BPT; b++) acc[a][b] = 0.0;
    double cs = 0.0;
    const int64_t n_tiles = (n_rows + GR_ROWS - 1) / GR_ROWS;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * GR_ROWS;
        for (int e = tid; e < GR_ROWS * LP; e += 256) {
            const int r = e / LP, c = e % LP;
            const int64_t row = r0 + r;
            double v = 0.0;
            if (row < n_rows) {
                if (MODE == 0) {
                    v = (double)Yf[row * LP + c];
                } else {
                    v = Zd[row * LP + c] - (colsum_mu[row] * inv_n_mu) * ssum[c];
                    Zd[row * LP + c] = v;
                }
            }
            tile[r][c] = v;
        }
        __syncthreads();
        for (int r = grp * 16; r < grp * 16 + 16; r++) {
            double a[BPT], b[BPT];
#pragma unroll
            for (int x = 0; x < BPT; x++) {
                a[x] = tile[r][ti * BPT + x];
                b[x] = tile[r][tj * BPT + x];
            }
#pragma unroll
            for (int x = 0; x < BPT; x++)
#pragma unroll
                for (int y = 0; y < BPT; y++) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
        if (tid < LP) {
            for (int r = 0; r < GR_ROWS; r++) cs += tile[r][tid];
        }
        __syncthreads();
    }
#pragma unroll
    for (int x = 0; x < BPT; x++)
#pragma unroll
        for (int y = 0; y < BPT; y++) atomicAdd(gram + (ti * BPT + x) * LP + tj * BPT + y, acc[x][y]);
    if (tid < LP) atomicAdd(csum + tid, cs);
}

// Centre the Gram matrix (optional), factor it, Gc = R^T R (R upper triangular, right-looking Cholesky with the
// trailing update spread over the whole CTA) and invert the factor.  Output: R^-1 (LP x LP, zero outside the
// L x L upper triangle).  One CTA of 512 threads.
__global__ void __launch_bounds__(512) k_chol(const double *__restrict__ gram, const double *__restrict__ csum,
                                              double n_rows, int centre, int L, int LP, double *__restrict__ rout,
                                              double *__restrict__ flag) {
    __shared__ double R[kMaxLP][kMaxLP + 1];
    const int tid = threadIdx.x;
    for (int e = tid; e < L * L; e += blockDim.x) {
        const int i = e / L, j = e % L;
        double g = gram[i * LP + j];
        if (centre) g -= csum[i] * csum[j] / n_rows;
        R[i][j] = g;
    }
    __syncthreads();
    for (int k = 0; k < L; k++) {
        const double d = R[k][k];
        __syncthreads();
        double piv;
        if (!(d > 0.0) || !isfinite(d)) {
            if (tid == 0) flag[0] = 1.0;
            piv = 1.0;
        } else {
            piv = sqrt(d);
        }
        const double inv = 1.0 / piv;
        for (int j = k + tid; j < L; j += blockDim.x) R[k][j] = (j == k) ? piv : R[k][j] * inv;
        __syncthreads();
        const int m = L - k - 1;  // trailing block is m x m, only j >= i is needed
        for (int e = tid; e < m * m; e += blockDim.x) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j >= i) R[i][j] -= R[k][i] * R[k][j];
        }
        __syncthreads();
    }
    // R^-1 (upper triangular) by back substitution, one column per group of 8 lanes; column j of the inverse
    // is kept in row j of the (otherwise unused) lower triangle: X[p][j] lives in R[j][p] for p < j.
    __shared__ double invd[kMaxLP];
    if (tid < L) invd[tid] = 1.0 / R[tid][tid];
    __syncthreads();
    {
        const int j = tid >> 3, sub = tid & 7;
        const unsigned gmask = 0xFFu << (8 * ((tid & 31) >> 3));
        if (j < L) {
            for (int i = j - 1; i >= 0; i--) {
                double part = 0.0;
                for (int p = i + 1 + sub; p <= j; p += 8) part += R[i][p] * (p == j ? invd[j] : R[j][p]);
                part += __shfl_xor_sync(gmask, part, 4);
                part += __shfl_xor_sync(gmask, part, 2);
                part += __shfl_xor_sync(gmask, part, 1);
                if (sub == 0) R[j][i] = -part * invd[i];
                __syncwarp(gmask);
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < LP * LP; e += blockDim.x) {
        const int i = e / LP, j = e % LP;
        double v = 0.0;
        if (i < L && j < L) v = i < j ? R[j][i] : (i == j ? invd[i] : 0.0);
        rout[e] = v;
    }
}

// Orthonormalise with the inverse Cholesky factor (float64 arithmetic, float32 result).
//   MODE 0 (tall): Y <- (Y - csum/n) R^-1 in place, ssum += column sums of the new Y.
//   MODE 1 (small): Qt[j][g] = (Z R^-1)[g][j] (transposed), and Zacc is cleared for the next pass.
// When `bt` is given the result is also written as TF32 hi/lo operand tiles for the next tcgen05 GEMM.
template <int LP, int MODE>
__global__ void __launch_bounds__(128) k_apply(float *__restrict__ Yf, double *__restrict__ Zd, int64_t n_rows,
                                               int L, const double *__restrict__ rinv,
                                               const double *__restrict__ csum, double inv_n,
                                               double *__restrict__ ssum, float *__restrict__ Qt, int ld,
                                               uint8_t *__restrict__ bt) {
    __shared__ double Rs[LP][LP];
    __shared__ double ms[LP];
    for (int e = threadIdx.x; e < LP * LP; e += blockDim.x) Rs[e / LP][e % LP] = rinv[e];
    if (threadIdx.x < LP) ms[threadIdx.x] = (MODE == 0) ? csum[threadIdx.x] * inv_n : 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_pad = (n_rows + 31) / 32 * 32;  // keep warps converged for the shuffles
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_pad; row += stride) {
        const bool ok = row < n_rows;
        double y[LP];
#pragma unroll
        for (int i = 0; i < LP; i++) {
            if (MODE == 0)
                y[i] = ok ? (double)Yf[row * LP + i] - ms[i] : 0.0;
            else
                y[i] = ok ? Zd[row * LP + i] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < LP; j++) {
            double o = 0.0;
#pragma unroll
            for (int i = 0; i <= j; i++) o = fma(y[i], Rs[i][j], o);
            const float of = (float)o;
            if (bt != nullptr && ok && j < L) {  // operand of the next tcgen05 GEMM: TF32-exact high part + remainder
                const float hi = __uint_as_float(__float_as_uint(of) & 0xffffe000u);
                *reinterpret_cast<float *>(bt + dd_tc_b_offset(row, j, 0)) = hi;
                *reinterpret_cast<float *>(bt + dd_tc_b_offset(row, j, 1)) = of - hi;
            }
            if (MODE == 0) {
                if (ok) Yf[row * LP + j] = of;
                double s = ok ? (double)of : 0.0;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (lane == 0 && j < L) atomicAdd(ssum + j, s);
            } else {
                if (ok) {
                    Qt[(int64_t)j * ld + row] = of;
                    Zd[row * LP + j] = 0.0;
                }
            }
        }
    }
}

// Symmetric eigen-decomposition of the L x L Gram matrix by parallel-ordered cyclic Jacobi (float64),
// eigenvalues sorted in decreasing order.  One CTA, kJacobiThreads threads, TWO barriers per round: the n / 2 rotations
// of a round are computed from the diagonal 2 x 2 blocks, then every 2 x 2 block of A gets its column rotation followed by
// its row rotation in one go (one thread per block), and the columns of V are rotated alongside.  The round-robin schedule
// is the closed form of "position 0 stays, positions 1 .. n-1 shift by one per round": pos_r[i] = ((i - 1 - r) mod (n - 1)) + 1.
constexpr int kJacobiThreads = 512;
__global__ void __launch_bounds__(kJacobiThreads) k_jacobi(const double *__restrict__ gram, int L, int LP,
                                                           double *__restrict__ evec, double *__restrict__ eval) {
    extern __shared__ double jac_sm[];
    double (*Am)[kMaxLP + 1] = reinterpret_cast<double (*)[kMaxLP + 1]>(jac_sm);
    double (*Vm)[kMaxLP + 1] = reinterpret_cast<double (*)[kMaxLP + 1]>(jac_sm + kMaxLP * (kMaxLP + 1));
    __shared__ double cs_c[kMaxLP / 2], cs_s[kMaxLP / 2];
    __shared__ int pp[kMaxLP / 2], qq[kMaxLP / 2];
    __shared__ int offmax;
    const int tid = threadIdx.x;
    const int n = (L + 1) & ~1;  // even number of players; index L (if padded) is a bye
    for (int e = tid; e < n * n; e += kJacobiThreads) {
        const int i = e / n, j = e % n;
        Am[i][j] = (i < L && j < L) ? gram[i * LP + j] : 0.0;
        Vm[i][j] = (i == j) ? 1.0 : 0.0;
    }
    if (tid == 0) offmax = 0;
    __syncthreads();
    const int half = n / 2, m = n - 1;
    for (int sweep = 0; sweep < 30; sweep++) {
        for (int round = 0; round < n - 1; round++) {
            if (tid < half) {
                // players at positions tid and n - 1 - tid of this round
                const int i0 = tid, i1 = n - 1 - tid;
                int p = i0 == 0 ? 0 : ((i0 - 1 - round) % m + m) % m + 1;
                int q = ((i1 - 1 - round) % m + m) % m + 1;
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < L) {
                    const double apq = Am[p][q];
                    const double scale = fabs(Am[p][p]) + fabs(Am[q][q]);
                    if (fabs(apq) > 1e-300 && fabs(apq) > 1e-17 * scale) {
                        const double tau = (Am[q][q] - Am[p][p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                        if (fabs(apq) > 1e-14 * scale) offmax = 1;  // benign race: any writer sets it
                    }
                }
                pp[tid] = p; qq[tid] = q; cs_c[tid] = c; cs_s[tid] = s;
            }
            __syncthreads();
            // A <- J^T A J block by block: columns (p2, q2) first, then rows (p1, q1) -- the arithmetic of two separate passes
            for (int e = tid; e < half * half; e += kJacobiThreads) {
                const int t1 = e / half, t2 = e % half;
                const int p1 = pp[t1], q1 = qq[t1], p2 = pp[t2], q2 = qq[t2];
                const double c1 = cs_c[t1], s1 = cs_s[t1], c2 = cs_c[t2], s2 = cs_s[t2];
                const double app = Am[p1][p2], apq = Am[p1][q2], aqp = Am[q1][p2], aqq = Am[q1][q2];
                const double bpp = c2 * app - s2 * apq, bpq = s2 * app + c2 * apq;
                const double bqp = c2 * aqp - s2 * aqq, bqq = s2 * aqp + c2 * aqq;
                Am[p1][p2] = c1 * bpp - s1 * bqp;
                Am[q1][p2] = s1 * bpp + c1 * bqp;
                Am[p1][q2] = c1 * bpq - s1 * bqq;
                Am[q1][q2] = s1 * bpq + c1 * bqq;
            }
            // V <- V J
            for (int e = tid; e < half * n; e += kJacobiThreads) {
                const int t = e / n, i = e % n;
                const int p = pp[t], q = qq[t];
                const double c = cs_c[t], s = cs_s[t];
                const double vip = Vm[i][p], viq = Vm[i][q];
                Vm[i][p] = c * vip - s * viq;
                Vm[i][q] = s * vip + c * viq;
            }
            __syncthreads();
        }
        const bool done = offmax == 0;
        __syncthreads();
        if (done) break;
        if (tid == 0) offmax = 0;
        __syncthreads();  // the first round of the next sweep may set it again
    }
    // sort eigenvalues (descending) by rank counting; ties by index
    if (tid < L) {
        const double d = Am[tid][tid];
        int rank = 0;
        for (int k = 0; k < L; k++) {
            const double dk = Am[k][k];
            rank += (dk > d) || (dk == d && k < tid);
        }
        eval[rank] = d;
        for (int i = 0; i < L; i++) evec[i * LP + rank] = Vm[i][tid];
    }
}

constexpr size_t kJacobiSmem = sizeof(double) * 2 * kMaxLP * (kMaxLP + 1);

// svd_flip(u_based_decision=False): W = Z Uhat = V S; the sign of component j is the sign of the largest-|.|
// entry of column j of W (first gene wins ties).  One thread per gene computes its row of W; the arg-max per
// column is a 64-bit atomicMax over keys  [ |w| (float64 bits, low 21 mantissa bits dropped) | ~gene | sign ].
template <int LP>
__global__ void __launch_bounds__(128) k_signs_w(const double *__restrict__ Zd, const float *__restrict__ Qt, int ld,
                                                 int n_genes, int L, int C, const double *__restrict__ evec,
                                                 unsigned long long *__restrict__ keys) {
    __shared__ double Us[LP][LP + 1];
    for (int e = threadIdx.x; e < LP * LP; e += 128) Us[e / LP][e % LP] = evec[e];
    __syncthreads();
    const int g = blockIdx.x * 128 + threadIdx.x;
    const bool ok = g < n_genes;
    double z[LP];
#pragma unroll
    for (int i = 0; i < LP; i++)  // rows of Z (float64, gene-major) or of Q given as Q^T (float32, LP x ld)
        z[i] = ok ? (Qt != nullptr ? (double)Qt[(int64_t)i * ld + g] : Zd[(int64_t)g * LP + i]) : 0.0;
    const int lane = threadIdx.x & 31;
    for (int j = 0; j < C; j++) {
        double w = 0.0;
#pragma unroll
        for (int i = 0; i < LP; i++) w = fma(z[i], Us[i][j], w);
        unsigned long long key = 0ull;
        if (ok) {
            const unsigned long long mag = ((unsigned long long)__double_as_longlong(fabs(w)) >> 21) << 21;
            key = mag | ((unsigned long long)(0xFFFFFu - (unsigned)g) << 1) | (w < 0.0 ? 1ull : 0ull);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
            key = o > key ? o : key;
        }
        if (lane == 0) atomicMax(keys + j, key);
    }
}

// X_pca (A x KP float, zero padded beyond C) = Yq T,  T[i][c] = Uhat[i][c] s_c sign_c  (U = Q Uhat, scaled
// by the singular values and flipped as sklearn does).  Block 0 also publishes the singular values.
template <int LP, int KP>
__global__ void __launch_bounds__(128) k_embed(const float *__restrict__ Yq, int64_t n_rows, int L, int C,
                                               const double *__restrict__ evec, const double *__restrict__ eval,
                                               const unsigned long long *__restrict__ keys, float *__restrict__ emb,
                                               double *__restrict__ sing_out, int scale_by_s) {
    __shared__ double Ts[LP][KP];
    for (int e = threadIdx.x; e < LP * KP; e += blockDim.x) {
        const int i = e / KP, c = e % KP;
        double v = 0.0;
        if (i < L && c < C)
            v = evec[i * LP + c] * (scale_by_s ? sqrt(fmax(eval[c], 0.0)) : 1.0) * ((keys[c] & 1ull) ? -1.0 : 1.0);
        Ts[i][c] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < C) sing_out[threadIdx.x] = sqrt(fmax(eval[threadIdx.x], 0.0));
    __syncthreads();
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    double y[LP];
#pragma unroll
    for (int i = 0; i < LP; i++) y[i] = (double)Yq[row * LP + i];
#pragma unroll
    for (int c4 = 0; c4 < KP; c4 += 4) {
        float o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < LP; i++) s = fma(y[i], Ts[i][c4 + q], s);
            o[q] = (float)s;
        }
        *reinterpret_cast<float4 *>(emb + row * KP + c4) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// float32 column means (sklearn's mean_) from the float64 column sums
__global__ void k_mu(const double *__restrict__ colsum, int n_genes, int ld, double inv_n, float *__restrict__ mu) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < ld) mu[g] = g < n_genes ? (float)(colsum[g] * inv_n) : 0.f;
}

// ------------------------------------------------------------------------------------------------
template <int LP>
int run_pca(dd_handle *h, int n_power_iter) {
    const int64_t A = h->A;
    const int ld = (int)h->ld;
    const int L = h->L, C = h->C, KP = h->KP;
    double *sm = h->d_small;
    static dd_once_per_device attr_set;  // function attributes are per device
    attr_set.run(h->device, [&] {
        cudaFuncSetAttribute(k_gemm_dq<LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm1_smem<LP>());
        cudaFuncSetAttribute(k_gemm_dty<LP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm2_smem<LP>());
        cudaFuncSetAttribute(k_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJacobiSmem);
    });
    const int grid1 = (int)((A + G1_BM - 1) / G1_BM);
    const int gblocks = (ld + G2_BG - 1) / G2_BG;
    int splits = std::max(1, (h->num_sms * 4 + gblocks - 1) / gblocks);
    int rows_per_split = (int)((A + splits - 1) / splits);
    rows_per_split = std::max(G2_BK, (rows_per_split + G2_BK - 1) / G2_BK * G2_BK);
    splits = (int)((A + rows_per_split - 1) / rows_per_split);
    const int tall_grid = h->num_sms * 2;
    const double inv_A = 1.0 / (double)h->A_glob;  // A = this rank's rows, A_glob = rows over all ranks
    const bool use_tc = LP == 40 && dd_tc_pca_enabled();
    if (dd_sharded(h) && !use_tc)
        return dd_fail(h, DD_ERR_UNSUPPORTED, "pca: cell-block sharding needs the tcgen05 path (n_components + 10 <= 40)");
    if (use_tc) {
        DD_TRY(dd_tc_prepare(h));
        const int64_t q_bytes = (int64_t)(ld / 32) * 12288, y_bytes = ((A + 31) / 32) * 12288;
        if (q_bytes > h->cap_qb || y_bytes > h->cap_yb) {
            DD_TRY(dd_reserve(h, &h->d_qb, &h->cap_qb, q_bytes));
            DD_TRY(dd_reserve(h, &h->d_yb, &h->cap_yb, y_bytes));
            // pad columns (n >= L) and pad rows are never written: they must read as zero
            DD_CUDA(h, cudaMemsetAsync(h->d_qb, 0, (size_t)h->cap_qb, h->stream));
            DD_CUDA(h, cudaMemsetAsync(h->d_yb, 0, (size_t)h->cap_yb, h->stream));
        }
        DD_CUDA(h, cudaMemsetAsync(h->d_yb + (y_bytes - 12288), 0, 12288, h->stream));  // rows >= A of the last chunk
        DD_CUDA(h, cudaMemcpyAsync(h->d_qb, h->d_omega_b, (size_t)q_bytes, cudaMemcpyDeviceToDevice, h->stream));
    }

    DD_CUDA(h, cudaMemsetAsync(sm, 0, sizeof(double) * SMALL_DOUBLES, h->stream));
    DD_CUDA(h, cudaMemsetAsync(h->d_Zacc, 0, sizeof(double) * (size_t)ld * LP, h->stream));
    DD_TRY(dd_dev_colstats(h, false));  // mu = colsum / A

    if (use_tc) {
        // tcgen05 path: D is centred on the fly inside the GEMMs (x - mu_g, float32, exactly sklearn's X -= mean_),
        // so no correction terms exist.  During the power iterations Y goes straight from the first product
        // into the second one (as hi/lo operand tiles written by the GEMM epilogue): normalising the tall panel
        // only re-scales the span (checked against the float64 oracle); the small side is
        // orthonormalised every iteration, the tall side once, for the final projection.
        DD_TRY(dd_reserve(h, &h->d_mu, &h->cap_mu, (int64_t)ld));
        DD_LAUNCH(h, "mu", k_mu, (ld + 255) / 256, 256, 0, h->d_colsum, (int)h->G, ld, inv_A, h->d_mu);
        // Cell-block sharding: every sum over cells is all-reduced (column sums above, D^T Y and the Gram matrix of
        // the tall panel here); the L x L factorisations are replicated and rank 0's result is broadcast, so that
        // every rank multiplies by bit-identical small matrices.
        for (int it = 0; it <= n_power_iter; it++) {
            const bool last = it == n_power_iter;
            DD_TRY(dd_tc_gemm_dq(h, /*write_y=*/last, /*write_tiles=*/!last));
            if (last) {  // Yq = orth(Y): sklearn's qr(A @ Q)
                DD_CUDA(h, cudaMemsetAsync(sm + OFF_GRAM, 0, sizeof(double) * kMaxLP * kMaxLP, h->stream));
                DD_CUDA(h, cudaMemsetAsync(sm + OFF_CSUM, 0, sizeof(double) * 2 * kMaxLP, h->stream));
                DD_LAUNCH(h, "gram_tall", (k_gram<LP, 0>), tall_grid, 256, 0, h->d_Y, nullptr, A, nullptr, 0.0, nullptr,
                          sm + OFF_GRAM, sm + OFF_CSUM);
                DD_TRY(dd_comm_allreduce_f64(h, sm + OFF_GRAM, kMaxLP * kMaxLP));
                DD_TRY(dd_comm_allreduce_f64(h, sm + OFF_CSUM, kMaxLP));
                DD_LAUNCH(h, "chol", k_chol, 1, 512, 0, sm + OFF_GRAM, sm + OFF_CSUM, (double)h->A_glob, 1, L, LP,
                          sm + OFF_RINV, sm + OFF_FLAG);
                DD_TRY(dd_comm_bcast(h, sm + OFF_RINV, sizeof(double) * kMaxLP * kMaxLP, 0));
                DD_LAUNCH(h, "apply_tall", (k_apply<LP, 0>), tall_grid, 128, 0, h->d_Y, nullptr, A, L, sm + OFF_RINV,
                          sm + OFF_CSUM, inv_A, sm + OFF_SSUM, nullptr, 0, h->d_yb);
                DD_CUDA(h, cudaMemsetAsync(sm + OFF_SSUM, 0, sizeof(double) * kMaxLP, h->stream));  // no mu (x) s term
            }
            DD_TRY(dd_tc_gemm_dty(h));
            if (last && h->ev_gemms_done) {  // last pass over the dense matrix: the next iteration's build may overwrite it
                DD_CUDA(h, cudaEventRecord(h->ev_gemms_done, h->stream));
                h->gemms_done_recorded = true;
            }
            DD_TRY(dd_comm_allreduce_f64(h, h->d_Zacc, (int64_t)ld * LP));
            DD_CUDA(h, cudaMemsetAsync(sm + OFF_GRAM, 0, sizeof(double) * kMaxLP * kMaxLP, h->stream));
            DD_LAUNCH(h, "gram_small", (k_gram<LP, 1>), std::min<int>(tall_grid, (ld + GR_ROWS - 1) / GR_ROWS), 256, 0, nullptr,
                      h->d_Zacc, (int64_t)ld, h->d_colsum, inv_A, sm + OFF_SSUM, sm + OFF_GRAM, sm + OFF_EVAL /*unused sums*/);
            if (!last) {
                DD_LAUNCH(h, "chol", k_chol, 1, 512, 0, sm + OFF_GRAM, nullptr, 1.0, 0, L, LP, sm + OFF_RINV, sm + OFF_FLAG);
                DD_TRY(dd_comm_bcast(h, sm + OFF_RINV, sizeof(double) * kMaxLP * kMaxLP, 0));
                DD_LAUNCH(h, "apply_small", (k_apply<LP, 1>), (ld + 127) / 128, 128, 0, nullptr, h->d_Zacc, (int64_t)ld, L,
                          sm + OFF_RINV, nullptr, 0.0, nullptr, h->d_Qt, ld, h->d_qb);
            }
        }
    } else
    for (int it = 0; it <= n_power_iter; it++) {
        const bool last = it == n_power_iter;
        // Y = D Q
        DD_LAUNCH(h, "gemm_dq", k_gemm_dq<LP>, grid1, 128, gemm1_smem<LP>(), h->d_dense, h->d_Qt, h->d_Y, A, ld);
        // Y' = orth(Y - mean)
        DD_CUDA(h, cudaMemsetAsync(sm + OFF_GRAM, 0, sizeof(double) * kMaxLP * kMaxLP, h->stream));
        DD_CUDA(h, cudaMemsetAsync(sm + OFF_CSUM, 0, sizeof(double) * 2 * kMaxLP, h->stream));  // csum + ssum
        DD_LAUNCH(h, "gram_tall", (k_gram<LP, 0>), tall_grid, 256, 0, h->d_Y, nullptr, A, nullptr, 0.0, nullptr,
                  sm + OFF_GRAM, sm + OFF_CSUM);
        DD_LAUNCH(h, "chol", k_chol, 1, 512, 0, sm + OFF_GRAM, sm + OFF_CSUM, (double)A, 1, L, LP, sm + OFF_RINV,
                  sm + OFF_FLAG);
        DD_LAUNCH(h, "apply_tall", (k_apply<LP, 0>), tall_grid, 128, 0, h->d_Y, nullptr, A, L, sm + OFF_RINV,
                  sm + OFF_CSUM, inv_A, sm + OFF_SSUM, nullptr, 0, nullptr);
        // Z = Dc^T Y'
        DD_LAUNCH(h, "gemm_dty", k_gemm_dty<LP>, dim3(gblocks, splits), 128, gemm2_smem<LP>(), h->d_dense, h->d_Y,
                  h->d_Zacc, A, ld, rows_per_split);
        DD_CUDA(h, cudaMemsetAsync(sm + OFF_GRAM, 0, sizeof(double) * kMaxLP * kMaxLP, h->stream));
        DD_LAUNCH(h, "gram_small", (k_gram<LP, 1>), std::min<int>(tall_grid, (ld + GR_ROWS - 1) / GR_ROWS), 256, 0, nullptr,
                  h->d_Zacc, (int64_t)ld, h->d_colsum, inv_A, sm + OFF_SSUM, sm + OFF_GRAM, sm + OFF_EVAL /*unused sums*/);
        if (!last) {
            DD_LAUNCH(h, "chol", k_chol, 1, 512, 0, sm + OFF_GRAM, nullptr, 1.0, 0, L, LP, sm + OFF_RINV,
                      sm + OFF_FLAG);
            DD_LAUNCH(h, "apply_small", (k_apply<LP, 1>), (ld + 127) / 128, 128, 0, nullptr, h->d_Zacc, (int64_t)ld, L,
                      sm + OFF_RINV, nullptr, 0.0, nullptr, h->d_Qt, ld, nullptr);
        }
    }
    // svd(B) with B^T = Z:  B B^T = Z^T Z = gram
    DD_LAUNCH(h, "jacobi", k_jacobi, 1, kJacobiThreads, kJacobiSmem, sm + OFF_GRAM, L, LP, sm + OFF_EVEC, sm + OFF_EVAL);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(sm + OFF_KEYS);
    DD_LAUNCH(h, "signs_w", k_signs_w<LP>, (unsigned)((h->G + 127) / 128), 128, 0, h->d_Zacc, (const float *)nullptr, 0,
              (int)h->G, L, C, sm + OFF_EVEC, keys);
    // eigenvectors, eigenvalues, flag and sign keys: rank 0's copy everywhere (no-op when not sharded)
    DD_TRY(dd_comm_bcast(h, sm + OFF_EVEC, sizeof(double) * (SMALL_DOUBLES - OFF_EVEC), 0));
    // the embedding rows go straight to their place in the global (originals, then synthetics) order
    const bool sh = dd_sharded(h);  // unsharded: one part, local row == global row
    const int64_t part_rows[2] = {sh ? h->blk_n : A, sh ? h->blk_m : 0};
    const int64_t part_src[2] = {0, h->blk_n};
    const int64_t part_dst[2] = {sh ? h->blk_n0 : 0, h->N + h->blk_m0};
    for (int part = 0; part < 2; part++) {
        const int64_t rows = part_rows[part];
        if (rows <= 0) continue;
        const int egrid = (int)((rows + 127) / 128);
        const float *yq = h->d_Y + part_src[part] * LP;
        float *emb = h->d_emb + part_dst[part] * KP;
        if (KP == 32)
            DD_LAUNCH(h, "embed", (k_embed<LP, 32>), egrid, 128, 0, yq, rows, L, C, sm + OFF_EVEC, sm + OFF_EVAL, keys, emb,
                      sm + OFF_CSUM, 1);
        else
            DD_LAUNCH(h, "embed", (k_embed<LP, 64>), egrid, 128, 0, yq, rows, L, C, sm + OFF_EVEC, sm + OFF_EVAL, keys, emb,
                      sm + OFF_CSUM, 1);
    }
    if (dd_sharded(h)) {  // all-gather of the low-dimensional embedding (the one data-path exchange kNN needs)
        std::vector<int64_t> begin, count;
        std::vector<int> owner;
        for (int r = 0; r < h->world; r++) {
            const int64_t n0 = h->N * r / h->world, n1 = h->N * (r + 1) / h->world;
            const int64_t m0 = h->M * r / h->world, m1 = h->M * (r + 1) / h->world;
            begin.push_back(n0); count.push_back(n1 - n0); owner.push_back(r);
            begin.push_back(h->N + m0); count.push_back(m1 - m0); owner.push_back(r);
        }
        DD_TRY(dd_comm_gather_ranges(h, h->d_emb, (int64_t)sizeof(float) * KP, (int)begin.size(), begin.data(),
                                     count.data(), owner.data()));
    }
    return DD_OK;
}

// Fewer rows than columns (A < G): sklearn's randomized_svd works on the transposed problem M = Dc^T
// (sklearn/utils/extmath.py:589-592, 618-620): Omega is A x L, the loop is
//     Q_G = normalise(Dc^T Q_A);  Q_A = normalise(Dc Q_G)        (n_iter times)
//     Q_G = qr(Dc^T Q_A);  B = Q_G^T Dc^T = (Dc Q_G)^T;  svd(B) = Uhat S Vt;  U_int = Q_G Uhat
// and the roles are swapped back at the end: scores U = Vt^T = Y Uhat / S with Y = Dc Q_G, components = U_int^T.
// Hence X_pca = U S = Y Uhat (sign from the largest-|.| entry of each column of Q_G Uhat).  Same kernels as the
// regular path; only the gene side (G x L, the small one) is orthonormalised, in float64.
int run_pca_transposed(dd_handle *h, int n_power_iter) {
    constexpr int LP = 40;
    const int64_t A = h->A;
    const int ld = (int)h->ld;
    const int L = h->L, C = h->C, KP = h->KP;
    double *sm = h->d_small;
    const int tall_grid = h->num_sms * 2;
    const double inv_A = 1.0 / (double)A;
    DD_TRY(dd_tc_prepare(h));
    const int64_t q_bytes = (int64_t)(ld / 32) * 12288, y_bytes = ((A + 31) / 32) * 12288;
    if (q_bytes > h->cap_qb || y_bytes > h->cap_yb) {
        DD_TRY(dd_reserve(h, &h->d_qb, &h->cap_qb, q_bytes));
        DD_TRY(dd_reserve(h, &h->d_yb, &h->cap_yb, y_bytes));
        DD_CUDA(h, cudaMemsetAsync(h->d_qb, 0, (size_t)h->cap_qb, h->stream));
        DD_CUDA(h, cudaMemsetAsync(h->d_yb, 0, (size_t)h->cap_yb, h->stream));
    }
    DD_CUDA(h, cudaMemsetAsync(sm, 0, sizeof(double) * SMALL_DOUBLES, h->stream));
    DD_CUDA(h, cudaMemsetAsync(h->d_Zacc, 0, sizeof(double) * (size_t)ld * LP, h->stream));
    DD_CUDA(h, cudaMemcpyAsync(h->d_yb, h->d_omega_b, (size_t)y_bytes, cudaMemcpyDeviceToDevice, h->stream));  // Q_A = Omega
    DD_TRY(dd_dev_colstats(h, false));
    DD_TRY(dd_reserve(h, &h->d_mu, &h->cap_mu, (int64_t)ld));
    DD_LAUNCH(h, "mu", k_mu, (ld + 255) / 256, 256, 0, h->d_colsum, (int)h->G, ld, inv_A, h->d_mu);
    for (int it = 0; it <= n_power_iter; it++) {
        const bool last = it == n_power_iter;
        DD_TRY(dd_tc_gemm_dty(h));  // Z = Dc^T Q_A
        DD_CUDA(h, cudaMemsetAsync(sm + OFF_GRAM, 0, sizeof(double) * kMaxLP * kMaxLP, h->stream));
        DD_LAUNCH(h, "gram_small", (k_gram<LP, 1>), std::min<int>(tall_grid, (ld + GR_ROWS - 1) / GR_ROWS), 256, 0, nullptr,
                  h->d_Zacc, (int64_t)ld, h->d_colsum, inv_A, sm + OFF_SSUM, sm + OFF_GRAM, sm + OFF_EVAL /*unused sums*/);
        DD_LAUNCH(h, "chol", k_chol, 1, 512, 0, sm + OFF_GRAM, nullptr, 1.0, 0, L, LP, sm + OFF_RINV, sm + OFF_FLAG);
        DD_LAUNCH(h, "apply_small", (k_apply<LP, 1>), (ld + 127) / 128, 128, 0, nullptr, h->d_Zacc, (int64_t)ld, L,
                  sm + OFF_RINV, nullptr, 0.0, nullptr, h->d_Qt, ld, h->d_qb);  // Q_G (orthonormal), Zacc cleared
        DD_TRY(dd_tc_gemm_dq(h, /*write_y=*/last, /*write_tiles=*/!last));     // Y = Dc Q_G
    }
    // svd(B), B^T = Y:  B B^T = Y^T Y
    DD_CUDA(h, cudaMemsetAsync(sm + OFF_GRAM, 0, sizeof(double) * kMaxLP * kMaxLP, h->stream));
    DD_LAUNCH(h, "gram_tall", (k_gram<LP, 0>), tall_grid, 256, 0, h->d_Y, nullptr, A, nullptr, 0.0, nullptr, sm + OFF_GRAM,
              sm + OFF_EVAL /*unused sums*/);
    DD_LAUNCH(h, "jacobi", k_jacobi, 1, kJacobiThreads, kJacobiSmem, sm + OFF_GRAM, L, LP, sm + OFF_EVEC, sm + OFF_EVAL);
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(sm + OFF_KEYS);
    DD_LAUNCH(h, "signs_w", k_signs_w<LP>, (unsigned)((h->G + 127) / 128), 128, 0, (const double *)nullptr, h->d_Qt, ld,
              (int)h->G, L, C, sm + OFF_EVEC, keys);
    const int egrid = (int)((A + 127) / 128);
    if (KP == 32)
        DD_LAUNCH(h, "embed", (k_embed<LP, 32>), egrid, 128, 0, h->d_Y, A, L, C, sm + OFF_EVEC, sm + OFF_EVAL, keys,
                  h->d_emb, sm + OFF_CSUM, 0);
    else
        DD_LAUNCH(h, "embed", (k_embed<LP, 64>), egrid, 128, 0, h->d_Y, A, L, C, sm + OFF_EVEC, sm + OFF_EVAL, keys,
                  h->d_emb, sm + OFF_CSUM, 0);
    return DD_OK;
}

}  // namespace

// Two embedding buffers of `rows` x KP floats (grow-only): h->d_emb points at the one in use.  The fit loop flips between
// them so that the kNN of iteration i (its own stream) reads one while the PCA of iteration i + 1 writes the other.
int dd_emb_reserve(dd_handle *h, int64_t rows, int32_t KP) {
    if (h->KP != KP || rows > h->cap_emb || !h->d_emb_base) {
        if (h->d_emb_base) cudaFree(h->d_emb_base);
        h->d_emb_base = h->d_emb = nullptr;
        h->cap_emb = 0;
        DD_CUDA(h, cudaMalloc(&h->d_emb_base, sizeof(float) * 2 * rows * KP));
        h->cap_emb = rows;
        h->emb_stride = rows * KP;
        h->d_emb = h->d_emb_base;
    }
    h->KP = KP;
    return DD_OK;
}

// Randomized PCA of the current dense matrix.  omega_host (G x n_random, row-major float32) may be
// NULL to reuse the matrix uploaded by the previous call.
int dd_dev_pca(dd_handle *h, int32_t n_comp, int32_t n_random, int32_t n_power_iter, const float *omega_host) {
    if (!h->dense_valid) return dd_fail(h, DD_ERR_ARG, "pca: no dense matrix (call dd_normalise_log first)");
    if (n_comp < 1 || n_random < n_comp || n_power_iter < 0) return dd_fail(h, DD_ERR_ARG, "pca: bad n_comp / n_random / n_power_iter");
    if (n_random > kMaxLP) return dd_fail(h, DD_ERR_UNSUPPORTED, "pca: n_components + 10 > 64 is outside the B200 hot path");
    if (h->G > 0xFFFFF) return dd_fail(h, DD_ERR_UNSUPPORTED, "pca: more than 2^20 - 1 genes");
    if (n_random > h->G || n_random > h->A_glob)
        return dd_fail(h, DD_ERR_UNSUPPORTED, "pca: n_components + 10 exceeds the matrix dimensions");
    const bool transposed = h->A_glob < h->G;  // sklearn works on the transposed problem: Omega is A x n_random
    if (transposed && dd_sharded(h))
        return dd_fail(h, DD_ERR_UNSUPPORTED, "pca: cell-block sharding with fewer augmented cells than genes");
    if (transposed && (n_random > 40 || !dd_tc_pca_enabled()))
        return dd_fail(h, DD_ERR_UNSUPPORTED,
                       "pca: fewer augmented cells than genes needs the tcgen05 path (n_components + 10 <= 40)");
    const int LP = n_random <= 40 ? 40 : 64;
    const int KP = n_comp <= 32 ? 32 : 64;
    const int64_t A = h->A, ld = h->ld;
    if (A > h->cap_pca_rows || ld > h->cap_pca_cols || LP > h->cap_LP) {
        const int64_t rows = std::max(A, h->cap_pca_rows), cols = std::max(ld, h->cap_pca_cols);
        const int lp = std::max(LP, (int)h->cap_LP);
        for (void *p : {(void *)h->d_Qt, (void *)h->d_Y, (void *)h->d_Zacc, (void *)h->d_small})
            if (p) cudaFree(p);
        h->d_Qt = nullptr; h->d_Y = nullptr; h->d_Zacc = nullptr; h->d_small = nullptr;
        h->cap_pca_rows = h->cap_pca_cols = 0; h->cap_LP = 0;
        DD_CUDA(h, cudaMalloc(&h->d_Qt, sizeof(float) * 2 * lp * cols));  // Qt and the pristine Omega^T
        DD_CUDA(h, cudaMalloc(&h->d_Y, sizeof(float) * rows * lp));
        DD_CUDA(h, cudaMalloc(&h->d_Zacc, sizeof(double) * cols * lp));
        DD_CUDA(h, cudaMalloc(&h->d_small, sizeof(double) * SMALL_DOUBLES));
        h->cap_pca_rows = rows; h->cap_pca_cols = cols; h->cap_LP = lp;
        h->L = 0;  // forces an Omega upload
    }
    float *omega_dev = h->d_Qt + (size_t)h->cap_LP * h->cap_pca_cols;
    if (omega_host && transposed) {
        std::vector<uint8_t> packed;
        dd_tc_pack_omega(omega_host, A, n_random, dd_round_up(A, 32), packed);  // rows of Omega are cells here
        DD_TRY(dd_reserve(h, &h->d_omega_b, &h->cap_omega_b, (int64_t)packed.size()));
        DD_CUDA(h, cudaMemcpyAsync(h->d_omega_b, packed.data(), packed.size(), cudaMemcpyHostToDevice, h->stream));
        DD_CUDA(h, cudaStreamSynchronize(h->stream));
    } else if (omega_host) {
        std::vector<float> qt((size_t)LP * ld, 0.f);
        for (int64_t g = 0; g < h->G; g++)
            for (int j = 0; j < n_random; j++) qt[(size_t)j * ld + g] = omega_host[g * n_random + j];
        DD_CUDA(h, cudaMemcpyAsync(omega_dev, qt.data(), sizeof(float) * LP * ld, cudaMemcpyHostToDevice, h->stream));
        std::vector<uint8_t> packed;
        if (LP == 40 && dd_tc_pca_enabled()) {
            dd_tc_pack_omega(omega_host, h->G, n_random, ld, packed);
            DD_TRY(dd_reserve(h, &h->d_omega_b, &h->cap_omega_b, (int64_t)packed.size()));
            DD_CUDA(h, cudaMemcpyAsync(h->d_omega_b, packed.data(), packed.size(), cudaMemcpyHostToDevice, h->stream));
        }
        DD_CUDA(h, cudaStreamSynchronize(h->stream));  // qt / packed are temporaries
    } else if (h->L != n_random || h->LP != LP) {
        return dd_fail(h, DD_ERR_ARG, "pca: omega is NULL but no matching test matrix was uploaded before");
    }
    h->L = n_random; h->LP = LP; h->C = n_comp;
    if (!transposed)
        DD_CUDA(h, cudaMemcpyAsync(h->d_Qt, omega_dev, sizeof(float) * LP * ld, cudaMemcpyDeviceToDevice, h->stream));
    DD_TRY(dd_emb_reserve(h, h->A_glob, KP));  // the embedding is global: every rank holds all A_glob rows
    int rc = transposed ? run_pca_transposed(h, n_power_iter)
                        : ((LP == 40) ? run_pca<40>(h, n_power_iter) : run_pca<64>(h, n_power_iter));
    if (rc != DD_OK) return rc;
    h->emb_rows = h->A_glob;
    h->emb_valid = true;
    return DD_OK;
}

// Checks the factorisation flag (synchronises the stream).
int dd_pca_check(dd_handle *h) {
    double flag = 0.0;
    DD_CUDA(h, cudaMemcpyAsync(&flag, h->d_small + OFF_FLAG, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (flag != 0.0) {
        h->emb_valid = false;
        return dd_fail(h, DD_ERR_UNSUPPORTED, "pca: rank-deficient range (Cholesky breakdown); the matrix has fewer than n_components + 10 independent directions");
    }
    return DD_OK;
}

// Queue an asynchronous copy of the factorisation flag into (pinned) host memory.
int dd_pca_flag_copy(dd_handle *h, double *host_flag) {
    DD_CUDA(h, cudaMemcpyAsync(host_flag, h->d_small + OFF_FLAG, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return DD_OK;
}

extern "C" int dd_pca(dd_handle *h, int32_t n_comp, int32_t n_random, int32_t n_power_iter, const float *omega,
                      float *emb_out, double *singular_values_out) {
    if (!h) return dd_fail(nullptr, DD_ERR_ARG, "dd_pca: null handle");
    if (!omega) return dd_fail(h, DD_ERR_ARG, "dd_pca: null omega");
    DD_CUDA(h, cudaSetDevice(h->device));
    DD_TRY(dd_stage_begin(h));
    DD_TRY(dd_dev_pca(h, n_comp, n_random, n_power_iter, omega));
    DD_TRY(dd_stage_end(h, "pca"));
    DD_TRY(dd_pca_check(h));
    if (emb_out)
        DD_CUDA(h, cudaMemcpy2DAsync(emb_out, sizeof(float) * n_comp, h->d_emb, sizeof(float) * h->KP,
                                     sizeof(float) * n_comp, h->emb_rows, cudaMemcpyDeviceToHost, h->stream));
    if (singular_values_out)
        DD_CUDA(h, cudaMemcpyAsync(singular_values_out, h->d_small + OFF_CSUM, sizeof(double) * n_comp,
                                   cudaMemcpyDeviceToHost, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    return DD_OK;
}

extern "C" int dd_upload_embedding(dd_handle *h, int64_t n_rows, int32_t n_comp, const float *emb) {
    if (!h || !emb || n_rows <= 0 || n_comp <= 0) return dd_fail(h, DD_ERR_ARG, "dd_upload_embedding: bad arguments");
    if (n_comp > 64) return dd_fail(h, DD_ERR_UNSUPPORTED, "dd_upload_embedding: more than 64 components");
    DD_CUDA(h, cudaSetDevice(h->device));
    const int KP = n_comp <= 32 ? 32 : 64;
    DD_TRY(dd_emb_reserve(h, n_rows, KP));
    h->KP = KP; h->C = n_comp;
    DD_CUDA(h, cudaMemsetAsync(h->d_emb, 0, sizeof(float) * n_rows * KP, h->stream));
    DD_CUDA(h, cudaMemcpy2DAsync(h->d_emb, sizeof(float) * KP, emb, sizeof(float) * n_comp, sizeof(float) * n_comp,
                                 n_rows, cudaMemcpyHostToDevice, h->stream));
    DD_CUDA(h, cudaStreamSynchronize(h->stream));
    h->emb_rows = n_rows;
    h->emb_valid = true;
    return DD_OK;
}
