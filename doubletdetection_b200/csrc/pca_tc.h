// pca_tc.h -- interface between pca.cu and the tcgen05 GEMMs of pca_tc.cu.
#pragma once
#include <stdint.h>

#include <vector>

struct dd_handle;

// Byte offset of element (k, n) of the small GEMM operand (Q: k = gene, Y': k = row; n = column < 48) inside
// the canonical K-major UMMA tiles: per 32-k chunk one 12 KB tile = [hi part | lo part], each part 48 x 32
// floats stored as 8 x 16-byte core matrices (LBO = 128 B along k, SBO = 1024 B along n).
__host__ __device__ inline size_t dd_tc_b_offset(int64_t k, int n, int part) {
    const int64_t c = k >> 5;
    const int kk = (int)(k & 31);
    return (size_t)c * 12288 + (size_t)part * 6144 + (size_t)(n >> 3) * 1024 + (size_t)(kk >> 2) * 128 + (size_t)(n & 7) * 16 +
           (size_t)(kk & 3) * 4;
}

bool dd_tc_pca_enabled();
int dd_tc_prepare(dd_handle *h);
void dd_tc_free(dd_handle *h);
int dd_tc_gemm_dq(dd_handle *h, bool write_y, bool write_tiles);
int dd_tc_gemm_dty(dd_handle *h);
void dd_tc_pack_omega(const float *omega, int64_t n_genes, int n_random, int64_t ld, std::vector<uint8_t> &out);
