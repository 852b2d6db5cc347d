// louvain_host.h -- pieces shared by the host community-detection code (louvain.cpp, leiden.cpp): the seeded
// generator, the CSR graph with self-loop weights, aggregation and the final label order.  Specification:
// oracle/louvain_ref.py.  Not part of the ABI.
#pragma once

#include <stdint.h>

#include <vector>

namespace ddlv {

struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
};

struct Graph {
    int32_t n = 0;
    std::vector<int64_t> indptr;
    std::vector<int32_t> indices;
    std::vector<double> weights;  // empty = all ones
    std::vector<double> selfw;
    double w(int64_t e) const { return weights.empty() ? 1.0 : weights[e]; }
    // The sweeps visit rows in a (pseudo)random order that is known a few steps ahead: the row's offset first, its
    // adjacency / weight lines once the offset has arrived (15-20 entries = 1 + 2-3 cache lines).  No semantic effect.
    void prefetch_offset(int32_t v) const { __builtin_prefetch(indptr.data() + v); }
    void prefetch_row(int32_t v) const {
        const int64_t e = indptr[v];
        __builtin_prefetch(indices.data() + e);
        if (weights.size() > 1) {
            const char *wl = reinterpret_cast<const char *>(weights.data() + e);
            __builtin_prefetch(wl);
            __builtin_prefetch(wl + 64);  // a hint: an address past the end is harmless
        }
    }
};

// aggregate g by comm (ids in [0, n)); node2new renumbers communities by first appearance over node index
void aggregate(const Graph &g, const std::vector<int32_t> &comm, Graph &out, std::vector<int32_t> &node2new);

// first-appearance ids, then by decreasing size (ties: smaller first-appearance id first)
void labels_by_size(std::vector<int32_t> &membership, int32_t *labels_out, int32_t *n_comm_out);

}  // namespace ddlv
