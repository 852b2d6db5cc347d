/* Deterministic Louvain specification, plain C (TEST INFRASTRUCTURE -- see oracle/__init__.py).
 *
 * Same algorithm, statement for statement, as oracle/louvain_ref.py (which documents it and cites
 * the reference call site doubletdetection.py:337-342); exists so that the oracle finishes in
 * seconds on 10^4..10^5-node graphs and can serve as the CPU baseline.  Straightforward data
 * structures on purpose -- this is the checker, not the product (the product's implementation is
 * doubletdetection_b200/csrc/louvain.cpp and must produce the same labels).
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC -o oracle/_build/liblouvain_ref.so oracle/louvain_ref.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } sm64;
static uint64_t sm64_next(sm64 *r) {
    r->s += 0x9E3779B97F4A7C15ULL;
    uint64_t z = r->s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

typedef struct {
    int64_t n;
    int64_t *indptr;
    int64_t *indices;
    double *weights;
    double *selfw;
} graph;

static void graph_free(graph *g) {
    free(g->indptr); free(g->indices); free(g->weights); free(g->selfw);
}

/* one level of local moving; comm[] out; returns 1 if any node moved */
static int one_level(const graph *g, double gamma, double two_m, sm64 *rng, int64_t *comm) {
    int64_t n = g->n;
    double *k = (double *)malloc(sizeof(double) * n);
    double *tot = (double *)malloc(sizeof(double) * n);
    double *neigh_w = (double *)malloc(sizeof(double) * n);
    char *seen = (char *)calloc(n, 1);
    char *inq = (char *)malloc(n);
    int64_t *queue = (int64_t *)malloc(sizeof(int64_t) * n); /* ring buffer, capacity n */
    int64_t *cands = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) {
        double s = g->selfw[i];
        for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) s += g->weights[e];
        k[i] = s; tot[i] = s; comm[i] = i; queue[i] = i; inq[i] = 1;
    }
    for (int64_t i = n - 1; i >= 1; i--) {
        int64_t j = (int64_t)(sm64_next(rng) % (uint64_t)(i + 1));
        int64_t t = queue[i]; queue[i] = queue[j]; queue[j] = t;
    }
    int64_t head = 0, count = n;
    int moved_any = 0;
    while (count > 0) {
        int64_t i = queue[head];
        head = (head + 1 == n) ? 0 : head + 1;
        count--;
        inq[i] = 0;
        int64_t ci = comm[i];
        double ki = k[i];
        int64_t nc = 0;
        cands[nc++] = ci; seen[ci] = 1; neigh_w[ci] = 0.0;
        for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) {
            int64_t c = comm[g->indices[e]];
            if (!seen[c]) { seen[c] = 1; neigh_w[c] = 0.0; cands[nc++] = c; }
            neigh_w[c] += g->weights[e];
        }
        tot[ci] -= ki;
        int64_t best = ci;
        double best_gain = neigh_w[ci] - ((gamma * ki) * tot[ci]) / two_m;
        for (int64_t t = 1; t < nc; t++) {
            int64_t c = cands[t];
            double gn = neigh_w[c] - ((gamma * ki) * tot[c]) / two_m;
            if (gn > best_gain) { best = c; best_gain = gn; }
        }
        for (int64_t t = 0; t < nc; t++) seen[cands[t]] = 0;
        tot[best] += ki;
        if (best != ci) {
            comm[i] = best;
            moved_any = 1;
            for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) {
                int64_t j = g->indices[e];
                if (comm[j] != best && !inq[j]) {
                    inq[j] = 1;
                    int64_t tail = head + count; if (tail >= n) tail -= n;
                    queue[tail] = j; count++;
                }
            }
        }
    }
    free(k); free(tot); free(neigh_w); free(seen); free(inq); free(queue); free(cands);
    return moved_any;
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* aggregate g by comm -> out; node2new[] out */
static void aggregate(const graph *g, const int64_t *comm, graph *out, int64_t *node2new) {
    int64_t n = g->n;
    int64_t *new_id = (int64_t *)malloc(sizeof(int64_t) * n);
    for (int64_t i = 0; i < n; i++) new_id[i] = -1;
    int64_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        if (new_id[comm[i]] < 0) new_id[comm[i]] = nc++;
        node2new[i] = new_id[comm[i]];
    }
    /* members in ascending node order: counting sort by new id */
    int64_t *mstart = (int64_t *)calloc(nc + 1, sizeof(int64_t));
    int64_t *members = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) mstart[node2new[i] + 1]++;
    for (int64_t a = 0; a < nc; a++) mstart[a + 1] += mstart[a];
    int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (nc > 0 ? nc : 1));
    memcpy(fill, mstart, sizeof(int64_t) * nc);
    for (int64_t i = 0; i < n; i++) members[fill[node2new[i]]++] = i;

    int64_t cap = g->indptr[n] > 0 ? g->indptr[n] : 1;
    out->n = nc;
    out->indptr = (int64_t *)malloc(sizeof(int64_t) * (nc + 1));
    out->indices = (int64_t *)malloc(sizeof(int64_t) * cap);
    out->weights = (double *)malloc(sizeof(double) * cap);
    out->selfw = (double *)malloc(sizeof(double) * (nc > 0 ? nc : 1));
    double *acc = (double *)malloc(sizeof(double) * (nc > 0 ? nc : 1));
    char *seen = (char *)calloc(nc > 0 ? nc : 1, 1);
    int64_t *touched = (int64_t *)malloc(sizeof(int64_t) * (nc > 0 ? nc : 1));
    int64_t nnz = 0;
    out->indptr[0] = 0;
    for (int64_t a = 0; a < nc; a++) {
        double s = 0.0;
        int64_t nt = 0;
        for (int64_t p = mstart[a]; p < mstart[a + 1]; p++) {
            int64_t i = members[p];
            s += g->selfw[i];
            for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) {
                int64_t b = node2new[g->indices[e]];
                if (b == a) { s += g->weights[e]; }
                else {
                    if (!seen[b]) { seen[b] = 1; acc[b] = 0.0; touched[nt++] = b; }
                    acc[b] += g->weights[e];
                }
            }
        }
        out->selfw[a] = s;
        qsort(touched, nt, sizeof(int64_t), cmp_i64);
        for (int64_t t = 0; t < nt; t++) {
            int64_t b = touched[t];
            out->indices[nnz] = b; out->weights[nnz] = acc[b]; nnz++;
            seen[b] = 0;
        }
        out->indptr[a + 1] = nnz;
    }
    free(new_id); free(mstart); free(members); free(fill); free(acc); free(seen); free(touched);
}

/* First level by synchronous coloured rounds (unweighted graph); same statement as
 * oracle/louvain_ref.py:level0_parallel.  comm[] out (community id = a node id). */
#define N_COLOURS 8
#define MAX_ROUNDS 32
static uint64_t colour_of(uint64_t seed, int64_t i) {
    uint64_t z = seed + (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return z % N_COLOURS;
}

/* weights == NULL: unit weights (two_m = number of entries); else the fixed-point flavour: wq = rint(w * 2^32) in int64,
 * every sum exact, two_m = sum of wq (the argument is ignored). */
static void level0_parallel(const graph *g, const double *weights, double gamma, double two_m, uint64_t seed, int64_t *comm) {
    int64_t n = g->n, nnz = g->indptr[n];
    int64_t *wq = (int64_t *)malloc(sizeof(int64_t) * (nnz > 0 ? nnz : 1));
    int64_t *k = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *tot = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *size = (int64_t *)malloc(sizeof(int64_t) * n);
    int64_t *desired = (int64_t *)malloc(sizeof(int64_t) * n);
    unsigned char *col = (unsigned char *)malloc(n);
    int64_t *cnt = (int64_t *)calloc(n, sizeof(int64_t)); /* w(i, c) scratch, indexed by community */
    if (weights) {
        int64_t s = 0;
        for (int64_t e = 0; e < nnz; e++) { wq[e] = (int64_t)rint(weights[e] * 4294967296.0); s += wq[e]; }
        two_m = (double)s;
    } else {
        for (int64_t e = 0; e < nnz; e++) wq[e] = 1;
    }
    for (int64_t i = 0; i < n; i++) {
        int64_t s = 0;
        for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) s += wq[e];
        k[i] = s;
        tot[i] = k[i]; size[i] = 1; comm[i] = i; col[i] = (unsigned char)colour_of(seed, i);
    }
    for (int round = 0; round < MAX_ROUNDS && two_m > 0.0; round++) {
        int64_t moved = 0;
        for (int c = 0; c < N_COLOURS; c++) {
            /* decide from the frozen state */
            for (int64_t i = 0; i < n; i++) {
                desired[i] = -1;
                if (col[i] != c || g->indptr[i + 1] == g->indptr[i]) continue;
                int64_t ci = comm[i];
                double gk = gamma * (double)k[i];
                for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) cnt[comm[g->indices[e]]] += wq[e];
                double gain_stay = (double)cnt[ci] - (gk * (double)(tot[ci] - k[i])) / two_m;
                int64_t best = -1; double best_gain = 0.0;
                for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) {
                    int64_t cc = comm[g->indices[e]];
                    if (cc == ci) continue;
                    double gn = (double)cnt[cc] - (gk * (double)tot[cc]) / two_m;
                    if (best < 0 || gn > best_gain || (gn == best_gain && cc < best)) { best = cc; best_gain = gn; }
                }
                for (int64_t e = g->indptr[i]; e < g->indptr[i + 1]; e++) cnt[comm[g->indices[e]]] = 0;
                if (best >= 0 && best_gain > gain_stay && !(size[ci] == 1 && size[best] == 1 && best > ci))
                    desired[i] = best;
            }
            /* apply simultaneously */
            for (int64_t i = 0; i < n; i++) {
                if (desired[i] < 0) continue;
                int64_t ci = comm[i], b = desired[i];
                comm[i] = b; tot[ci] -= k[i]; tot[b] += k[i]; size[ci]--; size[b]++; moved++;
            }
        }
        if (moved <= (n >> 9)) break; /* at most n / 512 moves: the level is settled */
    }
    free(wq); free(k); free(tot); free(size); free(desired); free(col); free(cnt);
}

typedef struct { int64_t size; int64_t fa; } comm_rank;
static int cmp_rank(const void *a, const void *b) {
    const comm_rank *x = (const comm_rank *)a, *y = (const comm_rank *)b;
    if (x->size != y->size) return (x->size < y->size) - (x->size > y->size); /* decreasing size */
    return (x->fa > y->fa) - (x->fa < y->fa);
}

/* indptr/indices: symmetric CSR without self loops; weights may be NULL (all 1).
 * labels_out: n int64, 0 = largest community.  Returns number of communities. */
static int64_t louvain_impl(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                            double resolution, uint64_t seed, int64_t *labels_out, int parallel0);

int64_t louvain_ref(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                    double resolution, uint64_t seed, int64_t *labels_out) {
    return louvain_impl(n, indptr, indices, weights, resolution, seed, labels_out, 0);
}

/* kNN-pipeline flavour: parallel first level (unweighted graphs), sequential levels above. */
int64_t louvain_ref_parallel0(int64_t n, const int64_t *indptr, const int64_t *indices, double resolution,
                              uint64_t seed, int64_t *labels_out) {
    return louvain_impl(n, indptr, indices, NULL, resolution, seed, labels_out, 1);
}

/* the same with weights (PhenoGraph's Jaccard graph): fixed-point first level, double weights above it */
int64_t louvain_ref_parallel0_w(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                                double resolution, uint64_t seed, int64_t *labels_out) {
    return louvain_impl(n, indptr, indices, weights, resolution, seed, labels_out, 1);
}

static int64_t louvain_impl(int64_t n, const int64_t *indptr, const int64_t *indices, const double *weights,
                            double resolution, uint64_t seed, int64_t *labels_out, int parallel0) {
    graph g;
    int64_t nnz = indptr[n];
    g.n = n;
    g.indptr = (int64_t *)malloc(sizeof(int64_t) * (n + 1));
    g.indices = (int64_t *)malloc(sizeof(int64_t) * (nnz > 0 ? nnz : 1));
    g.weights = (double *)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    g.selfw = (double *)calloc(n > 0 ? n : 1, sizeof(double));
    memcpy(g.indptr, indptr, sizeof(int64_t) * (n + 1));
    memcpy(g.indices, indices, sizeof(int64_t) * nnz);
    double two_m = 0.0;
    for (int64_t e = 0; e < nnz; e++) { g.weights[e] = weights ? weights[e] : 1.0; two_m += g.weights[e]; }
    int64_t *membership = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) membership[i] = i;
    sm64 rng; rng.s = seed;
    if (two_m > 0.0 && parallel0) {
        int64_t *comm = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
        level0_parallel(&g, weights, resolution, two_m, seed, comm);
        graph ng;
        int64_t *node2new = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
        aggregate(&g, comm, &ng, node2new);
        for (int64_t i = 0; i < n; i++) membership[i] = node2new[membership[i]];
        free(comm); free(node2new);
        graph_free(&g);
        g = ng;
    }
    if (two_m > 0.0) {
        for (int level = 0; level < 64; level++) {
            int64_t *comm = (int64_t *)malloc(sizeof(int64_t) * (g.n > 0 ? g.n : 1));
            int moved = one_level(&g, resolution, two_m, &rng, comm);
            if (!moved) { free(comm); break; }
            graph ng;
            int64_t *node2new = (int64_t *)malloc(sizeof(int64_t) * (g.n > 0 ? g.n : 1));
            aggregate(&g, comm, &ng, node2new);
            for (int64_t i = 0; i < n; i++) membership[i] = node2new[membership[i]];
            free(comm); free(node2new);
            graph_free(&g);
            g = ng;
        }
    }
    graph_free(&g);
    /* relabel: first-appearance ids, then by decreasing size */
    int64_t *fa_of = (int64_t *)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) fa_of[i] = -1;
    int64_t nc = 0;
    for (int64_t i = 0; i < n; i++) {
        if (fa_of[membership[i]] < 0) fa_of[membership[i]] = nc++;
        membership[i] = fa_of[membership[i]];
    }
    comm_rank *rk = (comm_rank *)malloc(sizeof(comm_rank) * (nc > 0 ? nc : 1));
    for (int64_t c = 0; c < nc; c++) { rk[c].size = 0; rk[c].fa = c; }
    for (int64_t i = 0; i < n; i++) rk[membership[i]].size++;
    qsort(rk, nc, sizeof(comm_rank), cmp_rank);
    int64_t *newlab = (int64_t *)malloc(sizeof(int64_t) * (nc > 0 ? nc : 1));
    for (int64_t r = 0; r < nc; r++) newlab[rk[r].fa] = r;
    for (int64_t i = 0; i < n; i++) labels_out[i] = newlab[membership[i]];
    free(fa_of); free(rk); free(newlab); free(membership);
    return nc;
}
