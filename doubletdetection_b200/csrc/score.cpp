// score.cpp -- per-community scoring of _one_fit (doubletdetection.py:344-383): synthetic fraction and
// hypergeometric log-survival p-value of every community that holds at least one original cell,
// broadcast back to the cells.  Host code: O(A) work per iteration on the labels the clustering
// produced.  dd_hypergeom_logsf follows scipy.stats.hypergeom.logsf (scipy/stats/_discrete_distns.py
// :669-718 and the rv_discrete.logsf wrapper): the shorter tail is summed with logsumexp over the
// log-pmf, the other one comes from log1p(-exp(.)).
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <limits>
#include <unordered_map>
#include <vector>

#include "dd_internal.h"

namespace {

inline double log_choose(double n, double k) { return lgamma(n + 1.0) - lgamma(k + 1.0) - lgamma(n - k + 1.0); }

// log pmf of hypergeom(tot, good, draw) at k (k inside the support)
inline double log_pmf(int64_t k, int64_t tot, int64_t good, int64_t draw) {
    const int64_t bad = tot - good;
    return log_choose((double)good, (double)k) + log_choose((double)bad, (double)(draw - k)) -
           log_choose((double)tot, (double)draw);
}

// logsumexp of the log-pmf over k in [lo, hi] intersected with the support
double log_sum_pmf(int64_t lo, int64_t hi, int64_t tot, int64_t good, int64_t draw) {
    const int64_t bad = tot - good;
    lo = std::max<int64_t>(lo, std::max<int64_t>(0, draw - bad));
    hi = std::min<int64_t>(hi, std::min(good, draw));
    if (lo > hi) return -std::numeric_limits<double>::infinity();
    double mx = -std::numeric_limits<double>::infinity();
    std::vector<double> t((size_t)(hi - lo + 1));
    for (int64_t k = lo; k <= hi; k++) {
        t[k - lo] = log_pmf(k, tot, good, draw);
        mx = std::max(mx, t[k - lo]);
    }
    double s = 0.0;
    for (double v : t) s += exp(v - mx);
    return mx + log(s);
}

}  // namespace

extern "C" double dd_hypergeom_logsf(int64_t k, int64_t M, int64_t n, int64_t N) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const double ninf = -std::numeric_limits<double>::infinity();
    if (!(M > 0 && n >= 0 && N >= 0 && n <= M && N <= M)) return nan;  // _argcheck
    const int64_t a = std::max<int64_t>(N - (M - n), 0), b = std::min(n, N);
    if (k < a) return 0.0;
    if (k >= b) return ninf;
    const double lhs = ((double)k + 0.5) * ((double)M + 0.5), rhs = ((double)n - 0.5) * ((double)N - 0.5);
    if (lhs < rhs) {
        // fewer terms below k: log(1 - cdf)
        const double logcdf = log_sum_pmf(0, k, M, n, N);
        return log1p(-exp(logcdf));
    }
    return log_sum_pmf(k + 1, N, M, n, N);
}

extern "C" int dd_score(int64_t n_cells, int64_t n_synth, const int32_t *labels, double *scores_out,
                        double *log_p_out) {
    if (n_cells < 0 || n_synth < 0 || (n_cells + n_synth > 0 && !labels) || (n_cells > 0 && (!scores_out || !log_p_out))) {
        dd_set_global_error("dd_score: bad arguments");
        return DD_ERR_ARG;
    }
    struct Cnt {
        int64_t orig = 0, synth = 0;
        double score = 0.0, logp = 0.0;
    };
    std::unordered_map<int32_t, Cnt> comm;
    const int64_t n_aug = n_cells + n_synth;
    int32_t min_id = n_aug > 0 ? labels[0] : 0;
    for (int64_t i = 0; i < n_aug; i++) {
        Cnt &c = comm[labels[i]];
        if (i < n_cells) c.orig++; else c.synth++;
        if (labels[i] < min_id) min_id = labels[i];
    }
    for (auto &kv : comm) {
        Cnt &c = kv.second;
        if (c.orig == 0) continue;  // only communities with original cells are scored (:361)
        c.score = (double)c.synth / (double)(c.synth + c.orig);
        c.logp = dd_hypergeom_logsf(c.synth, n_aug, n_synth, c.synth + c.orig);
    }
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int64_t i = 0; i < n_cells; i++) {
        const Cnt &c = comm[labels[i]];
        scores_out[i] = c.score;
        log_p_out[i] = c.logp;
        if (min_id < 0 && labels[i] == -1) {  // :379-381
            scores_out[i] = nan;
            log_p_out[i] = nan;
        }
    }
    return DD_OK;
}
