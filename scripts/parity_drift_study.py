"""CPU study: how far do communities / final labels move when the PCA embedding moves by the amount that
separates the GPU path (5e-6 from the float64 truth) from sklearn's float32 run (1e-4 from it)?

Runs the oracle twice on the same seeded input -- once with sklearn's float32 PCA (the oracle proper), once with the
float64 restatement cast to float32 (a stand-in for the GPU embedding) -- and prints per-iteration agreement.  The
thresholds of tests/test_gpu_e2e_parity.py come from here and from the same comparison made on the GPU.

    python scripts/parity_drift_study.py c2 25
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sklearn.metrics import adjusted_rand_score  # noqa: E402

from oracle import datasets, louvain_c, pca_f64, reference_path, upstream  # noqa: E402

SHAPES = {"c2": (10000, 3000), "c2s": (4000, 1000), "c3": (100000, 3000)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    n_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    n, g = SHAPES[name]
    if len(sys.argv) > 3 and sys.argv[3] == "doublets":
        counts, truth = datasets.structured_counts_with_doublets(n, g, seed=1234)
    else:
        counts, truth = datasets.structured_counts(n, g, seed=1234), None
    kw = dict(n_iters=n_iters, random_state=0, louvain_fn=louvain_c.louvain, keep_stages=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = reference_path.OracleClassifier(**kw).fit(counts)
        real = upstream.tl_pca

        def f64_pca(X, n_comps, random_state=0, svd_solver="auto"):
            emb, _, _ = pca_f64.randomized_pca_f64(X, n_comps, random_state=random_state)
            return np.ascontiguousarray(emb, dtype=np.float32), None

        upstream.tl_pca = f64_pca
        try:
            b = reference_path.OracleClassifier(**kw).fit(counts)
        finally:
            upstream.tl_pca = real
    same = (a.communities_ == b.communities_).all(axis=1)
    print(f"[{name}] iterations with identical communities: {int(same.sum())}/{n_iters}")
    for i in range(n_iters):
        fa = np.concatenate([a.communities_[i], a.synth_communities_[i]])
        fb = np.concatenate([b.communities_[i], b.synth_communities_[i]])
        ari = adjusted_rand_score(fa, fb)
        call_a, call_b = a.all_log_p_values_[i] <= np.log(1e-7), b.all_log_p_values_[i] <= np.log(1e-7)
        print(f"  iter {i:2d}: n_comm {int(fa.max()) + 1:3d} / {int(fb.max()) + 1:3d}  ARI {ari:.4f}  "
              f"per-iteration calls equal {np.mean(call_a == call_b):.5f}  called {int(call_a.sum())} / {int(call_b.sum())}")
    for p, v in ((1e-7, 0.9), (1e-16, 0.5)):
        la, lb = a.predict(p, v), b.predict(p, v)
        print(f"  predict(p_thresh={p}, voter_thresh={v}): labels equal {np.mean(la == lb):.5f}  "
              f"doublets {int(np.nansum(la))} / {int(np.nansum(lb))}"
              + ("" if truth is None else f"  recall of planted doublets {np.mean(la[truth] == 1):.3f} / {np.mean(lb[truth] == 1):.3f}"))
    sa, sb = np.ma.filled(a.doublet_score(), np.nan), np.ma.filled(b.doublet_score(), np.nan)
    print(f"  doublet_score: max rel diff {np.nanmax(np.abs(sa - sb) / np.maximum(np.abs(sa), 1e-300)):.3e}  "
          f"corr {np.corrcoef(sa, sb)[0, 1]:.6f}")


if __name__ == "__main__":
    main()
