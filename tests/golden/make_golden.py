"""Generate the committed golden vectors by running the REFERENCE'S OWN CODE
(/root/reference/doubletdetection/doubletdetection.py, unmodified) on seeded inputs.

Runs only in the build container (the reference is not on the GPU box).  The packages the
reference imports but the image lacks (scanpy, anndata, phenograph) are replaced by the stub
modules of ``oracle.refshim``; everything that is the reference's own arithmetic -- prologue/HVG,
``rng.choice`` parents, CSR row-pair add, normalise/log, scoring, predict, doublet_score -- is the
real thing.  Usage:  ``python tests/golden/make_golden.py``  (from the repo root).
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import datasets, refshim  # noqa: E402
from oracle import louvain_c  # noqa: E402


def run_case(name, counts, predict_kw, capture_iter0=True, **clf_kw):
    record = {}
    mod = refshim.load_reference(louvain_fn=louvain_c.louvain, record=record)
    clf = mod.BoostClassifier(**clf_kw)
    synths = []
    orig_one_fit = clf._one_fit

    def one_fit_capture():
        out = orig_one_fit()
        synths.append(clf._raw_synthetics.copy())
        return out

    clf._one_fit = one_fit_capture
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        clf.fit(counts)
        labels = np.asarray(clf.predict(**predict_kw), dtype=np.float64)
        score = clf.doublet_score()
    out = dict(
        parents=np.asarray(clf.parents_, dtype=np.int64),
        communities=clf.communities_,
        synth_communities=clf.synth_communities_,
        all_scores=clf.all_scores_,
        all_log_p_values=clf.all_log_p_values_,
        labels=labels,
        doublet_score=np.ma.filled(np.ma.asarray(score, dtype=np.float64), np.nan),
        doublet_score_mask=np.ma.getmaskarray(np.ma.asarray(score)),
    )
    if hasattr(clf, "voting_average_"):
        out["voting_average"] = clf.voting_average_
    if hasattr(clf, "suggested_score_cutoff_"):
        out["suggested_score_cutoff"] = np.float64(clf.suggested_score_cutoff_)
    if hasattr(clf, "top_var_genes_"):
        out["top_var_genes"] = np.asarray(clf.top_var_genes_)
    if capture_iter0:
        s0 = synths[0]
        out["synth0_indptr"] = s0.indptr
        out["synth0_indices"] = s0.indices
        out["synth0_data"] = s0.data
        pca_in = record["pca_input"][0]
        out["pca_input0"] = np.asarray(pca_in.toarray() if hasattr(pca_in, "toarray") else pca_in, dtype=np.float32)
        out["n_counts0"] = np.asarray(record["n_counts"][0])
        out["X_pca0"] = record["X_pca"][0]
        out["knn_indices0"] = record["knn_indices"][0].astype(np.int32)
        out["knn_distances0"] = record["knn_distances"][0]
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


def main():
    if not refshim.reference_available():
        raise SystemExit("reference not present; goldens can only be generated in the build container")
    c1 = datasets.poisson_counts(500, 100, seed=0)
    ref_test_predict = dict(p_thresh=1e-16, voter_thresh=0.5)  # tests/test_package.py:14
    # BASELINE.json configs[0]: 500 x 100 Poisson, n_iters=3, louvain
    run_case("c1_louvain", c1, ref_test_predict, n_iters=3, clustering_algorithm="louvain")
    # the reference test's own flags (tests/test_package.py:11-13)
    run_case("c1_louvain_scaled", c1, ref_test_predict, n_iters=2, clustering_algorithm="louvain",
             standard_scaling=True)
    # HVG selection + replace=True + non-default seed
    hv = datasets.poisson_counts(450, 400, seed=7, lam=0.7) * (np.arange(400) % 5 + 1)[None, :]
    run_case("hvg_replace", hv, dict(), n_iters=2, clustering_algorithm="louvain", n_top_var_genes=150,
             replace=True, boost_rate=0.6, random_state=11)
    # n_iters == 1 branch of predict (bool labels, suggested cutoff)
    run_case("single_iter", c1, dict(), capture_iter0=False, n_iters=1, clustering_algorithm="louvain")
    # cluster-structured counts, small (sparse CSR input)
    st = datasets.structured_counts(1500, 300, seed=1234)
    run_case("structured_1500x300", st, dict(p_thresh=1e-3, voter_thresh=0.5), n_iters=3,
             clustering_algorithm="louvain")
    # the other two branches of the reference's own test (tests/test_package.py:17-28), through the reference's real
    # control flow (:317-343) over the restated phenograph / umap + leiden calls
    run_case("c1_phenograph_scaled", c1, ref_test_predict, capture_iter0=False, n_iters=2,
             clustering_algorithm="phenograph", standard_scaling=True)
    run_case("c1_leiden_scaled", c1, ref_test_predict, capture_iter0=False, n_iters=2, clustering_algorithm="leiden",
             standard_scaling=True, random_state=123)
    st2 = datasets.structured_counts(900, 200, seed=77)
    run_case("structured_900x200_phenograph", st2, dict(p_thresh=1e-3, voter_thresh=0.5), capture_iter0=False, n_iters=2,
             clustering_algorithm="phenograph", clustering_kwargs={"prune": False})
    # the sparse branch: pseudocount == 1 keeps the matrix sparse (np.log1p, :296-297) and sc.tl.pca gets
    # svd_solver="arpack" (:308) -- through the reference's real control flow; PCA = the installed sklearn's arpack solver
    st3 = datasets.structured_counts(1200, 260, seed=5)
    run_case("structured_1200x260_pc1", st3, dict(p_thresh=1e-3, voter_thresh=0.5), n_iters=2, clustering_algorithm="louvain",
             pseudocount=1)
    # ... and with standard_scaling: sc.pp.scale densifies the sparse matrix, :308 then picks svd_solver="auto"
    run_case("structured_1200x260_pc1_scaled", st3, dict(p_thresh=1e-3, voter_thresh=0.5), n_iters=2,
             clustering_algorithm="louvain", pseudocount=1, standard_scaling=True)


if __name__ == "__main__":
    main()
