#!/bin/bash
# compute-sanitizer over the round-2 kernels at small sizes: HVG (hvg.cu), cluster-ordered kNN (knn_prune.cu + listed kernel),
# weighted Louvain level (louvain_gpu_w.cu), the fused Jacobi, the fit loop with two pipelines.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_sanitize_r2.sh'
mkdir -p gpurun_out
log=gpurun_out/r2p_sanitize.log
: > $log
cat > gpurun_out/_san_r2.py <<'PY'
import sys, warnings, numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
from doubletdetection_b200 import BoostClassifier, _capi
what = sys.argv[1]
h = _capi.Handle(0)
rs = np.random.default_rng(0)
if what == "hvg":
    m = sp.random(3000, 2500, density=0.05, random_state=1, format="csr", dtype=np.float32)
    m.data = np.ceil(m.data * 9).astype(np.float32)
    h.upload_counts(m)
    v = h.hvg_variances()
    h.select_genes(np.argsort(v)[-700:])
    print("hvg ok", h.download_counts().shape)
elif what == "knn":
    for n, k in ((3000, 10), (6000, 31)):
        emb = (rs.normal(size=(n, 30)) + rs.integers(0, 5, size=(n, 1)) * 3.0).astype(np.float32)
        h.upload_embedding(emb)
        h.set_knn_mode(1); a, _ = h.knn(k)
        h.set_knn_mode(2); b, _ = h.knn(k); c, _ = h.knn(k)
        print("knn", n, k, "equal", bool((a == b).all() and (a == c).all()))
elif what == "fit":
    counts = rs.poisson(1.0, (900, 200))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for algo in ("louvain", "phenograph"):
            clf = BoostClassifier(n_iters=4, clustering_algorithm=algo, n_jobs=2).fit(counts)
            print(algo, "ok", float(np.nanmean(clf.doublet_score())))
        clf = BoostClassifier(n_iters=2, clustering_algorithm="louvain", pseudocount=1).fit(counts)
        print("pc1 ok", float(np.nanmean(clf.doublet_score())))
h.close()
PY
for what in hvg knn fit; do
    for tool in memcheck racecheck; do
        echo "=== compute-sanitizer --tool $tool $what" | tee -a $log
        timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python gpurun_out/_san_r2.py $what 2>&1 | tail -15 | tee -a $log
    done
done
grep -c "ERROR SUMMARY: 0 errors" $log; grep "ERROR SUMMARY" $log
