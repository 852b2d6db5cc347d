"""GPU timing of the prologue's highly-variable-gene selection (doubletdetection.py:165-176) on the device against the
reference's scipy lines on the host, same matrix.    python scripts/hvg_bench.py [n_cells] [n_genes] [density] [n_top]"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from doubletdetection_b200 import _capi  # noqa: E402

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
n_genes = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
density = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
n_top = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
rs = np.random.default_rng(0)
t0 = time.perf_counter()
m = sp.random(n_cells, n_genes, density=density, random_state=1, format="csr", dtype=np.float32)
m.data = np.ceil(m.data * rs.integers(1, 40, size=n_genes)[m.indices]).astype(np.float32)  # gene-dependent scale
m.sort_indices()
print(f"matrix {m.shape}, nnz {m.nnz} ({time.perf_counter() - t0:.1f} s to generate)", flush=True)

t0 = time.perf_counter()
var = (np.array(m.power(2).mean(axis=0)) - (np.array(m.mean(axis=0))) ** 2)[0]
t_var = time.perf_counter() - t0
top = np.argsort(var)[-n_top:]
t0 = time.perf_counter()
sub = m.tocsc()[:, top].tocsr()
t_sub = time.perf_counter() - t0
print(f"host (scipy, the reference's lines): variances {1e3 * t_var:.1f} ms, column subset {1e3 * t_sub:.1f} ms", flush=True)

h = _capi.Handle(0)
h.set_kernel_timing(True)
for rep in range(3):
    t0 = time.perf_counter()
    h.upload_counts(m)
    t_up = time.perf_counter() - t0
    t0 = time.perf_counter()
    got = h.hvg_variances()
    t_dev_var = time.perf_counter() - t0
    ms_var = h.last_stage_ms("hvg")
    t0 = time.perf_counter()
    h.select_genes(np.argsort(got)[-n_top:])
    t_dev_sub = time.perf_counter() - t0
    print(f"device rep {rep}: upload {1e3 * t_up:.1f} ms; variances {1e3 * t_dev_var:.1f} ms wall ({ms_var:.2f} ms of kernels); "
          f"column subset {1e3 * t_dev_sub:.1f} ms wall; bit-identical variances: {bool((got.view(np.uint32) == var.view(np.uint32)).all())}",
          flush=True)
rep = h.kernel_timing_report()
for k, (ms, cnt) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
    if k.startswith(("hvg", "sel", "row_sums")):
        print(f"    {k:14s} {ms / cnt:8.3f} ms per launch x{cnt}")
got_sub = h.download_counts()
print("subset identical to scipy's:", bool((got_sub.indptr == sub.indptr).all() and (got_sub.indices == sub.indices).all()
                                            and (got_sub.data == sub.data).all()))
h.close()
