"""Seeded synthetic count matrices (TEST INFRASTRUCTURE shared by tests/ and bench.py's CPU legs;
the product's bench generates its inputs with the same formulas, re-stated in bench.py).

SURVEY.md section 8(d): the reference defines no inputs beyond ``np.random.poisson(size=(500, 100))``
(tests/test_package.py:8), so c1 is that with a fixed seed and c2-c5 are cluster-structured counts.
"""

import numpy as np
import scipy.sparse as sp_sparse


def poisson_counts(n_cells=500, n_genes=100, seed=0, lam=1.0):
    """c1: ``np.random.default_rng(seed).poisson(lam, (n_cells, n_genes))`` (int64, dense)."""
    return np.random.default_rng(seed).poisson(lam, (n_cells, n_genes))


def structured_counts(n_cells, n_genes, seed=1234, n_types=8, chunk=20000):
    """c2-c5: K cell types with log-normal gene profiles and per-cell depth; float32 CSR,
    density ~11 %, ~324 nnz/row at 3000 genes."""
    rs = np.random.default_rng(seed)
    base = rs.lognormal(-3.0, 1.2, n_genes)
    prof = base * np.exp(rs.normal(0, 0.8, (n_types, n_genes)))
    types = rs.integers(0, n_types, n_cells)
    depth = rs.lognormal(0, 0.3, n_cells)
    blocks = []
    for s in range(0, n_cells, chunk):
        e = min(s + chunk, n_cells)
        lam = prof[types[s:e]] * depth[s:e, None]
        blocks.append(sp_sparse.csr_matrix(rs.poisson(lam).astype(np.float32)))
    X = sp_sparse.vstack(blocks).tocsr()
    X.sort_indices()
    return X


def structured_counts_with_doublets(n_cells, n_genes, seed=1234, doublet_frac=0.08, n_types=8):
    """The structured counts with real doublets planted among the cells: the last ``int(doublet_frac * n_cells)`` rows are
    replaced by the sum of two randomly chosen earlier cells (so most of them are heterotypic).  ``structured_counts`` has
    no doublets and the classifier rightly calls none on it, which makes label parity trivial; this variant is what the
    end-to-end label comparisons run on.  Returns (float32 CSR, bool[n_cells] is_doublet)."""
    X = structured_counts(n_cells, n_genes, seed=seed, n_types=n_types)
    n_dbl = int(doublet_frac * n_cells)
    rs = np.random.default_rng([seed, 99])
    n_single = n_cells - n_dbl
    pa = rs.integers(0, n_single, n_dbl)
    pb = rs.integers(0, n_single, n_dbl)
    Y = sp_sparse.vstack([X[:n_single], X[pa] + X[pb]]).tocsr()
    Y.sort_indices()
    truth = np.zeros(n_cells, dtype=bool)
    truth[n_single:] = True
    return Y.astype(np.float32), truth
