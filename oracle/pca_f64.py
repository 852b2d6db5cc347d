"""Randomized-SVD PCA restated in float64 (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows sklearn 1.9.0 (what ``sc.tl.pca`` -> doubletdetection.py:309-314 bottoms out in):
``PCA._fit_truncated`` (sklearn/decomposition/_pca.py:726-758), ``_randomized_svd``
(sklearn/utils/extmath.py:560-633), ``_randomized_range_finder`` (:313-385), ``svd_flip``
(:970-981) -- same Omega (``RandomState(seed).normal(size=(G, n_comp+10))``), same iteration
count, same sign convention, but every product in float64.  This is the "truth" for the 1e-4
embedding tolerance: sklearn's own float32 run differs from it by ~1e-4, so the
GPU path is compared against this and the sklearn-f32 distance is reported beside it.
"""

import numpy as np
from scipy import linalg


def auto_solver(n_samples, n_features, n_components):
    """sklearn PCA ``svd_solver="auto"`` policy (_pca.py:524-536)."""
    if n_features <= 1000 and n_samples >= 10 * n_features:
        return "covariance_eigh"
    if max(n_samples, n_features) <= 500:
        return "full"
    if 1 <= n_components < 0.8 * min(n_samples, n_features):
        return "randomized"
    return "full"


def auto_n_iter(n_samples, n_features, n_components):
    """extmath.py:584-587."""
    return 7 if n_components < 0.1 * min(n_samples, n_features) else 4


def omega(n_rows, n_components, random_state, n_oversamples=10):
    """The Gaussian test matrix sklearn draws (extmath.py:323-333), float64."""
    rs = np.random.RandomState(random_state)
    return rs.normal(size=(n_rows, n_components + n_oversamples))


def randomized_pca_f64(X, n_components, random_state=0, normalizer="LU", n_iter=None):
    """Returns (embedding A x C float64, singular values, Vt)."""
    X = np.asarray(X, dtype=np.float64)
    n_samples, n_features = X.shape
    mean = X.mean(axis=0)
    Xc = X - mean
    M = Xc
    transpose = n_samples < n_features
    if transpose:
        M = M.T
    if n_iter is None:
        n_iter = auto_n_iter(n_samples, n_features, n_components)
    Q = omega(M.shape[1], n_components, random_state)
    if normalizer == "LU":
        norm = lambda Y: linalg.lu(Y, permute_l=True, check_finite=False)[0]
    else:
        norm = lambda Y: linalg.qr(Y, mode="economic", check_finite=False)[0]
    if n_iter <= 2:
        norm = lambda Y: Y
    for _ in range(n_iter):
        Q = norm(M @ Q)
        Q = norm(M.T @ Q)
    Q, _ = linalg.qr(M @ Q, mode="economic", check_finite=False)
    B = Q.T @ M
    Uhat, s, Vt = linalg.svd(B, full_matrices=False, lapack_driver="gesdd")
    U = Q @ Uhat
    if transpose:
        U, Vt = Vt.T, U.T  # back to input convention: U (samples), Vt (components)
    # svd_flip(u_based_decision=False): sign from the max-|.| entry of each Vt row
    idx = np.argmax(np.abs(Vt), axis=1)
    signs = np.sign(Vt[np.arange(Vt.shape[0]), idx])
    U = U * signs[np.newaxis, :]
    Vt = Vt * signs[:, np.newaxis]
    emb = U[:, :n_components] * s[:n_components]
    return emb, s[:n_components], Vt[:n_components]


def exact_pca_f64(X, n_components):
    """The float64 truth for sklearn's EXACT branches ("covariance_eigh" / "full", _pca.py:560-640): top principal
    components of the centred matrix by a float64 SVD, signs by ``svd_flip(u_based_decision=False)``.
    Returns (embedding A x C float64, singular values)."""
    X = np.asarray(X, dtype=np.float64)
    Xc = X - X.mean(axis=0)
    U, s, Vt = linalg.svd(Xc, full_matrices=False)
    idx = np.argmax(np.abs(Vt), axis=1)
    signs = np.sign(Vt[np.arange(Vt.shape[0]), idx])
    U = U * signs[np.newaxis, :]
    return U[:, :n_components] * s[:n_components], s[:n_components]
