"""Dry run of the classifier-level GPU tests on the CPU: the same test functions, with ``_capi.Handle`` replaced by the
oracle-backed stand-in of tests/test_classifier_host_logic.py.  It proves nothing about the kernels -- it keeps the GPU
tests' own logic (inputs, fixtures, tolerances, oracle calls) from rotting between GPU sessions, which matters when a
session ends without hardware access (a wrong parent draw in a new GPU test was caught exactly this way)."""

import sys

import pytest

from test_classifier_host_logic import OracleHandle


@pytest.fixture()
def oracle_backed(monkeypatch):
    from doubletdetection_b200 import _capi

    monkeypatch.setattr(_capi, "Handle", OracleHandle)
    # one pipelined loop: the stand-in runs the oracle's sklearn / BLAS calls, and two Python threads inside OpenBLAS at the
    # same time change the rounding of its LATER calls in this process (the float32 goldens of test_oracle_golden.py are
    # bit-exact only for an undisturbed BLAS); the product's two loops run CUDA work, not BLAS
    monkeypatch.setenv("DD_PIPELINES", "1")
    OracleHandle.calls = []
    return _capi


def test_end_to_end_and_determinism_tests(oracle_backed):
    import test_gpu_parity as T

    for name in ("c1_louvain", "c1_louvain_scaled", "hvg_replace", "single_iter"):
        T.test_classifier_end_to_end_vs_golden(name)
    T.test_classifier_is_deterministic_and_stream_continues()
    T.test_default_constructor_fits()


def test_phenograph_tests(oracle_backed):
    import test_gpu_parity as T
    import test_gpu_zz_leiden as Z

    T.test_classifier_phenograph_vs_oracle(dict(clustering_kwargs={"prune": False}))
    for name in ("c1_phenograph_scaled", "structured_900x200_phenograph"):
        Z.test_classifier_phenograph_vs_reference_golden(name)


def test_leiden_tests(oracle_backed):
    import test_gpu_zz_leiden as Z

    Z.test_classifier_leiden_vs_oracle()
    Z.test_reference_package_test_mirrored()


class StageHandle(OracleHandle):
    """The stage-by-stage entry points ``__graft_entry__.smoke()`` uses, evaluated by the oracle."""

    def create_doublets(self, parents):
        import numpy as np

        self._parents = np.asarray(parents)
        from oracle import reference_path

        self._synth = reference_path.create_doublets(self.raw, self._parents)

    def download_synthetics(self):
        return self._synth

    def median_lib_size(self):
        import numpy as np

        lib = np.asarray(self.raw.sum(axis=1)).ravel()
        return float(np.median(np.concatenate([lib, np.asarray(self._synth.sum(axis=1)).ravel()])))

    def normalise_log(self, median, pseudocount):
        import numpy as np
        from sklearn.utils.sparsefuncs_fast import inplace_csr_row_normalize_l1

        from oracle import reference_path

        lib = np.asarray(self.raw.sum(axis=1)).ravel()
        normed = self.raw.copy()
        inplace_csr_row_normalize_l1(normed)
        self._dense, _, _ = reference_path.normalise(self._synth, lib, normed, pseudocount)

    def download_dense(self):
        return self._dense

    def pca(self, n_comp, omega, n_power_iter):
        # the float64 restatement: sklearn's own float32 run is 1.4e-4 from it on this input, above the 1e-4 the device
        # path is held to, so it cannot stand in for the device here
        import numpy as np

        from oracle import pca_f64

        emb, _, _ = pca_f64.randomized_pca_f64(self._dense, n_comp, random_state=0)
        return np.asarray(emb, dtype=np.float32), None

    def kernel_launches(self):
        return 0


def test_smoke_entry_point_logic(monkeypatch, capsys):
    """__graft_entry__.smoke() is what the driver runs first on the GPU box: its own logic (inputs, oracle calls, assertions)
    is exercised here with the oracle standing in for the device."""
    from conftest import ROOT

    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    from doubletdetection_b200 import _capi

    monkeypatch.setattr(_capi, "Handle", StageHandle)
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
