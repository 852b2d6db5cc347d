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
