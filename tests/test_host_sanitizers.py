"""The host half of libdd_b200 (umap weights, Leiden, Louvain incl. the fixed-point first level, PhenoGraph finish, scoring)
under AddressSanitizer + UndefinedBehaviorSanitizer: the three .cpp files are compiled as plain C++ next to a small driver
and run on seeded kNN lists (with duplicate points and degenerate sizes).  These functions run on the fit loop's worker threads
in production, so memory errors there would corrupt results silently."""

import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import upstream


@pytest.mark.timeout(600)
def test_host_clustering_code_is_clean_under_asan_ubsan(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    rs = np.random.default_rng(3)
    n, k = 4000, 10
    pts = (rs.normal(size=(n, 8)) + rs.integers(0, 6, size=(n, 1)) * 2.4).astype(np.float32)
    pts[50:56] = pts[50]  # coincident cells: rho = 0 rows in smooth_knn_dist
    idx, dist = upstream.knn_brute(pts, k)
    idx.astype(np.int32).tofile(tmp_path / "idx.bin")
    dist.astype(np.float32).tofile(tmp_path / "dist.bin")
    upstream.knn_brute(pts, 31)[0].astype(np.int32).tofile(tmp_path / "idx31.bin")
    csrc = os.path.join(ROOT, "doubletdetection_b200", "csrc")
    exe = str(tmp_path / "driver")
    cmd = [gxx, "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
           "-std=c++17", "-ffp-contract=off", "-I/usr/local/cuda/include", "-x", "c++",
           os.path.join(ROOT, "tests", "host_sanitize_driver.cpp"), os.path.join(csrc, "leiden.cpp"),
           os.path.join(csrc, "louvain.cpp"), os.path.join(csrc, "score.cpp"), "-o", exe]
    build = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    if build.returncode != 0 and ("asan" in build.stderr or "sanitize" in build.stderr):
        pytest.skip("sanitizer runtime not available: " + build.stderr[-300:])
    assert build.returncode == 0, build.stderr[-3000:]
    run = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0"))
    assert run.returncode == 0, (run.stdout + run.stderr)[-4000:]
    assert "ERROR" not in run.stderr and "runtime error" not in run.stderr, run.stderr[-4000:]
    lines = run.stdout.splitlines()
    assert len(lines) == 11 and all(" rc=0" in ln for ln in lines), run.stdout
