"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the oracle port of the
reference's CPU path) runs and prints ONE JSON line with the keys the driver reads, the synthetic inputs are the
seeded ones of oracle/datasets.py, and the roofline arithmetic matches DESIGN.md's algorithmic bytes."""

import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import datasets  # noqa: E402


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "augmented-cells/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["metric"].startswith("augmented-cells/sec through BoostClassifier.fit")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload",
                          "c1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_inputs_are_the_oracle_datasets():
    for name in ("c1", "c2"):
        wl = bench.WORKLOADS[name]
        a = bench.make_counts(wl)
        b = (datasets.poisson_counts(wl["n_cells"], wl["n_genes"], seed=0) if wl["kind"] == "poisson"
             else datasets.structured_counts(wl["n_cells"], wl["n_genes"], seed=1234))
        import scipy.sparse as sp

        b = sp.csr_matrix(b)
        b.sort_indices()
        assert a.shape == b.shape and a.nnz == b.nnz
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)


def test_roofline_arithmetic():
    peaks = {"hbm_gbs": 6500.0, "bf16_tflops_sustained": 1400.0}
    n, m, g, nnz, nnz_par = 100000, 25000, 3000, 32.4e6, 16.2e6
    a = n + m
    report = {"dense_rows": (10 * 0.75, 10), "tc_gemm_dq": (3.0, 10), "knn_tc": (48.0, 10), "unknown": (1.0, 1)}
    r = bench.kernel_rooflines(report, n, m, g, nnz, nnz_par, peaks)
    assert set(r) == {"dense_rows", "tc_gemm_dq", "knn_tc"}
    dense_bytes = (nnz + nnz_par) * 8 + a * g * 4  # CSR entries of originals + both parents, dense matrix once
    assert r["dense_rows"]["algorithmic_bytes"] == dense_bytes and r["dense_rows"]["bound"] == "hbm"
    assert np.isclose(r["dense_rows"]["achieved"], dense_bytes / 0.75e-3 / 1e9)
    assert np.isclose(r["dense_rows"]["frac"], r["dense_rows"]["achieved"] / 6500.0)
    assert r["tc_gemm_dq"]["algorithmic_bytes"] == (a * g + g * 40 + a * 40) * 4
    assert r["knn_tc"]["bound"] == "tensor" and r["knn_tc"]["algorithmic_flops"] == 2.0 * a * a * 32
    assert np.isclose(r["knn_tc"]["achieved"], 2.0 * a * a * 32 / 4.8e-3 / 1e12)


def test_product_arm_line_with_a_stand_in_handle(monkeypatch, capsys):
    """bench.py's own arm end to end on the CPU: CUDA and the device handle are replaced by stand-ins (the numbers mean
    nothing), so that a typo in the measurement code cannot surface for the first time on the GPU box.  Checks every key
    the contract names."""
    import torch

    from doubletdetection_b200 import _capi

    class Handle:
        def __init__(self, device=0):
            self.launches = 0

        def upload_counts(self, csr):
            self.n_cells, self.n_genes = csr.shape

        def share_counts(self, src):
            self.n_cells, self.n_genes = src.n_cells, src.n_genes

        def counts_all_finite(self):
            return True

        # stage-wise calls of the "rooflines_alone" leg
        def create_doublets(self, parents):
            pass

        def median_lib_size(self):
            return 1.0

        def normalise_log(self, median, pseudocount):
            pass

        def pca(self, n_comp, omega, n_power_iter):
            pass

        def knn(self, k):
            pass

        def fit_iterations(self, parents, omega, **kw):
            n_iters, n_synth = parents.shape[:2]
            self.launches += 100
            rs = np.random.default_rng(0)
            return dict(scores=rs.random((n_iters, self.n_cells)), log_p=-30 * rs.random((n_iters, self.n_cells)),
                        communities=np.zeros((n_iters, self.n_cells), np.int32),
                        synth_communities=np.zeros((n_iters, n_synth), np.int32),
                        stage_ms=dict(host_cluster_score=1.0, normalise=1.0, scale=0.0, pca=1.0, knn=1.0, cluster_gpu_d2h=1.0,
                                      device_total=5.0, wall=6.0))

        def set_kernel_timing(self, on):
            pass

        def kernel_launches(self):
            return self.launches

        def last_stage_ms(self, stage):
            return 19.0 if stage == "lv_rounds" else -1.0

        def kernel_timing_report(self):
            return {"knn_tc": (10.0, 10), "tc_gemm_dq": (3.0, 10), "dense_rows": (7.0, 10), "lv_rounds_graph": (30.0, 10)}

        def close(self):
            pass

    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(_capi, "Handle", Handle)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "c1", "--steps", "2", "--warmup", "3"])
    for var in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(var, raising=False)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["unit"] == "augmented-cells/s" and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"])
    assert d["roofline"]["bound"] in ("hbm", "tensor") and d["roofline_kernel"] == "lv_rounds_graph"  # most summed time of ALL kernels
    assert np.isclose(d["roofline"]["frac"], d["roofline"]["achieved"] / d["roofline"]["peak"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] == "port"
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    assert d["rooflines_alone"] is not None
    assert d["config"]["pipelines_per_gpu"] == 2  # two pipelined loops per GPU (DD_PIPELINES), 100 launches per loop and step
    assert d["gpu_launches"] == 400 and set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
