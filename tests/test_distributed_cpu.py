"""world_size-2 gloo test (CPU) of the multi-GPU host logic: iteration sharding and the collection of
per-iteration results (no data-path collective exists on this path)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from doubletdetection_b200 import _capi, iteration_shard
from doubletdetection_b200.classifier import _allgather_iterations, broadcast_token, merge_owned_iterations


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_iters, n_cells, n_synth, q, everywhere=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.default_rng(7)  # same "truth" on every rank
        truth = dict(
            scores=rs.random((n_iters, n_cells)),
            log_p=-rs.random((n_iters, n_cells)) * 40,
            communities=rs.integers(0, 30, (n_iters, n_cells)).astype(np.int32),
            synth_communities=rs.integers(0, 30, (n_iters, n_synth)).astype(np.int32),
        )
        truth["log_p"][1, 3] = -np.inf  # values that do not survive arithmetic reductions
        truth["log_p"][n_iters - 1, 5] = np.nan
        truth["scores"][n_iters - 1, 5] = np.nan
        it0, it1 = iteration_shard(n_iters, rank, world)
        mine = {k: np.zeros_like(v) for k, v in truth.items()}
        for k in truth:
            mine[k][it0:it1] = truth[k][it0:it1]
        mine["stage_ms"] = {"pca": 1.0 + rank, "knn": 5.0 - rank}
        merged = _allgather_iterations(dist, mine, n_iters, device=0, everywhere=everywhere)
        if everywhere or rank == 0:
            ok = all(np.array_equal(merged[k], truth[k], equal_nan=True) for k in truth)
        else:  # gather-on-rank-0 mode: the other ranks keep exactly what they computed
            ok = all(np.array_equal(merged[k], mine[k], equal_nan=True) for k in truth)
        ok = ok and merged["stage_ms"] == {"knn": 5.0, "pca": float(world)}
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_iteration_sharding_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    for n_iters, everywhere in ((5, True), (6, False)):  # uneven and even blocks; all-gather and gather-to-0
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_iters, 40, 10, q, everywhere)) for r in range(world)]
        for p in procs:
            p.start()
        results = [q.get(timeout=100) for _ in range(world)]
        for p in procs:
            p.join(timeout=30)
        assert sorted(results) == [(0, True), (1, True)]


def _token_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 creates the NCCL rendezvous token of the cell-sharding communicator (host-only call) and ships it
        # over the existing process group; every rank must end up with the same 128 bytes
        token = _capi.comm_unique_id() if rank == 0 else None
        got = broadcast_token(dist, token, device=0)
        q.put((rank, got))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_cell_sharding_token_broadcast_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_token_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert len(results[0]) == _capi.COMM_ID_BYTES and results[0] == results[1]
    assert any(results[0])  # a real token, not the zero padding


def test_cell_blocks_partition_rows():
    # the rule shared by the library (dd_set_block) and the host: contiguous, disjoint, covering
    for n, world in ((100000, 8), (25000, 8), (7, 3), (5, 8), (1000000, 8)):
        blocks = [_capi.block_of(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks[:-1], blocks[1:]))
        sizes = [e - b for b, e in blocks]
        assert max(sizes) - min(sizes) <= 1


def _merge_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_iters, n_cells, n_synth = 5, 30, 7
        rs = np.random.default_rng(11)
        truth = dict(scores=rs.random((n_iters, n_cells)), log_p=-rs.random((n_iters, n_cells)) * 50,
                     communities=rs.integers(-1, 20, (n_iters, n_cells)).astype(np.int32),
                     synth_communities=rs.integers(-1, 20, (n_iters, n_synth)).astype(np.int32))
        truth["log_p"][0, 1] = -np.inf
        truth["log_p"][3, 2] = np.nan
        truth["scores"][3, 2] = np.nan
        mine = {k: np.zeros_like(v) for k, v in truth.items()}
        for it in range(n_iters):  # the library fills iteration i on rank i % world only
            if it % world == rank:
                for k in truth:
                    mine[k][it] = truth[k][it]
        mine["stage_ms"] = {"pca": 2.0 + rank, "knn": 1.0}
        merged = merge_owned_iterations(dist, mine, device=0)
        ok = all(np.array_equal(merged[k], truth[k], equal_nan=True) for k in truth)
        ok = ok and merged["stage_ms"] == {"knn": 1.0, "pca": 1.0 + world}
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_cell_sharding_result_merge_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, True), (1, True)]


def _fit_worker(rank, world, port, mode, q):
    """BoostClassifier(distributed=mode).fit under gloo with the oracle-backed stand-in for the device handle
    (tests/test_classifier_host_logic.py): the iteration blocks of the two ranks, gathered, must equal one process."""
    import warnings

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from test_classifier_host_logic import OracleHandle

        from doubletdetection_b200 import BoostClassifier
        from oracle import datasets

        _capi.Handle = OracleHandle
        counts = datasets.poisson_counts(500, 100, seed=0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            clf = BoostClassifier(n_iters=3, clustering_algorithm="louvain", distributed=mode).fit(counts)
            # predict / doublet_score BEFORE anything touches the (n_iters, N) arrays: they must work from this rank's rows
            # plus three all-reduced N-vectors (the rows are still pending), and give the complete answer on EVERY rank
            assert clf._pending is not None
            labels = np.asarray(clf.predict(p_thresh=1e-16, voter_thresh=0.5), dtype=np.float64)
            score = np.ma.filled(np.ma.asarray(clf.doublet_score(), dtype=np.float64), np.nan)
            assert clf._pending is not None
        calls = list(OracleHandle.calls)
        q.put((rank, calls[-1]["iter_begin"], calls[-1]["iter_end"], clf.all_scores_, clf.all_log_p_values_, clf.communities_,
               np.asarray(clf.parents_, dtype=np.int64), labels, score))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("mode", [True, "allgather"])
def test_classifier_iteration_sharding_end_to_end_gloo_world2(mode):
    from conftest import load_golden

    g = load_golden("c1_louvain")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fit_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {r[0]: r[1:] for r in (q.get(timeout=200) for _ in range(world))}
    for p in procs:
        p.join(timeout=30)
    assert (results[0][0], results[0][1]) == (0, 1) and (results[1][0], results[1][1]) == (1, 3)  # blocks of 3 iterations
    for rank in (0, 1):  # labels and scores are complete on every rank in both modes (all-reduced per-cell sums)
        np.testing.assert_array_equal(results[rank][6], g["labels"])
        np.testing.assert_allclose(results[rank][7], g["doublet_score"], rtol=1e-12, equal_nan=True)
    for rank in ((0, 1) if mode == "allgather" else (0,)):  # complete (n_iters, N) arrays: everywhere / on rank 0
        _, _, scores, logp, comm, parents = results[rank][:6]
        np.testing.assert_array_equal(parents, g["parents"])  # every rank draws ALL iterations' parents (one stream)
        np.testing.assert_array_equal(scores, g["all_scores"])
        np.testing.assert_allclose(logp, g["all_log_p_values"], rtol=1e-12)
        np.testing.assert_array_equal(comm, g["communities"])
    if mode is True:  # rank 1 keeps its own rows only
        np.testing.assert_array_equal(results[1][2][1:], g["all_scores"][1:])
        assert not results[1][2][0].any()


def _cells_failure_worker(rank, world, port, q):
    """distributed="cells": an iteration's failure is seen by the rank that owns it only (advisor finding, round 1) -- the
    status must be made collective so that EVERY rank raises instead of one raising and the others hanging in the merge."""
    import warnings

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from doubletdetection_b200 import BoostClassifier
        from oracle import datasets

        class FailingHandle:
            def __init__(self, device=0):
                self._comm_ready = True  # pretend the NCCL communicator exists

            def upload_counts(self, csr):
                pass

            def share_counts(self, src):
                pass

            def counts_all_finite(self):
                return True

            def shard_cells(self, on):
                pass

            def fit_iterations(self, parents, omega, **kw):
                if rank == 1:
                    raise NotImplementedError("libdd_b200: pca: rank-deficient range (Cholesky breakdown)")
                n_iters, n_synth = parents.shape[:2]
                return dict(scores=np.zeros((n_iters, 600)), log_p=np.zeros((n_iters, 600)),
                            communities=np.zeros((n_iters, 600), np.int32),
                            synth_communities=np.zeros((n_iters, n_synth), np.int32), stage_ms={"wall": 1.0})

        _capi.Handle = FailingHandle
        counts = datasets.structured_counts(600, 200, seed=2)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                BoostClassifier(n_iters=2, clustering_algorithm="louvain", distributed="cells").fit(counts)
            q.put((rank, "no error"))
        except NotImplementedError as e:
            q.put((rank, "own: " + str(e)))
        except RuntimeError as e:
            q.put((rank, "other: " + str(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_cell_sharding_failure_is_collective_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_cells_failure_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
    assert results[1].startswith("own: ") and "Cholesky" in results[1]
    assert results[0].startswith("other: ") and "another rank" in results[0]
