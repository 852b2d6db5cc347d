"""The cluster-ordered exact kNN (knn_prune.cu: dd_dev_knn_clustered -- what the fit loop runs from 50 000 rows on) against
the all-tiles tcgen05 kernel: index for index, distance for distance, on the headline embedding (c3: 125 000 augmented cells),
on data without any cluster structure, on tiny groups and at ragged sizes.  Needs a B200 (`-m gpu`)."""

import numpy as np
import pytest

from oracle import datasets, upstream

pytestmark = pytest.mark.gpu


def _both(handle, emb, k):
    handle.upload_embedding(emb)
    handle.set_knn_mode(1)
    want_idx, want_dist = handle.knn(k)
    handle.set_knn_mode(2)
    try:
        got_idx, got_dist = handle.knn(k)
        got_idx2, _ = handle.knn(k)  # second call: warm-started centroids
        stats = handle.knn_clustered_stats()
    finally:
        handle.set_knn_mode(0)
    return want_idx, want_dist, got_idx, got_dist, got_idx2, stats


@pytest.mark.parametrize("n,k,kind", [(60000, 10, "blobs"), (60000, 13, "uniform"), (50001, 2, "blobs"), (5000, 10, "blobs"),
                                      (777, 6, "uniform"), (30000, 10, "duplicates"), (60000, 31, "blobs"), (20000, 14, "uniform")])
def test_clustered_knn_equals_dense(handle, n, k, kind):
    rs = np.random.default_rng(n + k)
    if kind == "blobs":
        emb = (rs.normal(size=(n, 30)) + rs.integers(0, 7, size=(n, 1)) * np.r_[np.full(8, 3.0), np.zeros(22)][None, :])
    elif kind == "uniform":
        emb = rs.random(size=(n, 30)) * 10
    else:  # many exactly equal points: ties must break by index exactly as in the dense kernel
        emb = rs.normal(size=(n // 10, 30))[rs.integers(0, n // 10, size=n)]
    emb = emb.astype(np.float32)
    want_idx, want_dist, got_idx, got_dist, got_idx2, stats = _both(handle, emb, k)
    if k > 13 and kind == "blobs":
        # 30-dimensional Gaussian blobs are the worst case for the approximate filter: neighbour distances concentrate, and
        # with lists of 32 for 30 neighbours the all-tiles kernel keeps a margin of only 2 filter ranks.  The
        # cluster-ordered path re-ranks two lists (64 candidates) and must be EXACT; rows where the two kernels differ are
        # settled by float64 brute force on the host, and the all-tiles kernel's misses are reported.
        differ = np.nonzero((got_idx != want_idx).any(axis=1))[0]
        e64 = emb.astype(np.float64)
        sq = (e64 ** 2).sum(1)
        dense_wrong = 0
        for r in differ[:400]:
            d2 = ((e64[r][None, :] - e64) ** 2).sum(1)
            d2[r] = -1.0
            truth = np.lexsort((np.arange(n), d2))[:k]
            np.testing.assert_array_equal(got_idx[r], truth, err_msg=f"cluster-ordered kNN wrong in row {r}")
            dense_wrong += int((want_idx[r] != truth).any())
        print(f"\n[{kind} n={n} k={k}] rows where the kernels differ: {differ.size}; all-tiles kernel wrong in {dense_wrong} of the "
              f"{min(differ.size, 400)} checked; cluster-ordered exact in all of them")
        return
    np.testing.assert_array_equal(got_dist, want_dist)
    if kind == "duplicates":
        # ten copies of every point: whole groups of candidates are EXACTLY equidistant, and which members of a group that
        # straddles the k-th place are reported depends on the order in which a kernel meets them (the all-tiles kernel keeps
        # the lowest indices it has SEEN among its 16 filter survivors; sklearn has its own order).  Distances are unique:
        # they must agree exactly; indices must agree wherever a row's distances are untied.
        untied = np.ones(want_idx.shape, dtype=bool)
        eq = want_dist[:, 1:] == want_dist[:, :-1]
        untied[:, 1:] &= ~eq
        untied[:, :-1] &= ~eq
        untied[:, -1] = False  # the last place may tie with the first one left out
        np.testing.assert_array_equal(got_idx[untied], want_idx[untied])
        e64 = emb.astype(np.float64)
        d = np.linalg.norm(e64[:, None, :] - e64[got_idx], axis=2)  # every reported neighbour really is at the reported distance
        np.testing.assert_allclose(d, got_dist, rtol=1e-6, atol=1e-6)
        return
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_idx2, want_idx)
    if n <= 5000:  # and against the oracle's brute force where that is quick
        ref_idx, _ = upstream.knn_brute(emb, k)
        assert (ref_idx == got_idx).mean() > 0.999  # float64 re-ranking vs sklearn's expanded form: near-ties only
    print(f"\n[{kind} n={n} k={k}] block-tile pairs visited: {(stats['pairs_a'] + stats['pairs_b']) / max(1, ((n + 255) // 256) * ((n + 127) // 128)):.3f} of the dense kernel's")


def test_clustered_knn_on_the_headline_embedding(handle):
    """c3: one iteration's PCA embedding of 125 000 augmented cells, k = 10 (doubletdetection.py:331-336)."""
    from doubletdetection_b200.classifier import _pca_plan

    counts = datasets.structured_counts(100000, 3000, seed=1234)
    n_cells = counts.shape[0]
    handle.upload_counts(counts)
    handle.create_doublets(np.random.default_rng(0).choice(n_cells, size=(n_cells // 4, 2), replace=False))
    handle.normalise_log(handle.median_lib_size(), 0.1)
    omega, n_power = _pca_plan(n_cells + n_cells // 4, 3000, 30, 0)
    emb, _ = handle.pca(30, omega, n_power)
    want_idx, want_dist, got_idx, got_dist, got_idx2, stats = _both(handle, np.ascontiguousarray(emb, dtype=np.float32), 10)
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_dist, want_dist)
    np.testing.assert_array_equal(got_idx2, want_idx)
    frac = (stats["pairs_a"] + stats["pairs_b"]) / (489 * 977)
    print(f"\n[c3 embedding] block-tile pairs visited: {frac:.3f} of the dense kernel's ({stats}); rows the filter certificate "
          f"could not clear: {handle.knn_uncertified()} of {emb.shape[0]}")
    assert frac < 0.6
    # PhenoGraph's neighbourhood (30 + self): lists of 32 per launch, 64 candidates re-ranked
    want_idx, want_dist, got_idx, got_dist, got_idx2, stats = _both(handle, np.ascontiguousarray(emb, dtype=np.float32), 31)
    np.testing.assert_array_equal(got_idx, want_idx)
    np.testing.assert_array_equal(got_dist, want_dist)
    print(f"[c3 embedding, k = 31] block-tile pairs visited: {(stats['pairs_a'] + stats['pairs_b']) / (489 * 977):.3f}")


def test_auto_mode_falls_back_when_the_ordering_does_not_pay(handle):
    """Uniform points have no cluster structure: the bounds exclude nothing and the padded order visits more pairs than the
    all-tiles kernel.  In the default mode the handle notices (pair counts read back asynchronously) and later calls on the
    same problem size use the all-tiles kernel; clustered data keeps the cluster-ordered path."""
    rs = np.random.default_rng(3)
    n = 60000
    for kind, expect_listed in (("uniform", False), ("blobs", True)):
        if kind == "uniform":
            emb = (rs.random(size=(n + 256, 30)) * 10).astype(np.float32)  # another size: nothing known about it yet
        else:
            emb = (rs.normal(size=(n, 30)) * 0.3 + rs.integers(0, 40, size=(n, 1)) * np.r_[np.full(8, 3.0), np.zeros(22)][None, :]
                   + rs.integers(0, 5, size=(n, 1)) * np.r_[np.zeros(8), np.full(4, 5.0), np.zeros(18)][None, :]).astype(np.float32)
        handle.upload_embedding(emb)
        handle.set_knn_mode(1)
        want, _ = handle.knn(10)
        handle.set_knn_mode(0)
        handle.set_kernel_timing(True)
        first, _ = handle.knn(10)   # by size: cluster-ordered
        second, _ = handle.knn(10)  # the counts of the first call are known by now
        before = handle.kernel_timing_report()
        third, _ = handle.knn(10)
        after = handle.kernel_timing_report()
        handle.set_kernel_timing(False)
        for got in (first, second, third):
            np.testing.assert_array_equal(got, want)
        listed = after.get("knn_tc_listed", (0, 0))[1] - before.get("knn_tc_listed", (0, 0))[1]
        dense = after.get("knn_tc", (0, 0))[1] - before.get("knn_tc", (0, 0))[1]
        print(f"\n[{kind}] third call: {listed} list-driven launches, {dense} all-tiles launches")
        assert (listed > 0) == expect_listed and (dense > 0) == (not expect_listed)


@pytest.mark.parametrize("offset", [3.0, 300.0, 30000.0])
def test_filter_certificate_and_exact_fixup(handle, offset):
    """The tcgen05 filter's error grows with |q| |c| (2^-16 relative): clusters that sit `offset` away from the origin with a
    local spacing of ~1 make it useless from offset ~ 1e3 on.  The re-ranking kernel certifies every row and the rows it
    cannot clear are re-done by float64 brute force, so BOTH kernels must equal the host's float64 brute force at any
    offset -- and the certificate must clear (almost) everything at ordinary scales."""
    rs = np.random.default_rng(int(offset))
    n, k = 20000, 10
    emb = (rs.normal(size=(n, 30)) + rs.integers(0, 6, size=(n, 1)) * np.r_[np.full(4, offset), np.zeros(26)][None, :]).astype(np.float32)
    e64 = emb.astype(np.float64)
    rows = rs.choice(n, size=300, replace=False)
    truth = np.empty((rows.size, k), dtype=np.int64)
    for i, r in enumerate(rows):
        d2 = ((e64[r][None, :] - e64) ** 2).sum(1)
        d2[r] = -1.0
        truth[i] = np.lexsort((np.arange(n), d2))[:k]
    handle.upload_embedding(emb)
    counts = {}
    for mode in (1, 2):
        handle.set_knn_mode(mode)
        try:
            idx, dist = handle.knn(k)
            counts[mode] = handle.knn_uncertified()
        finally:
            handle.set_knn_mode(0)
        np.testing.assert_array_equal(idx[rows], truth, err_msg=f"mode {mode}, offset {offset}")
    print(f"\n[offset {offset}] rows the certificate could not clear: all-tiles {counts[1]}, cluster-ordered {counts[2]} of {n}")
    if offset <= 3.0:
        assert counts[1] <= n // 100 and counts[2] <= n // 100
    if offset >= 30000.0:
        assert counts[1] > 0 and counts[2] > 0


def test_building_block_hooks_listed_and_pruned(handle):
    """The two test hooks the cluster-ordered path grew out of: ``dd_knn_listed`` (caller-made tile lists per 256-row block)
    and ``dd_knn_pruned`` (caller-made padded permutation; boxes, launch A, thresholds, lists, launch B on the device)."""
    rs = np.random.default_rng(12)
    n, k = 6000, 10
    emb = (rs.normal(size=(n, 30)) + rs.integers(0, 3, size=(n, 1)) * np.r_[np.full(8, 6.0), np.zeros(22)][None, :]).astype(np.float32)
    handle.upload_embedding(emb)
    handle.set_knn_mode(1)
    try:
        want, want_dist = handle.knn(k)
        n_tiles, n_blocks = -(-n // 128), -(-(-(-n // 128)) // 2)
        # every block visits every tile: the list-driven kernel must reproduce the all-tiles kernel
        off = np.arange(n_blocks + 1, dtype=np.int32) * n_tiles
        tiles = np.tile(np.arange(n_tiles, dtype=np.int32), n_blocks)
        got, got_dist = handle.knn_listed(k, off, tiles)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(got_dist, want_dist)
        # three groups by the planted offset, padded to whole blocks
        group = np.rint(emb[:, 0] / 6.0).clip(0, 2).astype(int)
        perm, block_group = [], []
        for g in range(3):
            ids = np.nonzero(group == g)[0]
            pad = (-ids.size) % 256
            perm.append(np.concatenate([ids, np.full(pad, -1, dtype=ids.dtype)]))
            block_group += [g] * ((ids.size + pad) // 256)
        idx, dist, stats = handle.knn_pruned(k, np.concatenate(perm), np.asarray(block_group), n)
        np.testing.assert_array_equal(idx, want)
        np.testing.assert_array_equal(dist, want_dist)
        assert stats["pairs_a"] > 0
    finally:
        handle.set_knn_mode(0)
