"""Worker of the cell-block sharding test: run under torchrun with >= 2 ranks, one GPU each.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/sharded_worker.py [n_cells n_genes]

Every rank checks its share against an UNSHARDED handle on its own GPU and against the float64 oracle:
dense block bit-exact, embedding within 1e-4, kNN lists identical to the single-GPU kNN of the same
embedding, fit results identical on all ranks and in agreement with the single-GPU fit.  Prints one
"SHARDED_OK rank=r" line per rank on success."""

import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from doubletdetection_b200 import BoostClassifier, _capi  # noqa: E402
from doubletdetection_b200.classifier import _pca_plan, broadcast_token  # noqa: E402
from oracle import datasets, pca_f64  # noqa: E402


def main():
    n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    n_genes = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    counts = datasets.structured_counts(n_cells, n_genes, seed=1234)
    n_synth = n_cells // 4
    n_aug = n_cells + n_synth
    parents = np.random.default_rng(5).choice(n_cells, size=(n_synth, 2), replace=False)
    omega, npi = _pca_plan(n_aug, n_genes, 30, 0)

    # ---- reference on this GPU: the unsharded path
    ref = _capi.Handle(dev)
    ref.upload_counts(counts)
    ref.create_doublets(parents)
    med = ref.median_lib_size()
    ref.normalise_log(med, 0.1)
    dense_ref = ref.download_dense()
    emb_ref, sv_ref = ref.pca(30, omega, npi)

    # ---- sharded handle
    h = _capi.Handle(dev)
    h.upload_counts(counts)
    token = broadcast_token(dist, _capi.comm_unique_id() if rank == 0 else None, dev)
    h.comm_init(rank, world, token)
    h.shard_cells(True)
    h.create_doublets(parents)
    info = h.comm_info()
    n0, n1 = _capi.block_of(n_cells, rank, world)
    m0, m1 = _capi.block_of(n_synth, rank, world)
    assert (info["first_cell"], info["n_cells"], info["first_synth"], info["n_synth"]) == (n0, n1 - n0, m0, m1 - m0), info
    assert h.median_lib_size() == med
    h.normalise_log(med, 0.1)
    local = h.download_dense(0, info["n_cells"] + info["n_synth"])
    want = np.vstack([dense_ref[n0:n1], dense_ref[n_cells + m0:n_cells + m1]])
    assert np.array_equal(local, want), "dense block differs from the unsharded build"

    emb, sv = h.pca(30, omega, npi)
    assert emb.shape == (n_aug, 30)
    truth, _, _ = pca_f64.randomized_pca_f64(dense_ref, 30, random_state=0)
    scale = np.abs(truth).max()
    err_truth = np.abs(emb - truth).max() / scale
    err_ref = np.abs(emb - emb_ref).max() / scale
    assert err_truth < 1e-4, f"sharded embedding vs float64 oracle: {err_truth}"
    assert err_ref < 2e-5, f"sharded vs unsharded embedding: {err_ref}"
    # every rank holds the same gathered embedding
    t = torch.from_numpy(emb.copy()).cuda()
    t0 = t.clone()
    dist.broadcast(t0, src=0)
    assert torch.equal(t, t0), "ranks disagree on the gathered embedding"

    idx, dd = h.knn(10)
    ref.upload_embedding(emb)
    idx_ref, dd_ref = ref.knn(10)
    assert np.array_equal(idx, idx_ref), "sharded kNN lists differ from the single-GPU kNN of the same embedding"
    assert np.array_equal(dd, dd_ref)
    # the cluster-ordered form under sharding (what the fit loop runs from 50 000 rows on): rank 0's ordering is broadcast, the
    # 256-row blocks of the permuted order are dealt round-robin, the ranks' result buffers are summed
    for k in (10, 31):
        h.set_knn_mode(1)
        ref.set_knn_mode(1)
        idx_ref, dd_ref = ref.knn(k)
        h.set_knn_mode(2)
        idx2, dd2 = h.knn(k)
        idx3, dd3 = h.knn(k)  # warm-started centroids
        h.set_knn_mode(0)
        ref.set_knn_mode(0)
        assert np.array_equal(idx2, idx_ref) and np.array_equal(idx3, idx_ref), f"sharded cluster-ordered kNN differs (k = {k})"
        assert np.array_equal(dd2, dd_ref) and np.array_equal(dd3, dd_ref)

    # scaled variant: the column statistics are all-reduced
    ref.create_doublets(parents)
    ref.normalise_log(med, 0.1)
    ref.standard_scale(15.0)
    sc_ref = ref.download_dense()
    h.normalise_log(med, 0.1)
    h.standard_scale(15.0)
    sc = h.download_dense(0, info["n_cells"] + info["n_synth"])
    want = np.vstack([sc_ref[n0:n1], sc_ref[n_cells + m0:n_cells + m1]])
    np.testing.assert_allclose(sc, want, rtol=1e-6, atol=1e-6)
    h.close()
    ref.close()

    # ---- the public API: distributed="cells" against the single-GPU fit
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        kw = dict(n_iters=4, clustering_algorithm="louvain", random_state=0, n_jobs=2, device=dev)
        single = BoostClassifier(**kw).fit(counts)
        sharded = BoostClassifier(distributed="cells", **kw).fit(counts)
        lab_single = single.predict(p_thresh=1e-3, voter_thresh=0.5)
        lab = sharded.predict(p_thresh=1e-3, voter_thresh=0.5)
        # the same fit with the cluster-ordered kNN forced (it is the default only from 50 000 rows on): identical results
        forced = BoostClassifier(distributed="cells", **kw)
        forced._native().set_knn_mode(2)
        forced.fit(counts)
        forced._native().set_knn_mode(0)
    assert np.array_equal(forced.communities_, sharded.communities_), "cluster-ordered kNN changed the sharded fit"
    assert np.array_equal(forced.all_log_p_values_, sharded.all_log_p_values_)
    assert np.array_equal(np.asarray(single.parents_), np.asarray(sharded.parents_))
    agree = float(np.mean(lab == lab_single))
    assert agree >= 0.995, f"sharded fit labels agree with the single-GPU fit on only {agree:.4f} of the cells"
    for name in ("all_scores_", "all_log_p_values_", "communities_", "synth_communities_"):
        a = torch.from_numpy(np.ascontiguousarray(getattr(sharded, name)).view(np.int64).copy()).cuda()
        b = a.clone()
        dist.broadcast(b, src=0)
        assert torch.equal(a, b), f"ranks disagree on {name}"
    same_comm = float(np.mean(sharded.communities_ == single.communities_))
    print(f"SHARDED_OK rank={rank} world={world} emb_err_vs_f64={err_truth:.2e} emb_err_vs_single={err_ref:.2e} "
          f"label_agreement={agree:.4f} identical_community_entries={same_comm:.4f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
