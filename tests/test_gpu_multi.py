"""GPU suite (-m gpu, needs >= 2 GPUs, otherwise skipped): source-sharded registration over NCCL.
Two processes, one per GPU, each holding half of the source and the whole target, must reproduce the
single-GPU run: identical correspondences, iteration counts, and error trajectory / transform within FP64
summation noise (the shards' moment sums are added in a different order)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rank(rank, world, uid, W, metric, q, peer=1):
    os.environ["ICPB_PEER"] = str(peer)
    sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
    import icp_b200 as ib
    import icp_dist
    import icp_synth
    D, M = icp_synth.p2p_clouds(W)
    lo, hi = icp_dist.shard_bounds(D.shape[0], rank, world)
    ctx = ib.Context(rank, rank, world, uid)
    ctx.set_target(M)
    ctx.set_source(D[lo:hi])
    mode = ib.DIST_SQRT if metric else ib.DIST_SQ
    if metric:
        ctx.estimate_normals(4)
    launches0 = ctx.launch_count()
    err, res = ctx.run(ib.default_params(metric=metric, dist_mode=mode, max_iter=50))
    launches = ctx.launch_count() - launches0
    q.put((rank, lo, hi, err, res.iterations, res.iterations_run, list(res.R), list(res.t), ctx.correspondences(),
           ctx.dist_info()["peer_exchange"], launches))
    ctx.close()


@pytest.mark.parametrize("metric,W,peer", [(0, 128, 1), (1, 128, 1), (0, 300, 1), (0, 128, 0), (1, 128, 0)])
def test_two_gpus_reproduce_one_gpu(ib, metric, W, peer):
    """peer=1: sums exchanged inside K2/K7/K4 over peer memory (the default); peer=0: ncclAllReduce between kernels."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import icp_synth
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    uid = ib.nccl_unique_id()
    procs = [mpc.Process(target=_rank, args=(r, 2, uid, W, metric, q, peer)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=300) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single GPU
    D, M = icp_synth.p2p_clouds(W)
    with ib.Context(0) as ctx:
        ctx.set_target(M); ctx.set_source(D)
        mode = ib.DIST_SQRT if metric else ib.DIST_SQ
        if metric:
            ctx.estimate_normals(4)
        err, res = ctx.run(ib.default_params(metric=metric, dist_mode=mode, max_iter=50))
        idx = ctx.correspondences()
    for rank, lo, hi, e, it, run, R, t, sub_idx, peer_on, launches in out:
        assert peer_on == bool(peer), "fused peer-memory exchange %s" % ("not active" if peer else "active despite ICPB_PEER=0")
        # fused: 3 kernels per iteration as on one GPU (+ the key reset); NCCL path: 5 of ours per iteration + 2 allreduces
        assert launches <= (3 if peer else 5) * (run + 4) + 8, (launches, run)       # up to sync_every = 4 iterations are enqueued past the stop flag
        assert (it, run) == (res.iterations, res.iterations_run)
        k = res.iterations + 2
        assert np.all(np.abs(e[:k] - err[:k]) <= 1e-6 * np.abs(err[:k]) + 1e-7)
        assert np.abs(np.array(R) - np.array(res.R[:])).max() < 1e-9 and np.abs(np.array(t) - np.array(res.t[:])).max() < 1e-9
        assert np.array_equal(sub_idx, idx[lo:hi])
    assert out[0][3].tobytes() == out[1][3].tobytes() and out[0][6] == out[1][6], "both ranks hold identical bits"
