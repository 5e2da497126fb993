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


# ---- the C/C++ in-process multi-GPU path (icpb_group_*): no torch, no launcher ------------------------------------
def _gpu_count(ib):
    import ctypes
    n = ctypes.c_int(0)
    ib.lib.icpb_device_count(ctypes.byref(n))
    return n.value


@pytest.mark.parametrize("world,metric,W,block,peer", [
    (2, 0, 128, 2048, 1), (2, 1, 128, 2048, 1), (2, 0, 300, 0, 1), (2, 0, 128, 2048, 0), (2, 1, 128, 0, 0),
    (4, 0, 300, 2048, 1), (8, 0, 300, 2048, 1), (8, 1, 200, 2048, 1), (8, 0, 300, 0, 0), (8, 0, 1000, 2048, 1)])
def test_group_reproduces_one_gpu(ib, world, metric, W, block, peer):
    """world GPUs driven from ONE process through icpb_group_* (sources dealt in blocks of 2048 round-robin — what the
    bench and the executables use — or in contiguous shards; fused peer-memory exchange or ncclCommInitAll + allreduce)
    against one GPU: identical correspondences for every source point, identical iteration counts, trajectory and
    transform within FP64 summation-order noise."""
    if _gpu_count(ib) < world:
        pytest.skip("needs %d GPUs" % world)
    import icp_synth
    D, M = icp_synth.p2p_clouds(W)
    mode = ib.DIST_SQRT if metric else ib.DIST_SQ
    prm = ib.default_params(metric=metric, dist_mode=mode, max_iter=64)
    with ib.Context(0) as ctx:
        ctx.set_target(M); ctx.set_source(D)
        if metric:
            ctx.estimate_normals(4)
        err, res = ctx.run(prm)
        idx, P1 = ctx.correspondences(), ctx.get_source()
    os.environ["ICPB_PEER"] = str(peer)
    try:
        with ib.Group(world) as g:
            assert g.info()["peer_exchange"] == bool(peer)
            g.set_target(M); g.set_source(D, block)
            if metric:
                g.estimate_normals(4)
            l0 = g.launch_count(1)
            e, r = g.run(prm)
            launches = g.launch_count(1) - l0
            gi, gP = g.correspondences(), g.get_source()
    finally:
        del os.environ["ICPB_PEER"]
    assert (r.iterations, r.iterations_run) == (res.iterations, res.iterations_run)
    assert launches <= (3 if peer else 5) * (r.iterations_run + 4) + 8          # fused: the 3 kernels of one GPU per iteration
    k = res.iterations + 2
    assert np.all(np.abs(e[:k] - err[:k]) <= 1e-6 * np.abs(err[:k]) + 1e-7)
    assert np.abs(np.array(r.R[:]) - np.array(res.R[:])).max() < 1e-9 and np.abs(np.array(r.t[:]) - np.array(res.t[:])).max() < 1e-9
    assert np.array_equal(gi, idx), "correspondences differ from the single-GPU run"
    assert np.abs(gP - P1).max() <= 1e-5


def test_group_of_one_is_a_plain_context(ib):
    import icp_synth
    D, M = icp_synth.p2p_clouds(64)
    with ib.Group(1) as g, ib.Context(0) as ctx:
        g.set_target(M); g.set_source(D)
        e, r = g.run(ib.default_params())
        ctx.set_target(M); ctx.set_source(D)
        e1, r1 = ctx.run(ib.default_params())
        assert np.array_equal(e, e1) and list(r.R) == list(r1.R) and np.array_equal(g.correspondences(), ctx.correspondences())
        assert g.info() == {"ndev": 1, "peer_exchange": False, "devices": [0]}


def test_group_executable_matches_single_gpu_stdout(ib):
    """apps/icp_point_to_point --gpus N (host C++ only): same printed trajectory as one GPU."""
    import subprocess
    n = _gpu_count(ib)
    if n < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "apps", "icp_point_to_point")
    keep = lambda out: [l for l in out.splitlines() if not l.startswith("Elapsed") and not l.startswith("[report]")]
    one = subprocess.run([exe, "--width", "200"], capture_output=True, text=True, check=True).stdout
    many = subprocess.run([exe, "--width", "200", "--gpus", str(min(n, 8))], capture_output=True, text=True, check=True).stdout
    assert keep(one) == keep(many) and "ICP converged successfully!" in many
