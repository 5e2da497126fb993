"""CPU suite: the multi-GPU host logic with world_size 2 over gloo.

On GPUs the engine shards source points over ranks, replicates the target and combines 16 FP64 moment
sums with ncclAllReduce (csrc/dist.cpp). Here the same plan runs on CPU: icp_dist's shard bounds + the
oracle as compute + torch.distributed(gloo) as the collective. It checks that (a) shards partition the
cloud, (b) summed shard moments equal the global moments, (c) every rank derives the same R,T bits and the
same RMS as the unsharded computation.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
    import torch
    import torch.distributed as dist
    import oracle as orc
    import icp_dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    D, M = orc.synth_p2p(40)
    n = D.shape[0]
    lo, hi = icp_dist.shard_bounds(n, rank, world)
    P = D[lo:hi].copy()
    out = []
    for it in range(3):
        idx = orc.match(P, M, 0)
        mom = torch.from_numpy(orc.moments(P, M, idx))
        dist.all_reduce(mom)
        R, T = orc.rt_from_moments(mom.numpy())
        P = orc.transform(P, R.astype(np.float32), T.astype(np.float32))
        e = torch.tensor([orc.rms(P, M, idx) ** 2 * (hi - lo)], dtype=torch.float64)
        dist.all_reduce(e)
        out.append((mom.numpy().copy(), R.copy(), T.copy(), float(np.sqrt(e.item() / n))))
    uid = icp_dist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
    q.put((rank, lo, hi, out, uid))
    dist.destroy_process_group()


def test_sharded_moments_allreduce_world2(orc):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, out0, uid0), (r1, lo1, hi1, out1, uid1) = res
    assert (lo0, hi1) == (0, 1600) and hi0 == lo1, "shards partition the source cloud"
    assert uid0 == uid1 == bytes(range(128)), "the NCCL unique id reaches every rank unchanged"
    # single-process reference
    D, M = orc.synth_p2p(40)
    P = D.copy()
    for it in range(3):
        idx = orc.match(P, M, 0)
        mom = orc.moments(P, M, idx)
        R, T = orc.rt_from_moments(mom)
        P = orc.transform(P, R.astype(np.float32), T.astype(np.float32))
        e = orc.rms(P, M, idx)
        for out in (out0, out1):
            m_, R_, T_, e_ = out[it]
            assert np.allclose(m_, mom, rtol=1e-13, atol=1e-10)
            assert np.abs(R_ - R).max() < 1e-12 and np.abs(T_ - T).max() < 1e-12
            assert abs(e_ - e) < 1e-9
        assert np.array_equal(out0[it][1], out1[it][1]) and np.array_equal(out0[it][2], out1[it][2]), "identical bits on every rank"


def test_shard_bounds_cover_everything():
    sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
    import icp_dist
    for n in (1, 7, 1000, 1000000):
        for w in (1, 2, 3, 4, 8):
            b = [icp_dist.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_interleaved_shards_partition_the_cloud():
    """Blocks of 2048 sources dealt round-robin (bench.py's default): a partition, block-coherent, balanced to a block."""
    sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
    import icp_dist
    for n in (1, 2047, 2048, 2049, 100000, 1000000):
        for w in (1, 2, 3, 8):
            parts = [icp_dist.shard_indices(n, r, w) for r in range(w)]
            allidx = np.concatenate(parts)
            assert allidx.size == n and np.array_equal(np.sort(allidx), np.arange(n))
            sizes = [p.size for p in parts]
            assert max(sizes) - min(sizes) <= 2048
            for r, p in enumerate(parts):
                assert np.all(np.diff(p) > 0)
                assert np.all((p // 2048) % w == r)


def test_weighted_deal_is_a_partition_and_proportional():
    sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
    import icp_dist
    assert np.array_equal(icp_dist.deal_blocks(16, [1, 1, 1, 1]), np.arange(16) % 4)      # equal speeds: round-robin
    w = [1.0, 0.96, 1.04, 1.0, 0.9, 1.1, 1.0, 1.0]
    owner = icp_dist.deal_blocks(489, w)
    counts = np.bincount(owner, minlength=8)
    assert counts.sum() == 489 and np.abs(counts - 489 * np.array(w) / sum(w)).max() <= 1.0
    for r in range(8):                       # spread over the whole cloud, not bunched
        mine = np.nonzero(owner == r)[0]
        assert np.diff(mine).max() <= 2 * 8 / min(w)
    n = 1000000
    parts = [icp_dist.shard_indices_weighted(n, r, w) for r in range(8)]
    allidx = np.concatenate(parts)
    assert allidx.size == n and np.array_equal(np.sort(allidx), np.arange(n))
