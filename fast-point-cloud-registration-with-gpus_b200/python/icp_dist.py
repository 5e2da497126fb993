"""Host-side helpers for the source-sharded multi-GPU run (one process per GPU, torch.distributed as
plumbing: rendezvous and the broadcast of the NCCL unique id; the data-path collective itself is the
ncclAllReduce inside libicp_b200.so)."""


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of `n` source points for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_indices(n, rank, world, block=2048):
    """Interleaved shard: source blocks of `block` consecutive points are dealt round-robin, rank r taking blocks
    r, r + world, r + 2*world, ... Every rank then sees the same mix of regions of the cloud, so the per-rank cost of
    the matching step (which depends on how many sub-tiles a region cannot exclude) is balanced; with contiguous
    shards the slowest rank was 5-6 % behind the mean at 2-8 GPUs. Moment sums do not care how points are dealt.
    `block` = the matching kernels' source block (8 sources x 256 threads), which keeps warps spatially coherent."""
    import numpy as np
    nblocks = (n + block - 1) // block
    mine = np.arange(rank, nblocks, world, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def deal_blocks(nblocks, weights):
    """Deal `nblocks` source blocks to len(weights) ranks in proportion to `weights` (measured rank speeds), spreading
    each rank's blocks evenly over the cloud: block b goes to the rank that is furthest behind its quota
    (largest-deficit rule; ties to the lowest rank). Deterministic, so every rank computes the same deal from the
    same all-gathered weights. Equal weights reproduce the round-robin deal of shard_indices."""
    import numpy as np
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    got = np.zeros(len(w))
    owner = np.empty(nblocks, dtype=np.int64)
    for b in range(nblocks):
        r = int(np.argmax(w * (b + 1) - got - 1e-12 * np.arange(len(w))))
        owner[b] = r
        got[r] += 1
    return owner


def shard_indices_weighted(n, rank, weights, block=2048):
    """Like shard_indices, with the blocks dealt by deal_blocks(weights): faster GPUs get proportionally more blocks."""
    import numpy as np
    nblocks = (n + block - 1) // block
    mine = np.nonzero(deal_blocks(nblocks, weights) == rank)[0].astype(np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def broadcast_bytes(payload, size, src=0):
    """Broadcast `size` bytes from rank `src` to every rank of the default process group."""
    import torch
    import torch.distributed as dist
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8)
    else:
        t = torch.zeros(size, dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src)
    return bytes(t.cpu().tolist())
