"""Host-side helpers for the source-sharded multi-GPU run (one process per GPU, torch.distributed as
plumbing: rendezvous and the broadcast of the NCCL unique id; the data-path collective itself is the
ncclAllReduce inside libicp_b200.so)."""


def shard_bounds(n, rank, world):
    """Contiguous shard [lo, hi) of `n` source points for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


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
