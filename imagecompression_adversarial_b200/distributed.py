"""Multi-GPU plumbing: independent images shard across ranks (one process per GPU), no collective inside
the loop; one final gather of the per-image metrics (attack_rd.py:654-688 is a sequential for-loop in the
reference -- the sharding is new capability)."""
import torch


def shard_indices(n_items, rank, world):
    """Round-robin partition: rank r takes items r, r+world, ... (images[r::world])."""
    return list(range(rank, n_items, world))


def gather_metrics(local, n_items, rank, world, group=None):
    """``local``: float tensor [n_local, K] of per-image metrics of this rank's shard (device of the
    backend: CUDA for nccl, CPU for gloo).  Returns on every rank the [n_items, K] table in image order."""
    if world == 1:
        return local
    import torch.distributed as dist
    k = local.shape[1]
    n_max = (n_items + world - 1) // world
    pad = torch.full((n_max, k), float("nan"), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    out = torch.empty(n_items, k, dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][:len(idx)]
    return out
