"""Multi-GPU sharding of independent MPC instances.

Instances are independent OCPs, so a batch is split into contiguous shards, one per rank / GPU, with NO collective on
the data path.  The only exchange is the optional gather of results (trajectories + per-instance summaries) with one
`all_gather` per array — NCCL over NVLink on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_range(batch, rank, world):
    """Contiguous, balanced [lo, hi) of `batch` instances owned by `rank`."""
    base, rem = divmod(int(batch), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_arrays(local, batch, group=None, device=None):
    """all_gather per-rank shards (dict of numpy arrays with the instance axis first) into full-batch arrays on every
    rank.  Shards may be ragged (batch % world != 0): they are padded to the largest shard for the collective."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    mx = max(sizes)
    out = {}
    for name, arr in local.items():
        arr = np.ascontiguousarray(arr)
        assert arr.shape[0] == sizes[rank], (name, arr.shape, sizes[rank])
        pad = np.zeros((mx,) + arr.shape[1:], dtype=arr.dtype)
        pad[: arr.shape[0]] = arr
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t, group=group)
        out[name] = np.concatenate([b.cpu().numpy()[: sizes[r]] for r, b in enumerate(bufs)], axis=0)
    return out
