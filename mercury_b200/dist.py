"""Multi-GPU plumbing: one process per GPU, frames sharded by index, tables broadcast once.

SURVEY.md 8e / north_star: the RX path has no exchange step -- OFDM frames are independent -- so ranks only share
the read-only table blob (rank 0 builds it, everyone else receives it through torch.distributed: NCCL over
NVLink on the GPU box, gloo in the CPU tests) and then process disjoint contiguous frame ranges with zero
per-batch collectives.
"""
import numpy as np


def shard_range(n_frames, rank, world_size):
    """Contiguous [start, stop) of `n_frames` owned by `rank`; sizes differ by at most one frame."""
    base, rem = divmod(int(n_frames), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_tables(blob, src=0, device=None, group=None):
    """Rank `src` passes its blob (uint8 ndarray); every rank returns the same bytes. Uses the default process group."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        n[0] = int(blob.size)
    dist.broadcast(n, src, group=group)
    buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    if rank == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(blob, np.uint8)))
    dist.broadcast(buf, src, group=group)
    return buf.cpu().numpy()
