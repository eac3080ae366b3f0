"""Frame-batch sharding across the GPUs of one box (BASELINE config 5).

Frames, stereo halves and sequences are independent: the extract + describe path shards by
contiguous frame ranges with NO data-path collective.  The one exchange step is the
all-gather of every rank's keyframe descriptor block (fixed stride Kmax x 32 bytes per frame
plus an int32 count per frame) over NCCL/NVLink, after which each rank brute-force matches
its local frames against all gathered keyframes.

Host logic only (torch.distributed with whatever backend the process group uses: nccl on
the GPU box, gloo in the CPU tests); the kernels are behind ORBextractor / ORBmatcher.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous, balanced [begin, end) of `n_items` units for `rank`; a stereo pair (two
    consecutive frames) never straddles two ranks when n_items is counted in pairs."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_descriptors(desc, counts, group=None):
    """desc: torch uint8 [F_local, Kmax, 32], counts: torch int32 [F_local] (same F_local on
    every rank).  Returns (all_desc [world*F_local, Kmax, 32], all_counts [world*F_local]),
    rank-major, via all_gather_into_tensor (one flat NCCL all-gather over NVSwitch)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    out_d = torch.empty((world * desc.shape[0],) + tuple(desc.shape[1:]), dtype=desc.dtype, device=desc.device)
    out_c = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(out_d, desc.contiguous(), group=group)
    dist.all_gather_into_tensor(out_c, counts.contiguous(), group=group)
    return out_d, out_c


def cross_shard_pairs(n_local, rank, world, neighbours=1):
    """(local frame, global keyframe) pairs each rank matches after the all-gather: every local
    frame against the same-index frame of the next `neighbours` ranks (ring), so the work is
    balanced and every pair crosses a shard boundary."""
    pairs = []
    for f in range(n_local):
        for d in range(1, neighbours + 1):
            other = (rank + d) % world
            pairs.append((f, other * n_local + f))
    return np.asarray(pairs, np.int64).reshape(-1, 2)
