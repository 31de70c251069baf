"""Host-side plumbing of the data-parallel (gradient-sum) training mode.

One process per GPU.  The path has exactly one exchange step per training
step: the sum over ranks of the flat per-rank delta buffer
(-lr * sum_{b in shard} grad_b, all layers concatenated), after which every
rank applies the same W += delta (Nnet::BatchTrain semantics over the global
batch, nnet/nnet.h:624-669).  torch.distributed is used for rendezvous only.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of the global batch owned by `rank`
    (contiguous N/G samples per GPU; remainders go to low ranks)."""
    base, rem = divmod(n_samples, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def exchange_unique_id(dist, make_id, nbytes: int, device=None) -> bytes:
    """Rank 0 creates the communicator id (pb_comm_unique_id), everyone gets it."""
    import torch
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == 0:
        raw = make_id()
        assert len(raw) == nbytes
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())
