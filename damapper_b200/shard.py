"""Multi-GPU host logic of the mapping path: reads blocks are independent
(damapper.c:825-914), so rank r of N maps blocks r, r+N, ...; the only exchange is the sorted
reference index (16 B x reference k-mers + the two sentinels, per orientation), built on one rank
and broadcast.  Backend-agnostic: NCCL on device buffers in bench.py, gloo on CPU tensors in the
tests."""
from __future__ import annotations

from typing import List, Optional


def assign_blocks(nblocks: int, world: int, rank: int) -> List[int]:
    """Reads blocks mapped by `rank`: round-robin, every block exactly once over the ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %d of %d" % (rank, world))
    return list(range(rank, nblocks, world))


def broadcast_index(dist, torch, payload: Optional["torch.Tensor"], device, src: int = 0):
    """Broadcast a reference index (uint8 tensor of (len+2)*16 bytes) from `src`.

    Returns (tensor, nrecords).  `payload` is only read on `src`; the other ranks pass None."""
    rank = dist.get_rank()
    n = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        if payload.dtype != torch.uint8 or payload.numel() % 16 != 0 or payload.numel() < 32:
            raise ValueError("index payload must be uint8, (len+2)*16 bytes")
        n[0] = payload.numel() // 16 - 2
    dist.broadcast(n, src)
    ln = int(n.item())
    buf = payload if rank == src else torch.empty((ln + 2) * 16, dtype=torch.uint8, device=device)
    dist.broadcast(buf, src)
    return buf, ln
