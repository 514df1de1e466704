"""Multi-GPU plumbing: one process per GPU, images sharded over ranks, NCCL only at the edges.

Images in a batch never interact (no BatchNorm, every reduction is
per-sample), so the data path has no collective.  What remains is (1) a one-off broadcast of
the 1.5 M parameters from the rank that read the checkpoint and (2) optionally gathering the
outputs on one rank.  Both go through torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of ``n_items`` for ``rank`` (first ranks get the
    remainder).  Empty when there are more ranks than items."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """One flat broadcast of every parameter and buffer (6.05 MB for Wave-Mamba)."""
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n


@torch.no_grad()
def sharded_forward(forward, batch: Optional[torch.Tensor], device: torch.device, src: int = 0,
                    gather: bool = True) -> Optional[torch.Tensor]:
    """Run ``forward`` on this rank's slice of a batch that lives on rank ``src``.

    batch: (B,3,H,W) on rank ``src`` (None elsewhere).  Returns the (B,3,H,W) result on rank
    ``src`` when ``gather`` (None elsewhere), else this rank's slice.  Edge traffic only:
    one scatter of the inputs and one gather of the outputs.
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = [None]
    if rank == src:
        meta = [tuple(batch.shape)]
    dist.broadcast_object_list(meta, src=src)
    B, C, H, W = meta[0]
    begin, end = shard_range(B, rank, world)
    mine = torch.empty(end - begin, C, H, W, device=device)
    if rank == src:
        chunks = []
        for r in range(world):
            b0, b1 = shard_range(B, r, world)
            chunks.append(batch[b0:b1].to(device).contiguous())
    else:
        chunks = None
    # scatter needs equal-size chunks in some backends; use point-to-point for ragged shards
    if rank == src:
        for r in range(world):
            if r == src:
                mine.copy_(chunks[r])
            elif chunks[r].numel():
                dist.send(chunks[r], dst=r)
    elif mine.numel():
        dist.recv(mine, src=src)
    out = forward(mine) if mine.shape[0] else mine
    if not gather:
        return out
    if rank == src:
        result = torch.empty(B, C, H, W, device=device)
        for r in range(world):
            b0, b1 = shard_range(B, r, world)
            if r == src:
                result[b0:b1] = out
            elif b1 > b0:
                buf = torch.empty(b1 - b0, C, H, W, device=device)
                dist.recv(buf, src=r)
                result[b0:b1] = buf
        return result
    if out.numel():
        dist.send(out.contiguous(), dst=src)
    return None
