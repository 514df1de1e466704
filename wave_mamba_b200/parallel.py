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
    tensors = list(module.parameters()) + list(module.buffers())
    if not tensors or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        flat = torch.cat([t.reshape(-1).float() for t in tensors])
        dist.broadcast(flat, src=src)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))      # in-place, bumps ``_version``
            off += n
    from . import ops
    ops.clear_pack_cache()                             # pre-packed conv weights are stale now


_DTYPES = [torch.float32, torch.uint8, torch.float16, torch.bfloat16, torch.float64, torch.int32]


def _exchange(ops_list):
    """Post a group of point-to-point transfers (one NCCL group: they run concurrently over NVLink)."""
    if not ops_list:
        return []
    return dist.batch_isend_irecv(ops_list)


@torch.no_grad()
def sharded_forward(forward, batch: Optional[torch.Tensor], device: torch.device, src: int = 0,
                    gather: bool = True, stats: Optional[dict] = None) -> Optional[torch.Tensor]:
    """Run ``forward`` on this rank's slice of a batch that lives on rank ``src``.

    batch: (B, ...) tensor on rank ``src`` (host or device; None elsewhere), any dtype in _DTYPES;
    ``forward`` maps a (b, ...) slice to a tensor of the same shape and dtype.  Returns the (B, ...)
    result on ``src``'s device when ``gather`` (None elsewhere), else this rank's slice.

    Edge traffic only: the inputs are scattered and the outputs gathered with grouped
    point-to-point transfers (``batch_isend_irecv``: one NCCL group each way, ragged shards allowed);
    the root starts on its own shard while its sends are in flight.  ``stats`` (rank ``src``, CUDA):
    appends (start, end) event pairs under "scatter" / "gather".
    """
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = torch.zeros(10, dtype=torch.int64, device=device)
    if rank == src:
        shape = list(batch.shape)
        if len(shape) > 8:
            raise ValueError("sharded_forward: at most 8 dimensions")
        meta[0] = len(shape)
        meta[1] = _DTYPES.index(batch.dtype)
        meta[2:2 + len(shape)] = torch.tensor(shape, dtype=torch.int64)
    dist.broadcast(meta, src=src)
    meta = meta.tolist()
    shape, dtype = meta[2:2 + meta[0]], _DTYPES[meta[1]]
    B, rest = shape[0], shape[1:]
    begin, end = shard_range(B, rank, world)
    timed = stats is not None and rank == src and device.type == "cuda"

    def mark():
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        return ev

    # ---- scatter ----------------------------------------------------------------------------
    t0 = mark() if timed else None
    if rank == src:
        whole = batch.to(device, non_blocking=True).contiguous()
        sends = []
        for r in range(world):
            b0, b1 = shard_range(B, r, world)
            if r != src and b1 > b0:
                sends.append(dist.P2POp(dist.isend, whole[b0:b1], r))
        reqs = _exchange(sends)
        mine = whole[begin:end]
    else:
        mine = torch.empty([end - begin] + rest, dtype=dtype, device=device)
        reqs = _exchange([dist.P2POp(dist.irecv, mine, src)] if end > begin else [])
        for q in reqs:
            q.wait()
        reqs = []
    if timed:
        stats.setdefault("scatter", []).append((t0, mark()))
    out = forward(mine) if end > begin else mine
    for q in reqs:      # the root's sends overlapped its own forward
        q.wait()
    if not gather:
        return out
    # ---- gather -----------------------------------------------------------------------------
    t1 = mark() if timed else None
    result = None
    if rank == src:
        result = torch.empty(shape, dtype=out.dtype, device=device)
        recvs = []
        for r in range(world):
            b0, b1 = shard_range(B, r, world)
            if r != src and b1 > b0:
                recvs.append(dist.P2POp(dist.irecv, result[b0:b1], r))
        reqs = _exchange(recvs)
        result[begin:end] = out
    else:
        reqs = _exchange([dist.P2POp(dist.isend, out.contiguous(), src)] if end > begin else [])
    for q in reqs:
        q.wait()
    if timed:
        stats.setdefault("gather", []).append((t1, mark()))
    return result


@torch.no_grad()
def sharded_enhance_u8(net, batch: Optional[torch.Tensor], device: torch.device, window: int = 128,
                       out: Optional[torch.Tensor] = None, src: int = 0,
                       stats: Optional[dict] = None) -> Optional[torch.Tensor]:
    """BASELINE configs[3]: a batch of uint8 BGR images (B,H,W,3) held by rank ``src`` (host, ideally
    pinned) is enhanced by all ranks: H2D on ``src`` -> NCCL scatter -> ``enhance_bgr_u8`` on every
    rank's shard -> NCCL gather -> (optionally) D2H into ``out`` on ``src``.  uint8 crosses PCIe and
    NVLink (3 bytes per pixel each way).  Returns the result on ``src`` (``out`` if given)."""
    from .imageio import enhance_bgr_u8
    res = sharded_forward(lambda t: enhance_bgr_u8(net, t, window=window), batch, device, src=src,
                          gather=True, stats=stats)
    if res is not None and out is not None:
        out.copy_(res, non_blocking=True)
        return out
    return res


class ShardedEnhancePipeline:
    """``sharded_enhance_u8`` for a stream of batches (BASELINE configs[3]: every batch of B uint8 images is
    held by rank ``src``): the upload + NCCL scatter of batch i+1 and the NCCL gather + download of batch i-1
    run on side streams -- and on two separate NCCL communicators, so that a scatter never queues behind the
    gather that waits for the forward -- while batch i is in the network.  Collective by construction: every
    rank calls ``submit`` / ``flush`` the same number of times.

        pipe = ShardedEnhancePipeline(net, device, window=128)
        for batch, out in batches:              # pinned (B,H,W,3) uint8 on rank src, None elsewhere
            pipe.submit(batch, out, shape=(B, H, W, 3))
        pipe.flush()                            # every ``out`` on rank src is complete
    """

    def __init__(self, net, device: torch.device, window: int = 128, src: int = 0, depth: int = 2):
        if not dist.is_initialized():
            raise RuntimeError("ShardedEnhancePipeline needs an initialised torch.distributed process group")
        self.net, self.device, self.window, self.src = net, torch.device(device), int(window), int(src)
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.depth = max(int(depth), 2)
        self.pg_up = dist.new_group(backend=dist.get_backend())        # scatter
        self.pg_down = dist.new_group(backend=dist.get_backend())      # gather
        self.up = torch.cuda.Stream(self.device)
        self.down = torch.cuda.Stream(self.device)
        self._stage = [None] * self.depth       # src: the whole batch on the device; others: this rank's slice
        self._result = [None] * self.depth      # src: gathered results
        self._stage_free = [None] * self.depth
        self._done = [None] * self.depth
        self._n = 0

    @torch.no_grad()
    def submit(self, batch: Optional[torch.Tensor], out: Optional[torch.Tensor], shape) -> None:
        from .imageio import enhance_bgr_u8
        B = int(shape[0])
        rest = tuple(int(v) for v in shape[1:])
        slot = self._n % self.depth
        self._n += 1
        begin, end = shard_range(B, self.rank, self.world)
        is_src = self.rank == self.src
        want = (B,) + rest if is_src else (end - begin,) + rest
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            if self._stage[slot] is None or tuple(self._stage[slot].shape) != want:
                self._stage[slot] = torch.empty(want, dtype=torch.uint8, device=self.device)
                if is_src:
                    self._result[slot] = torch.empty(want, dtype=torch.uint8, device=self.device)
                self.up.wait_stream(cur)
                self.down.wait_stream(cur)
            stage = self._stage[slot]
            # ---- upload + scatter on the `up` stream ---------------------------------------------------
            with torch.cuda.stream(self.up):
                if self._stage_free[slot] is not None:
                    self.up.wait_event(self._stage_free[slot])
                if is_src:
                    if batch is None or tuple(batch.shape) != want or batch.dtype != torch.uint8:
                        raise ValueError(f"rank {self.src} must pass the (B,H,W,3) uint8 batch")
                    stage.copy_(batch, non_blocking=True)
                    sends = []
                    for r in range(self.world):
                        b0, b1 = shard_range(B, r, self.world)
                        if r != self.src and b1 > b0:
                            sends.append(dist.P2POp(dist.isend, stage[b0:b1], r, group=self.pg_up))
                    reqs = _exchange(sends)
                else:
                    reqs = _exchange([dist.P2POp(dist.irecv, stage, self.src, group=self.pg_up)] if end > begin else [])
                for q in reqs:
                    q.wait()
                arrived = self.up.record_event()
            # ---- this rank's shard on the current stream ------------------------------------------------
            cur.wait_event(arrived)
            mine = stage[begin:end] if is_src else stage
            res = enhance_bgr_u8(self.net, mine, window=self.window) if end > begin else mine
            ready = cur.record_event()
            # the staging buffer is free once the forward has consumed it AND (on src) the sends have left
            self._stage_free[slot] = ready
            res.record_stream(self.down)
            # ---- gather + download on the `down` stream -------------------------------------------------
            with torch.cuda.stream(self.down):
                self.down.wait_event(ready)
                if self._done[slot] is not None:
                    self.down.wait_event(self._done[slot])
                if is_src:
                    result = self._result[slot]
                    recvs = []
                    for r in range(self.world):
                        b0, b1 = shard_range(B, r, self.world)
                        if r != self.src and b1 > b0:
                            recvs.append(dist.P2POp(dist.irecv, result[b0:b1], r, group=self.pg_down))
                    reqs = _exchange(recvs)
                    result[begin:end].copy_(res)
                    for q in reqs:
                        q.wait()
                    if out is not None:
                        out.copy_(result, non_blocking=True)
                else:
                    reqs = _exchange([dist.P2POp(dist.isend, res.contiguous(), self.src, group=self.pg_down)]
                                     if end > begin else [])
                    for q in reqs:
                        q.wait()
                self._done[slot] = self.down.record_event()

    def flush(self, block: bool = True) -> None:
        cur = torch.cuda.current_stream(self.device)
        for ev in self._done:
            if ev is not None:
                cur.wait_event(ev)
        if block:
            for ev in self._done:
                if ev is not None:
                    ev.synchronize()
