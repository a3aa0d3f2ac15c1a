"""Multi-GPU pieces of the hot path (new work: the reference has no distributed code,
SURVEY.md section 2 'Parallelism strategies').

* Data-parallel calibration: every rank runs ``estimate_ranges`` over its shard of the batches;
  at the end of the block the running ranges of all quantizers are packed into ONE buffer and
  all-reduced once with MIN (mins and negated maxes travel together), then every rank recomputes
  identical (scale, offset).  The payload is a few KB, so this is latency-bound: one collective,
  not one per quantizer ('Distributed communication backend').
* Layer-sharded weight quantization: independent units, unit ``i`` belongs to rank ``i mod world``;
  no collective on the data path.
Works with any ``torch.distributed`` backend: NCCL over NVLink on the B200 box, gloo in CPU tests.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def _active(group) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def pack_ranges(ranges: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> Tuple[torch.Tensor, List[int]]:
    """[min_0 | -max_0 | min_1 | -max_1 ...] in float32 (exact for fp32/bf16/fp16 ranges)."""
    sizes = [mn.numel() for mn, _ in ranges]
    parts: List[torch.Tensor] = []
    for mn, mx in ranges:
        parts.append(mn.reshape(-1).float())
        parts.append(-mx.reshape(-1).float())
    return (torch.cat(parts) if parts else torch.empty(0)), sizes


def unpack_ranges_(packed: torch.Tensor, ranges: Sequence[Tuple[torch.Tensor, torch.Tensor]]) -> None:
    pos = 0
    for mn, mx in ranges:
        n = mn.numel()
        mn.copy_(packed[pos:pos + n].reshape(mn.shape).to(mn.dtype))
        mx.copy_((-packed[pos + n:pos + 2 * n]).reshape(mx.shape).to(mx.dtype))
        pos += 2 * n


def all_reduce_ranges(ranges: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                      flags: Optional[Sequence[Optional[torch.Tensor]]] = None, group=None) -> None:
    """In-place MIN/MAX all-reduce of running ranges (+ MAX of the +-inf flags)."""
    if not ranges or not _active(group):
        return
    packed, _ = pack_ranges(ranges)
    dist.all_reduce(packed, op=dist.ReduceOp.MIN, group=group)
    unpack_ranges_(packed, ranges)
    live = [f for f in (flags or []) if f is not None]
    if live:
        stacked = torch.stack([f.reshape(()) for f in live])
        dist.all_reduce(stacked, op=dist.ReduceOp.MAX, group=group)
        for f, v in zip(live, stacked):
            f.copy_(v.reshape(f.shape))


def all_reduce_minmax_buffers(buffers: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                              flags: Optional[Sequence[torch.Tensor]] = None, group=None) -> None:
    """In-place all-reduce of contiguous (min_buffer, max_buffer) pairs: one MIN and one MAX collective
    per pair (the estimator keeps all running ranges of a dtype in one pair), plus one MAX for the flags."""
    if not _active(group):
        return
    for mn, mx in buffers:
        dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    for f in flags or []:
        dist.all_reduce(f, op=dist.ReduceOp.MAX, group=group)


def shard_units(num_units: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> List[int]:
    """Indices of the independent units (layers, weight tensors) this rank owns: ``i % world == rank``."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    return list(range(rank, num_units, world_size))
