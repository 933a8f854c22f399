"""Multi-GPU data parallelism of the render operator (the reference is single-GPU).

One process per GPU.  Scene, BVH, sampler tables and material are replicated; the dispatch grid is
sharded: the image's 8-row tiles are dealt round-robin to the ranks (``zdr_params.shard_index /
shard_count``), so every pixel keeps its whole sample loop on one GPU and the forward image is
bit-identical to the single-GPU one.  Two collectives, both plain sums over NVLink (NCCL):

* forward:  ``all_reduce(image)``      — ranks hold disjoint rows, the rest is zero, so the sum is exact;
* backward: ``all_reduce(d_material)`` — every rank scatters the texture gradient of its own rows
  (16 MiB for a 1024^2 texture).

Second axis (``axis="samples"``), for images too small to tile over the ranks (light-stage sized jobs): every rank
renders ALL pixels but only its contiguous range of sample indices (``zdr_params.it0 / it1``).  Each rank's image is
then a partial mean (sum of its samples / spp); the all-reduce completes it.  The sum order differs from one GPU's,
so the image equals the single-GPU one to float accuracy instead of bit for bit.

The host logic is backend-agnostic so that it can be exercised on CPU with gloo (tests/): a backend
is a pair of callables ``forward(material, res, spp, seed, shard, its=None) -> image`` and
``backward(grad_output, d_material, material, res, spp, seed, shard, its=None)``; ``shard`` = (index, count) of the
tile sharding or None, ``its`` = (first, last+1) sample indices or None.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_of_rank(group=None) -> Tuple[int, int]:
    """(shard_index, shard_count) of this process; (0, 1) when torch.distributed is not initialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def rows_of_shard(height: int, shard: Tuple[int, int], row0: int = 0, tile: int = 8):
    """The image rows shard (index, count) renders: tiles i of `tile` rows with i % count == index."""
    index, count = shard
    rows = []
    for y in range(row0, height):
        if ((y - row0) // tile) % count == index:
            rows.append(y)
    return rows


def samples_of_shard(spp: int, shard: Tuple[int, int]) -> Tuple[int, int]:
    """The contiguous range of sample indices shard (index, count) renders; ranges tile [0, spp) and differ by <= 1."""
    index, count = shard
    return spp * index // count, spp * (index + 1) // count


class ShardedRenderOperator(torch.autograd.Function):
    """RenderOperator (render.py:201-223) over a process group."""

    @staticmethod
    def forward(ctx, material, fwd: Callable, bwd: Callable, group, res, spp, seed, axis="tiles"):
        if axis not in ("tiles", "samples"):
            raise ValueError(f"unknown sharding axis {axis!r}")
        shard = shard_of_rank(group)
        ctx.save_for_backward(material)
        ctx.bwd, ctx.group, ctx.shard, ctx.args = bwd, group, shard, (res, spp, seed)
        ctx.its = samples_of_shard(spp, shard) if (axis == "samples" and shard[1] > 1) else None
        if ctx.its is not None:
            image = fwd(material.detach(), res, spp, seed, None, its=ctx.its)
        else:
            image = fwd(material.detach(), res, spp, seed, shard)
        if shard[1] > 1:
            dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)
            image[..., 3] = 1.0          # alpha of the rows other ranks rendered (sum of one 1 and zeros already)
        return image

    @staticmethod
    def backward(ctx, grad_output):
        material, = ctx.saved_tensors
        res, spp, seed = ctx.args
        d_material = torch.zeros(material.shape, dtype=material.dtype, device=material.device)
        if ctx.its is not None:
            ctx.bwd(grad_output, d_material, material.detach(), res, spp, seed, None, its=ctx.its)
        else:
            ctx.bwd(grad_output, d_material, material.detach(), res, spp, seed, ctx.shard)
        if ctx.shard[1] > 1:
            dist.all_reduce(d_material, op=dist.ReduceOp.SUM, group=ctx.group)
        return d_material, None, None, None, None, None, None, None


def sharded_render(material, fwd: Callable, bwd: Callable, *, res, spp, seed=0, group: Optional[object] = None,
                   axis: str = "tiles"):
    return ShardedRenderOperator.apply(material, fwd, bwd, group, res, spp, seed, axis)
