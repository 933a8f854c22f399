"""Multi-GPU data parallelism of the render operator (the reference is single-GPU).

One process per GPU.  Scene, BVH, sampler tables and material are replicated; the dispatch grid is
sharded: the image's 8-row tiles are dealt round-robin to the ranks (``zdr_params.shard_index /
shard_count``), so every pixel keeps its whole sample loop on one GPU and the forward image is
bit-identical to the single-GPU one.  Two collectives, both plain sums over NVLink (NCCL):

* forward:  ``all_gather`` of the rows each rank owns (tile axis: 1/N of the image per rank instead of an all-reduce over
  the whole image — ranks hold disjoint rows, nothing has to be summed);
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


_row_index_cache = {}


def _row_indices(height: int, count: int, device):
    """Per rank: the image rows it owns, padded to the longest list (index tensors, cached)."""
    key = (height, count, str(device))
    if key not in _row_index_cache:
        rows = [rows_of_shard(height, (r, count)) for r in range(count)]
        n_max = max(len(r) for r in rows)
        own = [torch.tensor(r + [r[-1] if r else 0] * (n_max - len(r)), dtype=torch.long, device=device) for r in rows]
        dst = torch.tensor([y for r in rows for y in r], dtype=torch.long, device=device)
        src = torch.tensor([k * n_max + i for k, r in enumerate(rows) for i in range(len(r))], dtype=torch.long, device=device)
        _row_index_cache[key] = (own, dst, src, n_max)
    return _row_index_cache[key]


def gather_tiles(image: torch.Tensor, shard: Tuple[int, int], group=None) -> torch.Tensor:
    """Completes a tile-sharded image: every rank contributes only the rows it rendered (1/N of the image) and receives
    the others' — an all-gather of disjoint rows, where the first version all-reduced the whole image (rows of the
    other ranks as zeros).  Bit-identical to the single-GPU image: nothing is added."""
    index, count = shard
    H = image.shape[0]
    own, dst, src, n_max = _row_indices(H, count, image.device)
    mine = image.index_select(0, own[index]).contiguous()
    parts = [torch.empty_like(mine) for _ in range(count)]
    dist.all_gather(parts, mine, group=group)
    full = torch.empty_like(image)
    full.index_copy_(0, dst, torch.cat(parts, 0).index_select(0, src))
    return full


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
            if ctx.its is None:
                image = gather_tiles(image, shard, group)            # disjoint rows: gather, nothing to sum
            else:
                dist.all_reduce(image, op=dist.ReduceOp.SUM, group=group)      # partial means over sample ranges
                image[..., 3] = 1.0
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
