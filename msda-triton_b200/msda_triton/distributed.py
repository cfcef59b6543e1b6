"""Multi-GPU partitioning of the MSDA path (one process per GPU, ``torch.distributed``; NCCL over NVLink on B200).

The reference has no distributed code at all; this module is the B200-box addition BASELINE.json asks
for.  Two partitionings, both built on the fact that every output row (b, q, h) is independent and only ``grad_img``
couples rows of the same image (``/root/reference/src/msda_triton/kernels.py:16-21`` and ``:549-553``):

* **Batch sharding** (default): rank r owns ``b in [r*B/n, (r+1)*B/n)`` of all tensors.  No collective in forward or
  backward -- use :func:`shard_batch` and call the operator as usual.
* **Query sharding** (large encoder query sets, or B < n): the ranks of a group share an image.  Each rank holds a
  pixel shard of ``img`` and a query shard of ``sampling_points`` / ``attention_weights``.  Forward all-gathers the
  pixel shards (:func:`gather_pixels`), backward is the exact transpose: ONE reduce-scatter (sum) of ``grad_img``.
  ``grad_sampling_points`` / ``grad_attention_weights`` are query-local and need nothing.

All collectives go through ``torch.distributed`` (NCCL on GPUs; the same code runs on gloo for CPU tests, where
``reduce_scatter`` is emulated with ``all_reduce`` + slice because gloo lacks it).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .frontend import multiscale_deformable_attention


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) of `total` items for `rank` (first `total % world` ranks get one more)."""
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's images of a batch-leading tensor (no copy)."""
    b, e = shard_range(t.shape[0], rank, world)
    return t[b:e]


def shard_queries(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's queries of a [B, Q, ...] tensor (no copy)."""
    b, e = shard_range(t.shape[1], rank, world)
    return t[:, b:e]


def pixel_chunk(num_pixels: int, world: int) -> int:
    """Pixels per rank when the pyramid is cut into `world` equal chunks (the last one zero-padded)."""
    return (num_pixels + world - 1) // world


def shard_pixels(img: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's equal-size pixel chunk [B, chunk, H, D] of a full pyramid [B, Npix, H, D] (zero-padded tail)."""
    npix = img.shape[1]
    chunk = pixel_chunk(npix, world)
    b, e = rank * chunk, min((rank + 1) * chunk, npix)
    out = img.new_zeros((img.shape[0], chunk) + tuple(img.shape[2:]))
    if e > b:
        out[:, : e - b] = img[:, b:e]
    return out


def _backend(group) -> str:
    return dist.get_backend(group)


class _GatherPixels(torch.autograd.Function):
    """all-gather of pixel shards along dim 1; backward = reduce-scatter (sum) of the gradient.

    Zero-copy on both sides: the gathered pyramid is laid out ``[B, world * chunk, H, D]`` (image-major, the chunks of
    one image back to back = pixel order), so image b is ONE contiguous all-gather destination / reduce-scatter source
    and no transposed or zero-padded staging copy of the 45 MB tensors is made (round 1 made four).  The pyramid keeps
    its ``world * chunk - Npix`` padding rows; the kernels accept an image with more rows than the level table names."""

    @staticmethod
    def forward(ctx, shard: torch.Tensor, group):
        world = dist.get_world_size(group)
        ctx.group, ctx.world = group, world
        shard = shard.contiguous()
        B, chunk = shard.shape[0], shard.shape[1]
        full = shard.new_empty((B, world * chunk) + tuple(shard.shape[2:]))
        for b in range(B):
            dist.all_gather_into_tensor(full[b], shard[b], group=group)
        return full

    @staticmethod
    def backward(ctx, grad_full: torch.Tensor):
        world, group = ctx.world, ctx.group
        grad_full = grad_full.contiguous()
        B, chunk = grad_full.shape[0], grad_full.shape[1] // world
        out = grad_full.new_empty((B, chunk) + tuple(grad_full.shape[2:]))
        if _backend(group) == "gloo":
            dist.all_reduce(grad_full, group=group)                           # gloo has no reduce_scatter
            r = dist.get_rank(group)
            out.copy_(grad_full[:, r * chunk:(r + 1) * chunk])
        else:
            for b in range(B):
                dist.reduce_scatter_tensor(out[b], grad_full[b], op=dist.ReduceOp.SUM, group=group)
        return out, None


def gather_pixels(img_shard: torch.Tensor, num_pixels: int, group=None, keep_padding: bool = False) -> torch.Tensor:
    """[B, chunk, H, D] pixel shards -> the full pyramid on every rank of `group` (differentiable: the backward pass
    reduce-scatters grad_img over NVLink, message = B * world * chunk * H * D * elem_size bytes).

    keep_padding=False: ``[B, Npix, H, D]`` (a view of the padded buffer when world * chunk > Npix).
    keep_padding=True : ``[B, world * chunk, H, D]`` contiguous -- what :func:`query_sharded_msda` hands to the CUDA
    kernels so that nothing is copied."""
    full = _GatherPixels.apply(img_shard, group)
    return full if keep_padding else full[:, :num_pixels]


def query_sharded_msda(
    img_shard: torch.Tensor,
    num_pixels: int,
    img_shapes: torch.Tensor,
    sampling_points_local: torch.Tensor,
    attention_weights_local: torch.Tensor,
    padding_mode: str,
    align_corners: bool,
    group=None,
) -> torch.Tensor:
    """MSDA for the ranks of `group` that share images: pixel-sharded value in, query-sharded output out.

    forward : all-gather(img shards) -> local queries sample the full pyramid        (no other collective)
    backward: local partial grad_img -> reduce-scatter(sum) back to pixel shards     (the only backward collective)
    """
    on_gpu = img_shard.is_cuda
    # CUDA: the kernels read the padded gather buffer in place and the backward writes grad_img straight into the
    # reduce-scatter source; the torch route (CPU tensors) splits the pyramid by level sizes and gets the exact view
    img_full = gather_pixels(img_shard, num_pixels, group, keep_padding=on_gpu)
    return multiscale_deformable_attention(
        img_full, img_shapes, sampling_points_local, attention_weights_local, padding_mode, align_corners)


def all_reduce_grad_img_(grad_img: torch.Tensor, group=None) -> torch.Tensor:
    """For the replicated-``img`` variant of query sharding (every rank keeps the full value and wants the full
    gradient): in-place sum of the per-rank partial grad_img."""
    dist.all_reduce(grad_img, op=dist.ReduceOp.SUM, group=group)
    return grad_img
