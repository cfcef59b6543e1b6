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

The NCCL route goes through ``torch.distributed`` (the same code runs on gloo for CPU tests, where ``reduce_scatter``
is emulated with ``all_reduce`` + slice because gloo lacks it).  On one NVLink domain :class:`PeerPixelExchange` replaces
both collectives with the library's own kernels over peer memory (``csrc/msda_peer.cu``): the all-gather pulls the
peers' shards straight into the pyramid the forward kernel reads, and the backward kernel accumulates its partial
``grad_img`` directly in the symmetric buffer the peers' reduce-scatter kernels read -- no staging copies, one launch per
collective, flags instead of host-side synchronisation.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, kernels
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


# ---------------------------------------------------------------------------------------------------------------------
# NVLink peer-memory route (one process per GPU of one node; torch symmetric memory provides the peer mappings)
# ---------------------------------------------------------------------------------------------------------------------
class PeerPixelExchange:
    """Buffers and flags for query-sharded MSDA over peer memory, for a fixed problem size (fp32).

    Allocates, in symmetric memory: the gathered pyramid ``[B, world * chunk, H, D]`` the peers PUSH their shards into
    in the all-gather, this rank's partial ``grad_img`` (same shape) the peers pull from in the reduce-scatter, and a
    flag block.  Create it once (collective call: every rank of ``group``) and
    reuse it for every step; the ranks must issue the same sequence of :func:`peer_query_sharded_msda` calls.  The
    collectives keep their call counts on the device, so a whole step can be captured into a CUDA graph
    (``torch.cuda.graph``) and replayed -- every rank replaying the same number of times."""

    def __init__(self, batch: int, num_pixels: int, heads: int, channels: int, group=None,
                 device: Optional[torch.device] = None):
        import torch.distributed._symmetric_memory as symm
        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 16:
            raise ValueError("PeerPixelExchange: at most 16 ranks (one NVLink domain)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device, self.num_pixels = dev, num_pixels
        self.chunk = pixel_chunk(num_pixels, self.world)
        self.shape_shard = (batch, self.chunk, heads, channels)
        self.shape_full = (batch, self.world * self.chunk, heads, channels)
        if (self.chunk * heads * channels * 4) % 16:
            raise ValueError("PeerPixelExchange: a pixel chunk must be a multiple of 16 bytes")
        name = self.group.group_name
        self.full = symm.empty(self.shape_full, dtype=torch.float32, device=dev)
        self.partial = symm.empty(self.shape_full, dtype=torch.float32, device=dev)
        self.flags = symm.empty((4 * 16,), dtype=torch.int32, device=dev)
        self.flags.zero_()
        self._handles = [symm.rendezvous(t, name) for t in (self.full, self.partial, self.flags)]
        self.counters = torch.zeros(8, dtype=torch.int32, device=dev)   # CTA arrival counters + the two call counts
        arr = ctypes.c_void_p * self.world
        self._ptrs = [arr(*[int(p) for p in h.buffer_ptrs]) for h in self._handles]   # keep the host arrays alive
        self.ctx = _lib.MsdaPeerCtx(self.world, self.rank,
                                    ctypes.cast(self._ptrs[0], ctypes.POINTER(ctypes.c_void_p)),
                                    ctypes.cast(self._ptrs[1], ctypes.POINTER(ctypes.c_void_p)),
                                    ctypes.cast(self._ptrs[2], ctypes.POINTER(ctypes.c_void_p)),
                                    ctypes.c_void_p(self.counters.data_ptr()))
        torch.cuda.synchronize(dev)
        self._handles[2].barrier(channel=0)      # every rank's flags are zero before anybody signals
        torch.cuda.synchronize(dev)

    def all_gather(self, shard: torch.Tensor) -> torch.Tensor:
        """[B, chunk, H, D] pixel shard of this rank -> the padded pyramid [B, world * chunk, H, D] (self.full)."""
        if tuple(shard.shape) != self.shape_shard or shard.dtype != torch.float32 or not shard.is_contiguous():
            raise ValueError(f"PeerPixelExchange.all_gather: expected a contiguous fp32 {self.shape_shard} shard")
        per_image = self.chunk * self.shape_shard[2] * self.shape_shard[3] * 4
        rc = _lib.get_lib().msda_peer_all_gather(
            shard.data_ptr(), ctypes.byref(self.ctx), self.shape_shard[0], per_image,
            torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "msda_peer_all_gather")
        return self.full

    def reduce_scatter(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Sum over the ranks of their partial grad_img (self.partial), this rank's pixel chunk: [B, chunk, H, D]."""
        if out is None:
            out = torch.empty(self.shape_shard, dtype=torch.float32, device=self.device)
        per_image = self.chunk * self.shape_shard[2] * self.shape_shard[3]
        rc = _lib.get_lib().msda_peer_reduce_scatter(
            out.data_ptr(), ctypes.byref(self.ctx), self.shape_shard[0], per_image,
            torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "msda_peer_reduce_scatter")
        return out


class _PeerQueryShardedMsda(torch.autograd.Function):
    """all-gather (peer pull) + forward kernel; backward kernel into the symmetric partial + reduce-scatter (peer pull)."""

    @staticmethod
    def forward(ctx, img_shard, img_shapes, pts, aw, padding_mode, align_corners, ex: PeerPixelExchange):
        full = ex.all_gather(img_shard.contiguous())
        out = kernels.b200_multi_scale_deformable_attention_fwd(full, img_shapes, pts, aw, padding_mode, align_corners)
        ctx.save_for_backward(img_shapes, pts, aw)
        ctx.ex, ctx.mode = ex, (padding_mode, align_corners)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        img_shapes, pts, aw = ctx.saved_tensors
        ex = ctx.ex
        need_img, _, need_pts, need_aw = ctx.needs_input_grad[:4]
        # the gathered pyramid is still in ex.full (the forward of this step wrote it; nothing else touches it)
        _, gpts, gaw = kernels.b200_multi_scale_deformable_attention_bwd(
            grad_out, ex.full, img_shapes, pts, aw, ctx.mode[0], ctx.mode[1],
            needs=(True, need_pts, need_aw), deterministic=False, grads=(ex.partial, None, None))
        gshard = ex.reduce_scatter()
        return (gshard if need_img else None), None, gpts, gaw, None, None, None


def peer_query_sharded_msda(exchange: PeerPixelExchange, img_shard, img_shapes, sampling_points_local,
                            attention_weights_local, padding_mode: str, align_corners: bool) -> torch.Tensor:
    """:func:`query_sharded_msda` over NVLink peer memory (fp32, CUDA): same inputs / outputs / gradients."""
    return _PeerQueryShardedMsda.apply(img_shard, img_shapes, sampling_points_local, attention_weights_local,
                                       padding_mode, align_corners, exchange)
