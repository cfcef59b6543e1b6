"""Public surface of the package: same names, signatures, layouts and state_dict keys as rziga/msda-triton
(``/root/reference/src/msda_triton/frontend.py``), with the CUDA route backed by libmsda_b200.so.

Name map (reference line -> here):
  multiscale_deformable_attention          frontend.py:145-172 -> multiscale_deformable_attention (device dispatch)
  triton_multiscale_deformable_attention   frontend.py:71-105  -> b200_multiscale_deformable_attention (+ alias)
  _TritonMultiscaleDeformableAttentionFunction  frontend.py:108-142 -> _B200MsdaFunction
  native_multiscale_deformable_attention   frontend.py:15-68   -> native_multiscale_deformable_attention (CPU tensors)
  MultiscaleDeformableAttention            frontend.py:175-292 -> MultiscaleDeformableAttention

Dispatch rule.  The reference tries its GPU kernel and silently re-runs ANY failure on the slow torch route
(``except Exception``, frontend.py:170).  Here the route is chosen by the device of ``img`` and nothing is caught:
CUDA tensors always run the hand-written sm_100a kernels (errors propagate; a missing library raises), CPU tensors
run the torch route the reference documents for ``device="cpu"`` (README.md:132).
"""
from __future__ import annotations

import os
from typing import Literal

import torch
import torch.nn.functional as F
from torch import nn
from torch.amp import custom_bwd, custom_fwd
from torch.autograd.function import once_differentiable

from . import kernels

PaddingMode = Literal["border", "zeros"]


def _fused_module_enabled() -> bool:
    return os.environ.get("MSDA_B200_FUSED_MODULE", "1") != "0"


def _fused_value_proj_enabled() -> bool:
    return os.environ.get("MSDA_B200_FUSED_VALUE_PROJ", "1") != "0"

# dtypes of the CUDA route: the reference's three (frontend.py:84) plus bf16, which its Triton helper rejects
# (kernels.py:40-41) and therefore sends down the torch route.
CUDA_DTYPES = (torch.float16, torch.bfloat16, torch.float32, torch.float64)


# ---------------------------------------------------------------------------------------------------------------------
# CPU-tensor route
# ---------------------------------------------------------------------------------------------------------------------
def native_multiscale_deformable_attention(
    img: torch.Tensor,
    img_shapes: torch.Tensor,
    sampling_points: torch.Tensor,
    attention_weights: torch.Tensor,
    padding_mode: PaddingMode,
    align_corners: bool,
) -> torch.Tensor:
    """Torch-operator route (``F.grid_sample`` per level), differentiable through autograd.

    Serves CPU tensors.  It is never entered for CUDA inputs by :func:`multiscale_deformable_attention`; calling it
    directly with CUDA tensors works (the reference's tests and benchmark do, as the "torch" comparison line).
    """
    batch, _, heads, channels = img.shape
    queries, points = sampling_points.shape[1], sampling_points.shape[4]
    level_hw = [(int(h), int(w)) for h, w in img_shapes.tolist()]
    grids = sampling_points.mul(2).sub(1)                     # grid_sample wants [-1, 1]
    per_level = img.split([h * w for h, w in level_hw], dim=1)
    sampled = []
    for lvl, (h, w) in enumerate(level_hw):
        # [B, h*w, H, C] -> [B*H, C, h, w]
        feat = per_level[lvl].permute(0, 2, 3, 1).reshape(batch * heads, channels, h, w)
        # [B, N, H, P, 2] -> [B*H, N, P, 2]
        grid = grids[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(batch * heads, queries, points, 2)
        val = F.grid_sample(feat, grid, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)
        # [B*H, C, N, P] -> [B, N, H, P, C]
        sampled.append(val.reshape(batch, heads, channels, queries, points).permute(0, 3, 1, 4, 2))
    sampled = torch.stack(sampled, dim=3)                     # [B, N, H, L, P, C]
    return (attention_weights.unsqueeze(-1) * sampled).sum(dim=(3, 4))


# ---------------------------------------------------------------------------------------------------------------------
# CUDA route
# ---------------------------------------------------------------------------------------------------------------------
class _B200MsdaFunction(torch.autograd.Function):
    """Autograd boundary.  Contract identical to the reference's Function (frontend.py:108-142): under autocast the
    float inputs are cast to fp32 and the op runs in fp32; the four input tensors are saved (no intermediates, so
    extra memory = outputs only); backward is once-differentiable and returns grads for (img, None, points, weights,
    None, None).  Additionally ``ctx.needs_input_grad`` is honoured so unneeded gradients are not computed."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners):
        ctx.save_for_backward(img, img_shapes, sampling_points, attention_weights)
        ctx.padding_mode = padding_mode
        ctx.align_corners = align_corners
        return kernels.b200_multi_scale_deformable_attention_fwd(
            img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners)

    @staticmethod
    @once_differentiable
    @custom_bwd(device_type="cuda")
    def backward(ctx, out_grad):
        img, img_shapes, sampling_points, attention_weights = ctx.saved_tensors
        needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[2], ctx.needs_input_grad[3])
        img_grad, points_grad, weights_grad = kernels.b200_multi_scale_deformable_attention_bwd(
            out_grad, img, img_shapes, sampling_points, attention_weights, ctx.padding_mode, ctx.align_corners,
            needs=needs)
        return img_grad, None, points_grad, weights_grad, None, None


class _B200ModuleCoreFunction(torch.autograd.Function):
    """The module's core between its projections (frontend.py:253-289 of the reference) as ONE kernel each way:
    softmax over L*K, sampling-point arithmetic and the MSDA operator, without materialising sampling_points /
    attention_weights.  Same AMP contract as the operator (fp32 under autocast)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, value, img_shapes, projection, reference_points, padding_mode, align_corners):
        ctx.save_for_backward(value, img_shapes, projection, reference_points)
        ctx.padding_mode = padding_mode
        ctx.align_corners = align_corners
        return kernels.b200_module_core_fwd(value, img_shapes, projection, reference_points, padding_mode, align_corners)

    @staticmethod
    @once_differentiable
    @custom_bwd(device_type="cuda")
    def backward(ctx, out_grad):
        value, img_shapes, projection, reference_points = ctx.saved_tensors
        needs = (ctx.needs_input_grad[0], ctx.needs_input_grad[2], ctx.needs_input_grad[3])
        gvalue, gproj, gref = kernels.b200_module_core_bwd(
            out_grad, value, img_shapes, projection, reference_points, ctx.padding_mode, ctx.align_corners, needs=needs)
        return gvalue, None, gproj, gref, None, None


class _B200ValueProjCoreFunction(torch.autograd.Function):
    """``img_input_proj`` (frontend.py:259 of the reference: ``value = Linear(img)``) AND the module core as one autograd
    node, for 16-bit parameters.  Forward: the same cuBLAS GEMM torch would run, then the fused core kernel.  Backward:
    the core kernel's rounding pass (fp32 accumulation image -> 16-bit grad_value) also returns the column sums of what
    it rounds, which ARE the bias gradient of the projection -- torch computes them with a separate reduction kernel
    over the 16-bit tensor (~100 us for the B=8 x 22 223-pixel decoder pyramid) -- and the two remaining gradients are
    the two GEMMs autograd would issue.  Used outside autocast only (under autocast the core runs in fp32 and has no
    rounding pass)."""

    @staticmethod
    def forward(ctx, img, weight, bias, img_shapes, projection, reference_points, heads, padding_mode, align_corners):
        batch, num_pixels, _ = img.shape
        hidden = weight.shape[0]
        value = torch.nn.functional.linear(img, weight, bias).view(batch, num_pixels, heads, hidden // heads)
        ctx.save_for_backward(img, weight, value, img_shapes, projection, reference_points)
        ctx.padding_mode = padding_mode
        ctx.align_corners = align_corners
        return kernels.b200_module_core_fwd(value, img_shapes, projection, reference_points, padding_mode, align_corners)

    @staticmethod
    @once_differentiable
    def backward(ctx, out_grad):
        img, weight, value, img_shapes, projection, reference_points = ctx.saved_tensors
        need_img, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        need_value = need_img or need_w or need_b
        res = kernels.b200_module_core_bwd(
            out_grad, value, img_shapes, projection, reference_points, ctx.padding_mode, ctx.align_corners,
            needs=(need_value, ctx.needs_input_grad[4], ctx.needs_input_grad[5]), value_colsum=need_b)
        gvalue, gproj, gref = res[:3]
        gimg = gweight = gbias = None
        if need_value:
            g2 = gvalue.view(-1, weight.shape[0])
            if need_img:
                gimg = torch.matmul(g2, weight).view(img.shape)
            if need_w:
                gweight = torch.matmul(g2.t(), img.reshape(-1, img.shape[-1]))
            if need_b:
                gbias = res[3].reshape(-1).to(weight.dtype)
        return gimg, gweight, gbias, None, gproj, gref, None, None, None


def fused_value_proj_core(img, weight, bias, img_shapes, projection, reference_points, heads: int,
                          padding_mode: PaddingMode, align_corners: bool) -> torch.Tensor:
    """``core(Linear(img; weight, bias).view(B, I, heads, C), ...)`` with the bias gradient taken from the core's own
    rounding pass; returns ``[B, N, H, C]``.  16-bit dtypes, see kernels.module_value_colsum_supported."""
    return _B200ValueProjCoreFunction.apply(img, weight, bias, img_shapes, projection, reference_points, int(heads),
                                            padding_mode, bool(align_corners))


def fused_module_core(value, img_shapes, projection, reference_points, padding_mode: PaddingMode,
                      align_corners: bool) -> torch.Tensor:
    """``projection`` is the query projection viewed as ``[B, N, H, L, P, 3]`` (offset x, offset y, attention logit),
    ``value`` the projected pyramid ``[B, I, H, C]``; returns ``[B, N, H, C]``."""
    return _B200ModuleCoreFunction.apply(value, img_shapes, projection, reference_points, padding_mode,
                                         bool(align_corners))


def b200_multiscale_deformable_attention(
    img: torch.Tensor,
    img_shapes: torch.Tensor,
    sampling_points: torch.Tensor,
    attention_weights: torch.Tensor,
    padding_mode: PaddingMode,
    align_corners: bool,
) -> torch.Tensor:
    """CUDA route: validation (same two ``ValueError`` classes as frontend.py:84-95) then the autograd Function."""
    dtype, device = img.dtype, img.device
    uniform = (dtype in CUDA_DTYPES and sampling_points.dtype == dtype and attention_weights.dtype == dtype
               and device.type == "cuda" and img_shapes.device == device and sampling_points.device == device
               and attention_weights.device == device)
    if not uniform:   # slow path: find out what to complain about, or promote mixed dtypes
        for name, tensor in (("img", img), ("sampling_points", sampling_points),
                             ("attention_weights", attention_weights)):
            if tensor.dtype not in CUDA_DTYPES:
                raise ValueError(f"Dtype of `{name}` should be in {list(CUDA_DTYPES)}, but got {tensor.dtype}.")
        devices = [t.device for t in (img, img_shapes, sampling_points, attention_weights)]
        if any(d.type != "cuda" for d in devices):
            raise ValueError(f"Expected all inputs to be on gpu, but got {devices}.")
        if len({d.index for d in devices}) != 1:
            raise ValueError(f"Expected all inputs to be on the same gpu, but got {devices}.")
        if not torch.is_autocast_enabled("cuda"):
            # the kernels take one storage dtype; mixed inputs are promoted (under autocast custom_fwd casts to fp32)
            common = torch.promote_types(torch.promote_types(img.dtype, sampling_points.dtype),
                                         attention_weights.dtype)
            img, sampling_points, attention_weights = (t.to(common) for t in (img, sampling_points, attention_weights))
    if padding_mode not in ("border", "zeros"):
        raise ValueError(f"`padding_mode` should be 'border' or 'zeros', but got {padding_mode!r}.")
    if torch.compiler.is_compiling():
        # traced programs use the torch.library custom op (fake kernels + autograd formula, no graph break)
        from .ops import multiscale_deformable_attention_op
        if torch.is_autocast_enabled("cuda"):
            img, sampling_points, attention_weights = (t.float() for t in (img, sampling_points, attention_weights))
        return multiscale_deformable_attention_op(
            img, img_shapes, sampling_points, attention_weights, padding_mode, bool(align_corners))
    return _B200MsdaFunction.apply(img, img_shapes, sampling_points, attention_weights, padding_mode, bool(align_corners))


# The reference's tests and benchmark import the CUDA route under this name (tests/test_msda.py:8-12).
triton_multiscale_deformable_attention = b200_multiscale_deformable_attention


def multiscale_deformable_attention(
    img: torch.Tensor,
    img_shapes: torch.Tensor,
    sampling_points: torch.Tensor,
    attention_weights: torch.Tensor,
    padding_mode: PaddingMode,
    align_corners: bool,
) -> torch.Tensor:
    """Multiscale deformable attention (Deformable DETR, arXiv:2010.04159), differentiable in ``img``,
    ``sampling_points`` and ``attention_weights``.

    Shapes (B batch, I pyramid pixels, H heads, C channels per head, N queries, L levels, P points per level):

    * ``img``                ``[B, I, H, C]`` -- the levels of the feature pyramid flattened and concatenated along
      ``I = sum(h_l * w_l)``.
    * ``img_shapes``         ``[L, 2]`` integer tensor, one ``(height, width)`` row per level, same order as in ``img``.
    * ``sampling_points``    ``[B, N, H, L, P, 2]`` -- ``(x, y)`` in units of the level size: ``(0, 0)`` is the top-left
      and ``(1, 1)`` the bottom-right corner of every level; values outside ``[0, 1]`` are handled by ``padding_mode``.
    * ``attention_weights``  ``[B, N, H, L, P]``.
    * ``padding_mode``       ``"border"`` clamps out-of-range samples to the nearest pixel, ``"zeros"`` makes the
      out-of-range bilinear corners contribute 0.
    * ``align_corners``      same meaning as in ``torch.nn.functional.grid_sample``.

    Returns ``[B, N, H, C]``: for every (batch, query, head) the attention-weighted sum of the ``L * P`` bilinear samples.

    CUDA tensors run the hand-written sm_100a kernels; CPU tensors run the torch route.
    """
    if img.device.type == "cuda":
        if img_shapes.device != img.device:
            img_shapes = img_shapes.to(img.device, non_blocking=True)   # 16*L bytes; the kernels read it on device
        return b200_multiscale_deformable_attention(
            img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners)
    return native_multiscale_deformable_attention(
        img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners)


# ---------------------------------------------------------------------------------------------------------------------
# Module
# ---------------------------------------------------------------------------------------------------------------------
class MultiscaleDeformableAttention(nn.Module):
    """Deformable-attention layer: query / pyramid / output projections around the MSDA operator (Figure 2 of
    arXiv:2010.04159).

    The constructor signature and the parameter names (``img_input_proj``, ``query_input_proj``,
    ``query_output_proj``) are those of the reference module (frontend.py:199-223), so its checkpoints load with
    ``strict=True``.

    ``emb_dim`` is the width of the inputs and of the output, ``hidden_dim`` the width the pyramid is projected to
    (``hidden_dim // num_heads`` channels per head; a ``ValueError`` is raised when it is not a multiple of
    ``num_heads``), ``num_levels`` / ``num_points`` the number of pyramid levels and of sampling points per level,
    ``padding_mode`` / ``align_corners`` as in :func:`multiscale_deformable_attention`.
    """

    def __init__(
        self,
        emb_dim: int,
        hidden_dim: int,
        num_levels: int,
        num_heads: int,
        num_points: int,
        padding_mode: PaddingMode,
        align_corners: bool,
    ):
        super().__init__()
        if hidden_dim % num_heads != 0:
            raise ValueError(
                f"Hidden dimension ({hidden_dim=}) should be divisible by number of heads ({num_heads=}).")
        self.num_levels = num_levels
        self.num_heads = num_heads
        self.num_points = num_points
        self.hidden_dim = hidden_dim
        self.padding_mode = padding_mode
        self.align_corners = align_corners
        self.img_input_proj = nn.Linear(emb_dim, hidden_dim)
        # per (head, level, point): 2 offset coordinates + 1 attention logit
        self.query_input_proj = nn.Linear(emb_dim, num_heads * num_levels * num_points * 3)
        self.query_output_proj = nn.Linear(hidden_dim, emb_dim)

    def forward(
        self,
        img: torch.Tensor,
        img_shapes: torch.Tensor,
        queries: torch.Tensor,
        reference_points: torch.Tensor,
    ) -> torch.Tensor:
        """``img`` ``[B, I, emb_dim]`` is the flattened pyramid, ``img_shapes`` ``[L, 2]`` its ``(height, width)`` rows,
        ``queries`` ``[B, N, emb_dim]``; ``reference_points`` is ``[B, N, 2]`` (an ``(x, y)`` anchor per query) or
        ``[B, N, 4]`` (a ``(cx, cy, w, h)`` box per query), normalised like the sampling points.  Returns
        ``[B, N, emb_dim]``."""
        batch, num_pixels, _ = img.shape
        num_queries = queries.shape[1]
        heads, levels, points = self.num_heads, self.num_levels, self.num_points

        # offsets and logits come out of ONE projection, interleaved as (..., point, 3) (frontend.py:253-257)
        projected = self.query_input_proj(queries).reshape(batch, num_queries, heads, levels, points, 3)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError(
                f"`reference_points` should have the last dim either 2 or 4, but got {reference_points.shape[-1]}.")

        # 16-bit parameters on CUDA, training: the value projection and the core as ONE autograd node, so that the bias
        # gradient of img_input_proj comes out of the core's rounding pass instead of a reduction kernel of its own
        proj_w, proj_b = self.img_input_proj.weight, self.img_input_proj.bias
        if (img.is_cuda and projected.is_cuda and reference_points.is_cuda
                and img.dtype in (torch.float16, torch.bfloat16) and proj_b is not None
                and proj_w.dtype == img.dtype == projected.dtype == reference_points.dtype
                and (proj_b.requires_grad and torch.is_grad_enabled())
                and not torch.is_autocast_enabled("cuda") and not torch.compiler.is_compiling()
                and _fused_module_enabled() and _fused_value_proj_enabled() and not kernels.is_deterministic()
                and img.dim() == 3 and img.is_contiguous()):
            shape4 = (batch, num_pixels, heads, self.hidden_dim // heads)
            if (kernels.module_value_colsum_shape_supported(img.dtype, heads, shape4[3])
                    and kernels.module_core_shape_supported(img.dtype, shape4, projected, reference_points)):
                if img_shapes.device != img.device:
                    img_shapes = img_shapes.to(img.device, non_blocking=True)
                out = fused_value_proj_core(img, proj_w, proj_b, img_shapes, projected, reference_points, heads,
                                            self.padding_mode, self.align_corners)
                return self.query_output_proj(out.reshape(batch, num_queries, self.hidden_dim))

        value = self.img_input_proj(img).reshape(batch, num_pixels, heads, self.hidden_dim // heads)

        # CUDA fast path: softmax, sampling-point arithmetic and the operator in one kernel (no materialised
        # sampling_points / attention_weights).  Set MSDA_B200_FUSED_MODULE=0 to take the composed path below.
        # Operand precision follows the composed path below (= the reference, frontend.py:268-283): under autocast the
        # operator runs in fp32 (custom_fwd(cast_inputs=float32)) on sampling points computed as fp32 anchor + offsets,
        # so the fused core gets fp32 operands too -- the anchors are never rounded to the 16-bit autocast dtype.
        # Mixed dtypes outside autocast (e.g. a bf16 module fed fp32 anchors) take the composed path.
        autocast = value.is_cuda and torch.is_autocast_enabled("cuda")
        fusable = autocast or reference_points.dtype == value.dtype
        if (value.is_cuda and fusable and _fused_module_enabled()
                and not kernels.is_deterministic()):   # the bit-reproducible grad_img lives on the composed path
            ref = reference_points
            if autocast:
                value, projected, ref = value.float(), projected.float(), reference_points.float()
            if img_shapes.device != value.device:
                img_shapes = img_shapes.to(value.device, non_blocking=True)
            if torch.compiler.is_compiling():
                # traced programs: the same kernels behind torch.library custom ops (no graph break)
                if kernels.module_core_supported_static(value, projected, ref):
                    from .ops import module_core_op
                    out = module_core_op(value, img_shapes, projected, ref, self.padding_mode, self.align_corners)
                    return self.query_output_proj(out.reshape(batch, num_queries, self.hidden_dim))
            elif kernels.module_core_supported(value, projected, ref):
                out = fused_module_core(value, img_shapes, projected, ref, self.padding_mode, self.align_corners)
                return self.query_output_proj(out.reshape(batch, num_queries, self.hidden_dim))

        offsets, logits = projected[..., :2], projected[..., 2]
        attention_weights = logits.reshape(batch, num_queries, heads, levels * points).softmax(dim=-1)
        attention_weights = attention_weights.reshape(batch, num_queries, heads, levels, points)

        anchor = reference_points[:, :, None, None, None, :]
        coords = reference_points.shape[-1]
        if coords == 2:
            # NOTE: offsets are (x, y) while img_shapes rows are (h, w); the reference divides as-is
            # (frontend.py:272-276) and drop-in parity keeps that.
            sampling_points = anchor + offsets / img_shapes[:, None, :]
        elif coords == 4:
            sampling_points = anchor[..., :2] + offsets * anchor[..., 2:] / (2 * points)

        out = multiscale_deformable_attention(
            value, img_shapes, sampling_points, attention_weights, self.padding_mode, self.align_corners)
        return self.query_output_proj(out.reshape(batch, num_queries, self.hidden_dim))
