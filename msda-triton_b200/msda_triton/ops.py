"""``torch.library`` registration of the MSDA kernels (the reference has none).

``torch.ops.msda_b200.forward`` / ``torch.ops.msda_b200.backward`` are opaque custom ops with fake (meta) kernels and
an autograd formula, so a model that calls :func:`multiscale_deformable_attention_op` traces through
``torch.compile(fullgraph=True)`` / ``torch.export`` without graph breaks and without running the CUDA code at trace
time.  The eager public API (``msda_triton.multiscale_deformable_attention``) keeps using the
``torch.autograd.Function`` whose AMP contract mirrors the reference; under ``torch.compile`` it routes here.

``torch.ops.msda_b200.module_forward`` / ``module_backward`` do the same for the fused module core
(softmax + sampling-point arithmetic + operator in one kernel), so a compiled ``MultiscaleDeformableAttention`` keeps
the fused fast path.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import kernels

__all__ = ["multiscale_deformable_attention_op", "module_core_op"]


@torch.library.custom_op("msda_b200::forward", mutates_args=(), device_types="cuda")
def _forward(img: torch.Tensor, img_shapes: torch.Tensor, sampling_points: torch.Tensor,
             attention_weights: torch.Tensor, padding_mode: str, align_corners: bool) -> torch.Tensor:
    return kernels.b200_multi_scale_deformable_attention_fwd(
        img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners)


@_forward.register_fake
def _(img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners):
    B, _, H, D = img.shape
    return img.new_empty((B, sampling_points.shape[1], H, D))


@torch.library.custom_op("msda_b200::backward", mutates_args=(), device_types="cuda")
def _backward(out_grad: torch.Tensor, img: torch.Tensor, img_shapes: torch.Tensor, sampling_points: torch.Tensor,
              attention_weights: torch.Tensor, padding_mode: str, align_corners: bool,
              needs: List[bool]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    gi, gp, ga = kernels.b200_multi_scale_deformable_attention_bwd(
        out_grad, img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners, needs=needs)
    # custom ops must return tensors: gradients that were not requested come back as empty placeholders
    empty = img.new_empty((0,))
    return (gi if gi is not None else empty, gp if gp is not None else empty, ga if ga is not None else empty)


@_backward.register_fake
def _(out_grad, img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners, needs):
    empty = img.new_empty((0,))
    return (torch.empty_like(img, memory_format=torch.contiguous_format) if needs[0] else empty,
            torch.empty_like(sampling_points, memory_format=torch.contiguous_format) if needs[1] else empty,
            torch.empty_like(attention_weights, memory_format=torch.contiguous_format) if needs[2] else empty)


def _setup_context(ctx, inputs, output):
    img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners = inputs
    ctx.save_for_backward(img, img_shapes, sampling_points, attention_weights)
    ctx.padding_mode = padding_mode
    ctx.align_corners = align_corners


def _autograd_backward(ctx, out_grad):
    img, img_shapes, sampling_points, attention_weights = ctx.saved_tensors
    needs = [bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[2]), bool(ctx.needs_input_grad[3])]
    gi, gp, ga = torch.ops.msda_b200.backward(
        out_grad.contiguous(), img, img_shapes, sampling_points, attention_weights, ctx.padding_mode,
        ctx.align_corners, needs)
    return (gi if needs[0] else None, None, gp if needs[1] else None, ga if needs[2] else None, None, None)


_forward.register_autograd(_autograd_backward, setup_context=_setup_context)


def multiscale_deformable_attention_op(img, img_shapes, sampling_points, attention_weights, padding_mode: str,
                                       align_corners: bool) -> torch.Tensor:
    """Same semantics as ``msda_triton.multiscale_deformable_attention`` for CUDA tensors of one common dtype,
    expressed as a ``torch.library`` custom op (traceable by torch.compile / torch.export)."""
    return torch.ops.msda_b200.forward(img, img_shapes, sampling_points, attention_weights, padding_mode,
                                       bool(align_corners))


# ---------------------------------------------------------------------------------------------------------------------
# fused module core
# ---------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("msda_b200::module_forward", mutates_args=(), device_types="cuda")
def _module_forward(value: torch.Tensor, img_shapes: torch.Tensor, projection: torch.Tensor,
                    reference_points: torch.Tensor, padding_mode: str, align_corners: bool) -> torch.Tensor:
    return kernels.b200_module_core_fwd(value, img_shapes, projection, reference_points, padding_mode, align_corners)


@_module_forward.register_fake
def _(value, img_shapes, projection, reference_points, padding_mode, align_corners):
    B, _, H, D = value.shape
    return value.new_empty((B, projection.shape[1], H, D))


@torch.library.custom_op("msda_b200::module_backward", mutates_args=(), device_types="cuda")
def _module_backward(out_grad: torch.Tensor, value: torch.Tensor, img_shapes: torch.Tensor, projection: torch.Tensor,
                     reference_points: torch.Tensor, padding_mode: str, align_corners: bool,
                     needs: List[bool]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    gv, gp, gr = kernels.b200_module_core_bwd(
        out_grad, value, img_shapes, projection, reference_points, padding_mode, align_corners, needs=needs)
    empty = value.new_empty((0,))
    return (gv if gv is not None else empty, gp if gp is not None else empty, gr if gr is not None else empty)


@_module_backward.register_fake
def _(out_grad, value, img_shapes, projection, reference_points, padding_mode, align_corners, needs):
    empty = value.new_empty((0,))
    return (torch.empty_like(value, memory_format=torch.contiguous_format) if needs[0] else empty,
            torch.empty_like(projection, memory_format=torch.contiguous_format) if needs[1] else empty,
            torch.empty_like(reference_points, memory_format=torch.contiguous_format) if needs[2] else empty)


def _module_setup_context(ctx, inputs, output):
    value, img_shapes, projection, reference_points, padding_mode, align_corners = inputs
    ctx.save_for_backward(value, img_shapes, projection, reference_points)
    ctx.padding_mode = padding_mode
    ctx.align_corners = align_corners


def _module_autograd_backward(ctx, out_grad):
    value, img_shapes, projection, reference_points = ctx.saved_tensors
    needs = [bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[2]), bool(ctx.needs_input_grad[3])]
    gv, gp, gr = torch.ops.msda_b200.module_backward(
        out_grad.contiguous(), value, img_shapes, projection, reference_points, ctx.padding_mode, ctx.align_corners,
        needs)
    return (gv if needs[0] else None, None, gp if needs[1] else None, gr if needs[2] else None, None, None)


_module_forward.register_autograd(_module_autograd_backward, setup_context=_module_setup_context)


def module_core_op(value, img_shapes, projection, reference_points, padding_mode: str,
                   align_corners: bool) -> torch.Tensor:
    """``msda_triton.frontend.fused_module_core`` as a ``torch.library`` custom op (same operands, same result)."""
    return torch.ops.msda_b200.module_forward(value, img_shapes, projection, reference_points, padding_mode,
                                              bool(align_corners))
