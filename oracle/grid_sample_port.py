"""Port of the reference's native-torch CPU route -- TEST / BASELINE INFRASTRUCTURE ONLY.

Restates ``native_multiscale_deformable_attention`` (``/root/reference/src/msda_triton/frontend.py:15-68``):
per pyramid level, view the level as an NCHW image batch of ``B*H`` images, call ``torch.nn.functional.grid_sample``
(bilinear) at ``2*p - 1`` (``frontend.py:34``), then weight and sum over (level, point) (``frontend.py:64-66``).
The arithmetic itself lives in PyTorch's ``grid_sample`` (third-party; torch 2.11.0 here, the reference's lock pins
2.6.0) and is differentiable through torch autograd, which is how the reference's CPU route gets its backward.

Used as (a) the second, independent check of the C oracle and (b) the ``cpu_baseline`` / ``--impl reference`` timing
leg of ``bench.py`` -- "the reference's native-torch grid_sample CPU fallback timed on the box's host cores".
The product package never imports this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def grid_sample_msda(img, img_shapes, sampling_points, attention_weights, padding_mode: str, align_corners: bool):
    B, _, H, D = img.shape
    _, Q, _, L, K, _ = sampling_points.shape
    hw = [(int(h), int(w)) for h, w in img_shapes.tolist()]
    grid_all = sampling_points * 2 - 1                                   # [0,1] -> [-1,1]
    start = 0
    samples = []
    for lvl, (h, w) in enumerate(hw):
        feat = img[:, start:start + h * w]                               # [B, h*w, H, D]
        start += h * w
        feat = feat.permute(0, 2, 3, 1).reshape(B * H, D, h, w)          # NCHW with N = B*H
        grid = grid_all[:, :, :, lvl].permute(0, 2, 1, 3, 4).reshape(B * H, Q, K, 2)
        smp = F.grid_sample(feat, grid, mode="bilinear", padding_mode=padding_mode, align_corners=align_corners)
        samples.append(smp.reshape(B, H, D, Q, K).permute(0, 3, 1, 4, 2))   # [B, Q, H, K, D]
    # as the reference: materialise [B, Q, H, L, K, D], weight, and reduce over (level, point) in one sum
    stacked = torch.stack(samples, dim=3)
    return (attention_weights[..., None] * stacked).sum(dim=(3, 4))


def forward_backward(img, img_shapes, sampling_points, attention_weights, out_grad, padding_mode, align_corners):
    """Returns (out, grad_img, grad_points, grad_weights) via torch autograd, like the reference's CPU route."""
    img = img.detach().clone().requires_grad_(True)
    pts = sampling_points.detach().clone().requires_grad_(True)
    aw = attention_weights.detach().clone().requires_grad_(True)
    out = grid_sample_msda(img, img_shapes, pts, aw, padding_mode, align_corners)
    out.backward(out_grad)
    return out.detach(), img.grad, pts.grad, aw.grad
