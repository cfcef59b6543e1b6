"""Optional glue for model libraries that carry their own pure-PyTorch deformable attention.

The reference README's "Detection Example" replaces the operator inside a Deformable-DETR-like Hugging Face model by
hand (README.md:25-37).  :func:`patch_transformers` does that replacement: every ``MultiScaleDeformableAttention`` class
of the listed ``transformers`` model families (same ``forward`` signature everywhere: value ``[B, I, H, C]``, spatial
shapes ``[L, 2]`` as (height, width), sampling locations ``[B, N, H, L, P, 2]`` as (x, y) in [0, 1], attention weights
``[B, N, H, L, P]``; ``grid_sample`` with ``padding_mode="zeros"``, ``align_corners=False``) is routed to
:func:`msda_triton.multiscale_deformable_attention` for CUDA tensors and keeps its original code for CPU tensors.
Nothing here is imported by the package itself; ``transformers`` is only needed when the function is called.
"""
from __future__ import annotations

import importlib
from typing import Dict, Iterable, List

from .frontend import multiscale_deformable_attention

HF_FAMILIES = ("deformable_detr", "grounding_dino", "mm_grounding_dino", "rt_detr", "omdet_turbo", "lw_detr")
_originals: Dict[type, object] = {}


def _forward(self, value, value_spatial_shapes, value_spatial_shapes_list, level_start_index, sampling_locations,
             attention_weights, im2col_step=None):
    if not value.is_cuda:
        return _originals[type(self)](self, value, value_spatial_shapes, value_spatial_shapes_list, level_start_index,
                                      sampling_locations, attention_weights, im2col_step)
    out = multiscale_deformable_attention(value, value_spatial_shapes, sampling_locations, attention_weights,
                                          "zeros", False)
    return out.flatten(2)   # [B, N, H, C] -> [B, N, H*C], the layout the Hugging Face operator returns


def patch_transformers(families: Iterable[str] = HF_FAMILIES) -> List[str]:
    """Routes the deformable-attention operator of the given ``transformers`` model families through this package.
    Returns the families that were patched (families missing from the installed ``transformers`` are skipped)."""
    done = []
    for family in families:
        try:
            module = importlib.import_module(f"transformers.models.{family}.modeling_{family}")
        except Exception:  # noqa: BLE001 -- family not present in this transformers version
            continue
        cls = getattr(module, "MultiScaleDeformableAttention", None)
        if cls is None:
            continue
        if cls not in _originals:
            _originals[cls] = cls.forward
            cls.forward = _forward
        done.append(family)
    return done


def unpatch_transformers() -> None:
    """Restores every operator replaced by :func:`patch_transformers`."""
    for cls, original in _originals.items():
        cls.forward = original
    _originals.clear()
