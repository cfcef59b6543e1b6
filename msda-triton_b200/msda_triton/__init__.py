"""msda_triton -- drop-in for rziga/msda-triton whose CUDA route is hand-written sm_100a (B200) CUDA.

Same import name and the same two public exports as the reference (``src/msda_triton/__init__.py:7-10``).
``__version__`` does not depend on pip metadata being installed (the reference's does, ``__init__.py:5``).
"""
from .frontend import MultiscaleDeformableAttention, multiscale_deformable_attention
from .kernels import is_deterministic, set_deterministic

__version__ = "0.1.1+b200.1"

__all__ = [
    "multiscale_deformable_attention",
    "MultiscaleDeformableAttention",
]
