"""ctypes front-end for the C oracle (``oracle/msda_oracle.c``) -- TEST INFRASTRUCTURE ONLY.

The product package never imports this module.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()``,
``bench.py`` (``cpu_baseline`` leg and ``--impl reference``).

API mirrors the reference's operator boundary (``/root/reference/src/msda_triton/kernels.py:351-358`` forward,
``:556-564`` backward) on numpy arrays:

    forward(img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners) -> out
    backward(out_grad, img, img_shapes, sampling_points, attention_weights, padding_mode, align_corners)
        -> (img_grad, sampling_points_grad, attention_weights_grad)

fp32 inputs are computed in fp32, fp64 in fp64 (like the reference kernels); any other float dtype is first
widened to fp64 (used as the "exact" comparison for fp16/bf16 storage tests).
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libmsda_oracle.so"
_lib = None

_PAD = {"zeros": 0, "border": 1}


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc); returns the .so path."""
    src = _HERE / "msda_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(_LIB_PATH))
        i64 = ctypes.c_int64
        vp = ctypes.c_void_p
        for sfx in ("f32", "f64"):
            f = getattr(lib, f"msda_oracle_fwd_{sfx}")
            f.restype = ctypes.c_int
            f.argtypes = [vp, vp, vp, vp, vp] + [i64] * 7 + [ctypes.c_int, ctypes.c_int]
            g = getattr(lib, f"msda_oracle_bwd_{sfx}")
            g.restype = ctypes.c_int
            g.argtypes = [vp] * 8 + [i64] * 7 + [ctypes.c_int, ctypes.c_int]
        lib.msda_oracle_level_table.restype = ctypes.c_int
        lib.msda_oracle_level_table.argtypes = [vp, vp, i64]
        _lib = lib
    return _lib


def _np(x, dtype=None):
    if hasattr(x, "detach"):  # torch tensor
        x = x.detach().cpu()
        if str(x.dtype) in ("torch.bfloat16", "torch.float16"):
            x = x.double()
        x = x.numpy()
    x = np.ascontiguousarray(x)
    if dtype is not None and x.dtype != dtype:
        x = x.astype(dtype)
    return x


def _prep(img, img_shapes, pts, aw):
    img = _np(img)
    work = img.dtype if img.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    img = _np(img, work)
    pts = _np(pts, work)
    aw = _np(aw, work)
    shapes = _np(img_shapes, np.int64)
    B, Npix, H, D = img.shape
    B2, Q, H2, L, K, two = pts.shape
    assert two == 2 and B2 == B and H2 == H and aw.shape == (B, Q, H, L, K) and shapes.shape == (L, 2)
    assert int((shapes[:, 0] * shapes[:, 1]).sum()) == Npix, "sum(h*w) must equal the pyramid length"
    return img, shapes, pts, aw, (B, Npix, H, D, Q, L, K), ("f32" if work == np.float32 else "f64")


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def forward(img, img_shapes, sampling_points, attention_weights, padding_mode: str, align_corners: bool):
    lib = _load()
    img, shapes, pts, aw, dims, sfx = _prep(img, img_shapes, sampling_points, attention_weights)
    B, Npix, H, D, Q, L, K = dims
    out = np.empty((B, Q, H, D), dtype=img.dtype)
    rc = getattr(lib, f"msda_oracle_fwd_{sfx}")(
        _p(out), _p(img), _p(shapes), _p(pts), _p(aw), *dims, _PAD[padding_mode], int(bool(align_corners)))
    assert rc == 0
    return out


def backward(out_grad, img, img_shapes, sampling_points, attention_weights, padding_mode: str, align_corners: bool):
    lib = _load()
    img, shapes, pts, aw, dims, sfx = _prep(img, img_shapes, sampling_points, attention_weights)
    gout = _np(out_grad, img.dtype)
    B, Npix, H, D, Q, L, K = dims
    assert gout.shape == (B, Q, H, D)
    gimg = np.empty_like(img)
    gpts = np.empty_like(pts)
    gaw = np.empty_like(aw)
    rc = getattr(lib, f"msda_oracle_bwd_{sfx}")(
        _p(gimg), _p(gpts), _p(gaw), _p(gout), _p(img), _p(shapes), _p(pts), _p(aw), *dims,
        _PAD[padding_mode], int(bool(align_corners)))
    assert rc == 0
    return gimg, gpts, gaw


def level_table(img_shapes):
    """Rows of (h, w, offset): what the device-side preprocessing must produce (kernels.py:60-62)."""
    lib = _load()
    shapes = _np(img_shapes, np.int64)
    table = np.empty((shapes.shape[0], 3), dtype=np.int64)
    lib.msda_oracle_level_table(_p(table), _p(shapes), shapes.shape[0])
    return table


def set_threads(n: int) -> None:
    _load().msda_oracle_set_threads(int(n))


def max_threads() -> int:
    return int(_load().msda_oracle_max_threads())
