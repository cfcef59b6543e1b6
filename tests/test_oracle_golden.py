"""Pins the C oracle against the committed golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py: reference Triton kernels under the CPU interpreter + reference native grid_sample route)."""
import itertools

import numpy as np
import pytest

from oracle import msda_oracle as oracle

from conftest import GOLDEN

FILES = sorted(p for p in GOLDEN.glob("*.npz") if not p.name.endswith("float16.npz") and not p.name.startswith("module_"))
MODES = list(itertools.product(("zeros", "border"), (False, True)))

# oracle and reference compute in the same precision with the same per-element op order; only the (l,k) summation
# order differs (sequential vs tree), hence a few ulps.
TOL = {np.dtype(np.float32): dict(rtol=1e-5, atol=2e-6), np.dtype(np.float64): dict(rtol=1e-12, atol=1e-13)}


def _close(a, b, what, **tol):
    np.testing.assert_allclose(a, b, err_msg=what, **tol)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.stem)
@pytest.mark.parametrize("pm,ac", MODES)
def test_oracle_matches_reference_triton_kernels(path, pm, ac):
    g = np.load(path)
    tag = f"{pm}_{int(ac)}"
    args = (g["img"], g["img_shapes"], g["sampling_points"], g["attention_weights"], pm, ac)
    tol = TOL[g["img"].dtype]
    out = oracle.forward(*args)
    _close(out, g[f"triton_out_{tag}"], "out", **tol)
    gimg, gpts, gaw = oracle.backward(g["out_grad"], *args)
    _close(gimg, g[f"triton_gimg_{tag}"], "grad_img", **tol)
    _close(gaw, g[f"triton_gaw_{tag}"], "grad_attention_weights", **tol)
    scale = max(1.0, float(np.abs(g[f"triton_gpts_{tag}"]).max()))
    _close(gpts, g[f"triton_gpts_{tag}"], "grad_sampling_points", rtol=tol["rtol"], atol=tol["atol"] * scale)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.stem)
@pytest.mark.parametrize("pm,ac", MODES)
def test_oracle_matches_reference_native_route(path, pm, ac):
    """grid_sample un-normalises as ((2p-1)+1)*W/2 - .5 (ATen GridSampler.h:27-35): mathematically the same
    coordinate, rounded differently, so the comparison is to coordinate-rounding accuracy (fp32: ~1e-5 relative
    to the largest value), and cells may flip for points exactly on pixel boundaries ("edges" case:
    grad_sampling_points is compared only where the forward coordinate is not within 1e-6 of an integer)."""
    g = np.load(path)
    tag = f"{pm}_{int(ac)}"
    args = (g["img"], g["img_shapes"], g["sampling_points"], g["attention_weights"], pm, ac)
    f32 = g["img"].dtype == np.float32
    rtol, atol = (1e-4, 1e-4) if f32 else (1e-9, 1e-10)
    out = oracle.forward(*args)
    _close(out, g[f"native_out_{tag}"], "out", rtol=rtol, atol=atol)
    gimg, gpts, gaw = oracle.backward(g["out_grad"], *args)
    _close(gimg, g[f"native_gimg_{tag}"], "grad_img", rtol=rtol, atol=atol)
    _close(gaw, g[f"native_gaw_{tag}"], "grad_attention_weights", rtol=rtol, atol=atol)
    ref = g[f"native_gpts_{tag}"]
    shapes = g["img_shapes"].astype(np.float64)
    p = g["sampling_points"].astype(np.float64)
    wh = shapes[:, ::-1][None, None, None, :, None, :]  # (w, h) per level, broadcast over K
    coord = p * (wh - 1) if ac else p * wh - 0.5
    on_edge = (np.abs(coord - np.round(coord)) < 1e-6)
    keep = ~on_edge
    scale = max(1.0, float(np.abs(ref).max()))
    np.testing.assert_allclose(gpts[keep], ref[keep], rtol=rtol * 10, atol=atol * scale * 10)


def test_level_table():
    t = oracle.level_table(np.array([[100, 167], [50, 84], [25, 42], [13, 21]]))
    assert t[:, 2].tolist() == [0, 16700, 20900, 21950]
    assert int(t[-1, 2] + t[-1, 0] * t[-1, 1]) == 22223
