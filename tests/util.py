"""Seeded input generators and comparison helpers shared by the tests (and mirrored by bench.py)."""
from __future__ import annotations

import numpy as np
import torch

BENCH_PYRAMID = [(64, 64), (32, 32), (16, 16), (8, 8)]            # scripts/benchmark.py:27
DETR_PYRAMID = [(100, 167), (50, 84), (25, 42), (13, 21)]          # 800x1333 at strides 8..64


def make_inputs(B, Q, H, D, shapes, K, dtype=torch.float32, seed=0, points="unit", device="cpu", weights="softmax_k"):
    """img ~ N(0,1), points ~ U[0,1) ('unit'), U[-.25,1.25) ('wide') or N(.5,1.5) ('far'), weights = softmax(N(0,1))
    over K (the reference's fixture, tests/test_msda.py:41) or over L*K (module-like), grad_out ~ U[0,1)."""
    g = torch.Generator().manual_seed(seed)
    L = len(shapes)
    npix = sum(h * w for h, w in shapes)
    img = torch.randn(B, npix, H, D, generator=g, dtype=torch.float32)
    pts = torch.rand(B, Q, H, L, K, 2, generator=g, dtype=torch.float32)
    if points == "wide":
        pts = pts * 1.5 - 0.25
    elif points == "far":
        pts = torch.randn(B, Q, H, L, K, 2, generator=g, dtype=torch.float32) * 1.5 + 0.5
    logits = torch.randn(B, Q, H, L, K, generator=g, dtype=torch.float32)
    if weights == "softmax_k":
        aw = torch.softmax(logits, dim=-1)
    else:
        aw = torch.softmax(logits.reshape(B, Q, H, L * K), dim=-1).reshape(B, Q, H, L, K)
    go = torch.rand(B, Q, H, D, generator=g, dtype=torch.float32)
    shapes_t = torch.tensor(shapes, dtype=torch.int64)
    cast = lambda t: t.to(dtype).to(device)  # noqa: E731
    return cast(img), shapes_t.to(device), cast(pts), cast(aw), cast(go)


def to_np(t):
    t = t.detach().cpu()
    if t.dtype in (torch.float16, torch.bfloat16):
        t = t.double()
    return t.numpy()


def assert_close(test, ref, rtol, atol, what="", max_outliers=0):
    """|test-ref| <= atol + rtol*|ref| elementwise, allowing `max_outliers` violations (floor-cell flips in
    grad_sampling_points; 0 everywhere else)."""
    test = np.asarray(test, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert test.shape == ref.shape, f"{what}: shape {test.shape} vs {ref.shape}"
    err = np.abs(test - ref)
    bad = err > (atol + rtol * np.abs(ref))
    nbad = int(bad.sum())
    if nbad > max_outliers:
        idx = np.unravel_index(np.argmax(err * bad), err.shape)
        raise AssertionError(
            f"{what}: {nbad} / {err.size} elements outside rtol={rtol} atol={atol:.3g}; worst at {idx}: "
            f"test={test[idx]!r} ref={ref[idx]!r} err={err[idx]:.3e}; max|ref|={np.abs(ref).max():.3e}")
    return nbad


def load_module_golden(path, device="cpu", dtype=torch.float64):
    """Rebuilds OUR module from a golden file written by oracle/make_golden.py (reference module, seeded weights)."""
    from msda_triton import MultiscaleDeformableAttention
    g = np.load(path)
    emb, hidden, levels, heads, points, coords, ac = (int(v) for v in g["config"])
    module = MultiscaleDeformableAttention(emb, hidden, levels, heads, points, str(g["padding_mode"]), bool(ac))
    state = {k[len("param."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param.")}
    module.load_state_dict(state, strict=True)          # identical parameter names = drop-in checkpoints
    module = module.to(device=device, dtype=dtype)
    inputs = {k: torch.from_numpy(g[k]).to(device=device, dtype=dtype).requires_grad_(True)
              for k in ("img", "queries", "reference_points")}
    shapes = torch.from_numpy(g["img_shapes"]).to(device)
    return g, module, inputs, shapes


def check_module_against_golden(g, module, inputs, shapes, rtol, atol_scale, max_outliers=0):
    out = module(inputs["img"], shapes, inputs["queries"], inputs["reference_points"])
    out.double().square().sum().backward()
    assert_close(to_np(out), g["out"], rtol, atol_scale * np.abs(g["out"]).max(), "module out", max_outliers)
    for name, key in (("img", "grad_img"), ("queries", "grad_queries"), ("reference_points", "grad_reference_points")):
        ref = g[key]
        assert_close(to_np(inputs[name].grad), ref, rtol, atol_scale * np.abs(ref).max(), key, max_outliers)
    for name, p in module.named_parameters():
        ref = g[f"grad.{name}"]
        assert_close(to_np(p.grad), ref, rtol, atol_scale * np.abs(ref).max(), f"grad {name}", max_outliers)


class knobs:
    """with knobs(MSDA_B200_FORCE_GENERIC="1"): ...  -- sets library knobs in the environment and has the library
    re-read them (it reads the environment once, not per launch); restores both on exit."""

    def __init__(self, **env):
        self.env = {k: str(v) for k, v in env.items()}
        self.saved = {}

    def __enter__(self):
        import os
        from msda_triton import _lib
        for k, v in self.env.items():
            self.saved[k] = os.environ.get(k)
            os.environ[k] = v
        _lib.reload_tuning()
        return self

    def __exit__(self, *exc):
        import os
        from msda_triton import _lib
        for k, old in self.saved.items():
            if old is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old
        _lib.reload_tuning()
        return False
