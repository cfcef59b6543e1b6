"""CPU-side checks: the C-ABI library loads and exports every symbol include/msda_b200.h declares (no compute calls
without a GPU), the Python surface matches the reference's names / signatures / state_dict keys, the CPU-tensor route
agrees with the oracle, and the CUDA route refuses to run without CUDA tensors."""
import ctypes
import os
import inspect
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from util import BENCH_PYRAMID, assert_close, make_inputs, to_np

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def libpath():
    import importlib.util
    spec = importlib.util.spec_from_file_location("msda_b200_build", ROOT / "msda-triton_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_library()


def declared_symbols():
    text = (ROOT / "include" / "msda_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    assert declared_symbols() == sorted([
        "msda_abi_version", "msda_last_error", "msda_forward", "msda_backward_workspace_bytes", "msda_backward",
        "msda_level_table", "msda_probe_gather", "msda_probe_scatter", "msda_module_supported", "msda_module_forward",
        "msda_module_backward", "msda_reload_tuning", "msda_peer_all_gather", "msda_peer_reduce_scatter",
        "msda_module_colsum_supported", "msda_module_colsum_offset"])


def test_peer_entry_points_validate_without_a_gpu(libpath):
    """The peer-memory collectives reject a malformed context before touching the device."""
    from msda_triton import _lib
    lib = _lib.get_lib()
    ctx = _lib.MsdaPeerCtx(0, 0, None, None, None, None)          # world = 0
    assert lib.msda_peer_all_gather(None, ctypes.byref(ctx), 1, 16, None) < 0
    assert lib.msda_peer_reduce_scatter(None, ctypes.byref(ctx), 1, 4, None) < 0


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(str(libpath))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/msda_b200.h but not exported"
    lib.msda_abi_version.restype = ctypes.c_int
    assert lib.msda_abi_version() == 1
    lib.msda_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.msda_last_error(), bytes)


def test_argument_validation_needs_no_gpu(libpath):
    from msda_triton import _lib
    lib = _lib.get_lib()
    bad = _lib.MsdaProblem(1, 10, 2, 8, 4, 1, 1, 99, 0, 0, 0)       # unknown dtype code
    rc = lib.msda_forward(None, None, None, None, None, ctypes.byref(bad), None)
    assert rc == -2 and b"dtype" in lib.msda_last_error()
    bad = _lib.MsdaProblem(1, 10, 2, 8, 4, 1, 1, 0, 7, 0, 0)        # unknown padding mode
    assert lib.msda_forward(None, None, None, None, None, ctypes.byref(bad), None) == -4
    ok = _lib.MsdaProblem(2, 5440, 8, 32, 100, 4, 4, _lib.DTYPE_BF16, 0, 0, 0)
    assert lib.msda_backward_workspace_bytes(ctypes.byref(ok), 7) == 2 * 5440 * 8 * 32 * 4
    ok32 = _lib.MsdaProblem(2, 5440, 8, 32, 100, 4, 4, _lib.DTYPE_F32, 0, 0, 0)
    assert lib.msda_backward_workspace_bytes(ctypes.byref(ok32), 7) == 0
    # deterministic mode: exact row adds on the tuned fp32 shapes (a 64-bit bound per pyramid row + a float per unit) ...
    det = lib.msda_backward_workspace_bytes(ctypes.byref(ok32), 7 | 8)
    assert 8 * (2 * 5440 * 8) + 4 * (2 * 100 * 8) <= det < 2 << 20
    # ... the sorted-segment path everywhere else (24 bytes per bilinear corner)
    odd = _lib.MsdaProblem(2, 5440, 8, 24, 100, 4, 4, _lib.DTYPE_F32, 0, 0, 0)
    assert lib.msda_backward_workspace_bytes(ctypes.byref(odd), 7 | 8) >= 4 * 4 * (2 * 100 * 8 * 16 * 4)


def test_static_module_rule_agrees_with_library(libpath):
    """kernels.module_core_supported_static (used under torch.compile) restates msda_module_supported."""
    import ctypes
    import itertools
    import torch
    from msda_triton import _lib, kernels
    lib = _lib.get_lib()
    code = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3}
    cases = itertools.product(code, (16, 32, 64, 128), ((4, 4), (2, 8), (8, 2), (16, 1), (4, 3), (1, 16)), (2, 4),
                              ((2, 5440, 900), (1, 300000, 10), (1, 2 ** 18, 5)))
    checked = 0
    for dtype, D, (L, K), ref_dim, (B, npix, Q) in cases:
        H = 8
        value = torch.empty((B, npix, H, D), dtype=dtype, device="meta")
        proj = torch.empty((B, Q, H, L, K, 3), dtype=dtype, device="meta")
        ref = torch.empty((B, Q, ref_dim), dtype=dtype, device="meta")
        prob = _lib.MsdaProblem(B, npix, H, D, Q, L, K, code[dtype], 0, 0, 0)
        want = bool(lib.msda_module_supported(ctypes.byref(prob), ref_dim))
        # the static rule also requires CUDA tensors; lift that part for the comparison on meta tensors
        got = _static_rule_on_any_device(kernels, value, proj, ref)
        assert got == want, (dtype, D, L, K, ref_dim, B, npix, Q)
        checked += want
    assert checked > 0


def _static_rule_on_any_device(kernels, value, proj, ref):
    class _AsCuda:
        def __init__(self, t):
            self._t = t
            self.device = type("D", (), {"type": "cuda"})()
        def __getattr__(self, name):
            return getattr(self._t, name)
    return kernels.module_core_supported_static(_AsCuda(value), proj, ref)


def test_public_surface_matches_reference():
    import msda_triton
    from msda_triton import frontend
    assert set(msda_triton.__all__) == {"multiscale_deformable_attention", "MultiscaleDeformableAttention"}
    assert isinstance(msda_triton.__version__, str)
    for name in ("triton_multiscale_deformable_attention", "native_multiscale_deformable_attention",
                 "MultiscaleDeformableAttention", "multiscale_deformable_attention"):
        assert hasattr(frontend, name)
    want = ["img", "img_shapes", "sampling_points", "attention_weights", "padding_mode", "align_corners"]
    for fn in (frontend.multiscale_deformable_attention, frontend.triton_multiscale_deformable_attention,
               frontend.native_multiscale_deformable_attention):
        assert list(inspect.signature(fn).parameters) == want
    ctor = list(inspect.signature(frontend.MultiscaleDeformableAttention.__init__).parameters)[1:]
    assert ctor == ["emb_dim", "hidden_dim", "num_levels", "num_heads", "num_points", "padding_mode", "align_corners"]
    m = frontend.MultiscaleDeformableAttention(64, 32, 3, 4, 2, "zeros", False)
    assert sorted(m.state_dict()) == sorted(f"{p}.{w}" for p in ("img_input_proj", "query_input_proj", "query_output_proj")
                                            for w in ("weight", "bias"))
    assert m.query_input_proj.out_features == 4 * 3 * 2 * 3
    with pytest.raises(ValueError):
        frontend.MultiscaleDeformableAttention(64, 30, 3, 4, 2, "zeros", False)


def test_cuda_route_rejects_cpu_tensors_and_bad_dtypes():
    from msda_triton.frontend import triton_multiscale_deformable_attention as cuda_route
    img, s, pts, aw, _ = make_inputs(1, 4, 2, 8, [(4, 4)], 2)
    with pytest.raises(ValueError, match="gpu"):
        cuda_route(img, s, pts, aw, "zeros", False)
    with pytest.raises(ValueError, match="Dtype"):
        cuda_route(img.to(torch.int32), s, pts, aw, "zeros", False)


@pytest.mark.parametrize("pm", ["zeros", "border"])
@pytest.mark.parametrize("ac", [False, True])
def test_cpu_tensor_route_matches_oracle(pm, ac):
    """CPU tensors take the torch route the reference documents for device='cpu' (README.md:132)."""
    import msda_triton
    from oracle import msda_oracle
    img, s, pts, aw, go = make_inputs(2, 40, 4, 16, [(9, 7), (5, 4), (2, 3)], 3, dtype=torch.float64, seed=3, points="wide")
    a, b, c = (t.clone().requires_grad_(True) for t in (img, pts, aw))
    out = msda_triton.multiscale_deformable_attention(a, s, b, c, pm, ac)
    out.backward(go)
    ref_out = msda_oracle.forward(img, s, pts, aw, pm, ac)
    rgi, rgp, rga = msda_oracle.backward(go, img, s, pts, aw, pm, ac)
    assert_close(to_np(out), ref_out, 1e-9, 1e-10, "out")
    assert_close(to_np(a.grad), rgi, 1e-9, 1e-10, "grad_img")
    assert_close(to_np(c.grad), rga, 1e-9, 1e-10, "grad_weights")
    assert_close(to_np(b.grad), rgp, 1e-8, 1e-9 * np.abs(rgp).max(), "grad_points")


@pytest.mark.parametrize("coords", [2, 4])
def test_module_on_cpu(coords):
    """Mirrors the reference's only GPU-less test (tests/test_msda.py:154-168): D = 4, K = 8, randn reference points."""
    from msda_triton import MultiscaleDeformableAttention
    torch.manual_seed(0)
    channels, heads, levels, points = 256, 8, 4, 8
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img = torch.randn(2, npix, channels)
    queries = torch.randn(2, 50, channels)
    ref_pts = torch.randn(2, 50, coords)
    module = MultiscaleDeformableAttention(channels, channels // heads, levels, heads, points, "border", True)
    out = module(img, torch.tensor(BENCH_PYRAMID), queries, ref_pts)
    assert out.shape == (2, 50, channels) and torch.isfinite(out).all()
    with pytest.raises(ValueError):
        module(img, torch.tensor(BENCH_PYRAMID), queries, torch.randn(2, 50, 3))


def test_grid_sample_port_matches_oracle():
    """The CPU-baseline port (oracle/grid_sample_port.py) and the C oracle are two independent restatements."""
    from oracle import grid_sample_port, msda_oracle
    img, s, pts, aw, go = make_inputs(2, 30, 3, 8, [(6, 7), (3, 4)], 4, dtype=torch.float64, seed=8, points="wide")
    out, gi, gp, ga = grid_sample_port.forward_backward(img, s, pts, aw, go, "zeros", False)
    assert_close(to_np(out), msda_oracle.forward(img, s, pts, aw, "zeros", False), 1e-9, 1e-10, "out")
    rgi, rgp, rga = msda_oracle.backward(go, img, s, pts, aw, "zeros", False)
    assert_close(to_np(gi), rgi, 1e-9, 1e-10, "grad_img")
    assert_close(to_np(ga), rga, 1e-9, 1e-10, "grad_weights")


@pytest.mark.parametrize("name", ["module_ref2d_float64", "module_ref4d_float64", "module_ref2d_hd64_float64"])
def test_module_matches_reference_module_golden_cpu(name):
    """Our nn.Module, loaded with the REFERENCE module's state_dict, reproduces the reference module's output and all
    gradients (golden vectors from oracle/make_golden.py) on the CPU-tensor route."""
    from conftest import GOLDEN
    from util import check_module_against_golden, load_module_golden
    g, module, inputs, shapes = load_module_golden(GOLDEN / f"{name}.npz")
    check_module_against_golden(g, module, inputs, shapes, rtol=1e-9, atol_scale=1e-11)


def test_bench_reference_arm_emits_contract_json():
    """`bench.py --impl reference` (the reference's CPU route on the host cores) prints ONE JSON line with the keys the
    driver reads -- needs no GPU."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_transformers_patch_keeps_cpu_path_and_restores():
    """integrations.patch_transformers: CPU tensors keep Hugging Face's own operator (and agree with our CPU route);
    unpatch restores the class."""
    import torch
    try:
        from transformers.models.deformable_detr import modeling_deformable_detr as m
    except Exception as e:  # noqa: BLE001
        pytest.skip(f"transformers unavailable: {e}")
    import msda_triton
    from msda_triton import integrations
    from util import make_inputs
    shapes_list = [(6, 8), (3, 4)]
    img, shapes, pts, aw, _ = make_inputs(2, 9, 4, 8, shapes_list, 3, seed=2, weights="softmax_lk")
    level_start = torch.tensor([0, 48])
    op = m.MultiScaleDeformableAttention()
    before = op(img, shapes, shapes_list, level_start, pts, aw, 64)
    original = m.MultiScaleDeformableAttention.forward
    assert "deformable_detr" in integrations.patch_transformers(["deformable_detr", "no_such_family"])
    try:
        assert m.MultiScaleDeformableAttention.forward is not original
        patched = op(img, shapes, shapes_list, level_start, pts, aw, 64)
    finally:
        integrations.unpatch_transformers()
    assert m.MultiScaleDeformableAttention.forward is original
    assert torch.equal(patched, before)
    ours = msda_triton.multiscale_deformable_attention(img, shapes, pts, aw, "zeros", False).flatten(2)
    torch.testing.assert_close(ours, before, rtol=1e-5, atol=1e-6)


def test_missing_library_fails_loudly():
    """No CPU / torch fallback for the CUDA path: without libmsda_b200.so the first library use raises."""
    import subprocess
    import sys
    from pathlib import Path
    pkg = Path(__file__).resolve().parent.parent / "msda-triton_b200"
    code = (
        "import sys; sys.path.insert(0, sys.argv[1])\n"
        "import msda_triton\n"
        "from msda_triton import _lib\n"
        "try:\n"
        "    _lib.get_lib()\n"
        "except _lib.MsdaLibraryError as e:\n"
        "    assert 'no fallback' in str(e); print('raised')\n"
    )
    env = dict(os.environ, MSDA_B200_LIB="/nonexistent/libmsda_b200.so")
    out = subprocess.run([sys.executable, "-c", code, str(pkg)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "raised", out.stderr[-2000:]


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under msda-triton_b200/ imports, loads or mentions a path into it."""
    from pathlib import Path
    pkg = Path(__file__).resolve().parent.parent / "msda-triton_b200"
    offenders = []
    for path in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        text = path.read_text()
        if "import oracle" in text or "from oracle" in text or "msda_oracle" in text or "oracle/" in text:
            offenders.append(str(path))
    assert not offenders, offenders


def test_value_colsum_workspace_contract(libpath):
    """MSDA_BWD_VALUE_COLSUM: 16-bit storage only; the H*D column sums live behind the (256-byte padded) accumulation image."""
    from msda_triton import _lib
    lib = _lib.get_lib()
    for dtype, ok in ((2, 1), (1, 1), (0, 0)):
        prob = _lib.MsdaProblem(2, 85, 8, 32, 10, 4, 4, dtype, 0, 0, 0)
        assert lib.msda_module_colsum_supported(ctypes.byref(prob)) == ok
        accum = 4 * 2 * 85 * 8 * 32
        assert lib.msda_module_colsum_offset(ctypes.byref(prob)) == (accum + 255) // 256 * 256
        plain = lib.msda_backward_workspace_bytes(ctypes.byref(prob), 1)
        with_sums = lib.msda_backward_workspace_bytes(ctypes.byref(prob), 1 | _lib.BWD_VALUE_COLSUM)
        if ok:
            assert plain == accum and with_sums == (accum + 255) // 256 * 256 + 4 * 8 * 32
        else:
            assert plain == 0 and with_sums == 0
    wide = _lib.MsdaProblem(1, 85, 40, 64, 10, 4, 4, 2, 0, 0, 0)     # H*D = 2560 > 2048 columns
    assert lib.msda_module_colsum_supported(ctypes.byref(wide)) == 0
    # the flag on an fp32 problem is refused before anything touches the device
    prob = _lib.MsdaProblem(2, 85, 8, 32, 10, 4, 4, 0, 0, 0, 0)
    rc = lib.msda_module_backward(None, None, None, None, None, None, None, None, 2, ctypes.byref(prob),
                                  1 | _lib.BWD_VALUE_COLSUM, None, 0, None)
    assert rc == -4 and b"MSDA_BWD_VALUE_COLSUM" in lib.msda_last_error()      # MSDA_ERR_BAD_MODE
