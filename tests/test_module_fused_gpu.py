"""Fused module core (softmax + sampling-point arithmetic + MSDA in one kernel each way) against (a) the composed
torch prologue + CPU route in fp64 and (b) the same nn.Module taking the composed CUDA path."""
import itertools
import os

import numpy as np
import pytest
import torch

from util import BENCH_PYRAMID, DETR_PYRAMID, assert_close, to_np

pytestmark = pytest.mark.gpu


def composed_reference(value, shapes, proj, ref, pm, ac):
    """frontend.py:253-289 of the reference, spelled out on CPU in fp64 with torch autograd."""
    from msda_triton.frontend import native_multiscale_deformable_attention
    B, Q, H, L, K, _ = proj.shape
    off, logit = proj[..., :2], proj[..., 2]
    aw = logit.reshape(B, Q, H, L * K).softmax(-1).reshape(B, Q, H, L, K)
    anchor = ref[:, :, None, None, None, :]
    if ref.shape[-1] == 2:
        pts = anchor + off / shapes[:, None, :]
    else:
        pts = anchor[..., :2] + off * anchor[..., 2:] / (2 * K)
    return native_multiscale_deformable_attention(value, shapes, pts, aw, pm, ac)


@pytest.mark.parametrize("coords", [2, 4])
@pytest.mark.parametrize("pm,ac", list(itertools.product(("zeros", "border"), (False, True))))
@pytest.mark.parametrize("pyramid", [BENCH_PYRAMID, [(25, 42), (13, 21), (7, 11), (4, 6)]], ids=["square", "rect"])
@pytest.mark.parametrize("D", [32, 64])
def test_fused_core_matches_composed_fp64(coords, pm, ac, pyramid, D):
    """fp32 kernels against fp64 TRUTH: the coordinate arithmetic rounds differently (x * w - 0.5 carries half an ulp of
    x, i.e. up to 4e-6 of a pixel on a 64-pixel level, times the local slope of the image), so this comparison can only
    hold to ~1e-4 / 1e-5 and a few floor-cell flips in the gradients; the BASELINE bar (1e-5 / 1e-6, 1e-4) is checked at
    the SAME precision in test_fused_core_matches_composed_fp32_same_precision below."""
    from msda_triton import kernels
    from msda_triton.frontend import fused_module_core
    g = torch.Generator().manual_seed(31 + coords)
    B, Q, H, L, K = 2, 333, 8, 4, 4
    npix = sum(h * w for h, w in pyramid)
    value = torch.randn(B, npix, H, D, generator=g)
    proj = torch.randn(B, Q, H, L, K, 3, generator=g)
    proj[..., :2] *= 3.0                                   # offsets of a few pixels
    ref = torch.rand(B, Q, coords, generator=g)
    if coords == 4:
        ref[..., 2:] = ref[..., 2:] * 0.5 + 0.05
    go = torch.rand(B, Q, H, D, generator=g)
    shapes = torch.tensor(pyramid)

    a, b, c = (t.double().requires_grad_(True) for t in (value, proj, ref))
    want = composed_reference(a, shapes, b, c, pm, ac)
    want.backward(go.double())

    x, y, z = (t.cuda().requires_grad_(True) for t in (value, proj, ref))
    assert kernels.module_core_supported(x, y, z)
    got = fused_module_core(x, shapes.cuda(), y, z, pm, ac)
    got.backward(go.cuda())
    assert_close(to_np(got), to_np(want), 1e-4, 1e-5, "out")
    for t, r, what in ((x.grad, a.grad, "grad_value"), (y.grad, b.grad, "grad_projection"), (z.grad, c.grad, "grad_ref")):
        r = to_np(r)
        assert_close(to_np(t), r, 1e-3, 2e-5 * np.abs(r).max(), what, max_outliers=4)


@pytest.mark.parametrize("coords", [2, 4])
@pytest.mark.parametrize("pm,ac", [("zeros", False), ("border", True)])
@pytest.mark.parametrize("D", [32, 64])
def test_fused_core_matches_composed_fp32_same_precision(coords, pm, ac, D):
    """Same precision on both sides: the fused core against softmax + sampling-point arithmetic in torch fp32 on the GPU
    followed by the unfused operator (whose parity with the reference is pinned by the golden vectors).  BASELINE bar:
    forward rtol 1e-5 / atol 1e-6 (per element, relative to sum |w v|), backward rtol 1e-4 (atol 1e-5 max|ref|)."""
    from msda_triton import kernels, multiscale_deformable_attention
    from msda_triton.frontend import fused_module_core
    g = torch.Generator().manual_seed(77 + coords)
    B, Q, H, L, K = 2, 333, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    value = torch.randn(B, npix, H, D, generator=g).cuda()
    proj = torch.randn(B, Q, H, L, K, 3, generator=g)
    proj[..., :2] *= 3.0
    proj = proj.cuda()
    ref = torch.rand(B, Q, coords, generator=g)
    if coords == 4:
        ref[..., 2:] = ref[..., 2:] * 0.5 + 0.05
    ref = ref.cuda()
    go = torch.rand(B, Q, H, D, generator=g).cuda()
    shapes = torch.tensor(BENCH_PYRAMID, device="cuda")

    def composed(v, p, r):
        off, logit = p[..., :2], p[..., 2]
        aw = logit.reshape(B, Q, H, L * K).softmax(-1).reshape(B, Q, H, L, K)
        anchor = r[:, :, None, None, None, :]
        pts = anchor + off / shapes[:, None, :] if coords == 2 else anchor[..., :2] + off * anchor[..., 2:] / (2 * K)
        return multiscale_deformable_attention(v, shapes, pts, aw, pm, ac), pts.detach(), aw.detach()

    a, b, c = (t.clone().requires_grad_(True) for t in (value, proj, ref))
    want, pts, aw = composed(a, b, c)
    want.backward(go)
    x, y, z = (t.clone().requires_grad_(True) for t in (value, proj, ref))
    assert kernels.module_core_supported(x, y, z)
    got = fused_module_core(x, shapes, y, z, pm, ac)
    got.backward(go)
    S = to_np(multiscale_deformable_attention(value.abs(), shapes, pts, aw.abs(), pm, ac)).astype(np.float64)
    w = to_np(want).astype(np.float64)
    err = np.abs(to_np(got).astype(np.float64) - w)
    lim = 1e-5 * np.abs(w) + 1e-6 * np.maximum(1.0, S)
    assert (err <= lim).all(), f"out: {int((err > lim).sum())} elements outside 1e-5 / 1e-6; worst {float((err - lim).max()):.2e} over"
    for t, r, what in ((x.grad, a.grad, "grad_value"), (y.grad, b.grad, "grad_projection"), (z.grad, c.grad, "grad_ref")):
        r = to_np(r)
        assert_close(to_np(t), r, 1e-4, 1e-5 * np.abs(r).max(), what)


@pytest.mark.parametrize("coords", [2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("hidden", [256, 512], ids=["d32", "d64"])
def test_module_fused_equals_composed_cuda(coords, dtype, hidden):
    """Same module, same weights: fused fast path vs the composed CUDA path (MSDA_B200_FUSED_MODULE=0) run in fp32.
    (For bf16 the composed path itself rounds the sampling points to bf16 -- a quarter of a pixel on a 64-pixel level --
    so the yardstick is the fp32 composed module on the same bf16-rounded weights and inputs.)"""
    import copy
    from msda_triton import MultiscaleDeformableAttention
    torch.manual_seed(3)
    emb, heads, levels, points = 256, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img = torch.randn(2, npix, emb).to("cuda", dtype)
    queries = torch.randn(2, 200, emb).to("cuda", dtype)
    ref_pts = (torch.rand(2, 200, coords) * 0.8 + 0.1).to("cuda", dtype)
    shapes = torch.tensor(BENCH_PYRAMID, device="cuda")
    module = MultiscaleDeformableAttention(emb, hidden, levels, heads, points, "border", True).to("cuda", dtype)
    reference = copy.deepcopy(module).float()

    def run(mod, cast, fused):
        os.environ["MSDA_B200_FUSED_MODULE"] = "1" if fused else "0"
        try:
            i, q, r = (t.to(cast).clone().requires_grad_(True) for t in (img, queries, ref_pts))
            mod.zero_grad()
            out = mod(i, shapes, q, r)
            out.float().square().sum().backward()
            return [out.detach(), i.grad, q.grad, r.grad] + [p.grad.clone() for p in mod.parameters()]
        finally:
            os.environ.pop("MSDA_B200_FUSED_MODULE")

    got = run(module, dtype, fused=True)
    want = run(reference, torch.float32, fused=False)
    # Gradients that pass through the sampling OFFSETS are discontinuous at pixel boundaries, and bf16-quantised
    # reference points / offsets land on such boundaries often (few mantissa bits): fp32 and fp64 evaluation then pick
    # different cells for a small fraction of the points.  Hence: a tolerance for the bulk, plus an outlier budget.
    rtol, atol, frac = (2e-4, 2e-4, 1e-4) if dtype == torch.float32 else (5e-2, 3e-2, 6e-2)
    for a, b in zip(got, want):
        b = to_np(b)
        assert a.dtype == dtype
        assert_close(to_np(a), b, rtol, atol * max(1e-3, np.abs(b).max()), "fused vs composed",
                     max_outliers=max(8, int(frac * b.size)))


@pytest.mark.parametrize("amp", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("coords", [2, 4])
def test_module_under_autocast_fused_equals_composed(amp, coords):
    """AMP training (fp32 module and inputs under torch.autocast): the reference keeps the reference points in fp32
    (anchor + offsets promotes, frontend.py:268-283) and runs the operator in fp32 (custom_fwd cast_inputs).  The fused
    core must sample the SAME locations: its operands are widened to fp32, never narrowed to the autocast dtype --
    rounding an fp32 anchor to bf16 moves it by up to 2^-9, i.e. 0.1-0.3 px on the 64-167 px levels."""
    from msda_triton import MultiscaleDeformableAttention
    torch.manual_seed(11)
    emb, heads, levels, points = 256, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img = torch.randn(2, npix, emb, device="cuda")
    queries = torch.randn(2, 300, emb, device="cuda")
    ref_pts = torch.rand(2, 300, coords, device="cuda") * 0.8 + 0.1     # fp32 anchors with bits below bf16 precision
    shapes = torch.tensor(BENCH_PYRAMID, device="cuda")
    module = MultiscaleDeformableAttention(emb, emb, levels, heads, points, "border", True).cuda()

    def run(fused):
        os.environ["MSDA_B200_FUSED_MODULE"] = "1" if fused else "0"
        try:
            i, q, r = (t.clone().requires_grad_(True) for t in (img, queries, ref_pts))
            module.zero_grad()
            with torch.autocast("cuda", dtype=amp):
                out = module(i, shapes, q, r)
            out.float().square().sum().backward()
            return [out.detach().float(), i.grad, q.grad, r.grad]
        finally:
            os.environ.pop("MSDA_B200_FUSED_MODULE")

    got, want = run(True), run(False)
    # both paths see the same autocast-rounded projections and the same fp32 anchors: they differ only by fp32
    # summation order (and the rare floor-boundary tie in the offset gradients)
    for a, b, what in zip(got, want, ("out", "grad_img", "grad_queries", "grad_reference_points")):
        b = to_np(b)
        assert_close(to_np(a), b, 2e-3, 2e-3 * max(1e-3, np.abs(b).max()), f"autocast {what}",
                     max_outliers=max(8, int(2e-3 * b.size)))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("coords", [2, 4])
@pytest.mark.parametrize("D", [32, 64])
def test_fused_core_16bit_storage(dtype, coords, D):
    """16-bit storage, fp32 compute: out and grad_value (continuous in the inputs) within one storage rounding of the
    fp64 composed reference on the same rounded inputs; offset / reference gradients within the same bound except for
    the boundary-tie outliers described above (< 2 % of the offset gradients; each reference-point gradient sums 128 of
    them, so up to 10 % of those may contain one)."""
    from msda_triton.frontend import fused_module_core
    g = torch.Generator().manual_seed(5)
    B, Q, H, L, K = 2, 333, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    value = torch.randn(B, npix, H, D, generator=g).to(dtype)
    proj = torch.randn(B, Q, H, L, K, 3, generator=g)
    proj[..., :2] *= 3.0
    proj = proj.to(dtype)
    ref = torch.rand(B, Q, coords, generator=g).to(dtype)
    go = torch.rand(B, Q, H, D, generator=g).to(dtype)
    shapes = torch.tensor(BENCH_PYRAMID)
    a, b, c = (t.double().requires_grad_(True) for t in (value, proj, ref))
    want = composed_reference(a, shapes, b, c, "border", True)
    want.backward(go.double())
    x, y, z = (t.cuda().requires_grad_(True) for t in (value, proj, ref))
    got = fused_module_core(x, shapes.cuda(), y, z, "border", True)
    got.backward(go.cuda())
    eps = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    for t, r, what, frac in ((got, want, "out", 0.0), (x.grad, a.grad, "grad_value", 0.0),
                             (y.grad, b.grad, "grad_projection", 0.02), (z.grad, c.grad, "grad_ref", 0.10)):
        r = to_np(r)
        assert t.dtype == dtype
        assert_close(to_np(t), r, eps, eps * 2e-2 * np.abs(r).max(), f"{dtype} {what}", max_outliers=int(frac * r.size))


def test_fused_core_rejects_unsupported_and_falls_back():
    from msda_triton import MultiscaleDeformableAttention, kernels
    value = torch.randn(1, 85, 2, 4, device="cuda")           # head_dim 4: not covered by the fused kernels
    proj = torch.randn(1, 5, 2, 4, 8, 3, device="cuda")
    ref = torch.rand(1, 5, 2, device="cuda")
    assert not kernels.module_core_supported(value, proj, ref)
    with pytest.raises(ValueError):
        kernels.b200_module_core_fwd(value, torch.tensor([(8, 8), (4, 4), (2, 2), (1, 1)], device="cuda"), proj, ref,
                                     "zeros", False)
    module = MultiscaleDeformableAttention(64, 8, 4, 2, 8, "border", True).cuda()      # falls back to the composed path
    out = module(torch.randn(1, 85, 64, device="cuda"), torch.tensor([(8, 8), (4, 4), (2, 2), (1, 1)], device="cuda"),
                 torch.randn(1, 5, 64, device="cuda"), ref)
    assert out.shape == (1, 5, 64)


@pytest.mark.parametrize("name", ["module_ref2d_float64", "module_ref4d_float64", "module_ref2d_hd64_float64"])
@pytest.mark.parametrize("dtype,fused", [(torch.float64, False), (torch.float32, True), (torch.float32, False)],
                         ids=["f64-composed", "f32-fused", "f32-composed"])
def test_module_matches_reference_module_golden_cuda(name, dtype, fused):
    """The reference module's golden output / gradients (seeded weights, CPU, fp64) against our module on CUDA: the
    composed path in fp64 (generic kernels), and the fused and composed paths in fp32."""
    from conftest import GOLDEN
    from util import check_module_against_golden, load_module_golden
    os.environ["MSDA_B200_FUSED_MODULE"] = "1" if fused else "0"
    try:
        g, module, inputs, shapes = load_module_golden(GOLDEN / f"{name}.npz", device="cuda", dtype=dtype)
        if dtype == torch.float64:
            check_module_against_golden(g, module, inputs, shapes, rtol=1e-8, atol_scale=1e-10)
        else:
            check_module_against_golden(g, module, inputs, shapes, rtol=2e-4, atol_scale=2e-5, max_outliers=2)
    finally:
        os.environ.pop("MSDA_B200_FUSED_MODULE")


@pytest.mark.parametrize("coords", [2, 4])
def test_compiled_module_keeps_the_fused_core(coords):
    """torch.compile(fullgraph=True) of the nn.Module: the fused core runs behind torch.ops.msda_b200.module_forward /
    module_backward (no graph break, no composed fallback) and reproduces the eager fused path."""
    import copy
    from msda_triton import MultiscaleDeformableAttention
    torch.manual_seed(11)
    emb, heads, levels, points = 256, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img = torch.randn(2, npix, emb, device="cuda")
    queries = torch.randn(2, 150, emb, device="cuda")
    ref_pts = torch.rand(2, 150, coords, device="cuda") * 0.8 + 0.1
    shapes = torch.tensor(BENCH_PYRAMID, device="cuda")
    module = MultiscaleDeformableAttention(emb, emb, levels, heads, points, "zeros", False).cuda()
    twin = copy.deepcopy(module)

    def run(mod, fn):
        i, q, r = (t.clone().requires_grad_(True) for t in (img, queries, ref_pts))
        out = fn(i, shapes, q, r)
        out.square().sum().backward()
        return [out.detach(), i.grad, q.grad, r.grad] + [p.grad for p in mod.parameters()]

    want = run(module, module)
    calls = {"fwd": 0}
    from msda_triton import kernels
    real = kernels.b200_module_core_fwd

    def counting(*args, **kwargs):
        calls["fwd"] += 1
        return real(*args, **kwargs)

    kernels.b200_module_core_fwd = counting
    try:
        got = run(twin, torch.compile(twin, backend="aot_eager", fullgraph=True))
    finally:
        kernels.b200_module_core_fwd = real
    assert calls["fwd"] == 1                      # the compiled program went through the fused custom op
    for a, b in zip(got, want):
        b = to_np(b)
        assert_close(to_np(a), b, 2e-4, 2e-5 * max(1e-3, np.abs(b).max()), "compiled vs eager", max_outliers=8)


def test_module_core_custom_op_opcheck():
    value = torch.randn(1, 85, 8, 32, device="cuda", requires_grad=True)
    proj = torch.randn(1, 40, 8, 4, 4, 3, device="cuda", requires_grad=True)
    ref = torch.rand(1, 40, 2, device="cuda", requires_grad=True)
    shapes = torch.tensor([(8, 8), (4, 4), (2, 2), (1, 1)], device="cuda")
    import msda_triton.ops  # noqa: F401  (registers the ops)
    torch.library.opcheck(torch.ops.msda_b200.module_forward.default, (value, shapes, proj, ref, "border", True),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))


# ---------------------------------------------------------------------------------------------------------------------
# value projection + core as one autograd node (16-bit parameters): img_input_proj's bias gradient comes out of the
# core's rounding pass (MSDA_BWD_VALUE_COLSUM) instead of a reduction kernel of its own
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("D", [32, 64])
@pytest.mark.parametrize("needs_value", [True, False])
def test_rounding_pass_column_sums(dtype, D, needs_value):
    """The fourth result of the fused backward: fp32 sums over (batch, pixel) of the UNROUNDED grad_value."""
    from msda_triton import kernels
    g = torch.Generator().manual_seed(41)
    B, Q, H, L, K = 3, 211, 8, 4, 4
    pyramid = [(25, 42), (13, 21), (7, 11), (4, 6)]
    npix = sum(h * w for h, w in pyramid)
    value = torch.randn(B, npix, H, D, generator=g).to("cuda", dtype)
    proj = torch.randn(B, Q, H, L, K, 3, generator=g).to("cuda", dtype)
    ref = torch.rand(B, Q, 2, generator=g).to("cuda", dtype)
    go = torch.randn(B, Q, H, D, generator=g).to("cuda", dtype)
    shapes = torch.tensor(pyramid, device="cuda")
    assert kernels.module_value_colsum_supported(value)
    plain = kernels.b200_module_core_bwd(go, value, shapes, proj, ref, "zeros", False, needs=(True, True, True))
    res = kernels.b200_module_core_bwd(go, value, shapes, proj, ref, "zeros", False,
                                       needs=(needs_value, True, True), value_colsum=True)
    if not needs_value:
        assert len(res) == 3 and res[0] is None      # no grad_value, no sums
        return
    gvalue, gproj, gref, colsum = res
    assert torch.equal(gproj, plain[1])                                   # no atomics there: same bits
    for a, b in ((gvalue, plain[0]), (gref, plain[2])):                   # atomics: the order of the fp32 adds varies
        assert_close(to_np(a), to_np(b), 2.0 ** -7, 2.0 ** -7 * float(b.float().abs().max()), "repeat run")
    assert colsum.shape == (H, D) and colsum.dtype == torch.float32
    # yardstick: fp64 sum of the fp32 gradient of the same operands (no 16-bit rounding of the addends on either side)
    v32, p32, r32, g32 = (t.float() for t in (value, proj, ref, go))
    want = kernels.b200_module_core_bwd(g32, v32, shapes, p32, r32, "zeros", False, needs=(True, False, False))[0]
    want = want.double().sum((0, 1))
    assert_close(to_np(colsum), to_np(want), 1e-4, 1e-4 * float(want.abs().max()), "column sums")
    # and it is what summing the rounded tensor gives, up to that rounding
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    rounded = gvalue.double().sum((0, 1))
    bound = eps * gvalue.double().abs().sum((0, 1)) + 1e-6
    assert bool(((colsum.double() - rounded).abs() <= bound).all())


def test_rounding_pass_column_sums_without_queries():
    from msda_triton import kernels
    value = torch.randn(1, 85, 8, 32, device="cuda").bfloat16()
    proj = torch.empty(1, 0, 8, 4, 4, 3, device="cuda", dtype=torch.bfloat16)
    ref = torch.empty(1, 0, 2, device="cuda", dtype=torch.bfloat16)
    go = torch.empty(1, 0, 8, 32, device="cuda", dtype=torch.bfloat16)
    shapes = torch.tensor([(8, 8), (4, 4), (2, 2), (1, 1)], device="cuda")
    gvalue, _, _, colsum = kernels.b200_module_core_bwd(go, value, shapes, proj, ref, "border", True, value_colsum=True)
    assert float(gvalue.float().abs().max()) == 0.0 and float(colsum.abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=["bf16", "f16"])
@pytest.mark.parametrize("hidden", [256, 512], ids=["d32", "d64"])
@pytest.mark.parametrize("coords", [2, 4])
def test_module_value_projection_node_equals_separate_nodes(dtype, hidden, coords, monkeypatch):
    """Same bf16 / fp16 module and inputs with and without the fused value-projection node: identical forward, the same
    gradients (the two GEMMs are the ones autograd issues; the bias gradient is summed before instead of after the
    rounding to 16 bits)."""
    from msda_triton import MultiscaleDeformableAttention, frontend
    torch.manual_seed(5)
    emb, heads, levels, points = 256, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img = torch.randn(2, npix, emb).to("cuda", dtype)
    queries = torch.randn(2, 300, emb).to("cuda", dtype)
    ref_pts = (torch.rand(2, 300, coords) * 0.8 + 0.1).to("cuda", dtype)
    shapes = torch.tensor(BENCH_PYRAMID, device="cuda")
    module = MultiscaleDeformableAttention(emb, hidden, levels, heads, points, "zeros", False).to("cuda", dtype)
    gout = torch.randn(2, 300, emb).to("cuda", dtype)
    calls = []
    real = frontend.fused_value_proj_core
    monkeypatch.setattr(frontend, "fused_value_proj_core", lambda *a, **k: (calls.append(1), real(*a, **k))[1])

    def run(fused_value_proj):
        monkeypatch.setenv("MSDA_B200_FUSED_VALUE_PROJ", "1" if fused_value_proj else "0")
        i, q = (t.clone().requires_grad_(True) for t in (img, queries))
        module.zero_grad()
        out = module(i, shapes, q, ref_pts)
        out.backward(gout)
        return [out.detach(), i.grad, q.grad] + [p.grad.clone() for p in module.parameters()]

    names = ["out", "grad_img", "grad_queries"] + ["grad " + n for n, _ in module.named_parameters()]
    got = run(True)
    assert len(calls) == 1
    want = run(False)
    assert len(calls) == 1
    assert torch.equal(got[0], want[0])
    eps = 2.0 ** -8 if dtype == torch.bfloat16 else 2.0 ** -11
    for name, a, b in zip(names[1:], got[1:], want[1:]):
        b = to_np(b)
        assert a.dtype == dtype
        assert_close(to_np(a), b, 4 * eps, 4 * eps * np.abs(b).max(), name)

    # not taken: inference, frozen bias, autocast
    with torch.no_grad():
        module(img, shapes, queries, ref_pts)
    module.img_input_proj.bias.requires_grad_(False)
    module(img, shapes, queries, ref_pts)
    assert len(calls) == 1
