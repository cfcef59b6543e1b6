"""GPU tests of the public surface: autograd Function, AMP contract, module, reference-style usage."""
import itertools

import numpy as np
import pytest
import torch

from util import BENCH_PYRAMID, assert_close, make_inputs, to_np

pytestmark = pytest.mark.gpu
MODES = list(itertools.product(("zeros", "border"), (False, True)))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("pm,ac", MODES)
def test_autograd_backward_like_reference_suite(dtype, pm, ac):
    """Same structure as the reference's test_backward (tests/test_msda.py:121-151) on its fixture shape
    (B=4, H=8, C=32, L=4, N=1000, P=3), against the oracle instead of torch.compile'd HF code."""
    import msda_triton
    from msda_triton.frontend import triton_multiscale_deformable_attention
    from oracle import msda_oracle
    img, s, pts, aw, go = make_inputs(4, 1000, 8, 32, BENCH_PYRAMID, 3, dtype=dtype, seed=1)
    a, b, c = (t.cuda().requires_grad_(True) for t in (img, pts, aw))
    out = triton_multiscale_deformable_attention(a, s.cuda(), b, c, pm, ac)
    out.backward(go.cuda())
    ref = (msda_oracle.forward(img, s, pts, aw, pm, ac),) + msda_oracle.backward(go, img, s, pts, aw, pm, ac)
    tol = (1e-5, 1e-6, 1e-4, 1e-5) if dtype == torch.float32 else (1e-8, 1e-8, 1e-8, 1e-8)
    assert_close(to_np(out), ref[0], tol[0], tol[1], "out")
    for t, r, what in ((a.grad, ref[1], "grad_img"), (b.grad, ref[2], "grad_points"), (c.grad, ref[3], "grad_weights")):
        assert_close(to_np(t), r, tol[2], tol[3] * max(1.0, np.abs(r).max()), what)
    assert out.dtype == dtype and a.grad.shape == img.shape


def test_partial_requires_grad():
    import msda_triton
    img, s, pts, aw, go = make_inputs(2, 100, 8, 32, BENCH_PYRAMID, 4, seed=2)
    a = img.cuda().requires_grad_(True)
    out = msda_triton.multiscale_deformable_attention(a, s.cuda(), pts.cuda(), aw.cuda(), "zeros", False)
    out.backward(go.cuda())
    assert a.grad is not None and float(a.grad.abs().max()) > 0
    with torch.no_grad():
        out2 = msda_triton.multiscale_deformable_attention(a, s.cuda(), pts.cuda(), aw.cuda(), "zeros", False)
    assert torch.equal(out, out2) and not out2.requires_grad


@pytest.mark.parametrize("amp_dtype", [torch.float16, torch.bfloat16])
def test_autocast_runs_in_fp32(amp_dtype):
    """custom_fwd(cast_inputs=float32): under autocast the op computes and returns fp32 (frontend.py:111)."""
    import msda_triton
    from oracle import msda_oracle
    img, s, pts, aw, go = make_inputs(2, 64, 8, 32, BENCH_PYRAMID, 4, seed=6)
    a = img.cuda().to(amp_dtype).requires_grad_(True)
    with torch.amp.autocast(device_type="cuda", dtype=amp_dtype):
        out = msda_triton.multiscale_deformable_attention(a, s.cuda(), pts.cuda(), aw.cuda(), "zeros", False)
    assert out.dtype == torch.float32
    out.backward(go.cuda())
    assert a.grad.dtype == amp_dtype
    ref = msda_oracle.forward(a.detach().float().cpu(), s, pts, aw, "zeros", False)
    assert_close(to_np(out), ref, 1e-5, 1e-6, "autocast out")


def test_mixed_dtypes_are_promoted():
    import msda_triton
    img, s, pts, aw, _ = make_inputs(1, 16, 2, 32, BENCH_PYRAMID, 4, seed=6)
    out = msda_triton.multiscale_deformable_attention(img.cuda().half(), s.cuda(), pts.cuda(), aw.cuda().half(),
                                                      "border", True)
    assert out.dtype == torch.float32


@pytest.mark.parametrize("coords", [2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_module_cuda_matches_cpu_route(coords, dtype):
    """The module on CUDA (our kernels) equals the same module on CPU (torch route) -- Grounding-DINO decoder config
    (BASELINE.json configs[3]: emb 256, H=8, L=4, K=4, border, align_corners=True)."""
    from msda_triton import MultiscaleDeformableAttention
    torch.manual_seed(0)
    emb, heads, levels, points = 256, 8, 4, 4
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img = torch.randn(2, npix, emb)
    queries = torch.randn(2, 90, emb)
    ref_pts = torch.rand(2, 90, coords) * 0.8 + 0.1
    shapes = torch.tensor(BENCH_PYRAMID)
    module = MultiscaleDeformableAttention(emb, emb, levels, heads, points, "border", True)
    want = module(img, shapes, queries, ref_pts)
    m2 = module.to("cuda", dtype)
    got = m2(img.to("cuda", dtype), shapes.cuda(), queries.to("cuda", dtype), ref_pts.to("cuda", dtype))
    assert got.dtype == dtype
    tol = 2e-4 if dtype == torch.float32 else 6e-2
    assert_close(to_np(got), to_np(want), tol, tol, "module output")
    got.float().sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m2.parameters())


def test_readme_example():
    """README.md:123-148 of the reference, verbatim shapes."""
    from msda_triton import multiscale_deformable_attention
    batch, head_dim, num_queries, num_heads, num_points = 2, 32, 900, 8, 4
    img_shapes = [(64, 64), (32, 32), (16, 16), (8, 8)]
    num_pixels = sum(h * w for h, w in img_shapes)
    device = "cuda"
    img = torch.randn(batch, num_pixels, num_heads, head_dim, device=device)
    shapes = torch.tensor(img_shapes, device=device)
    pts = torch.rand(batch, num_queries, num_heads, len(img_shapes), num_points, 2, device=device)
    aw = torch.rand(batch, num_queries, num_heads, len(img_shapes), num_points, device=device)
    out = multiscale_deformable_attention(img, shapes, pts, aw, "zeros", False)
    assert out.shape == (batch, num_queries, num_heads, head_dim)


def test_cuda_graph_capture():
    """No allocation, sync or default-stream use inside the library: forward+backward capture into a CUDA graph."""
    from msda_triton import kernels as K
    img, s, pts, aw, go = (t.cuda() for t in make_inputs(2, 256, 8, 32, BENCH_PYRAMID, 4, seed=12))
    eager_out = K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "border", True)
    eager_g = K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "border", True, deterministic=True)
    torch.cuda.synchronize()
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "border", True)   # warm the allocator
    torch.cuda.current_stream().wait_stream(stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "border", True)
        g = K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "border", True, deterministic=True)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager_out)
    for a, b in zip(g, eager_g):
        assert torch.equal(a, b)


def test_cuda_graph_capture_of_a_paced_multi_wave_launch(monkeypatch):
    """Multi-wave launches zero an arrival counter on the stream (memset node) and pace their CTAs on it: the pair
    captures into a CUDA graph and replays (the graph reuses its counter slot)."""
    from msda_triton import kernels as K
    monkeypatch.setenv("MSDA_B200_SLICES_PER_WAVE", "1")           # 16 waves on a full persistent grid
    monkeypatch.setenv("MSDA_B200_WAVE_PACING", "2")               # pace although these waves are small
    from msda_triton import _lib
    _lib.reload_tuning()
    img, s, pts, aw, go = (t.cuda() for t in make_inputs(2, 3000, 8, 32, BENCH_PYRAMID, 4, seed=14))
    eager_out = K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "zeros", False)
    eager_g = K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "zeros", False)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "zeros", False)
        g = K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "zeros", False)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager_out)
    assert torch.equal(g[1], eager_g[1]) and torch.equal(g[2], eager_g[2])
    torch.testing.assert_close(g[0], eager_g[0], rtol=1e-4, atol=1e-4)


def test_host_pipeline_matches_device_path():
    """msda_triton.host.HostMsda (pinned host buffers, chunked H2D / kernels / D2H overlap) == plain device calls."""
    from msda_triton import kernels as K
    from msda_triton.host import HostMsda
    img, s, pts, aw, go = make_inputs(3, 500, 8, 32, BENCH_PYRAMID, 4, seed=19, points="wide")
    pin = [t.pin_memory() for t in (img, pts, aw, go)]
    h_out = torch.empty(3, 500, 8, 32).pin_memory()
    h_gi, h_gp, h_ga = (torch.empty_like(t).pin_memory() for t in (img, pts, aw))
    pipe = HostMsda(3, img.shape[1], 8, 32, 500, 4, 4)
    for _ in range(3):      # back-to-back calls overlap (H2D of call i+1 with D2H of call i)
        pipe.run(pin[0], s.cuda(), pin[1], pin[2], "zeros", False, h_out, pin[3], h_gi, h_gp, h_ga, deterministic=True)
    pipe.synchronize()
    d = [t.cuda() for t in (img, s, pts, aw, go)]
    out = K.b200_multi_scale_deformable_attention_fwd(d[0], d[1], d[2], d[3], "zeros", False)
    gi, gp, ga = K.b200_multi_scale_deformable_attention_bwd(d[4], d[0], d[1], d[2], d[3], "zeros", False, deterministic=True)
    assert torch.equal(h_out, out.cpu()) and torch.equal(h_gp, gp.cpu()) and torch.equal(h_ga, ga.cpu())
    assert torch.equal(h_gi, gi.cpu())
    with pytest.raises(ValueError):
        pipe.run(img, s.cuda(), pin[1], pin[2], "zeros", False, h_out)      # unpinned input


def test_torch_library_op_and_compile():
    """torch.ops.msda_b200.forward: opcheck (schema, fake kernel, autograd registration) and a fullgraph compile."""
    import msda_triton
    from msda_triton.ops import multiscale_deformable_attention_op
    img, s, pts, aw, go = (t.cuda() for t in make_inputs(2, 128, 8, 32, BENCH_PYRAMID, 4, seed=23))
    a, b, c = (t.clone().requires_grad_(True) for t in (img, pts, aw))
    torch.library.opcheck(torch.ops.msda_b200.forward.default, (a, s, b, c, "zeros", False),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))

    def model(img, pts, aw):
        out = msda_triton.multiscale_deformable_attention(img * 1.0, s, pts, aw, "zeros", False)
        return out.sum(dim=-1)

    eager = model(a, b, c)
    eager.backward(go.sum(-1))
    want = [eager.detach().clone(), a.grad.clone(), b.grad.clone(), c.grad.clone()]
    a.grad = b.grad = c.grad = None
    compiled = torch.compile(model, backend="aot_eager", fullgraph=True)
    got = compiled(a, b, c)
    got.backward(go.sum(-1))
    torch.testing.assert_close(got, want[0], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(b.grad, want[2], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(c.grad, want[3], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(a.grad, want[1], rtol=1e-4, atol=1e-4)
    out2 = multiscale_deformable_attention_op(img, s, pts, aw, "zeros", False)
    assert torch.equal(out2, msda_triton.multiscale_deformable_attention(img, s, pts, aw, "zeros", False))
