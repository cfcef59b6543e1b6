"""BASELINE.json configs at their STATED sizes that no other test pins:

  C5  training fwd+bwd at B=64 encoder shapes (Q = Npix = 22 223): three images of the batched result against the CPU
      oracle, every other image against the single-image CUDA result (the batched launch walks 64 L2-sized waves with
      pacing; a single image is one wave);
  C4  Grounding-DINO decoder through the nn.Module, exactly B=8, Q=900, emb 256, H=8, L=4, K=4, border /
      align_corners=True, bf16 parameters and inputs, on both pyramids (5 440 and 22 223 pixels): fused module core
      against the same module evaluated in fp64 on the CPU route (bf16 storage bound stated in the test).
Reference: /root/reference/src/msda_triton/frontend.py:145-172 (operator), :175-292 (module).
"""
import numpy as np
import pytest
import torch

from util import BENCH_PYRAMID, DETR_PYRAMID, assert_close, to_np

pytestmark = pytest.mark.gpu


def test_c5_b64_encoder_full_size():
    from msda_triton import kernels as K
    from oracle import msda_oracle
    B, H, D, Kp, pm, ac = 64, 8, 32, 4, "zeros", False
    pyr = DETR_PYRAMID
    L, npix = len(pyr), sum(h * w for h, w in pyr)
    Q = npix
    g = torch.Generator(device="cuda").manual_seed(2025)
    d = dict(device="cuda", generator=g)
    img = torch.randn(B, npix, H, D, **d)
    pts = torch.rand(B, Q, H, L, Kp, 2, **d) * 1.1 - 0.05
    aw = torch.softmax(torch.randn(B, Q, H, L * Kp, **d), -1).reshape(B, Q, H, L, Kp)
    go = torch.rand(B, Q, H, D, **d)
    shapes = torch.tensor(pyr, device="cuda")
    out = K.b200_multi_scale_deformable_attention_fwd(img, shapes, pts, aw, pm, ac)
    gi, gp, ga = K.b200_multi_scale_deformable_attention_bwd(go, img, shapes, pts, aw, pm, ac)
    torch.cuda.synchronize()
    # (a) three images against the oracle: first, one in the middle of the wave sequence, last
    for b in (0, 37, 63):
        sl = slice(b, b + 1)
        cpu = [t[sl].cpu() for t in (img, pts, aw, go)]
        ref_out = msda_oracle.forward(cpu[0], shapes.cpu(), cpu[1], cpu[2], pm, ac)
        rgi, rgp, rga = msda_oracle.backward(cpu[3], cpu[0], shapes.cpu(), cpu[1], cpu[2], pm, ac)
        assert_close(to_np(out[sl]), ref_out, 1e-5, 1e-6 * max(1.0, float(np.abs(to_np(img[sl])).max())), f"C5 image {b}: out")
        for t, r, n in ((gi, rgi, "grad_img"), (gp, rgp, "grad_points"), (ga, rga, "grad_weights")):
            assert_close(to_np(t[sl]), r, 1e-4, 1e-5 * np.abs(r).max(), f"C5 image {b}: {n}")
    # (b) every image against the single-image launch: no atomics in out / grad_points / grad_weights -> same bits
    for b in range(B):
        sl = slice(b, b + 1)
        o1 = K.b200_multi_scale_deformable_attention_fwd(img[sl], shapes, pts[sl], aw[sl], pm, ac)
        g1 = K.b200_multi_scale_deformable_attention_bwd(go[sl], img[sl], shapes, pts[sl], aw[sl], pm, ac)
        assert torch.equal(o1, out[sl]), f"image {b}: out differs from the single-image launch"
        assert torch.equal(g1[1], gp[sl]) and torch.equal(g1[2], ga[sl]), f"image {b}: grad_points / grad_weights"
        scale = float(g1[0].abs().max())
        # fp32 atomics in another order: a few ulp of the largest partial sums (rows take up to ~400 contributions)
        assert float((g1[0] - gi[sl]).abs().max()) <= 5e-6 * scale, f"image {b}: grad_img beyond atomic-order noise"


@pytest.mark.parametrize("pyramid", [BENCH_PYRAMID, DETR_PYRAMID], ids=["pyramid_5440", "pyramid_22223"])
def test_c4_gdino_decoder_module_bf16_exact_config(pyramid):
    from msda_triton import MultiscaleDeformableAttention
    torch.manual_seed(4)
    B, Q, emb, H, L, Kp = 8, 900, 256, 8, 4, 4
    npix = sum(h * w for h, w in pyramid)
    bf = torch.bfloat16
    # inputs and parameters are bf16 values; the fp64 CPU module sees exactly those values
    img = torch.randn(B, npix, emb).to(bf)
    queries = torch.randn(B, Q, emb).to(bf)
    ref_pts = (torch.rand(B, Q, 2) * 0.9 + 0.05).to(bf)
    gout = torch.rand(B, Q, emb).to(bf)
    shapes = torch.tensor(pyramid)
    mod = MultiscaleDeformableAttention(emb, emb, L, H, Kp, "border", True).to(bf)
    cpu = MultiscaleDeformableAttention(emb, emb, L, H, Kp, "border", True).double()
    cpu.load_state_dict({k: v.double() for k, v in mod.state_dict().items()})

    # The bf16 module STORES its three projection outputs and the operator output in bf16; the fp64 oracle rounds the same
    # tensors to bf16 values (straight-through gradient), so both sample the same positions (grad_sampling_points is
    # piecewise constant in the cell index: unrounded offsets would put ~1 % of the points into neighbouring cells) and
    # the comparison isolates what it should: fp32 compute on bf16 storage against exact arithmetic on the same values.
    def as_bf16_values(x):
        return x + (x.to(bf).double() - x).detach()
    cpu.img_input_proj.register_forward_hook(lambda m, i, o: as_bf16_values(o))
    cpu.query_input_proj.register_forward_hook(lambda m, i, o: as_bf16_values(o))
    cpu.query_output_proj.register_forward_pre_hook(lambda m, args: (as_bf16_values(args[0]),))
    a, b = img.double().requires_grad_(True), queries.double().requires_grad_(True)
    want = cpu(a, shapes, b, ref_pts.double())
    want.backward(gout.double())
    dev = mod.cuda()
    x, y = img.cuda().requires_grad_(True), queries.cuda().requires_grad_(True)
    got = dev(x, shapes.cuda(), y, ref_pts.cuda())
    got.backward(gout.cuda())
    assert got.dtype == bf
    # bf16 storage bound: every tensor on the way (projections, value, output) is rounded to 8 significant bits
    # (eps = 2^-7); a handful of roundings accumulate, hence 4 eps relative to the tensor's scale
    eps = 2.0 ** -7
    for name, t, r in (("out", got, want), ("grad_img", x.grad, a.grad), ("grad_queries", y.grad, b.grad)):
        r = to_np(r)
        # grad_queries runs through grad_sampling_points, which is piecewise constant in the cell index: the few points
        # that fp32 and fp64 position arithmetic put on different sides of a pixel boundary give isolated outliers
        # (budget 0.05 % of the elements; out and grad_img are continuous and get none)
        budget = r.size // 2000 if name == "grad_queries" else 0
        assert_close(to_np(t), r, 4 * eps, 4 * eps * np.abs(r).max(), f"C4 module bf16 {name}", max_outliers=budget)
    for (n, p_dev), (_, p_cpu) in zip(dev.named_parameters(), cpu.named_parameters()):
        r = to_np(p_cpu.grad)
        assert_close(to_np(p_dev.grad), r, 8 * eps, 8 * eps * max(np.abs(r).max(), 1e-30), f"C4 module bf16 grad {n}")
