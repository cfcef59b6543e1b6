"""Drop-in check on a real Deformable-DETR block (the reference README's "Detection Example", README.md:25-37, replaces
the operator inside a Hugging Face model and observes identical detections).  Here: the Hugging Face
``DeformableDetrEncoderLayer`` (random weights, no download) runs once with its own pure-PyTorch operator
(``grid_sample`` per level, padding_mode="zeros", align_corners=False) and once with this package's CUDA operator
patched in by ``msda_triton.integrations.patch_transformers``; output and all gradients must agree.  Skipped when ``transformers`` cannot build the layer offline."""
import pytest
import torch

pytestmark = pytest.mark.gpu

PYRAMID = [(24, 32), (12, 16), (6, 8), (3, 4)]


def _build_layer():
    try:
        from transformers import DeformableDetrConfig, ResNetConfig
        from transformers.models.deformable_detr import modeling_deformable_detr as m
        config = DeformableDetrConfig(
            backbone_config=ResNetConfig(out_features=["stage2", "stage3", "stage4"]), use_timm_backbone=False,
            use_pretrained_backbone=False, d_model=256, encoder_layers=1, decoder_layers=1, num_queries=30,
            encoder_attention_heads=8, num_feature_levels=4, encoder_n_points=4, dropout=0.0, attention_dropout=0.0,
            activation_dropout=0.0)
        torch.manual_seed(0)
        return m, m.DeformableDetrEncoderLayer(config)
    except Exception as e:  # noqa: BLE001 -- optional dependency / API drift
        pytest.skip(f"transformers DeformableDetrEncoderLayer unavailable offline: {type(e).__name__}: {e}")


def test_hf_deformable_detr_encoder_layer_with_our_operator():
    m, layer = _build_layer()
    layer = layer.cuda().train()
    with torch.no_grad():   # HF initialises the offset bias to a fixed grid; spread the offsets over a few pixels
        layer.self_attn.sampling_offsets.weight.normal_(0.0, 0.05)
    batch, npix = 2, sum(h * w for h, w in PYRAMID)
    g = torch.Generator(device="cuda").manual_seed(1)
    hidden = torch.randn(batch, npix, 256, device="cuda", generator=g)
    pos = torch.randn(batch, npix, 256, device="cuda", generator=g) * 0.1
    shapes = torch.tensor(PYRAMID, device="cuda")
    level_start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
    # encoder reference points: every pixel's own normalised centre, replicated for the 4 levels
    centres = []
    for h, w in PYRAMID:
        ys, xs = torch.meshgrid(torch.arange(h, device="cuda"), torch.arange(w, device="cuda"), indexing="ij")
        centres.append(torch.stack(((xs.reshape(-1) + 0.5) / w, (ys.reshape(-1) + 0.5) / h), -1))
    ref = torch.cat(centres)[None, :, None, :].expand(batch, npix, len(PYRAMID), 2).contiguous()
    grad_out = torch.randn(batch, npix, 256, device="cuda", generator=g)

    def run():
        x = hidden.clone().requires_grad_(True)
        layer.zero_grad()
        out = layer(x, attention_mask=None, spatial_position_embeddings=pos, reference_points=ref, spatial_shapes=shapes,
                    spatial_shapes_list=PYRAMID, level_start_index=level_start)
        out = out[0] if isinstance(out, tuple) else out
        out.backward(grad_out)
        return [out.detach(), x.grad] + [p.grad.clone() for p in layer.parameters()]

    want = run()

    from msda_triton import integrations, kernels
    calls = {"n": 0}
    real = kernels.b200_multi_scale_deformable_attention_fwd

    def counting(*args, **kwargs):
        calls["n"] += 1
        return real(*args, **kwargs)

    kernels.b200_multi_scale_deformable_attention_fwd = counting
    assert "deformable_detr" in integrations.patch_transformers()
    try:
        got = run()
    finally:
        integrations.unpatch_transformers()
        kernels.b200_multi_scale_deformable_attention_fwd = real
    assert calls["n"] == 1                                  # the layer went through the CUDA operator
    assert m.MultiScaleDeformableAttention.forward.__name__ == "forward"          # and the original is back
    for a, b in zip(got, want):
        torch.testing.assert_close(a, b, rtol=2e-4, atol=2e-4 * float(b.abs().max()))
