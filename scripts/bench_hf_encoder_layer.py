"""A Hugging Face Deformable-DETR encoder layer (d_model 256, 8 heads, 4 levels, 4 points; random weights) at the
800x1333 pyramid (22 223 pixels, B=2): forward+backward with HF's own pure-PyTorch operator (grid_sample per level)
versus this package's CUDA operator patched in by msda_triton.integrations.patch_transformers
(tests/test_hf_dropin_gpu.py checks that results agree).
Run on the GPU box."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "msda-triton_b200"))
from msda_triton import integrations  # noqa: E402
from transformers import DeformableDetrConfig, ResNetConfig  # noqa: E402
from transformers.models.deformable_detr import modeling_deformable_detr as m  # noqa: E402

PYRAMID = [(100, 167), (50, 84), (25, 42), (13, 21)]


def median_ms(fn, steps=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    config = DeformableDetrConfig(
        backbone_config=ResNetConfig(out_features=["stage2", "stage3", "stage4"]), use_timm_backbone=False,
        use_pretrained_backbone=False, d_model=256, encoder_layers=1, decoder_layers=1, num_queries=300,
        encoder_attention_heads=8, num_feature_levels=4, encoder_n_points=4, dropout=0.0, attention_dropout=0.0,
        activation_dropout=0.0)
    torch.manual_seed(0)
    for dt in (torch.float32, torch.bfloat16):
        layer = m.DeformableDetrEncoderLayer(config).to("cuda", dt).train()
        batch, npix = 2, sum(h * w for h, w in PYRAMID)
        hidden = torch.randn(batch, npix, 256, device="cuda", dtype=dt, requires_grad=True)
        pos = (torch.randn(batch, npix, 256, device="cuda") * 0.1).to(dt)
        shapes = torch.tensor(PYRAMID, device="cuda")
        level_start = torch.cat((shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]))
        ref = torch.rand(batch, npix, 4, 2, device="cuda").to(dt)
        gout = torch.randn(batch, npix, 256, device="cuda").to(dt)

        def step():
            out = layer(hidden, attention_mask=None, spatial_position_embeddings=pos, reference_points=ref,
                        spatial_shapes=shapes, spatial_shapes_list=PYRAMID, level_start_index=level_start)
            (out[0] if isinstance(out, tuple) else out).backward(gout)
            hidden.grad = None

        torch.cuda.reset_peak_memory_stats()
        native = median_ms(step)
        native_mem = torch.cuda.max_memory_allocated() / 2 ** 20
        integrations.patch_transformers(["deformable_detr"])
        try:
            torch.cuda.reset_peak_memory_stats()
            patched = median_ms(step)
            patched_mem = torch.cuda.max_memory_allocated() / 2 ** 20
        finally:
            integrations.unpatch_transformers()
        print(f"{dt}: encoder layer fwd+bwd  HF operator {native:.2f} ms (peak {native_mem:.0f} MB)  ->  this package "
              f"{patched:.2f} ms (peak {patched_mem:.0f} MB)  x{native / patched:.1f}", flush=True)


if __name__ == "__main__":
    main()
