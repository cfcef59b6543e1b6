"""Generate tests/golden/*.npz from the UNMODIFIED reference -- run in the build container only.

    python oracle/make_golden.py            # needs /root/reference (read-only); writes tests/golden/

Two reference executions are recorded for every seeded case and every (padding_mode, align_corners):

* ``triton_*``: the reference's own Triton kernels
  (``/root/reference/src/msda_triton/kernels.py:267-348`` forward, ``:396-553`` backward), executed on CPU by Triton's
  interpreter (``TRITON_INTERPRET=1``).  The kernels are launched exactly as the reference's wrappers launch them
  (``kernels.py:365-378`` / ``:575-590``: grid ``[N, B, H]``, next-power-of-two block sizes) but through the
  autotuner's inner ``.fn`` because the autotuner itself needs a GPU to time configs.  Interpreter arithmetic is numpy:
  same op order as the source, no FMA contraction.
* ``native_*``: the reference's torch route ``native_multiscale_deformable_attention`` (``frontend.py:15-68``) and its
  autograd gradients.

The reference cannot travel to the GPU box, so the vectors are committed; this script is the provenance.
"""
from __future__ import annotations

import importlib.util
import itertools
import os
import sys
import types
from pathlib import Path

os.environ["TRITON_INTERPRET"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402
import triton  # noqa: E402

REF = Path("/root/reference/src/msda_triton")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def load_reference():
    pkg = types.ModuleType("msda_triton")
    pkg.__path__ = [str(REF)]
    sys.modules["msda_triton"] = pkg  # bypasses __init__.py:5 (needs pip metadata)
    mods = {}
    for name in ("kernels", "frontend"):
        spec = importlib.util.spec_from_file_location(f"msda_triton.{name}", REF / f"{name}.py")
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"msda_triton.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["kernels"], mods["frontend"]


def ref_triton_fwd(k, img, shapes, pts, aw, pm, ac):
    B, I, H, C = img.shape
    _, N, _, L, P, _ = pts.shape
    out = img.new_empty(B, N, H, C)
    k.triton_multi_scale_deformable_attention_fwd_kernel.fn[N, B, H](
        out, img.contiguous(), pts.contiguous(), aw.contiguous(), shapes.contiguous(),
        B, I, C, N, H, L, P,
        triton.next_power_of_2(C), triton.next_power_of_2(L), triton.next_power_of_2(P), pm, ac)
    return out


def ref_triton_bwd(k, go, img, shapes, pts, aw, pm, ac):
    B, I, H, C = img.shape
    _, N, _, L, P, _ = pts.shape
    gi, gp, ga = torch.zeros_like(img), torch.zeros_like(pts), torch.zeros_like(aw)
    k.triton_multi_scale_deformable_attention_bwd_kernel.fn[N, B, H](
        gi, gp, ga, go.contiguous(), img.contiguous(), pts.contiguous(), aw.contiguous(), shapes.contiguous(),
        B, I, C, N, H, L, P,
        triton.next_power_of_2(C), triton.next_power_of_2(L), triton.next_power_of_2(P), pm, ac)
    return gi, gp, ga


def edge_points(shapes, K):
    """Points on exact pixel centres / edges / corners for both align modes (K per level, cycled)."""
    pts = []
    for (h, w) in shapes:
        cand = [(0.0, 0.0), (1.0, 1.0), (0.5, 0.5), (0.5 / w, 0.5 / h), ((w - 0.5) / w, (h - 0.5) / h),
                (1.0 / w, 1.0 / h), (1.0 / max(w - 1, 1), 1.0 / max(h - 1, 1)), (1.0, 0.0), (0.0, 1.0),
                (-0.5 / w, 0.25), (1.0 + 0.5 / w, 0.75), (0.25, -1.0 / h), (2.5, -1.5)]
        pts.append(cand)
    return pts


CASES = {
    # name: (B, Q, H, D, shapes, K, point_mode)
    "tiny_oob": (2, 5, 2, 8, [(6, 5), (3, 4)], 3, "wide"),
    "bench_like": (1, 12, 2, 32, [(8, 8), (4, 4), (2, 2), (1, 1)], 4, "unit"),
    "detr_like": (1, 6, 4, 32, [(7, 11), (4, 6), (2, 3), (1, 2)], 4, "unit"),
    "d4_k8_far": (2, 6, 2, 4, [(8, 8), (4, 4), (2, 2), (1, 1)], 8, "far"),
    "edges": (1, 13, 1, 16, [(5, 7), (4, 4), (1, 3)], 2, "edges"),
    "odd_d": (1, 4, 3, 6, [(5, 4), (3, 3)], 5, "wide"),
}


def make_inputs(name, dtype):
    B, Q, H, D, shapes, K, mode = CASES[name]
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    L = len(shapes)
    npix = sum(h * w for h, w in shapes)
    img = torch.randn(B, npix, H, D, generator=g, dtype=torch.float64)
    if mode == "unit":
        pts = torch.rand(B, Q, H, L, K, 2, generator=g, dtype=torch.float64)
    elif mode == "wide":
        pts = torch.rand(B, Q, H, L, K, 2, generator=g, dtype=torch.float64) * 1.5 - 0.25
    elif mode == "far":
        pts = torch.randn(B, Q, H, L, K, 2, generator=g, dtype=torch.float64) * 1.5 + 0.5
    elif mode == "edges":
        cand = edge_points(shapes, K)
        pts = torch.empty(B, Q, H, L, K, 2, dtype=torch.float64)
        for q in range(Q):
            for l in range(L):
                for kk in range(K):
                    pts[:, q, :, l, kk, :] = torch.tensor(cand[l][(q + kk * 5) % len(cand[l])], dtype=torch.float64)
    aw = torch.softmax(torch.randn(B, Q, H, L * K, generator=g, dtype=torch.float64), -1).reshape(B, Q, H, L, K)
    go = torch.rand(B, Q, H, D, generator=g, dtype=torch.float64)
    shapes_t = torch.tensor(shapes, dtype=torch.int64)
    return img.to(dtype), shapes_t, pts.to(dtype), aw.to(dtype), go.to(dtype)


def make_module_golden(f):
    """The reference nn.Module (frontend.py:175-292) with seeded weights on CPU: state_dict, inputs, output and the
    gradients of a scalar loss -- pins the module-level semantics (projection interleaving, softmax over L*K, the
    (h, w) normaliser of 2-d reference points, the 4-d reference-point formula)."""
    for coords, pm, ac, hidden in ((2, "zeros", False, 64), (4, "border", True, 64), (2, "border", False, 128)):
        torch.manual_seed(100 + coords + (0 if hidden == 64 else hidden))
        emb, levels, heads, points = 16, 4, 2, 4     # head_dim 32 (hidden 64) / 64 (hidden 128), L*K = 16: fused CUDA shapes
        shapes = [(9, 12), (5, 6), (3, 3), (2, 2)]
        npix = sum(h * w for h, w in shapes)
        module = f.MultiscaleDeformableAttention(emb, hidden, levels, heads, points, pm, ac).double()
        img = torch.randn(2, npix, emb, dtype=torch.float64, requires_grad=True)
        queries = torch.randn(2, 7, emb, dtype=torch.float64, requires_grad=True)
        ref = torch.rand(2, 7, coords, dtype=torch.float64)
        if coords == 4:
            ref[..., 2:] = ref[..., 2:] * 0.4 + 0.1
        ref.requires_grad_(True)
        out = module(img, torch.tensor(shapes), queries, ref)
        out.square().sum().backward()
        rec = {f"param.{k}": v.detach().numpy() for k, v in module.state_dict().items()}
        rec.update({f"grad.{k}": p.grad.numpy() for k, p in module.named_parameters()})
        rec.update(img=img.detach().numpy(), queries=queries.detach().numpy(), reference_points=ref.detach().numpy(),
                   img_shapes=np.array(shapes), out=out.detach().numpy(), grad_img=img.grad.numpy(),
                   grad_queries=queries.grad.numpy(), grad_reference_points=ref.grad.numpy(),
                   config=np.array([emb, hidden, levels, heads, points, coords, int(ac)]), padding_mode=np.array(pm))
        path = OUT / (f"module_ref{coords}d_float64.npz" if hidden == 64 else f"module_ref{coords}d_hd64_float64.npz")
        np.savez_compressed(path, **rec)
        print(f"wrote {path}  ({path.stat().st_size / 1024:.1f} KiB)")


def main():
    k, f = load_reference()
    OUT.mkdir(parents=True, exist_ok=True)
    make_module_golden(f)
    if "--module-only" in sys.argv:
        return
    for name in CASES:
        for dtype in (torch.float32, torch.float64, torch.float16):
            if dtype == torch.float16 and name not in ("bench_like", "tiny_oob"):
                continue
            img, shapes, pts, aw, go = make_inputs(name, dtype)
            rec = {"img": img.numpy(), "img_shapes": shapes.numpy(), "sampling_points": pts.numpy(),
                   "attention_weights": aw.numpy(), "out_grad": go.numpy()}
            for pm, ac in itertools.product(("zeros", "border"), (False, True)):
                tag = f"{pm}_{int(ac)}"
                out = ref_triton_fwd(k, img, shapes, pts, aw, pm, ac)
                gi, gp, ga = ref_triton_bwd(k, go, img, shapes, pts, aw, pm, ac)
                rec[f"triton_out_{tag}"] = out.numpy()
                rec[f"triton_gimg_{tag}"] = gi.numpy()
                rec[f"triton_gpts_{tag}"] = gp.numpy()
                rec[f"triton_gaw_{tag}"] = ga.numpy()
                if dtype != torch.float16:  # CPU grid_sample has no fp16 kernel worth pinning
                    a, b, c = (t.clone().requires_grad_(True) for t in (img, pts, aw))
                    nout = f.native_multiscale_deformable_attention(a, shapes, b, c, pm, ac)
                    nout.backward(go)
                    rec[f"native_out_{tag}"] = nout.detach().numpy()
                    rec[f"native_gimg_{tag}"] = a.grad.numpy()
                    rec[f"native_gpts_{tag}"] = b.grad.numpy()
                    rec[f"native_gaw_{tag}"] = c.grad.numpy()
            dn = str(dtype).split(".")[-1]
            path = OUT / f"{name}_{dn}.npz"
            np.savez_compressed(path, **rec)
            print(f"wrote {path}  ({path.stat().st_size / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
