"""GPU parity tests: the CUDA path (through the C ABI of libmsda_b200.so) against the CPU oracle and against the
golden vectors produced by the unmodified reference.  Tolerances are BASELINE.json's:
   fp32 forward   rtol 1e-5 / atol 1e-6
   fp32 backward  rtol 1e-4 (atol 1e-5 * max|ref| per tensor -- elements near zero need an absolute floor)
   fp64           1e-8 (the reference's own bar, tests/test_msda.py:23-26), in practice ~1e-13
   fp16 / bf16    storage bounds stated in test_16bit_storage.
"""
import itertools
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from util import BENCH_PYRAMID, DETR_PYRAMID, assert_close, knobs, make_inputs, to_np

pytestmark = pytest.mark.gpu

MODES = list(itertools.product(("zeros", "border"), (False, True)))


@pytest.fixture(scope="module")
def K():
    from msda_triton import kernels
    assert torch.cuda.is_available()
    return kernels


@pytest.fixture(scope="module")
def oracle():
    from oracle import msda_oracle
    return msda_oracle


def run_cuda(K, img, shapes, pts, aw, go, pm, ac, **kw):
    dev = "cuda"
    a, s, p, w, g = (t.to(dev) for t in (img, shapes, pts, aw, go))
    out = K.b200_multi_scale_deformable_attention_fwd(a, s, p, w, pm, ac)
    gi, gp, ga = K.b200_multi_scale_deformable_attention_bwd(g, a, s, p, w, pm, ac, **kw)
    torch.cuda.synchronize()
    return out, gi, gp, ga


def check_against(test, ref, dtype, what, gpts_outliers=0):
    out, gi, gp, ga = (to_np(t) for t in test)
    rout, rgi, rgp, rga = ref
    if dtype == torch.float32:
        assert_close(out, rout, 1e-5, 1e-6, f"{what} out")
        assert_close(gi, rgi, 1e-4, 1e-5 * np.abs(rgi).max(), f"{what} grad_img")
        assert_close(ga, rga, 1e-4, 1e-5 * np.abs(rga).max(), f"{what} grad_attention_weights")
        assert_close(gp, rgp, 1e-4, 1e-5 * np.abs(rgp).max(), f"{what} grad_sampling_points", gpts_outliers)
    else:
        assert_close(out, rout, 1e-8, 1e-8, f"{what} out")
        assert_close(gi, rgi, 1e-8, 1e-8, f"{what} grad_img")
        assert_close(ga, rga, 1e-8, 1e-8, f"{what} grad_attention_weights")
        assert_close(gp, rgp, 1e-8, 1e-8 * max(1.0, np.abs(rgp).max()), f"{what} grad_sampling_points", gpts_outliers)


# ---------------------------------------------------------------------------------------------------------------------
# golden vectors from the unmodified reference (Triton kernels under the CPU interpreter)
# ---------------------------------------------------------------------------------------------------------------------
GOLD = sorted(p for p in GOLDEN.glob("*.npz") if not p.name.endswith("float16.npz") and not p.name.startswith("module_"))


@pytest.mark.parametrize("path", GOLD, ids=lambda p: p.stem)
@pytest.mark.parametrize("pm,ac", MODES)
def test_golden_reference_kernels(K, path, pm, ac):
    g = np.load(path)
    tag = f"{pm}_{int(ac)}"
    t = {k: torch.from_numpy(g[k]) for k in ("img", "img_shapes", "sampling_points", "attention_weights", "out_grad")}
    test = run_cuda(K, t["img"], t["img_shapes"], t["sampling_points"], t["attention_weights"], t["out_grad"], pm, ac)
    ref = tuple(g[f"triton_{n}_{tag}"] for n in ("out", "gimg", "gpts", "gaw"))
    check_against(test, ref, t["img"].dtype, f"{path.stem} {tag}")


@pytest.mark.parametrize("path", sorted(GOLDEN.glob("*float16.npz")), ids=lambda p: p.stem)
@pytest.mark.parametrize("pm,ac", MODES)
def test_golden_fp16_reference_kernels(K, path, pm, ac):
    """The reference computes fp16 IN fp16 (coordinates included), we compute in fp32 from the same fp16 inputs, so
    agreement is bounded by the reference's own fp16 rounding: its test bar is atol=rtol=1e-1 (tests/test_msda.py:16-18).
    The coordinate rounding (x*w-0.5 in fp16) moves samples by up to 1/32 px, hence the looser gradient bars."""
    g = np.load(path)
    tag = f"{pm}_{int(ac)}"
    t = {k: torch.from_numpy(g[k]) for k in ("img", "img_shapes", "sampling_points", "attention_weights", "out_grad")}
    out, gi, gp, ga = run_cuda(K, t["img"], t["img_shapes"], t["sampling_points"], t["attention_weights"],
                               t["out_grad"], pm, ac)
    assert out.dtype == torch.float16
    assert_close(to_np(out), g[f"triton_out_{tag}"].astype(np.float64), 1e-1, 1e-1, "fp16 out")
    assert_close(to_np(gi), g[f"triton_gimg_{tag}"].astype(np.float64), 1e-1, 1e-1, "fp16 grad_img")
    assert_close(to_np(ga), g[f"triton_gaw_{tag}"].astype(np.float64), 1e-1, 1e-1, "fp16 grad_aw")


# ---------------------------------------------------------------------------------------------------------------------
# oracle parity on the reference's own fixture shapes (tests/test_msda.py:30-47, :162-167) and BASELINE configs
# ---------------------------------------------------------------------------------------------------------------------
CASES = {
    # name: (B, Q, H, D, shapes, K, points, weights)
    "ref_fixture_k3": (4, 1000, 8, 32, BENCH_PYRAMID, 3, "unit", "softmax_k"),        # generic kernel (L*K = 12)
    "ref_module_d4_k8": (4, 1000, 8, 4, BENCH_PYRAMID, 8, "far", "softmax_lk"),       # D=4, far out of bounds
    "readme_c1": (2, 900, 8, 32, BENCH_PYRAMID, 4, "unit", "softmax_lk"),             # tuned kernel
    "readme_c1_wide": (2, 900, 8, 32, BENCH_PYRAMID, 4, "wide", "softmax_k"),
    "detr_small": (1, 777, 8, 32, [(25, 42), (13, 21), (7, 11), (4, 6)], 4, "wide", "softmax_lk"),
    "d64": (2, 333, 4, 64, [(20, 30), (10, 15), (5, 8), (3, 4)], 4, "wide", "softmax_lk"),
    "odd_everything": (3, 61, 5, 6, [(9, 7), (5, 4), (2, 3)], 5, "wide", "softmax_lk"),
    "one_level_one_point": (2, 50, 3, 16, [(7, 9)], 1, "far", "softmax_lk"),
    "wide_channels": (1, 40, 2, 200, [(6, 6), (3, 3)], 2, "wide", "softmax_lk"),      # D > 32*4: channel chunks
    "many_points": (1, 30, 2, 8, [(6, 6), (3, 3), (2, 2)], 19, "wide", "softmax_lk"), # L*K = 57 > 32 lanes
    "lk8_two_levels": (2, 501, 8, 32, [(20, 30), (10, 15)], 4, "wide", "softmax_lk"),  # tuned kernels, L*K = 8
    "lk8_one_level": (2, 300, 8, 32, [(24, 31)], 8, "far", "softmax_lk"),             # tuned kernels, L = 1
    "lk32_k8": (2, 403, 8, 32, BENCH_PYRAMID, 8, "wide", "softmax_lk"),               # tuned forward, L*K = 32
    "lk20_five_levels": (2, 350, 8, 32, [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3)], 4, "wide", "softmax_lk"),  # padded 24
    "lk4_one_level": (2, 350, 8, 32, [(17, 23)], 4, "wide", "softmax_lk"),            # padded 8
    "lk28_k7": (1, 222, 8, 32, BENCH_PYRAMID, 7, "wide", "softmax_lk"),               # padded 32 (forward)
    "lk12_rtdetr": (2, 300, 8, 32, [(40, 40), (20, 20), (10, 10)], 4, "unit", "softmax_lk"),  # padded 16 (L=3, K=4)
    # backward with more than 16 points per unit: sub-units of 8 / 16 slots (forward: generic beyond 32 points)
    "lk24_k8": (2, 301, 8, 32, [(40, 40), (20, 20), (10, 10)], 8, "wide", "softmax_lk"),      # 3 x 8 exact
    "lk48_six_levels": (1, 203, 8, 32, [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3), (1, 2)], 8, "wide", "softmax_lk"),
    "lk30_k10": (1, 257, 8, 32, [(40, 40), (20, 20), (10, 10)], 10, "far", "softmax_lk"),     # 2 x 16, 2 dead slots
    "lk36_k9": (1, 199, 8, 32, BENCH_PYRAMID, 9, "wide", "softmax_lk"),                       # 5 x 8, 4 dead slots
    "lk64_d64": (1, 150, 4, 64, [(32, 40), (16, 20), (8, 10), (4, 5), (2, 3), (1, 2), (1, 1), (2, 2)], 8, "wide",
                 "softmax_lk"),                                                               # 4 x 16, 16 lanes
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("pm,ac", MODES)
def test_oracle_parity(K, oracle, name, dtype, pm, ac):
    B, Q, H, D, shapes, Kp, points, weights = CASES[name]
    img, s, pts, aw, go = make_inputs(B, Q, H, D, shapes, Kp, dtype=dtype, seed=7, points=points, weights=weights)
    test = run_cuda(K, img, s, pts, aw, go, pm, ac)
    ref = (oracle.forward(img, s, pts, aw, pm, ac),) + oracle.backward(go, img, s, pts, aw, pm, ac)
    check_against(test, ref, dtype, f"{name} {pm}/{ac}")


@pytest.mark.parametrize("pm,ac", MODES)
@pytest.mark.parametrize("Kp", [4, 8, 5], ids=["lk16", "lk32_split", "lk20_split_padded"])
def test_tuned_equals_generic(K, pm, ac, Kp):
    """The tuned (persistent, (b,h)-major) kernels and the generic kernels implement the same arithmetic per corner;
    only summation order differs."""
    img, s, pts, aw, go = make_inputs(2, 500, 8, 32, BENCH_PYRAMID, Kp, seed=3, points="wide")
    tuned = run_cuda(K, img, s, pts, aw, go, pm, ac)
    with knobs(MSDA_B200_FORCE_GENERIC="1"):
        generic = run_cuda(K, img, s, pts, aw, go, pm, ac)
    for a, b, what in zip(tuned, generic, ("out", "grad_img", "grad_points", "grad_weights")):
        b = to_np(b)
        assert_close(to_np(a), b, 1e-5, 2e-6 * max(1.0, np.abs(b).max()), what)


# ---------------------------------------------------------------------------------------------------------------------
# 16-bit storage
# ---------------------------------------------------------------------------------------------------------------------
def _bounds(dtype):
    # one rounding of the result to storage precision (eps/2 relative) + fp32 accumulation noise
    eps = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    return eps


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
@pytest.mark.parametrize("D,Kp", [(32, 4), (64, 4), (32, 3), (4, 8), (32, 2), (32, 1), (32, 8), (32, 5), (64, 8)])
@pytest.mark.parametrize("pm,ac", MODES)
def test_16bit_storage_bounds(K, oracle, dtype, D, Kp, pm, ac):
    """16-bit STORAGE, fp32 compute: against the fp64 oracle evaluated on the same (already rounded) inputs every
    output may differ by one storage rounding: |err| <= eps*|ref| + eps*1e-2*max|ref| with eps = 2^-10 (fp16) or
    2^-7 (bf16).  (The reference computes fp16 in fp16 and accepts 1e-1.)"""
    img, s, pts, aw, go = make_inputs(2, 300, 8, D, BENCH_PYRAMID, Kp, dtype=dtype, seed=11, points="wide",
                                      weights="softmax_lk")
    out, gi, gp, ga = run_cuda(K, img, s, pts, aw, go, pm, ac)
    assert out.dtype == dtype and gi.dtype == dtype and gp.dtype == dtype and ga.dtype == dtype
    rout = oracle.forward(img, s, pts, aw, pm, ac)
    rgi, rgp, rga = oracle.backward(go, img, s, pts, aw, pm, ac)
    eps = _bounds(dtype)
    for t, r, what in ((out, rout, "out"), (gi, rgi, "grad_img"), (gp, rgp, "grad_points"), (ga, rga, "grad_weights")):
        assert_close(to_np(t), r, eps, eps * 1e-2 * np.abs(r).max(), f"{dtype} {what}")


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: direct oracle comparison + size-independent properties
# ---------------------------------------------------------------------------------------------------------------------
FULL = {
    "bench_q10k": (4, 10000, 8, 32, BENCH_PYRAMID, 4),
    "detr_encoder": (2, 22223, 8, 32, DETR_PYRAMID, 4),
}


@pytest.mark.parametrize("name", list(FULL))
@pytest.mark.parametrize("pm,ac", [("border", True), ("zeros", False)])
def test_full_size_oracle_and_properties(K, oracle, name, pm, ac):
    B, Q, H, D, shapes, Kp = FULL[name]
    img, s, pts, aw, go = make_inputs(B, Q, H, D, shapes, Kp, seed=5, points="unit", weights="softmax_k")
    out, gi, gp, ga = run_cuda(K, img, s, pts, aw, go, pm, ac)
    ref = (oracle.forward(img, s, pts, aw, pm, ac),) + oracle.backward(go, img, s, pts, aw, pm, ac)
    # no floor-cell flips in grad_sampling_points: the coordinate arithmetic is bit-identical to the oracle's
    check_against((out, gi, gp, ga), ref, torch.float32, name, gpts_outliers=0)

    # property 1: linearity in img  (f(2*img + img2) = 2 f(img) + f(img2))
    img2 = torch.roll(img, 1, dims=1)
    dev = "cuda"
    f = lambda x: K.b200_multi_scale_deformable_attention_fwd(x.to(dev), s.to(dev), pts.to(dev), aw.to(dev), pm, ac)  # noqa: E731
    lhs = f(2 * img + img2)
    rhs = 2 * out + f(img2)
    assert_close(to_np(lhs), to_np(rhs), 1e-5, 1e-5, "linearity")

    # property 2 (border): bilinear weights sum to 1, so sum_pixels grad_img[b,:,h,d] = sum_q go[b,q,h,d] * sum_lk w
    if pm == "border":
        lhs = gi.double().sum(dim=1)                                                       # [B, H, D]
        rhs = (go.double().to(dev) * aw.double().to(dev).sum(dim=(3, 4))[..., None]).sum(dim=1)
        assert_close(to_np(lhs), to_np(rhs), 1e-5, 1e-4, "grad_img checksum")

    # property 3: grad_attention_weights is the directional derivative of out w.r.t. each weight:
    #   sum_{l,k} w * gaw = <go, out>  per unit
    lhs = (aw.double().to(dev) * ga.double()).sum(dim=(3, 4))
    rhs = (go.double().to(dev) * out.double()).sum(dim=-1)
    assert_close(to_np(lhs), to_np(rhs), 1e-4, 1e-4, "euler identity")


# ---------------------------------------------------------------------------------------------------------------------
# edge cases
# ---------------------------------------------------------------------------------------------------------------------
def test_empty_inputs(K):
    img, s, pts, aw, go = make_inputs(2, 0, 4, 32, BENCH_PYRAMID, 4)
    out, gi, gp, ga = run_cuda(K, img, s, pts, aw, go, "zeros", False)
    assert out.shape == (2, 0, 4, 32) and gp.shape == pts.shape and ga.shape == aw.shape
    assert gi.shape == img.shape and float(gi.abs().max()) == 0.0


def test_needs_subset(K, oracle):
    img, s, pts, aw, go = make_inputs(2, 200, 8, 32, BENCH_PYRAMID, 4, seed=2)
    ref = oracle.backward(go, img, s, pts, aw, "border", True)
    for needs in itertools.product((False, True), repeat=3):
        got = run_cuda(K, img, s, pts, aw, go, "border", True, needs=needs)[1:]
        for n, t, r, what in zip(needs, got, ref, ("grad_img", "grad_points", "grad_weights")):
            if n:
                assert_close(to_np(t), r, 1e-4, 1e-5 * np.abs(r).max(), f"needs={needs} {what}")
            else:
                assert t is None


def test_noncontiguous_and_misaligned(K, oracle):
    img, s, pts, aw, go = make_inputs(2, 100, 8, 32, BENCH_PYRAMID, 4, seed=9)
    dev = "cuda"
    img_nc = img.to(dev).permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)            # strided view
    flat = torch.empty(pts.numel() + 1, device=dev)
    flat[1:] = pts.to(dev).reshape(-1)
    pts_mis = flat[1:].view(pts.shape)                                                   # 4-byte aligned only
    assert not img_nc.is_contiguous() and pts_mis.data_ptr() % 16 != 0
    out = K.b200_multi_scale_deformable_attention_fwd(img_nc, s.to(dev), pts_mis, aw.to(dev), "zeros", False)
    gi, gp, ga = K.b200_multi_scale_deformable_attention_bwd(go.to(dev), img_nc, s.to(dev), pts_mis, aw.to(dev),
                                                             "zeros", False)
    ref = (oracle.forward(img, s, pts, aw, "zeros", False),) + oracle.backward(go, img, s, pts, aw, "zeros", False)
    check_against((out, gi, gp, ga), ref, torch.float32, "noncontiguous")
    assert gi.is_contiguous() and float(gi.abs().max()) > 0     # the reference returns zeros here (kernels.py:570-583)


def test_level_table_on_device(K, oracle):
    s = torch.tensor(DETR_PYRAMID, device="cuda")
    t = K.level_table(s, 22223).cpu().numpy()
    ref = oracle.level_table(np.array(DETR_PYRAMID))
    assert (t[:4, :3] == ref).all()
    assert t[4].tolist() == [22223, 22223, 1, 0]
    assert K.level_table(s, 22222).cpu().numpy()[4, 2] == 0


def test_optional_shape_validation(K):
    img, s, pts, aw, go = make_inputs(1, 8, 2, 32, BENCH_PYRAMID, 4)
    bad = s.clone()
    bad[0, 0] += 1
    os.environ["MSDA_B200_VALIDATE"] = "1"
    try:
        K.b200_multi_scale_deformable_attention_fwd(img.cuda(), s.cuda(), pts.cuda(), aw.cuda(), "zeros", False)
        with pytest.raises(ValueError, match="pixels"):
            K.b200_multi_scale_deformable_attention_fwd(img.cuda(), bad.cuda(), pts.cuda(), aw.cuda(), "zeros", False)
    finally:
        os.environ.pop("MSDA_B200_VALIDATE")


def test_int32_shapes_and_cpu_shapes(K, oracle):
    import msda_triton
    img, s, pts, aw, go = make_inputs(1, 64, 2, 32, BENCH_PYRAMID, 4, seed=4)
    ref = oracle.forward(img, s, pts, aw, "border", False)
    out = K.b200_multi_scale_deformable_attention_fwd(img.cuda(), s.to(torch.int32).cuda(), pts.cuda(), aw.cuda(),
                                                      "border", False)
    assert_close(to_np(out), ref, 1e-5, 1e-6, "int32 shapes")
    out = msda_triton.multiscale_deformable_attention(img.cuda(), s, pts.cuda(), aw.cuda(), "border", False)
    assert_close(to_np(out), ref, 1e-5, 1e-6, "cpu shapes through the dispatcher")


def test_bad_arguments_raise(K):
    img, s, pts, aw, go = make_inputs(1, 8, 2, 32, BENCH_PYRAMID, 4)
    with pytest.raises(ValueError):
        K.b200_multi_scale_deformable_attention_fwd(img.cuda(), s.cuda(), pts.cuda(), aw.cuda(), "reflect", False)
    with pytest.raises(ValueError):
        K.b200_multi_scale_deformable_attention_fwd(img.cuda(), s.cuda(), pts.cuda()[:, :, :1], aw.cuda(), "zeros", False)
    with pytest.raises(ValueError):
        K.b200_multi_scale_deformable_attention_fwd(img.cuda().half(), s.cuda(), pts.cuda(), aw.cuda(), "zeros", False)


# ---------------------------------------------------------------------------------------------------------------------
# deterministic (sorted-segment) grad_img
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,dtype", [("readme_c1_wide", torch.float32), ("ref_fixture_k3", torch.float32),
                                        ("odd_everything", torch.float64), ("ref_module_d4_k8", torch.float32),
                                        ("d64", torch.bfloat16)])
@pytest.mark.parametrize("pm,ac", [("zeros", False), ("border", True)])
def test_deterministic_mode(K, oracle, name, dtype, pm, ac):
    B, Q, H, D, shapes, Kp, points, weights = CASES[name]
    img, s, pts, aw, go = make_inputs(B, Q, H, D, shapes, Kp, dtype=dtype, seed=13, points=points, weights=weights)
    runs = [run_cuda(K, img, s, pts, aw, go, pm, ac, deterministic=True) for _ in range(3)]
    for r in runs[1:]:
        for a, b in zip(runs[0][1:], r[1:]):
            assert torch.equal(a, b), "deterministic mode must be bit-reproducible"
    ref = (oracle.forward(img, s, pts, aw, pm, ac),) + oracle.backward(go, img, s, pts, aw, pm, ac)
    if dtype in (torch.float32, torch.float64):
        check_against(runs[0], ref, dtype, f"deterministic {name}")
    else:
        eps = _bounds(dtype)
        assert_close(to_np(runs[0][1]), ref[1], eps, eps * 1e-2 * np.abs(ref[1]).max(), "deterministic bf16 grad_img")


def test_deterministic_full_size_bitwise(K):
    B, Q, H, D, shapes, Kp = FULL["bench_q10k"]
    img, s, pts, aw, go = make_inputs(B, Q, H, D, shapes, Kp, seed=5)
    a = run_cuda(K, img, s, pts, aw, go, "border", True, deterministic=True, needs=(True, False, False))[1]
    b = run_cuda(K, img, s, pts, aw, go, "border", True, deterministic=True, needs=(True, False, False))[1]
    c = run_cuda(K, img, s, pts, aw, go, "border", True, deterministic=False, needs=(True, False, False))[1]
    assert torch.equal(a, b)
    assert_close(to_np(a), to_np(c), 1e-4, 1e-5 * float(c.abs().max()), "deterministic vs atomic")


@pytest.mark.parametrize("pm,ac", [("zeros", False), ("border", True)])
def test_split_backward_variant(K, oracle, pm, ac):
    """The opt-in split backward (K1 without grad_img + scatter-only K2 with in-CTA binning) must match the oracle."""
    img, s, pts, aw, go = make_inputs(2, 700, 8, 32, BENCH_PYRAMID, 4, seed=17, points="wide", weights="softmax_lk")
    with knobs(MSDA_B200_BWD_SPLIT="1"):
        test = run_cuda(K, img, s, pts, aw, go, pm, ac)
        only_img = run_cuda(K, img, s, pts, aw, go, pm, ac, needs=(True, False, False))
    ref = (oracle.forward(img, s, pts, aw, pm, ac),) + oracle.backward(go, img, s, pts, aw, pm, ac)
    check_against(test, ref, torch.float32, "split backward")
    assert_close(to_np(only_img[1]), ref[1], 1e-4, 1e-5 * np.abs(ref[1]).max(), "split backward, grad_img only")


@pytest.mark.parametrize("Kp", [4, 3], ids=["tuned", "generic"])
def test_addressing_beyond_2_to_31_elements(K, Kp):
    """B=400 DETR-size images: img has 2.27e9 elements (9.1 GB fp32), so element offsets exceed int32.  The last image
    must give exactly what it gives when processed alone (same kernels, same per-unit arithmetic)."""
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~30 GB of free device memory")
    B, Q, H, D = 400, 8, 8, 32
    npix = sum(h * w for h, w in DETR_PYRAMID)
    g = torch.Generator(device="cuda").manual_seed(77)
    img = torch.randn(B, npix, H, D, device="cuda", generator=g)
    assert img.numel() > 2 ** 31
    s = torch.tensor(DETR_PYRAMID, device="cuda")
    pts = torch.rand(B, Q, H, 4, Kp, 2, device="cuda", generator=g)
    aw = torch.rand(B, Q, H, 4, Kp, device="cuda", generator=g)
    go = torch.rand(B, Q, H, D, device="cuda", generator=g)
    out = K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "zeros", False)
    gi, gp, ga = K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "zeros", False)
    for b in (0, B - 1):
        sl = slice(b, b + 1)
        out1 = K.b200_multi_scale_deformable_attention_fwd(img[sl], s, pts[sl], aw[sl], "zeros", False)
        gi1, gp1, ga1 = K.b200_multi_scale_deformable_attention_bwd(go[sl], img[sl], s, pts[sl], aw[sl], "zeros", False)
        assert torch.equal(out[sl], out1) and torch.equal(gp[sl], gp1) and torch.equal(ga[sl], ga1)
        torch.testing.assert_close(gi[sl], gi1, rtol=1e-5, atol=1e-6)     # atomics: order may differ
    del img, gi
    torch.cuda.empty_cache()


@pytest.mark.parametrize("slack", ["0", "1"])
def test_wave_pacing_many_waves(K, slack):
    """Multi-wave schedule with the wave pacing active (one (b,h) slice per wave -> 16 waves on a full persistent grid):
    same forward bits and the same gradients as the single-wave schedule."""
    img, s, pts, aw, go = make_inputs(2, 3000, 8, 32, BENCH_PYRAMID, 4, seed=13, points="wide")
    base = run_cuda(K, img, s, pts, aw, go, "zeros", False)
    # pace although these waves are small
    with knobs(MSDA_B200_SLICES_PER_WAVE="1", MSDA_B200_WAVE_PACING="2", MSDA_B200_PACE_SLACK=slack):
        paced = run_cuda(K, img, s, pts, aw, go, "zeros", False)
    assert torch.equal(paced[0], base[0])
    assert torch.equal(paced[2], base[2]) and torch.equal(paced[3], base[3])
    b = to_np(base[1])
    assert_close(to_np(paced[1]), b, 1e-5, 2e-6 * np.abs(b).max(), "grad_img")


@pytest.mark.parametrize("Kp", [4, 3], ids=["tuned", "generic"])
@pytest.mark.parametrize("pm", ["zeros", "border"])
def test_nonfinite_points_are_memory_safe(K, Kp, pm):
    """NaN / +-Inf / 1e30 sampling points (a diverged training step) must not fault or touch other units: indices are
    clamped in floating point before the integer cast (kernels.py:166-169), so only the poisoned queries change."""
    img, s, pts, aw, go = make_inputs(2, 257, 8, 32, BENCH_PYRAMID, Kp, seed=21, points="wide")
    clean = run_cuda(K, img, s, pts, aw, go, pm, False)
    bad = pts.clone()
    poison = [float("nan"), float("inf"), -float("inf"), 1e30, -1e30]
    for i, v in enumerate(poison):
        bad[0, 10 + i, :, :, :, i % 2] = v        # queries 10..14 of image 0
    out, gi, gp, ga = run_cuda(K, img, s, bad, aw, go, pm, False)
    keep = torch.ones(257, dtype=torch.bool)
    keep[10:15] = False
    assert torch.equal(out[1], clean[0][1]) and torch.equal(out[0, keep.cuda()], clean[0][0, keep.cuda()])
    assert torch.equal(gp[0, keep.cuda()], clean[2][0, keep.cuda()])
    assert torch.equal(ga[0, keep.cuda()], clean[3][0, keep.cuda()])
    b = to_np(clean[1][1])
    assert_close(to_np(gi[1]), b, 1e-5, 2e-6 * np.abs(b).max(), "grad_img of the clean image")


def test_concurrent_paced_launches_on_two_streams(K, monkeypatch):
    """Two multi-wave (paced) launches in flight on different streams: each persistent grid finds only part of the SMs
    free, so its CTAs are not co-resident -- pacing must degrade to a bounded wait (never a deadlock) and both results
    must equal the serial ones."""
    monkeypatch.setenv("MSDA_B200_SLICES_PER_WAVE", "1")
    monkeypatch.setenv("MSDA_B200_WAVE_PACING", "2")               # pace although these waves are small
    from msda_triton import _lib
    _lib.reload_tuning()
    sets = [[t.cuda() for t in make_inputs(2, 3000, 8, 32, BENCH_PYRAMID, 4, seed=30 + i, points="wide")] for i in range(2)]
    serial = []
    for img, s, pts, aw, go in sets:
        serial.append((K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "border", True),
                       K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "border", True)))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    results = [None, None]
    for rep in range(3):
        for i, (img, s, pts, aw, go) in enumerate(sets):
            with torch.cuda.stream(streams[i]):
                results[i] = (K.b200_multi_scale_deformable_attention_fwd(img, s, pts, aw, "border", True),
                              K.b200_multi_scale_deformable_attention_bwd(go, img, s, pts, aw, "border", True))
    torch.cuda.synchronize()
    for (out, g), (want_out, want_g) in zip(results, serial):
        assert torch.equal(out, want_out)
        assert torch.equal(g[1], want_g[1]) and torch.equal(g[2], want_g[2])
        b = to_np(want_g[0])
        assert_close(to_np(g[0]), b, 1e-5, 2e-6 * np.abs(b).max(), "grad_img")


@pytest.mark.parametrize("Kp", [4, 3], ids=["lk16", "lk12"])
def test_pyramid_aligned_to_16_but_not_32_bytes(K, oracle, Kp):
    """The C ABI asks for 16-byte alignment; the fp32 forward's 256-bit row loads need 32 and must step aside for a
    pyramid that only has 16 (a view 4 floats into a storage)."""
    img, s, pts, aw, go = make_inputs(2, 300, 8, 32, BENCH_PYRAMID, Kp, seed=17, points="wide")
    flat = torch.empty(img.numel() + 4, device="cuda")
    flat[4:] = img.cuda().reshape(-1)
    img16 = flat[4:].view(img.shape)
    assert img16.data_ptr() % 32 == 16 and img16.is_contiguous()
    out = K.b200_multi_scale_deformable_attention_fwd(img16, s.cuda(), pts.cuda(), aw.cuda(), "border", False)
    gi, gp, ga = K.b200_multi_scale_deformable_attention_bwd(go.cuda(), img16, s.cuda(), pts.cuda(), aw.cuda(),
                                                             "border", False)
    torch.cuda.synchronize()
    ref = (oracle.forward(img, s, pts, aw, "border", False),) + oracle.backward(go, img, s, pts, aw, "border", False)
    check_against((out, gi, gp, ga), ref, torch.float32, "16-byte aligned pyramid")
