"""GPU parity of the pair-aggregated backward (opt-in AGG instantiation of csrc/msda_bwd_tiled.cu, MSDA_B200_BWD_AGG=1;
measured slower than the plain kernel and therefore off by default): when two neighbouring queries
of a warp tile put a sampling point into the same cell, the even lane group adds both contributions with one row add and
the odd one adds nothing.  Covered: inputs where every / some / no pairs form, all four modes (clamped and masked
corners must match too), padding queries in the last tile, a pair whose partner is a padding query, non-finite grad_out,
and the self-attention shape (Q == Npix).  Reference semantics: /root/reference/src/msda_triton/kernels.py:542-553.
"""
import itertools

import numpy as np
import pytest
import torch

from util import BENCH_PYRAMID, assert_close, knobs, make_inputs, to_np

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


def bwd(K, img, s, pts, aw, go, pm, ac, **kw):
    a, sh, p, w, g = (t.cuda() for t in (img, s, pts, aw, go))
    out = K.b200_multi_scale_deformable_attention_bwd(g, a, sh, p, w, pm, ac, deterministic=False, **kw)
    torch.cuda.synchronize()
    return out


def check(test, ref, what):
    for t, r, name in zip(test, ref, ("grad_img", "grad_points", "grad_weights")):
        r = np.asarray(r)
        assert_close(to_np(t), r, 1e-4, 1e-5 * max(np.abs(r).max(), 1e-30), f"{what}: {name}")


def coherent_points(B, Q, H, L, Kp, spread, seed, wide=False):
    """Queries in groups of `spread` consecutive queries share their sampling points up to a tiny jitter (same cell almost
    always); spread = 1: independent points."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, (Q + spread - 1) // spread, H, L, Kp, 2, generator=g)
    if wide:
        base = base * 1.5 - 0.25
    pts = base.repeat_interleave(spread, dim=1)[:, :Q].clone()
    pts += (torch.rand(pts.shape, generator=g) - 0.5) * 1e-4
    return pts


@pytest.mark.parametrize("pm,ac", MODES)
@pytest.mark.parametrize("spread", [1, 2, 4, 3])
def test_pair_aggregation_matches_oracle(K, oracle, pm, ac, spread):
    B, Q, H, D = 2, 1203, 8, 32            # Q not a multiple of 4: padding queries, one of them a would-be partner
    img, s, _, aw, go = make_inputs(B, Q, H, D, BENCH_PYRAMID, 4, seed=81, weights="softmax_lk")
    pts = coherent_points(B, Q, H, 4, 4, spread, seed=82, wide=True)
    with knobs(MSDA_B200_BWD_AGG="1"):
        test = bwd(K, img, s, pts, aw, go, pm, ac)
    with knobs(MSDA_B200_BWD_AGG="0"):
        plain = bwd(K, img, s, pts, aw, go, pm, ac)
    ref = oracle.backward(go, img, s, pts, aw, pm, ac)
    check(test, ref, f"pair aggregation {pm}/{ac} spread={spread}")
    assert torch.equal(test[1], plain[1]) and torch.equal(test[2], plain[2])     # no atomics there: same bits


def test_mixed_levels_pair_only_on_the_coarse_ones(K, oracle):
    """Neighbouring queries one level-0 pixel apart with identical offsets: same cell on the coarse levels, not on level 0."""
    B, Q, H, D = 1, 2000, 8, 32
    img, s, _, aw, go = make_inputs(B, Q, H, D, BENCH_PYRAMID, 4, seed=83)
    g = torch.Generator().manual_seed(84)
    q = torch.arange(Q, dtype=torch.float32)
    ref_xy = torch.stack((((q % 64) + 0.5) / 64, ((q // 64) % 64 + 0.5) / 64), -1)              # raster over level 0
    offs = (torch.rand(1, 1, H, 4, 4, 2, generator=g) - 0.5) * 0.1
    pts = (ref_xy[None, :, None, None, None, :] + offs).expand(B, Q, H, 4, 4, 2).contiguous()
    for pm, ac in (("zeros", False), ("border", True)):
        with knobs(MSDA_B200_BWD_AGG="1"):
            test = bwd(K, img, s, pts, aw, go, pm, ac)
        check(test, oracle.backward(go, img, s, pts, aw, pm, ac), f"raster queries {pm}/{ac}")


def test_needs_subsets_and_nonfinite(K):
    B, Q, H, D = 1, 801, 8, 32
    img, s, _, aw, go = make_inputs(B, Q, H, D, BENCH_PYRAMID, 4, seed=85)
    pts = coherent_points(B, Q, H, 4, 4, 2, seed=86)
    go[0, 10, 3, 7] = float("inf")           # query 10 pairs with query 11
    with knobs(MSDA_B200_BWD_AGG="0"):
        base = bwd(K, img, s, pts, aw, go, "zeros", False)
    with knobs(MSDA_B200_BWD_AGG="1"):
        test = bwd(K, img, s, pts, aw, go, "zeros", False)
        only_img = bwd(K, img, s, pts, aw, go, "zeros", False, needs=(True, False, False))
        no_img = bwd(K, img, s, pts, aw, go, "zeros", False, needs=(False, True, True))
    assert torch.equal(torch.isfinite(test[0]), torch.isfinite(base[0]))
    assert only_img[1] is None and no_img[0] is None
    for x, y in ((no_img[1], base[1]), (no_img[2], base[2])):          # same bits, NaNs (from the inf) in the same places
        assert torch.equal(torch.isnan(x), torch.isnan(y)) and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))
    fin = torch.isfinite(base[0])
    assert torch.allclose(only_img[0][fin], base[0][fin], rtol=1e-4, atol=1e-5 * float(base[0][fin].abs().max()))


def test_self_attention_shape(K, oracle):
    """Q == Npix (one query per pixel), with and without the knob: the oracle's results either way."""
    npix = sum(h * w for h, w in BENCH_PYRAMID)
    img, s, _, aw, go = make_inputs(1, npix, 8, 32, BENCH_PYRAMID, 4, seed=87)
    pts = coherent_points(1, npix, 8, 4, 4, 2, seed=88)
    ref = oracle.backward(go, img, s, pts, aw, "border", True)
    check(bwd(K, img, s, pts, aw, go, "border", True), ref, "default kernel")
    with knobs(MSDA_B200_BWD_AGG="1"):
        check(bwd(K, img, s, pts, aw, go, "border", True), ref, "pair aggregation")
