"""GPU parity of the tensor-memory backward (csrc/msda_bwd_tmem.cu: the coarsest pyramid levels are accumulated in TMEM,
turns handed round a named-barrier ring, one row add per row at the flush) against the CPU oracle and the plain tuned
backward.

Places where a bug loses or doubles a contribution, each exercised here:
  * folded twin corners (border clamping, zeros-mode corners outside the level)  -> 'wide' / 'far' points, all modes;
  * records of one unit hitting the same rows (sequential read-modify-write)      -> tiny coarse levels, clustered points;
  * flushes when the ring crosses (b,h) slices / L2 waves, ragged last rounds     -> small Q with many slices, 1 slice per wave;
  * one level / two levels / none in tensor memory                                -> MSDA_B200_TMEM_LEVELS, odd pyramids.
Reference semantics: /root/reference/src/msda_triton/kernels.py:542-553 (corner grads + atomic adds).
"""
import itertools

import numpy as np
import pytest
import torch

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


def bwd(K, img, s, pts, aw, go, pm, ac, **kw):
    a, sh, p, w, g = (t.cuda() for t in (img, s, pts, aw, go))
    out = K.b200_multi_scale_deformable_attention_bwd(g, a, sh, p, w, pm, ac, deterministic=False, **kw)
    torch.cuda.synchronize()
    return out


def check(test, ref, what):
    for t, r, name in zip(test, ref, ("grad_img", "grad_points", "grad_weights")):
        r = np.asarray(r)
        assert_close(to_np(t), r, 1e-4, 1e-5 * max(np.abs(r).max(), 1e-30), f"{what}: {name}")


@pytest.mark.parametrize("pm,ac", MODES)
@pytest.mark.parametrize("points", ["unit", "wide", "far"])
def test_tmem_backward_matches_oracle(K, oracle, pm, ac, points):
    """Bench pyramid (two levels in tensor memory), Q not a multiple of 4 (padding queries), all four modes."""
    img, s, pts, aw, go = make_inputs(2, 1203, 8, 32, BENCH_PYRAMID, 4, seed=41, points=points, weights="softmax_lk")
    with knobs(MSDA_B200_BWD_TMEM="1"):
        test = bwd(K, img, s, pts, aw, go, pm, ac)
    ref = oracle.backward(go, img, s, pts, aw, pm, ac)
    check(test, ref, f"tmem backward {pm}/{ac}/{points}")


@pytest.mark.parametrize("levels", ["0", "1", "2"])
def test_tmem_backward_level_count(K, oracle, levels):
    """None / one / two of the coarsest levels in tensor memory."""
    img, s, pts, aw, go = make_inputs(2, 900, 8, 32, BENCH_PYRAMID, 4, seed=42, points="wide")
    with knobs(MSDA_B200_BWD_TMEM="1", MSDA_B200_TMEM_LEVELS=levels):
        test = bwd(K, img, s, pts, aw, go, "border", True)
    ref = oracle.backward(go, img, s, pts, aw, "border", True)
    check(test, ref, f"tmem backward levels={levels}")


@pytest.mark.parametrize("pyr", [[(8, 8), (4, 4), (2, 2), (1, 1)], [(40, 40), (20, 20), (3, 5), (2, 2)],
                                 [(30, 30), (15, 15), (18, 17), (1, 7)], [(9, 9), (8, 8), (4, 4), (2, 2)],
                                 [(64, 64), (32, 32), (22, 22), (4, 5)], [(20, 20), (10, 10), (21, 21), (8, 7)]],
                         ids=["width1_tail", "two_small_tails", "ragged", "rows_lt_16_first", "region_limit",
                              "second_does_not_fit"])
@pytest.mark.parametrize("pm,ac", [("zeros", False), ("border", True), ("border", False)])
def test_tmem_backward_odd_pyramids(K, oracle, pyr, pm, ac):
    """Tiny coarse levels (the points of a unit share rows all the time), levels that do not fit, width-1 levels."""
    img, s, pts, aw, go = make_inputs(3, 333, 8, 32, pyr, 4, seed=43, points="wide")
    with knobs(MSDA_B200_BWD_TMEM="1"):
        test = bwd(K, img, s, pts, aw, go, pm, ac)
    ref = oracle.backward(go, img, s, pts, aw, pm, ac)
    check(test, ref, f"tmem backward {pyr} {pm}/{ac}")


def test_tmem_backward_clustered_points(K, oracle):
    """All points of a unit within a pixel or two of each other: consecutive records update the same rows."""
    B, Q, H, D = 2, 600, 8, 32
    img, s, pts, aw, go = make_inputs(B, Q, H, D, BENCH_PYRAMID, 4, seed=44)
    g = torch.Generator().manual_seed(7)
    centre = torch.rand(B, Q, H, 1, 1, 2, generator=g)
    pts = centre + (torch.rand(B, Q, H, 4, 4, 2, generator=g) - 0.5) * 0.08
    pts[:, ::5] = centre[:, ::5]                       # every fifth query: all 16 points identical
    for pm, ac in (("zeros", False), ("border", True)):
        with knobs(MSDA_B200_BWD_TMEM="1"):
            test = bwd(K, img, s, pts, aw, go, pm, ac)
        ref = oracle.backward(go, img, s, pts, aw, pm, ac)
        check(test, ref, f"tmem backward clustered {pm}/{ac}")


@pytest.mark.parametrize("B,Q", [(24, 37), (40, 3), (5, 61), (1, 59)])
def test_tmem_backward_many_slices_per_cta(K, oracle, B, Q):
    """Few queries, many (b,h) slices: the ring crosses slices all the time (flush inside a turn), ragged last rounds."""
    img, s, pts, aw, go = make_inputs(B, Q, 8, 32, BENCH_PYRAMID, 4, seed=45, points="wide")
    with knobs(MSDA_B200_BWD_TMEM="1"):
        test = bwd(K, img, s, pts, aw, go, "zeros", False)
    ref = oracle.backward(go, img, s, pts, aw, "zeros", False)
    check(test, ref, "tmem backward, many slices per CTA")


@pytest.mark.parametrize("slack", ["0", "1"])
def test_tmem_backward_multi_wave_paced(K, slack):
    """One (b,h) slice per wave with the wave pacing forced: same gradients as the single-wave plain backward."""
    img, s, pts, aw, go = make_inputs(2, 3000, 8, 32, BENCH_PYRAMID, 4, seed=46, points="wide")
    with knobs(MSDA_B200_BWD_TMEM="0"):
        base = bwd(K, img, s, pts, aw, go, "border", True)
    with knobs(MSDA_B200_BWD_TMEM="1", MSDA_B200_SLICES_PER_WAVE="1", MSDA_B200_WAVE_PACING="2",
               MSDA_B200_PACE_SLACK=slack):
        test = bwd(K, img, s, pts, aw, go, "border", True)
    assert torch.equal(test[1], base[1]) and torch.equal(test[2], base[2])     # no atomics there: same bits
    b = to_np(base[0])
    assert_close(to_np(test[0]), b, 1e-5, 2e-6 * np.abs(b).max(), "grad_img")


def test_tmem_backward_detr_pyramid_and_needs(K, oracle):
    """DETR pyramid (only the 13x21 level fits a quarter) and the needs_input_grad subsets that keep grad_img."""
    img, s, pts, aw, go = make_inputs(1, 2001, 8, 32, DETR_PYRAMID, 4, seed=47, points="wide", weights="softmax_lk")
    ref = oracle.backward(go, img, s, pts, aw, "zeros", False)
    with knobs(MSDA_B200_BWD_TMEM="1"):
        test = bwd(K, img, s, pts, aw, go, "zeros", False)
        check(test, ref, "tmem backward DETR pyramid")
        gi, gp, ga = bwd(K, img, s, pts, aw, go, "zeros", False, needs=(True, True, False))
        assert ga is None
        check((gi, gp), ref[:2], "tmem backward needs=(img, points)")


def test_tmem_backward_nonfinite_grad_out_stays_in_its_rows(K):
    """An inf in grad_out reaches exactly the rows the plain kernel puts it in (masked corners add nothing, not 0 * inf)."""
    img, s, pts, aw, go = make_inputs(1, 400, 8, 32, BENCH_PYRAMID, 4, seed=49, points="wide")
    go[0, 17, 3, 5] = float("inf")
    with knobs(MSDA_B200_BWD_TMEM="0"):
        base = bwd(K, img, s, pts, aw, go, "zeros", False)
    with knobs(MSDA_B200_BWD_TMEM="1"):
        test = bwd(K, img, s, pts, aw, go, "zeros", False)
    assert torch.equal(torch.isfinite(test[0]), torch.isfinite(base[0]))


def test_tmem_backward_full_size_bench_shape(K, oracle):
    """BASELINE C2 at full size (B=4, Q=10 000): tensor-memory backward vs the oracle, plus the grad_img checksum identity
    sum(grad_img) == sum_units (sum_c grad_out) * (sum of the unit's corner weights)."""
    img, s, pts, aw, go = make_inputs(4, 10000, 8, 32, BENCH_PYRAMID, 4, seed=48)
    with knobs(MSDA_B200_BWD_TMEM="1"):
        test = bwd(K, img, s, pts, aw, go, "border", True)
    with knobs(MSDA_B200_BWD_TMEM="0"):
        plain = bwd(K, img, s, pts, aw, go, "border", True)
    ref = oracle.backward(go, img, s, pts, aw, "border", True)
    check(test, ref, "tmem backward, bench shape")
    # border mode: the corner weights of a point sum to 1, so sum(grad_img) = sum_u sum_p aw[u,p] * sum_c go[u,c]
    want = float((aw.double().sum(dim=(-1, -2)) * go.double().sum(-1)).sum())
    got = float(test[0].double().sum())
    assert abs(got - want) <= 1e-6 * abs(want), (got, want)
    assert torch.equal(test[1], plain[1]) and torch.equal(test[2], plain[2])
