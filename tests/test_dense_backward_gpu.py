"""GPU parity of the dense-level backward (DENSE instantiations of csrc/msda_bwd_tiled.cuh, MSDA_B200_BWD_DENSE=2|4): the owner
warps per CTA accumulate the coarsest pyramid level (at most 64 cells) in registers and add it to grad_img once per (b,h)
slice; the worker warps skip that level's row adds.  Covered: all four modes (clamped and masked corners, far-out-of-range
points), Q not a multiple of 4 (padding queries), slice crossings inside a CTA's range (many small slices), multi-wave
launches, pyramids whose dense level is not the last / not 8x8 / absent, every owner count, grad_img-only and
grad_img-less calls, non-finite grad_out and weights.  Reference semantics: /root/reference/src/msda_triton/kernels.py:542-553.
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
@pytest.mark.parametrize("nown,pf", [(2, 2), (2, 3), (4, 2), (4, 3)])
def test_dense_backward_matches_oracle(K, oracle, pm, ac, nown, pf):
    B, Q, H, D = 2, 1203, 8, 32            # Q not a multiple of 4: padding queries in the last tile of every slice
    img, s, pts, aw, go = make_inputs(B, Q, H, D, BENCH_PYRAMID, 4, seed=91, points="wide", weights="softmax_lk")
    with knobs(MSDA_B200_BWD_DENSE=nown, MSDA_B200_DENSE_PF=pf):
        test = bwd(K, img, s, pts, aw, go, pm, ac)
    with knobs(MSDA_B200_BWD_DENSE=0):
        plain = bwd(K, img, s, pts, aw, go, pm, ac)
    ref = oracle.backward(go, img, s, pts, aw, pm, ac)
    check(test, ref, f"dense backward {pm}/{ac} owners={nown}")
    assert torch.equal(test[1], plain[1]) and torch.equal(test[2], plain[2])     # no atomics there: same bits


@pytest.mark.parametrize("pyramid", [
    [(8, 8), (32, 32), (16, 16), (64, 64)],      # the dense level comes first
    [(20, 30), (10, 15), (5, 8), (3, 4)],        # 12 cells, odd widths
    [(9, 7), (7, 9), (6, 10), (8, 8)],           # several candidates: the largest (8x8 = 64) wins over 63 and 60
    [(30, 30), (4, 16), (16, 4), (2, 2)],        # 64 cells as 4x16
    DETR_PYRAMID,                                # no level of <= 64 cells: the owners only keep the barriers company
], ids=["first", "tiny", "ties", "4x16", "none"])
@pytest.mark.parametrize("pm,ac", [("zeros", False), ("border", True)])
def test_dense_level_selection(K, oracle, pyramid, pm, ac):
    img, s, pts, aw, go = make_inputs(1, 517, 8, 32, pyramid, 4, seed=92, points="far")
    with knobs(MSDA_B200_BWD_DENSE=4):
        test = bwd(K, img, s, pts, aw, go, pm, ac)
    check(test, oracle.backward(go, img, s, pts, aw, pm, ac), f"pyramid {pyramid} {pm}/{ac}")


def test_many_small_slices_and_waves(K, oracle):
    """B*H = 96 slices of 37 queries: every CTA range crosses slices (flushes in the middle of a range); then the same
    problem cut into one-slice waves with pacing forced."""
    img, s, pts, aw, go = make_inputs(12, 37, 8, 32, BENCH_PYRAMID, 4, seed=93, points="wide")
    ref = oracle.backward(go, img, s, pts, aw, "zeros", True)
    with knobs(MSDA_B200_BWD_DENSE=2):
        check(bwd(K, img, s, pts, aw, go, "zeros", True), ref, "many slices")
    with knobs(MSDA_B200_BWD_DENSE=4, MSDA_B200_SLICES_PER_WAVE=8, MSDA_B200_WAVE_PACING=2):
        check(bwd(K, img, s, pts, aw, go, "zeros", True), ref, "many slices, 12 waves")
    with knobs(MSDA_B200_BWD_DENSE=4, MSDA_B200_SLICES_PER_WAVE=1, MSDA_B200_WAVE_PACING=2):
        check(bwd(K, img, s, pts, aw, go, "zeros", True), ref, "many slices, 96 waves")


def test_tiny_query_counts(K, oracle):
    """Fewer queries per slice than owner warps; a single query."""
    for Q in (1, 3, 5):
        img, s, pts, aw, go = make_inputs(2, Q, 8, 32, BENCH_PYRAMID, 4, seed=94 + Q, points="wide")
        with knobs(MSDA_B200_BWD_DENSE=4):
            test = bwd(K, img, s, pts, aw, go, "border", False)
        check(test, oracle.backward(go, img, s, pts, aw, "border", False), f"Q={Q}")


def test_needs_subsets_and_nonfinite(K):
    B, Q, H, D = 1, 801, 8, 32
    img, s, pts, aw, go = make_inputs(B, Q, H, D, BENCH_PYRAMID, 4, seed=95, points="wide")
    go[0, 10, 3, 7] = float("inf")
    go[0, 500, 1, 0] = float("nan")
    aw[0, 77, 2, 3, 1] = float("nan")         # a weight of the dense level
    with knobs(MSDA_B200_BWD_DENSE=0):
        base = bwd(K, img, s, pts, aw, go, "zeros", False)
    with knobs(MSDA_B200_BWD_DENSE=4):
        test = bwd(K, img, s, pts, aw, go, "zeros", False)
        only_img = bwd(K, img, s, pts, aw, go, "zeros", False, needs=(True, False, False))
        no_img = bwd(K, img, s, pts, aw, go, "zeros", False, needs=(False, True, True))
    assert torch.equal(torch.isfinite(test[0]), torch.isfinite(base[0]))       # the same rows are poisoned, no others
    assert only_img[1] is None and no_img[0] is None
    for x, y in ((no_img[1], base[1]), (no_img[2], base[2])):
        assert torch.equal(torch.isnan(x), torch.isnan(y)) and torch.equal(torch.nan_to_num(x), torch.nan_to_num(y))
    fin = torch.isfinite(base[0])
    assert torch.allclose(test[0][fin], base[0][fin], rtol=1e-4, atol=1e-5 * float(base[0][fin].abs().max()))
    assert torch.allclose(only_img[0][fin], base[0][fin], rtol=1e-4, atol=1e-5 * float(base[0][fin].abs().max()))


def test_full_size_bench_shape(K):
    """B=4, Q=10 000: against the plain kernel (itself checked against the oracle in test_cuda_parity.py)."""
    img, s, pts, aw, go = make_inputs(4, 10000, 8, 32, BENCH_PYRAMID, 4, seed=96)
    with knobs(MSDA_B200_BWD_DENSE=0):
        base = bwd(K, img, s, pts, aw, go, "border", True)
    for nown in (2, 4):
        with knobs(MSDA_B200_BWD_DENSE=nown):
            test = bwd(K, img, s, pts, aw, go, "border", True)
        b = to_np(base[0])
        assert_close(to_np(test[0]), b, 1e-5, 2e-6 * np.abs(b).max(), f"grad_img, owners={nown}")
        assert torch.equal(test[1], base[1]) and torch.equal(test[2], base[2])
