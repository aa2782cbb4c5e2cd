"""The fp32 forms the kernels use in place of the reference's fp64 sub-expressions / IEEE divisions are the SAME
functions: evaluated side by side on the device (kb_selftest_arith, include/kb200.h) on random and adversarial
operands.  Reference expressions: utils/common.py:453, :467-470, :556-561, :639, :686, :255."""
import ctypes

import numpy as np
import pytest
import torch

from ken_burns_effect_b200 import _native as nat

pytestmark = pytest.mark.gpu


def _run(which, a, b, W=1024):
    ta = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    tb = torch.from_numpy(np.ascontiguousarray(b, dtype=np.float32)).cuda()
    bad = torch.zeros(1, dtype=torch.int64, device="cuda")
    nat.check(nat.lib().kb_selftest_arith(which, ctypes.c_void_p(ta.data_ptr()), ctypes.c_void_p(tb.data_ptr()), ta.numel(), W,
                                          ctypes.c_void_p(bad.data_ptr()),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "kb_selftest_arith")
    torch.cuda.synchronize()
    return int(bad.item())


def _finite_bits(rng, n):
    v = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    return v[np.isfinite(v)]


def test_shared_reciprocal_quantisation_equals_ieee_division():
    rng = np.random.default_rng(0)
    n = 1 << 22
    w = np.concatenate([rng.uniform(0, 8, n), rng.uniform(0, 1e-3, n), np.exp(rng.uniform(-40, 40, n))]).astype(np.float32)
    # numerators: colour * weight sums, exact multiples of k/255 that land on quantisation boundaries, tiny and huge values
    k = rng.integers(0, 256, 3 * n).astype(np.float32)
    a = np.concatenate([rng.uniform(0, 1, n) * w[:n], (k[:n] / np.float32(255.0)) * (w[n:2 * n] + np.float32(1e-7)),
                        np.exp(rng.uniform(-60, 60, n))]).astype(np.float32)
    assert _run(0, a, w) == 0
    fb = _finite_bits(rng, n)
    assert _run(0, fb, np.abs(rng.permutation(fb))) == 0


def test_exact_comparisons_equal_fp64_forms():
    rng = np.random.default_rng(1)
    n = 1 << 22
    zee = (1e6 - rng.uniform(0, 2000, n)).astype(np.float32)
    err = (zee + (rng.integers(-32, 33, n) * 0.0625 + rng.integers(0, 2, n)).astype(np.float32)).astype(np.float32)
    assert _run(1, err, zee) == 0
    a = rng.uniform(-100, 100, n).astype(np.float32)
    c = (a + 1.0 + rng.integers(-2, 3, n) * np.exp2(-rng.integers(0, 30, n).astype(np.float64))).astype(np.float32)
    assert _run(1, c, a) == 0
    fb = _finite_bits(rng, n)
    assert _run(1, fb, rng.permutation(fb)) == 0          # includes the magnitudes that take the literal fp64 path


def test_floor_and_round_without_conversions():
    rng = np.random.default_rng(2)
    n = 1 << 22
    v = np.concatenate([rng.uniform(-4194303, 4194303, n), rng.uniform(-2000, 2000, n),
                        rng.integers(-2000, 2001, n) + rng.integers(-2, 3, n) * np.exp2(-rng.integers(0, 26, n).astype(np.float64)),
                        rng.integers(-2000, 2001, n) + 0.5]).astype(np.float32)
    assert _run(2, v, v) == 0


@pytest.mark.parametrize("W", [2, 3, 48, 64, 255, 768, 1024, 2160, 3840, 4194302])
def test_pixel_coordinate_sum_equals_two_fp64_additions(W):
    rng = np.random.default_rng(3)
    n = 1 << 21
    v = np.concatenate([_finite_bits(rng, n), rng.uniform(-3000, 3000, n), rng.uniform(-1, 1, n) * 1e-6,
                        rng.uniform(0.0009, 0.0011, n)]).astype(np.float32)
    assert _run(3, v, v, W) == 0
