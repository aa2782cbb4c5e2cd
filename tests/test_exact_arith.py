"""CPU brute force of the arithmetic identities kb_common.cuh relies on (IEEE fp32/fp64 are the same on the host):
see the comment above `struct Proj` there.  The device-side counterpart is tests/test_gpu_arith.py."""
import numpy as np


def _finite_bits(rng, n):
    v = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    return v[np.isfinite(v)]


def test_pixel_coordinate_single_fp32_add():
    rng = np.random.default_rng(0)
    for W in (2, 3, 48, 255, 768, 1024, 3840, 4194302):
        v = np.concatenate([_finite_bits(rng, 1 << 20), rng.uniform(-3000, 3000, 1 << 20).astype(np.float32),
                            (rng.uniform(-1, 1, 1 << 20) * 1e-6).astype(np.float32)])
        with np.errstate(over='ignore'):
            lit = ((v.astype(np.float64) + 0.5 * W) - 0.5).astype(np.float32)
            mine = v + np.float32(0.5 * W - 0.5)
        assert np.array_equal(lit.view(np.int32), mine.view(np.int32)), W


def test_z_threshold():
    v = np.nextafter(np.float32(0.001), np.float32([0, 1]))
    v = np.concatenate([v, np.float32([0.001]), np.random.default_rng(1).uniform(0.0009, 0.0011, 1 << 16).astype(np.float32)])
    assert np.array_equal(v.astype(np.float64) < 0.001, ~(v >= np.float32(0.001)))


def _twosum_sign(a, b):
    d = a - b
    bb = d - a
    e = (a - (d - bb)) + ((-b) - bb)
    return d, e


def test_exact_comparison_via_twosum():
    rng = np.random.default_rng(2)
    n = 1 << 21
    zee = (1e6 - rng.uniform(0, 2000, n)).astype(np.float32)
    err = (zee + (rng.integers(-32, 33, n) * 0.0625 + rng.integers(0, 2, n)).astype(np.float32)).astype(np.float32)
    a2 = rng.uniform(-100, 100, n).astype(np.float32)
    c2 = (a2 + 1.0 + rng.integers(-2, 3, n) * np.exp2(-rng.integers(0, 30, n).astype(np.float64))).astype(np.float32)
    for a, b in ((err, zee), (c2, a2)):
        ok = (np.abs(b) >= 1) & (np.abs(b) <= 1e15) & (np.abs(a) <= 1e15)
        a, b = a[ok], b[ok]
        d, e = _twosum_sign(a, b)
        ge = np.where(d != 1, d > 1, e >= 0)
        le = np.where(d != 1, d < 1, e <= 0)
        assert np.array_equal(ge, a.astype(np.float64) >= b.astype(np.float64) + 1.0)
        assert np.array_equal(le, a.astype(np.float64) <= b.astype(np.float64) + 1.0)


def test_magic_number_floor_and_round():
    rng = np.random.default_rng(3)
    n = 1 << 21
    v = np.concatenate([rng.uniform(-4194303, 4194303, n), rng.uniform(-2000, 2000, n),
                        rng.integers(-2000, 2001, n) + 0.5,
                        rng.integers(-2000, 2001, n) + rng.integers(-2, 3, n) * np.exp2(-rng.integers(0, 26, n).astype(np.float64))]).astype(np.float32)
    magic = np.float32(12582912.0)
    r = v + magic
    i = (r.view(np.int32) - np.int32(0x4B400000)).astype(np.int64)
    f = r - magic
    adj = f > v
    fl_f, fl_i = np.where(adj, f - np.float32(1), f), np.where(adj, i - 1, i)
    assert np.array_equal(fl_f, np.floor(v)) and np.array_equal(fl_i, np.floor(v).astype(np.int64))
    diff = v - f
    ri = i + ((diff == 0.5) & (v > 0)) - ((diff == -0.5) & (v < 0))
    ref = np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), np.ceil(v.astype(np.float64) - 0.5)).astype(np.int64)
    assert np.array_equal(ri, ref)
