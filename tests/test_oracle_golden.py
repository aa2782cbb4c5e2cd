"""CPU: the oracle (oracle/kb_oracle.c) against outputs of the REFERENCE's own CUDA kernels.

tests/golden/ref_render_*.npz were produced on a B200 by tools/diag_degrid_race.py, which launches the
cubins that oracle/build_ref.py compiles from /root/reference/utils/common.py (the reference has no tests
or golden vectors of its own).  This pins the oracle without a GPU.
"""
import glob
import os

import numpy as np
import pytest

import oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_render_*.npz")))


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_kernels(path):
    assert GOLDEN, "golden fixtures missing"
    oracle.set_threads(1)
    g = np.load(path)
    pts, data = g["points"][None], g["data"][None]
    focal, baseline = float(g["focal"]), float(g["baseline"])
    H, W = g["zee_raw"].shape[-2:]
    # pass 1: bit-exact z-buffer
    zraw = oracle.splat_min(pts, H, W, focal, baseline)
    assert np.array_equal(zraw.view(np.int32), g["zee_raw"].view(np.int32))
    # pass 2: race-free degrid vs the reference's racy in-place pass: equal except a few race pixels, and
    # far closer than the sequential in-place order
    zj, zs = oracle.degrid(zraw, 0), oracle.degrid(zraw, 1)
    nj, ns = int((zj != g["zee_degrid"]).sum()), int((zs != g["zee_degrid"]).sum())
    assert nj <= max(16, int(0.004 * H * W)), nj
    assert nj < ns
    # pass 3 on the reference's own z-buffer: accumulators equal up to fp32 summation order
    out = oracle.splat_accum(pts, data, g["zee_degrid"], focal, baseline)
    assert np.array_equal(out == 0, g["out"] == 0)
    assert rel_l2(out, g["out"]) < 2e-6
    # epilogue (utils/common.py:686)
    render, existing = oracle.normalize(g["out"])
    assert rel_l2(render, g["render"]) < 1e-6
    assert np.array_equal(existing, g["existing"])


def test_oracle_thread_count_invariance():
    """More OpenMP threads change nothing but fp32 summation order."""
    g = np.load(GOLDEN[0])
    pts, data = g["points"][None], g["data"][None]
    focal, baseline = float(g["focal"]), float(g["baseline"])
    H, W = g["zee_raw"].shape[-2:]
    oracle.set_threads(1)
    a = oracle.render_pointcloud(pts, data, W, H, focal, baseline, want_zee=True)
    oracle.set_threads(4)
    b = oracle.render_pointcloud(pts, data, W, H, focal, baseline, want_zee=True)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    assert rel_l2(b[0], a[0]) < 1e-6
    oracle.set_threads(1)


def test_oracle_fill_properties():
    """fill_disocclusion restatement: valid pixels untouched, holes copy an existing valid pixel, idempotent
    on a hole-free image, and a single hole between two valid pixels takes the FARTHER one (:904-907)."""
    H, W = 24, 32
    rng = np.random.default_rng(0)
    img = rng.random((1, 4, H, W)).astype(np.float32)
    depth = rng.uniform(1, 10, (1, 1, H, W)).astype(np.float32)
    out = oracle.fill_disocclusion(img, depth)
    assert np.array_equal(out, img)
    depth[0, 0, 10, 10:14] = 0.0
    depth[0, 0, 10, 9], depth[0, 0, 10, 14] = 2.0, 50.0
    out, xy = oracle.fill_disocclusion(img, depth, want_xy=True)
    mask = depth[0, 0] <= 0
    assert np.array_equal(out[0][:, ~mask], img[0][:, ~mask])
    for x in range(10, 14):
        fx, fy = xy[0, 10, x]
        assert depth[0, 0, fy, fx] > 0
        assert np.array_equal(out[0, :, 10, x], img[0, :, fy, fx])
    # all-hole image: nothing to copy from, output equals input
    z = np.zeros_like(depth)
    assert np.array_equal(oracle.fill_disocclusion(img, z), img)


def test_oracle_empty_and_culled():
    oracle.set_threads(1)
    pts = np.zeros((1, 3, 10), np.float32)
    pts[0, 2] = 0.0005
    zee, idx = oracle.splat_min(pts, 8, 8, 4.0, 120, want_idx=True)
    assert (zee == 1e6).all() and (idx == -1).all()
    r, e = oracle.render_pointcloud(pts, np.ones((1, 4, 10), np.float32), 8, 8, 4.0, 120)
    assert not r.any() and not e.any()


def test_oracle_shift_points():
    """process_shift tensor half (utils/common.py:104-109): z/(z+1e-7) is exactly 1 for z >= 2, 0 for z = 0."""
    pts = np.array([[1.5, -2.0, 0.0], [0.25, 3.0, 0.0], [512.0, 3.0, 0.0]], np.float32)
    sh = np.array([0.5, -0.25, 2.0], np.float32)
    out = oracle.shift_points(pts, sh)
    assert np.array_equal(out[:, 0], np.array([2.0, 0.0, 514.0], np.float32))
    assert np.array_equal(out[:, 2], sh)          # z = 0: x*0 + shift
    assert out[0, 1] == np.float32(-2.0) * (np.float32(3.0) / (np.float32(3.0) + np.float32(1e-7))) + np.float32(0.5)


def test_mask_zee_index_order_semantics():
    """generate_mask's kernel in index order (oracle.mask_zee): strict '>' keeps the first of equal points, a nearer later
    point displaces an earlier one, point 0 is never cleared (utils/common.py:755-765)."""
    import oracle
    W, H, focal = 8, 8, 4.0
    pts = np.zeros((1, 3, W * H), np.float32)
    pts[0, 2] = 0.0005
    for n, z in ((0, 50.0), (3, 30.0), (5, 20.0), (9, 20.0), (12, 25.0)):
        pts[0, :, n] = (0.3 * z / focal, 0.3 * z / focal, z)
    raw = oracle.mask_zee(pts, H, W, focal, 120)
    assert raw[0, 0] == 1.0          # displaced, but pid > 0 protects point 0
    assert raw[0, 3] == 0.0          # displaced by 5
    assert raw[0, 5] == 1.0          # nearest, first of the two equal points
    assert raw[0, 9] == 0.0 and raw[0, 12] == 0.0
    assert raw.sum() == 2.0
