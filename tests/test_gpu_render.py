"""GPU parity tests of the render path (run with -m gpu on the B200 box).

Ground truth, strongest first:
  1. the REFERENCE's own kernels (oracle/_ref cubins built from /root/reference/utils/common.py),
  2. the CPU oracle (oracle/kb_oracle.c), itself pinned against 1. here and against tests/golden/ on CPU.
Bars: z-buffer after updateZee bit-exact; degrid differs from the reference only where its in-place race
can act (bounded count); rendered RGB/depth within 1e-3 relative L2 (north_star), measured ~1e-6.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import refgpu
from ken_burns_effect_b200.utils import common as kb
import kb_helpers as helpers

pytestmark = pytest.mark.gpu

# (W, H, focal, extra points, C) -- must exist in oracle/build_ref.py:SHAPES
CASES = [
    (64, 48, 32.0, 0, 4),
    (64, 48, 32.0, 517, 4),
    (256, 192, 128.0, 4099, 4),
    (256, 192, 101.37, 0, 4),
    (1024, 768, 512.0, 0, 4),
    (1024, 768, 512.0, 70001, 4),
    (3840, 2160, 1920.0, 0, 4),           # configs[3]
]


def _inputs(W, H, focal, extra, C, step=1.0, seed=1234):
    pts, rgb, dep, common = helpers.scene(W, H, focal, extra, seed)
    shifted, sh, f = helpers.shifted_cloud(pts, common, W, H, step)
    if C == 4:
        data = np.concatenate([rgb, dep], 0)
    else:
        rng = np.random.default_rng(7)
        data = np.concatenate([rgb, dep / dep.max(), rng.standard_normal((C - 4, rgb.shape[1])).astype(np.float32)], 0)
    return shifted[None], data[None].astype(np.float32), common


@pytest.mark.parametrize("W,H,focal,extra,C", CASES)
def test_render_vs_reference_kernels_and_oracle(W, H, focal, extra, C):
    if not refgpu.available():
        pytest.skip("oracle/_ref not built")
    oracle.set_threads(1 if W <= 256 else 0)
    pts, data, common = _inputs(W, H, focal, extra, C)
    tp = torch.from_numpy(pts).cuda()
    td = torch.from_numpy(data).cuda()
    # reference kernels on the GPU
    r_render, r_exist, r_zraw, r_zdeg, r_out = refgpu.render_pointcloud(tp, td, W, H, focal, 120, stages=True)
    # product
    render, exist, zraw, zdeg = kb.render_pointcloud(tp, td, W, H, focal, 120, return_zee=True)
    torch.cuda.synchronize()
    # CPU oracle
    o_render, o_exist, o_zraw, o_zdeg = oracle.render_pointcloud(pts, data, W, H, focal, 120, want_zee=True)

    # 1. z-buffer after the min pass: bit-exact, three ways
    assert np.array_equal(zraw.cpu().numpy().view(np.int32), r_zraw.cpu().numpy().view(np.int32))
    assert np.array_equal(o_zraw.view(np.int32), r_zraw.cpu().numpy().view(np.int32))
    # 2. degrid: product == oracle exactly (both race-free).  The reference updates in place while
    #    neighbours read (utils/common.py:556-567); on B200 it resolves like the race-free version except at
    #    a handful of pixels, about as many as differ between two runs of the reference itself
    #    (tools/diag_degrid_race.py, profiles/degrid_race_r01.md).
    assert np.array_equal(zdeg.cpu().numpy().view(np.int32), o_zdeg.view(np.int32))
    differs = (zdeg != r_zdeg)
    n_race = int(differs.sum().item())
    assert n_race <= max(16, int(0.004 * W * H)), f"degrid differs from the reference at {n_race} pixels"
    # 3. accumulate + normalise against the REFERENCE's own z-buffer (isolates the race): tight bound,
    #    identical hole set
    a_render, a_exist = kb.accumulate_with_zee(tp, td, r_zdeg, focal, 120)
    assert torch.equal(a_exist == 0, r_exist == 0)
    assert helpers.rel_l2(a_render.cpu().numpy(), r_render.cpu().numpy()) < 2e-5
    assert helpers.rel_l2(a_exist.cpu().numpy(), r_exist.cpu().numpy()) < 2e-5
    # 4. end to end, away from the pixels the race touched (a point splats onto a 2x2 footprint)
    keep = ~torch.nn.functional.max_pool2d(differs.float(), 5, 1, 2).bool()
    assert helpers.rel_l2((render * keep).cpu().numpy(), (r_render * keep).cpu().numpy()) < 2e-5
    assert torch.equal((exist == 0) & keep, (r_exist == 0) & keep)
    assert helpers.rel_l2((render * keep).cpu().numpy(), o_render * keep.cpu().numpy()) < 2e-5
    # 5. whole image incl. race pixels: one flipped hole moves the depth channel from 0 to ~3000, which
    #    alone is ~2e-3 relative L2 -- the reference differs from itself by that much between runs
    assert helpers.rel_l2(render[:, :3].cpu().numpy(), r_render[:, :3].cpu().numpy()) < 5e-2


def test_render_c68_vs_reference():
    """The 68-channel splat of Inpaint.pointcloud_inpainting (models/pointcloud_inpainting.py:201-206)."""
    if not refgpu.available():
        pytest.skip("oracle/_ref not built")
    W, H, focal, C = 256, 192, 128.0, 68
    pts, data, common = _inputs(W, H, focal, 0, C)
    tp, td = torch.from_numpy(pts).cuda(), torch.from_numpy(data).cuda()
    r_render, r_exist, r_zraw, r_zdeg, _ = refgpu.render_pointcloud(tp, td, W, H, focal, 120, stages=True)
    render, exist, zraw, zdeg = kb.render_pointcloud(tp, td, W, H, focal, 120, return_zee=True)
    assert torch.equal(zraw.view(torch.int32), r_zraw.view(torch.int32))
    a_render, a_exist = kb.accumulate_with_zee(tp, td, r_zdeg, focal, 120)
    assert helpers.rel_l2(a_render.cpu().numpy(), r_render.cpu().numpy()) < 2e-5
    assert helpers.rel_l2(a_exist.cpu().numpy(), r_exist.cpu().numpy()) < 2e-5
    keep = ~torch.nn.functional.max_pool2d((zdeg != r_zdeg).float(), 5, 1, 2).bool()
    assert helpers.rel_l2((render * keep).cpu().numpy(), (r_render * keep).cpu().numpy()) < 2e-5


def test_render_batch2_vs_reference():
    """B > 1 (the training-time callers of the reference batch their renders, utils/utils.py:303-337)."""
    if not refgpu.available():
        pytest.skip("oracle/_ref not built")
    W, H, focal = 128, 96, 64.0
    a, da, _ = _inputs(W, H, focal, 0, 4, step=1.0, seed=1)
    b, db, _ = _inputs(W, H, focal, 0, 4, step=0.0, seed=2)
    tp = torch.from_numpy(np.concatenate([a, b], 0)).cuda()
    td = torch.from_numpy(np.concatenate([da, db], 0)).cuda()
    r_render, r_exist, r_zraw, r_zdeg, _ = refgpu.render_pointcloud(tp, td, W, H, focal, 120, stages=True)
    render, exist, zraw, _ = kb.render_pointcloud(tp, td, W, H, focal, 120, return_zee=True)
    assert torch.equal(zraw.view(torch.int32), r_zraw.view(torch.int32))
    a_render, a_exist = kb.accumulate_with_zee(tp, td, r_zdeg, focal, 120)
    assert helpers.rel_l2(a_render.cpu().numpy(), r_render.cpu().numpy()) < 2e-5
    assert torch.equal(a_exist == 0, r_exist == 0)


@pytest.mark.parametrize("W,H,focal", [(64, 48, 32.0), (256, 192, 128.0), (1024, 768, 512.0)])
def test_fill_vs_reference_and_oracle(W, H, focal):
    if not refgpu.available():
        pytest.skip("oracle/_ref not built")
    pts, data, common = _inputs(W, H, focal, 0, 4)
    tp, td = torch.from_numpy(pts).cuda(), torch.from_numpy(data).cuda()
    render, exist = kb.render_pointcloud(tp, td, W, H, focal, 120)
    depth = render[:, 3:4] * (exist > 0).float()
    assert int((depth <= 0).sum()) > 0, "scene must have disocclusions"
    mine = kb.fill_disocclusion(render, depth)
    ref = refgpu.fill_disocclusion(render, depth)
    orc = oracle.fill_disocclusion(render.cpu().numpy(), depth.cpu().numpy())
    assert torch.equal(mine, ref)
    assert np.array_equal(mine.cpu().numpy(), orc)


def test_empty_and_culled_points():
    """All points behind the camera / outside the image: zee stays 1e6, render all zero (ragged input)."""
    W, H = 64, 48
    pts = torch.zeros(1, 3, 100, device="cuda")
    pts[:, 2] = 0.0005                      # z < 0.001 -> culled (:453)
    dat = torch.ones(1, 4, 100, device="cuda")
    render, exist, zraw, zdeg = kb.render_pointcloud(pts, dat, W, H, 32.0, 120, return_zee=True)
    assert float(exist.abs().sum()) == 0 and float(render.abs().sum()) == 0
    assert bool((zraw == 1000000.0).all())
    pts[:, 2] = 100.0
    pts[:, 0] = 1e6                         # far outside the frame
    render, exist = kb.render_pointcloud(pts, dat, W, H, 32.0, 120)
    assert float(exist.abs().sum()) == 0


def test_negative_error_branch():
    """Depths below focal*baseline/1e6 give a negative z-buffer key: exercises the CAS path of the min."""
    W, H, focal = 64, 48, 32.0
    rng = np.random.default_rng(3)
    N = 2000
    pts = np.zeros((1, 3, N), np.float32)
    pts[0, 2] = rng.uniform(0.0011, 0.01, N).astype(np.float32)
    pts[0, 0] = rng.uniform(-0.5, 0.5, N).astype(np.float32) * pts[0, 2]
    pts[0, 1] = rng.uniform(-0.5, 0.5, N).astype(np.float32) * pts[0, 2]
    tp = torch.from_numpy(pts).cuda()
    td = torch.rand(1, 4, N, device="cuda")
    _, _, zraw, _ = kb.render_pointcloud(tp, td, W, H, focal, 120, return_zee=True)
    o = oracle.splat_min(pts, H, W, focal, 120)
    assert np.array_equal(zraw.cpu().numpy().view(np.int32), o.view(np.int32))
    assert (o < 0).any()


def test_laplacian_kernel_vs_torch_conv():
    """kb_laplacian5 against the reference's formulation of spatial_filter(x, 'laplacian') (F.conv2d with the 5-tap kernel
    on a replicate-padded map, utils/common.py:398-409), evaluated by torch on the CPU."""
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 37, 53, generator=g) * 4
    ref = kb.spatial_filter(x, 'laplacian')                 # CPU tensors take the torch path
    got = kb.spatial_filter(x.cuda(), 'laplacian').cpu()
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) <= 4e-6           # five fp32 terms of magnitude <= 16
    # the kernel is asymmetric: a unit ramp along x gives -x - (x+1) - (x-1) + 4x - (x-1) = 1 away from the border
    ramp = torch.arange(53, dtype=torch.float32).view(1, 1, 1, 53).expand(1, 1, 9, 53).contiguous()
    lap = kb.spatial_filter(ramp.cuda(), 'laplacian').cpu()
    assert torch.equal(lap[..., 1:-1, 1:-1], torch.ones(1, 1, 7, 51))
