"""GPU parity of the fused per-frame loop (kb_render_frames) against the CPU oracle's restatement of
utils/common.py:238-257 (process_shift -> render -> fill -> uint8 -> getRectSubPix -> resize).

uint8 frames: the fp32 accumulation order differs between any two runs of the reference itself (float
atomicAdd), so a value sitting within 1 ulp of an integer boundary may truncate differently.  The bar:
  * < 1e-3 of the bytes differ at all;
  * outside the footprint of depth-tie holes NO byte differs by more than 2, and by 2 only where OpenCV's fixed-point
    resize ((x + 2) >> 2 after two truncated products) sums +1 on several of its four taps (<= 1e-5 of the bytes);
  * a byte may be off by more ONLY where the oracle itself says the pixel reads a filled hole whose two ray end points have
    rendered depths equal to 1e-5 relative: fill_disocclusion copies the FARTHER one (utils/common.py:904-907), so there the
    winner -- and a whole colour -- is decided by summation order, in the reference as much as here (oracle.frame_with_ties).
"""
import numpy as np
import pytest
import torch

import oracle
from ken_burns_effect_b200.utils import common as kb
import kb_helpers as helpers

pytestmark = pytest.mark.gpu


def _render_frames(pts, rgb, dep, common, W, H, steps, dolly=False, batch=16):
    st = helpers.settings(common, W, H, steps, dolly)
    poses = kb.kenburns_poses(st, common)
    f, t = st['objectFrom'], st['objectTo']
    cw, ch = max(f['intCropWidth'], t['intCropWidth']), max(f['intCropHeight'], t['intCropHeight'])
    r = kb.FrameRenderer(torch.from_numpy(pts).cuda(), torch.from_numpy(rgb).cuda(), torch.from_numpy(dep).cuda(),
                         W, H, common['dblBaseline'], cw, ch, batch=batch)
    out = torch.empty(len(poses), H, W, 3, dtype=torch.uint8, device="cuda")
    r.render_into(poses, out)
    torch.cuda.synchronize()
    return out.cpu().numpy(), poses, (cw, ch)


def _oracle_frames(pts, rgb, dep, common, W, H, poses, crop):
    """-> (frames [n,H,W,3], tie footprints bool [n,H,W])"""
    data = np.concatenate([rgb, dep], 0)
    frames, ties = [], []
    for sh, focal in poses:
        shifted = oracle.shift_points(pts, sh)
        f, t, _ = oracle.frame_with_ties(shifted, data, W, H, focal, common['dblBaseline'], crop[0], crop[1])
        frames.append(f)
        ties.append(t)
    return np.stack(frames), np.stack(ties)


def _compare(mine, ref_and_ties):
    ref, ties = ref_and_ties
    d = np.abs(mine.astype(np.int16) - ref.astype(np.int16))
    frac = float((d > 0).mean())
    outside = d * (~ties)[..., None]
    assert int(outside.max()) <= 2, f"a byte differs by {int(outside.max())} outside every depth-tie footprint"
    assert int((outside > 1).sum()) <= max(3, 1e-5 * d.size), f"{int((outside > 1).sum())} bytes off by 2 outside tie footprints"
    # (on a flat background most holes ARE ties, yet nearly all resolve identically: a pixel fed by one point has no summation order)
    n_tie = int((d > 1).sum())
    assert n_tie <= max(12, 1e-4 * d.size) and n_tie <= max(12, 0.01 * 3 * int(ties.sum())), \
        f"{n_tie} bytes differ by more than 1 (max {d.max()}); tie footprints cover {int(ties.sum())} pixels"
    assert frac < 1e-3, f"{frac:.2e} of bytes differ"
    assert helpers.rel_l2(mine, ref) < 1e-3


@pytest.mark.parametrize("W,H,focal,extra,dolly", [
    (64, 48, 32.0, 0, False),
    (256, 192, 128.0, 4099, False),
    (256, 192, 128.0, 0, True),
    (250, 190, 125.0, 33, False),       # W not a multiple of 4, even crop sizes -> sub-pixel getRectSubPix
])
def test_frames_small(W, H, focal, extra, dolly):
    oracle.set_threads(0)
    pts, rgb, dep, common = helpers.scene(W, H, focal, extra)
    steps = np.linspace(0.0, 1.0, 5).tolist()
    mine, poses, crop = _render_frames(pts, rgb, dep, common, W, H, steps, dolly, batch=3)
    ref = _oracle_frames(pts, rgb, dep, common, W, H, poses, crop)
    _compare(mine, ref)


def test_frames_dolly_dense_holes():
    """Dolly zoom at a size where the late poses have far more than 16384 holes (the foreground spreads apart): those poses
    take the thread-per-hole fill kernel (kf_fill_dense), the early ones the warp-per-hole kernel; both must give the
    oracle's frames."""
    oracle.set_threads(0)
    W, H, focal = 512, 384, 256.0
    pts, rgb, dep, common = helpers.scene(W, H, focal, 0)
    steps = [0.0, 0.5, 0.8, 1.0]
    mine, poses, crop = _render_frames(pts, rgb, dep, common, W, H, steps, dolly=True, batch=4)
    ref = _oracle_frames(pts, rgb, dep, common, W, H, poses, crop)
    _compare(mine, ref)
    # the regime really was exercised: count the holes of the last pose with the oracle's own render
    data = np.concatenate([rgb, dep], 0)
    r, e = oracle.render_pointcloud(oracle.shift_points(pts, poses[-1][0])[None], data[None], W, H, poses[-1][1], common['dblBaseline'])
    assert int(((r[0, 3] * (e[0, 0] > 0)) <= 0).sum()) > 16384


def test_frames_full_size_two_poses():
    oracle.set_threads(0)
    W, H, focal = 1024, 768, 512.0
    pts, rgb, dep, common = helpers.scene(W, H, focal, 70001)
    mine, poses, crop = _render_frames(pts, rgb, dep, common, W, H, [0.0, 1.0])
    ref = _oracle_frames(pts, rgb, dep, common, W, H, poses, crop)
    _compare(mine, ref)


def test_frames_4k_two_poses():
    """configs[3]: 3840x2160 (8.3 M points + stand-in inpainted points), both extremes of the default path."""
    oracle.set_threads(0)
    W, H, focal = 3840, 2160, 1920.0
    pts, rgb, dep, common = helpers.scene(W, H, focal, 300007)
    mine, poses, crop = _render_frames(pts, rgb, dep, common, W, H, [0.0, 1.0], batch=2)
    ref = _oracle_frames(pts, rgb, dep, common, W, H, poses, crop)
    _compare(mine, ref)


def test_frames_dolly_full_size():
    """configs[2]'s loop at 1024x768: dolly zoom, late poses are 30-65 % holes (thread-per-hole fill path)."""
    oracle.set_threads(0)
    W, H, focal = 1024, 768, 512.0
    pts, rgb, dep, common = helpers.scene(W, H, focal, 0)
    mine, poses, crop = _render_frames(pts, rgb, dep, common, W, H, [0.0, 0.6, 1.0], dolly=True, batch=3)
    ref = _oracle_frames(pts, rgb, dep, common, W, H, poses, crop)
    _compare(mine, ref)


def test_frames_batching_invariance():
    """Size-independent property at full size: the frames do not depend on how poses are batched, and the
    z-buffer/hole structure is deterministic -> at most summation-order noise between batchings."""
    W, H, focal = 1024, 768, 512.0
    pts, rgb, dep, common = helpers.scene(W, H, focal, 0)
    steps = np.linspace(0.0, 1.0, 7).tolist()
    a, _, _ = _render_frames(pts, rgb, dep, common, W, H, steps, batch=7)
    b, _, _ = _render_frames(pts, rgb, dep, common, W, H, steps, batch=2)
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    assert (d > 1).mean() < 2e-5 and (d > 0).mean() < 1e-3


def test_process_kenburns_host_path_vs_oracle():
    """The public process_kenburns (frames through the pinned-memory path: device staging buffers, copy stream, pooled host buffer)
    against the oracle's frames of the same cloud and poses."""
    oracle.set_threads(0)
    W, H, focal = 256, 192, 128.0
    pts, rgb, dep, common = helpers.scene(W, H, focal, 0)
    steps = np.linspace(0.0, 1.0, 4).tolist()
    st = helpers.settings(common, W, H, steps, dolly=True)   # dolly: no inpainting stage
    common = dict(common)
    P = W * H
    common['tensorRawImage'] = torch.from_numpy(rgb).cuda().view(1, 3, H, W)
    common['tensorRawDepth'] = torch.from_numpy(dep).cuda().view(1, 1, H, W)
    common['tensorRawDisparity'] = (focal * 120) / (common['tensorRawDepth'] + 1e-7)
    common['tensorRawPoints'] = torch.from_numpy(pts).cuda().view(1, 3, P)
    frames = kb.process_kenburns(st, common, None)
    assert len(frames) == 4 and frames[0].shape == (H, W, 3) and frames[0].dtype == np.uint8
    poses = kb.kenburns_poses(st, common)
    _compare(np.stack(frames), _oracle_frames(pts, rgb, dep, common, W, H, poses, kb.crop_size(st)))
