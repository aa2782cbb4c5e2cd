"""Seeded synthetic inputs for tests and bench.py: an image made of smooth
per-channel sinusoid gradients + random filled rectangles/ellipses + noise, and a matching disparity map
(background plane + the same shapes at larger disparity, blurred) that produces realistic disocclusions.
No dataset or checkpoint is needed (there is no network access)."""
import math

import cv2
import numpy as np


def synthetic_scene(W=1024, H=768, seed=1234, n_shapes=24, baseline=120.0):
    """-> (image uint8 [H,W,3] BGR, disparity float32 [H,W] in (0, baseline])."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.empty((H, W, 3), np.float32)
    for c in range(3):
        fx, fy, ph = rng.uniform(1.0, 4.0), rng.uniform(1.0, 4.0), rng.uniform(0, 2 * math.pi)
        img[:, :, c] = 127.5 + 90.0 * np.sin(2 * math.pi * (fx * xx / W + fy * yy / H) + ph)
    disp = np.full((H, W), 20.0, np.float32)
    s = min(W, H)
    for _ in range(n_shapes):
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        a, b = int(rng.integers(s // 24, s // 5)), int(rng.integers(s // 24, s // 5))
        col = tuple(float(v) for v in rng.integers(0, 256, 3))
        d = float(rng.uniform(40.0, baseline))
        if rng.random() < 0.5:
            cv2.rectangle(img, (cx - a, cy - b), (cx + a, cy + b), col, -1)
            cv2.rectangle(disp, (cx - a, cy - b), (cx + a, cy + b), d, -1)
        else:
            cv2.ellipse(img, (cx, cy), (a, b), 0, 0, 360, col, -1)
            cv2.ellipse(disp, (cx, cy), (a, b), 0, 0, 360, d, -1)
    img += rng.normal(0.0, 4.0, img.shape).astype(np.float32)
    img = np.clip(img, 0, 255).astype(np.uint8)
    disp = cv2.GaussianBlur(disp, (0, 0), 1.0)
    disp = disp / disp.max() * np.float32(baseline)
    return img, disp.astype(np.float32)


def default_zoom(W, H, dolly=False):
    """The crop windows kbe.py falls back to (kbe.py:128-140)."""
    if not dolly:
        frm = dict(dblCenterU=W / 2.15, dblCenterV=H / 2.15, intCropWidth=int(math.floor(0.90 * W)), intCropHeight=int(math.floor(0.90 * H)))
        to = dict(dblCenterU=W / 1.85, dblCenterV=H / 1.85, intCropWidth=int(math.floor(0.85 * W)), intCropHeight=int(math.floor(0.85 * H)))
    else:
        frm = dict(dblCenterU=W / 2, dblCenterV=H / 2, intCropWidth=int(math.floor(0.8 * W)), intCropHeight=int(math.floor(0.8 * H)))
        to = dict(dblCenterU=W / 2, dblCenterV=H / 2, intCropWidth=int(math.floor(0.3 * W)), intCropHeight=int(math.floor(0.3 * H)))
    return {'objectFrom': frm, 'objectTo': to}


def _standin_inpaint(pts, rgb, dep, W, H, focal, shift):
    """A numpy stand-in for one process_inpaint pass (utils/common.py:47-81) WITHOUT the CNN: warp the grid
    cloud by `shift`, find the disoccluded pixels of that view, give them the depth/colour of the farther of
    their nearest valid row neighbours (what a good inpainter would hallucinate: background), and
    back-project them into the reference frame.  Input generator only -- not part of any parity claim."""
    x, y, z = pts[0].astype(np.float64) + shift[0], pts[1].astype(np.float64) + shift[1], pts[2].astype(np.float64) + shift[2]
    ok = z > 1e-3
    u = np.rint(focal * x / np.where(ok, z, 1.0) + W / 2.0 - 0.5).astype(np.int64)
    v = np.rint(focal * y / np.where(ok, z, 1.0) + H / 2.0 - 0.5).astype(np.int64)
    ok &= (u >= 0) & (u < W) & (v >= 0) & (v < H)
    pix = (v * W + u)[ok]
    idx = np.nonzero(ok)[0]
    order = np.lexsort((z[idx], pix))
    pix_s, idx_s = pix[order], idx[order]
    first = np.ones(len(pix_s), bool)
    first[1:] = pix_s[1:] != pix_s[:-1]
    win = np.full(W * H, -1, np.int64)
    win[pix_s[first]] = idx_s[first]
    win = win.reshape(H, W)
    hole = win < 0
    # nearest valid neighbour to the left / right on the same row
    cols = np.arange(W)[None, :].repeat(H, 0)
    left = np.maximum.accumulate(np.where(hole, -1, cols), axis=1)
    right = np.minimum.accumulate(np.where(hole, W, cols)[:, ::-1], axis=1)[:, ::-1]
    rows = np.arange(H)[:, None].repeat(W, 1)
    lw = np.where(left >= 0, win[rows, np.clip(left, 0, W - 1)], -1)
    rw = np.where(right < W, win[rows, np.clip(right, 0, W - 1)], -1)
    zl = np.where(lw >= 0, z[np.clip(lw, 0, None)], -np.inf)
    zr = np.where(rw >= 0, z[np.clip(rw, 0, None)], -np.inf)
    src = np.where(zl >= zr, lw, rw)
    sel = hole & (src >= 0)
    hv, hu = np.nonzero(sel)
    s_idx = src[sel]
    D = z[s_idx].astype(np.float32)
    ul = (np.linspace(-0.5 * W + 0.5, 0.5 * W - 0.5, W, dtype=np.float32) * np.float32(1.0 / focal))
    vl = (np.linspace(-0.5 * H + 0.5, 0.5 * H - 0.5, H, dtype=np.float32) * np.float32(1.0 / focal))
    new_pts = np.stack([D * ul[hu], D * vl[hv], D], 0).astype(np.float32) - np.asarray(shift, np.float32)[:, None]
    return new_pts, rgb[:, s_idx], D[None, :]


def scene_cloud(W=1024, H=768, seed=1234, focal=None, baseline=120.0, extra_points=0, inpaint_standin=False):
    """A ready-to-render cloud in the layout process_kenburns keeps in objectCommon (numpy, CPU):
    points [3,N] (depth_to_points of depth = f*B/(disp+1e-7)), rgb [3,N] in [0,1], depth [1,N], plus the
    objectCommon scalars process_shift needs.  extra_points > 0 appends that many points resampled from the
    far background at jittered positions, standing in for the inpainted points of utils/common.py:75-80."""
    if focal is None:
        focal = max(W, H) / 2.0
    img, disp = synthetic_scene(W, H, seed, baseline=baseline)
    depth = (np.float32(focal * baseline) / (disp + np.float32(1e-7))).astype(np.float32)
    u = (np.linspace(-0.5 * W + 0.5, 0.5 * W - 0.5, W, dtype=np.float32) * np.float32(1.0 / focal))[None, :]
    v = (np.linspace(-0.5 * H + 0.5, 0.5 * H - 0.5, H, dtype=np.float32) * np.float32(1.0 / focal))[:, None]
    pts = np.stack([depth * u, depth * v, depth], 0).reshape(3, -1).astype(np.float32)
    rgb = (img[:, :, ::-1].astype(np.float32) / 255.0).transpose(2, 0, 1).reshape(3, -1)
    dep = depth.reshape(1, -1)
    if extra_points > 0:
        rng = np.random.default_rng(seed + 1)
        idx = np.sort(rng.integers(0, W * H, extra_points))
        ex = pts[:, idx].copy()
        scale = rng.uniform(1.02, 1.3, extra_points).astype(np.float32)   # pushed behind the surface
        ex *= scale[None, :]
        ex[0] += rng.normal(0, 0.5, extra_points).astype(np.float32) * ex[2] / np.float32(focal)
        pts = np.concatenate([pts, ex], 1)
        rgb = np.concatenate([rgb, rgb[:, idx]], 1)
        dep = np.concatenate([dep, ex[2:3]], 1)
    depthrange = cv2.minMaxLoc(depth[128:-128, 128:-128]) if (H > 256 and W > 256) else cv2.minMaxLoc(depth)
    if inpaint_standin:
        # the two extreme views of the default camera path, shift scaled by 1.1 like utils/common.py:218
        from . import common as kb
        cm = {'dblFocal': float(focal), 'dblBaseline': baseline, 'intWidth': W, 'intHeight': H,
              'objectDepthrange': depthrange}
        zoom = default_zoom(W, H)
        st = {'dblSteps': [0.0, 1.0], 'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'], 'dolly': False}
        grid_pts, grid_rgb, grid_dep = pts[:, :W * H], rgb[:, :W * H], dep[:, :W * H]
        for sh, _ in kb.kenburns_poses(st, cm):
            shift = (sh.astype(np.float32) * np.float32(1.1)).astype(np.float64)
            a, b, c = _standin_inpaint(grid_pts, grid_rgb, grid_dep, W, H, float(focal), shift)
            pts = np.concatenate([pts, a], 1)
            rgb = np.concatenate([rgb, b], 1)
            dep = np.concatenate([dep, c], 1)
    mn, mx, mnl, mxl = cv2.minMaxLoc(depth[128:-128, 128:-128]) if (H > 256 and W > 256) else cv2.minMaxLoc(depth)
    common = {
        'dblFocal': float(focal), 'dblBaseline': baseline, 'intWidth': W, 'intHeight': H,
        'objectDepthrange': (mn, mx, mnl, mxl),
    }
    return np.ascontiguousarray(pts), np.ascontiguousarray(rgb), np.ascontiguousarray(dep), common
