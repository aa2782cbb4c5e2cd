"""CPU oracle for the novel-view render path -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this
package.  The product package (ken_burns_effect_b200) must never import it; tests/test_abi.py::test_product_never_touches_the_oracle
greps for that.

`kb_oracle.c` restates the reference's CUDA kernels (utils/common.py:428-937) and its numpy/OpenCV frame
tail (utils/common.py:255-257) in plain C; this module is the numpy/ctypes face of it.  Parity is pinned
against the reference's own kernels compiled by build_ref.py (see tests/test_gpu_render.py, tests/test_gpu_mask.py and the
fixtures in tests/golden/).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i16p = ctypes.POINTER(ctypes.c_int16)
c_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "libkb_oracle.so")
    src = os.path.join(_HERE, "kb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libkb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.kbo_max_threads.restype = ctypes.c_int
    return _LIB


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(c_f32p)


def set_threads(n):
    lib().kbo_set_threads(int(n))


def max_threads():
    return int(lib().kbo_max_threads())


def shift_points(xyz, shift3):
    """process_shift tensor half (common.py:104-109). xyz [3,N] -> [3,N]."""
    xyz, px = _f32(xyz)
    s, ps = _f32(np.asarray(shift3).reshape(3))
    out = np.empty_like(xyz)
    lib().kbo_shift_points(px, ctypes.c_long(xyz.shape[-1]), ps, out.ctypes.data_as(c_f32p))
    return out


def splat_min(xyz, H, W, focal, baseline, want_idx=False):
    """updateZee (common.py:434-507). xyz [B,3,N] -> zee [B,1,H,W] (and pix_idx [B,N])."""
    xyz, px = _f32(xyz)
    B, _, N = xyz.shape
    zee = np.empty((B, 1, H, W), np.float32)
    idx = np.empty((B, N), np.int32) if want_idx else None
    lib().kbo_splat_min(px, B, ctypes.c_long(N), ctypes.c_double(focal), ctypes.c_double(baseline),
                        zee.ctypes.data_as(c_f32p), H, W,
                        idx.ctypes.data_as(c_i32p) if want_idx else None)
    return (zee, idx) if want_idx else zee


def mask_zee(xyz, H, W, focal, baseline):
    """The kernel of generate_mask (common.py:696-827) in index order. xyz [B,3,N] (already shifted) -> mask [B,N]."""
    xyz, px = _f32(xyz)
    B, _, N = xyz.shape
    mask = np.empty((B, N), np.float32)
    zee = np.empty((B, H, W), np.float32)
    lib().kbo_mask_zee(px, B, ctypes.c_long(N), ctypes.c_double(focal), ctypes.c_double(baseline), H, W,
                       mask.ctypes.data_as(c_f32p), zee.ctypes.data_as(c_f32p))
    return mask


def degrid(zee, mode=0):
    """updateDegrid (common.py:524-568). mode 0 = race-free (canonical), 1 = in-place raster order."""
    zee, pz = _f32(zee)
    B, _, H, W = zee.shape
    out = np.empty_like(zee)
    lib().kbo_degrid(pz, out.ctypes.data_as(c_f32p), B, H, W, mode)
    return out


def splat_accum(xyz, data, zee, focal, baseline):
    """updateOutput (common.py:585-669). -> out [B,C+1,H,W]."""
    xyz, px = _f32(xyz)
    data, pd = _f32(data)
    zee, pz = _f32(zee)
    B, C, N = data.shape
    _, _, H, W = zee.shape
    out = np.empty((B, C + 1, H, W), np.float32)
    lib().kbo_splat_accum(px, pd, B, ctypes.c_long(N), C, ctypes.c_double(focal), ctypes.c_double(baseline),
                          pz, out.ctypes.data_as(c_f32p), H, W)
    return out


def normalize(out):
    out, po = _f32(out)
    B, C1, H, W = out.shape
    render = np.empty((B, C1 - 1, H, W), np.float32)
    existing = np.empty((B, 1, H, W), np.float32)
    lib().kbo_normalize(po, B, C1 - 1, H, W, render.ctypes.data_as(c_f32p), existing.ctypes.data_as(c_f32p))
    return render, existing


def render_pointcloud(xyz, data, W, H, focal, baseline, degrid_mode=0, want_zee=False):
    """render_pointcloud (common.py:428-686): -> (render [B,C,H,W], existing [B,1,H,W])."""
    xyz, px = _f32(xyz)
    data, pd = _f32(data)
    B, C, N = data.shape
    render = np.empty((B, C, H, W), np.float32)
    existing = np.empty((B, 1, H, W), np.float32)
    zraw = np.empty((B, 1, H, W), np.float32)
    zout = np.empty((B, 1, H, W), np.float32)
    rc = lib().kbo_render_pointcloud(px, pd, B, ctypes.c_long(N), C, W, H, ctypes.c_double(focal),
                                     ctypes.c_double(baseline), degrid_mode,
                                     render.ctypes.data_as(c_f32p), existing.ctypes.data_as(c_f32p),
                                     zraw.ctypes.data_as(c_f32p), zout.ctypes.data_as(c_f32p))
    assert rc == 0
    if want_zee:
        return render, existing, zraw, zout
    return render, existing


def fill_disocclusion(inp, depth, want_xy=False):
    """fill_disocclusion (common.py:833-937)."""
    inp, pi = _f32(inp)
    depth, pd = _f32(depth)
    B, C, H, W = inp.shape
    out = np.empty_like(inp)
    xy = np.empty((B, H, W, 2), np.int32) if want_xy else None
    lib().kbo_fill(pi, pd, B, C, H, W, out.ctypes.data_as(c_f32p),
                   xy.ctypes.data_as(c_i32p) if want_xy else None)
    return (out, xy) if want_xy else out


def fill_ends(inp, depth):
    """fill_disocclusion that also reports, per hole pixel, the two end points of the winning ray -> (out, ends [B,H,W,4])."""
    inp, pi = _f32(inp)
    depth, pd = _f32(depth)
    B, C, H, W = inp.shape
    out = np.empty_like(inp)
    ends = np.empty((B, H, W, 4), np.int32)
    lib().kbo_fill_ends(pi, pd, B, C, H, W, out.ctypes.data_as(c_f32p), None, ends.ctypes.data_as(c_i32p))
    return out, ends


def frame_with_ties(xyz_shifted, rgbd, W, H, focal, baseline, crop_w, crop_h, rel_gap=1e-5):
    """frame() assembled from the stage functions, plus a bool [H,W] map of the OUTPUT pixels whose bilinear footprint
    (resize o getRectSubPix, common.py:256-257) touches a filled hole that was decided by a depth tie: a hole copies the
    FARTHER of the two end points of its shortest ray (:904-907); when their rendered depths agree to `rel_gap` the winner
    depends on the order of the fp32 atomicAdds of updateOutput (:641), in the reference as much as anywhere else."""
    render, existing = render_pointcloud(np.asarray(xyz_shifted)[None], np.asarray(rgbd)[None], W, H, focal, baseline)
    dmask = render[:, 3:4] * (existing > 0.0)
    filled, ends = fill_ends(render, dmask)
    u8 = to_uint8(filled[0])
    out = resize_linear(getrectsubpix(u8, crop_w, crop_h, W / 2.0, H / 2.0), W, H)
    e = ends[0]
    hole = e[..., 0] >= 0
    da = dmask[0, 0][np.clip(e[..., 1], 0, H - 1), np.clip(e[..., 0], 0, W - 1)]
    db = dmask[0, 0][np.clip(e[..., 3], 0, H - 1), np.clip(e[..., 2], 0, W - 1)]
    tie = hole & (np.abs(da - db) <= rel_gap * np.maximum(np.abs(da), np.abs(db)))
    # a tie at end point level also decides every hole that copies the same pair; mark the holes themselves
    # footprint of output pixel (X, Y): patch columns sx, sx+1 with sx = floor((X+.5)*cw/W-.5); render columns
    # floor(ox+sx) .. floor(ox+sx+1)+1 with ox = W/2-(cw-1)/2  -> a 3x3 (at most) block; dilate generously by 2
    ox, oy = W / 2.0 - (crop_w - 1) * 0.5, H / 2.0 - (crop_h - 1) * 0.5
    X = np.arange(W)
    Y = np.arange(H)
    sx = np.floor((X + 0.5) * crop_w / W - 0.5)
    sy = np.floor((Y + 0.5) * crop_h / H - 0.5)
    rx0 = np.clip(np.floor(ox + sx).astype(np.int64) - 1, 0, W - 1)
    ry0 = np.clip(np.floor(oy + sy).astype(np.int64) - 1, 0, H - 1)
    t = tie.astype(np.int32)
    ii = np.pad(t, ((1, 0), (1, 0))).cumsum(0).cumsum(1)                 # integral image
    rx1 = np.clip(rx0 + 4, 0, W)
    ry1 = np.clip(ry0 + 4, 0, H)
    cnt = ii[ry1][:, rx1] - ii[ry0][:, rx1] - ii[ry1][:, rx0] + ii[ry0][:, rx0]
    return out, cnt > 0, int(tie.sum())


def to_uint8(render):
    """common.py:255 for one sample: render [>=3,H,W] float -> uint8 [H,W,3]."""
    render, pr = _f32(render)
    _, H, W = render.shape
    out = np.empty((H, W, 3), np.uint8)
    lib().kbo_to_uint8(pr, H, W, out.ctypes.data_as(c_u8p))
    return out


def getrectsubpix(img, patch_w, patch_h, cx, cy):
    img = np.ascontiguousarray(img, np.uint8)
    sh, sw, _ = img.shape
    out = np.empty((patch_h, patch_w, 3), np.uint8)
    lib().kbo_getrectsubpix_8u3(img.ctypes.data_as(c_u8p), sh, sw, patch_w, patch_h,
                                ctypes.c_double(cx), ctypes.c_double(cy), out.ctypes.data_as(c_u8p))
    return out


def resize_tables(ssize, dsize, is_x):
    ofs = np.empty(dsize, np.int32)
    coef = np.empty((dsize, 2), np.int16)
    lib().kbo_resize_tables(ssize, dsize, int(is_x), ofs.ctypes.data_as(c_i32p), coef.ctypes.data_as(c_i16p))
    return ofs, coef


def resize_linear(img, dw, dh):
    img = np.ascontiguousarray(img, np.uint8)
    sh, sw, _ = img.shape
    out = np.empty((dh, dw, 3), np.uint8)
    lib().kbo_resize_linear_8u3(img.ctypes.data_as(c_u8p), sh, sw, dh, dw, out.ctypes.data_as(c_u8p))
    return out


def frame(xyz_shifted, rgbd, W, H, focal, baseline, crop_w, crop_h, degrid_mode=0):
    """One iteration of process_kenburns' loop body after process_shift (common.py:246-257)."""
    xyz, px = _f32(xyz_shifted)
    rgbd, pd = _f32(rgbd)
    N = xyz.shape[-1]
    out = np.empty((H, W, 3), np.uint8)
    rc = lib().kbo_frame(px, pd, ctypes.c_long(N), W, H, ctypes.c_double(focal), ctypes.c_double(baseline),
                         crop_w, crop_h, degrid_mode, out.ctypes.data_as(c_u8p))
    assert rc == 0
    return out


def median5_binary(x):
    x, px = _f32(x)
    H, W = x.shape[-2:]
    out = np.empty_like(x)
    lib().kbo_median5_binary(px, H, W, out.ctypes.data_as(c_f32p))
    return out
