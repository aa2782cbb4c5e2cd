"""Host-side mirror of the reference's view-synthesis operators (utils/common.py in pierlj/ken-burns-effect).

Same function names, argument meaning and return values as the reference so that `models/*` and
`utils/pipeline.py` style callers work unchanged, but every CUDA kernel the reference JIT-compiles from a
source string (utils/common.py:428-937) is replaced by a call into libkb200.so (include/kb200.h), and the
per-frame loop of process_kenburns (utils/common.py:222-260) is one fused multi-pose call.

There is no CPU path for the kernels: tensors must live on a CUDA device and the shared library must be
present, otherwise a RuntimeError is raised.
"""
import ctypes
import math

import numpy as np
import torch

from .. import _native as nat

# Poses rendered per fused launch group (kb_render_frames), bounded by KB_MAX_POSES.  Measured at 1024x768 (profiles/): 32 poses
# per launch render 4.8 % faster than 16 when the frames stay on the device, but frames bound for host memory are copied
# batch by batch on a second stream and overlap better in batches of 16 (end to end 19.7 k vs 18.6 k frames/s).
FRAME_BATCH = 32
FRAME_BATCH_TO_HOST = 16
# (A shorter FIRST host-bound batch -- 4 poses, so that the copy engine starts 0.4 ms earlier -- was measured: 20.7 k vs 20.9 k
# frames/s end to end, i.e. nothing; not kept.)


def _stream():
    # queried per call: the reference captures the stream once at import (utils/common.py:267-269)
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _need_cuda(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise RuntimeError("ken_burns_effect_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if t.dtype != torch.float32:
            raise RuntimeError("ken_burns_effect_b200 kernels are fp32-in/fp32-out like the reference")


# ---------------------------------------------------------------------------------------------------------
# geometry helpers (reference: utils/common.py:382-426)
# ---------------------------------------------------------------------------------------------------------

def depth_to_points(tensorDepth, dblFocal):
    """Pinhole back-projection, utils/common.py:382-392.  [B,1,H,W] -> [B,3,H,W].

    The pixel-centre ramps are built exactly like the reference (CPU torch.linspace, then * (1/f) in fp32,
    then moved to the depth's device) so that the products are bit-identical.
    """
    B, _, H, W = tensorDepth.shape
    inv_f = 1.0 / dblFocal
    u = (torch.linspace(-0.5 * W + 0.5, 0.5 * W - 0.5, W) * inv_f).view(1, 1, 1, W).to(tensorDepth)
    v = (torch.linspace(-0.5 * H + 0.5, 0.5 * H - 0.5, H) * inv_f).view(1, 1, H, 1).to(tensorDepth)
    return torch.cat([tensorDepth * u, tensorDepth * v, tensorDepth], 1)


_LAPLACE_TAPS = ((0, 1, -1.0), (0, 2, -1.0), (1, 1, 4.0), (1, 0, -1.0), (2, 0, -1.0))


def spatial_filter(tensorInput, strType):
    """utils/common.py:394-426: 'laplacian' (the reference's asymmetric 5-tap kernel, replicate padding),
    'median-3' / 'median-5' (reflect padding)."""
    if strType == 'laplacian' and tensorInput.is_cuda and tensorInput.dtype == torch.float32:
        x = tensorInput.contiguous()
        out = torch.empty_like(x)
        B, C, H, W = x.shape
        nat.check(nat.lib().kb_laplacian5(_ptr(x), _ptr(out), B * C, H, W, _stream()), "kb_laplacian5")
        return out
    if strType == 'laplacian':
        C = tensorInput.size(1)
        k = tensorInput.new_zeros(C, C, 3, 3)
        for c in range(C):
            for (r, s, val) in _LAPLACE_TAPS:
                k[c, c, r, s] = val
        padded = torch.nn.functional.pad(tensorInput, [1, 1, 1, 1], mode='replicate')
        return torch.nn.functional.conv2d(padded, k)

    if strType in ('median-3', 'median-5'):
        r = 1 if strType == 'median-3' else 2
        if strType == 'median-5' and tensorInput.is_cuda and tensorInput.size(1) == 1 \
                and bool(((tensorInput == 0) | (tensorInput == 1)).all()):
            # binary map: the 25-way median is (5x5 box count >= 13); one kernel instead of a
            # [B,1,H,W,25] unfold (79 MB at 1024x768) + sort
            return median5_binary(tensorInput)
        n = 2 * r + 1
        x = torch.nn.functional.pad(tensorInput, [r, r, r, r], mode='reflect')
        x = x.unfold(2, n, 1).unfold(3, n, 1).contiguous()
        x = x.view(x.size(0), x.size(1), x.size(2), x.size(3), n * n)
        return x.median(-1, False)[0]
    return None


def median5_binary(tensorMask):
    """spatial_filter(mask, 'median-5') for a {0,1} mask [B,1,H,W] (utils/common.py:417-421)."""
    _need_cuda(tensorMask)
    x = tensorMask.contiguous()
    out = torch.empty_like(x)
    B, _, H, W = x.shape
    nat.check(nat.lib().kb_median5_binary(_ptr(x), _ptr(out), B, H, W, _stream()), "kb_median5_binary")
    return out


# ---------------------------------------------------------------------------------------------------------
# render_pointcloud / fill_disocclusion (reference: utils/common.py:428-686, :833-937)
# ---------------------------------------------------------------------------------------------------------

def render_pointcloud(tensorInput, tensorData, intWidth, intHeight, dblFocal, dblBaseline, return_zee=False):
    """Z-buffered bilinear splat.  tensorInput [B,3,N] (already shifted points), tensorData [B,C,N] ->
    (tensorRender [B,C,H,W], tensorExisting [B,1,H,W]); same contract as utils/common.py:428."""
    _need_cuda(tensorInput, tensorData)
    pts = tensorInput.contiguous()
    dat = tensorData.contiguous()
    B, C, N = dat.shape
    if pts.shape != (B, 3, N):
        raise RuntimeError(f"render_pointcloud: points {tuple(pts.shape)} do not match data {tuple(dat.shape)}")
    H, W = int(intHeight), int(intWidth)
    L = nat.lib()
    render = torch.empty(B, C, H, W, device=pts.device, dtype=torch.float32)
    existing = torch.empty(B, 1, H, W, device=pts.device, dtype=torch.float32)
    ws = torch.empty(L.kb_render_workspace_bytes(B, C, H, W), device=pts.device, dtype=torch.uint8)
    nat.check(L.kb_render_pointcloud(_ptr(pts), _ptr(dat), B, N, C, W, H, float(dblFocal), float(dblBaseline),
                                     _ptr(render), _ptr(existing), _ptr(ws), _stream()), "kb_render_pointcloud")
    if return_zee:
        P = H * W
        z = ws[:8 * B * P].view(torch.float32)
        return render, existing, z[:B * P].view(B, 1, H, W).clone(), z[B * P:].view(B, 1, H, W).clone()
    return render, existing


def render_rows(tensorInput, data_rows, intWidth, intHeight, dblFocal, dblBaseline):
    """render_pointcloud (utils/common.py:428-686) for per-point data given as NHWC rows [B, N, C] (a view whose last stride is
    1 and whose row stride is a multiple of 4 floats).  Returns (accum [B,H,W,Cp] channels-last accumulators with the weight in
    channel C, weight [B,1,H,W]); kb_normalize_rows / normalize_rows() finishes them in place."""
    _need_cuda(tensorInput, data_rows)
    pts = tensorInput.contiguous()
    B, N, C = data_rows.shape
    if pts.shape != (B, 3, N) or data_rows.stride(2) != 1 or data_rows.stride(1) % 4 or data_rows.stride(0) != N * data_rows.stride(1):
        raise RuntimeError(f"render_rows: points {tuple(pts.shape)} / rows {tuple(data_rows.shape)} strides {data_rows.stride()}")
    H, W = int(intHeight), int(intWidth)
    L = nat.lib()
    P = H * W
    zz = torch.empty(2, B, H, W, device=pts.device, dtype=torch.float32)
    nat.check(L.kb_splat_min(_ptr(pts), B, N, None, float(dblFocal), float(dblBaseline), _ptr(zz[0]), H, W, None, _stream()), "kb_splat_min")
    nat.check(L.kb_degrid(_ptr(zz[0]), _ptr(zz[1]), B, H, W, _stream()), "kb_degrid")
    Cp = L.kb_accum_channels(C)
    acc = torch.empty(B, H, W, Cp, device=pts.device, dtype=torch.float32)
    nat.check(L.kb_splat_accum_rows(_ptr(pts), _ptr(data_rows), data_rows.stride(1), B, N, C, None, float(dblFocal),
                                    float(dblBaseline), _ptr(zz[1]), _ptr(acc), H, W, _stream()), "kb_splat_accum_rows")
    weight = torch.empty(B, 1, H, W, device=pts.device, dtype=torch.float32)
    nat.check(L.kb_accum_weight(_ptr(acc), B, C, H, W, _ptr(weight), _stream()), "kb_accum_weight")
    return acc, weight


def normalize_rows(acc, C, mask):
    """In place: acc[..., :C] = acc[..., :C] / (acc[..., C] + 1e-7) * mask, acc[..., C] = mask  (mask [B,1,H,W] or [B,H,W])."""
    B, H, W, _ = acc.shape
    m = mask.reshape(B, H, W).contiguous()
    nat.check(nat.lib().kb_normalize_rows(_ptr(acc), B, C, H, W, _ptr(m), _stream()), "kb_normalize_rows")
    return acc


def accumulate_with_zee(tensorInput, tensorData, tensorZee, dblFocal, dblBaseline):
    """Passes 3+4 of render_pointcloud (updateOutput + epilogue, utils/common.py:585-686) against a z-buffer
    supplied by the caller ([B,1,H,W], already degridded).  Lets a test isolate the accumulation from the
    reference's racy degrid pass."""
    _need_cuda(tensorInput, tensorData, tensorZee)
    pts, dat, zee = tensorInput.contiguous(), tensorData.contiguous(), tensorZee.contiguous()
    B, C, N = dat.shape
    H, W = zee.shape[-2:]
    L = nat.lib()
    Cp = L.kb_accum_channels(C)
    acc = torch.empty(B, H, W, Cp, device=pts.device, dtype=torch.float32)
    render = torch.empty(B, C, H, W, device=pts.device, dtype=torch.float32)
    existing = torch.empty(B, 1, H, W, device=pts.device, dtype=torch.float32)
    nat.check(L.kb_splat_accum(_ptr(pts), _ptr(dat), B, N, C, None, float(dblFocal), float(dblBaseline), _ptr(zee),
                               _ptr(acc), H, W, _stream()), "kb_splat_accum")
    nat.check(L.kb_normalize(_ptr(acc), B, C, H, W, _ptr(render), _ptr(existing), _stream()), "kb_normalize")
    return render, existing


def generate_mask(tensorInput, tensorShift, intWidth, intHeight, dblFocal, dblBaseline):
    """utils/common.py:689-830: which points of the grid cloud [B,3,H*W] stay visible after the camera shift -> [B,1,H,W]
    (1 = the point owns the z-buffer cell it projects to), median-5 filtered.  The reference's kernel races; this is its
    deterministic index-order outcome (include/kb200.h: kb_generate_mask).  Training-time helper (utils/utils.py:285-287)."""
    pts = (tensorInput + tensorShift).contiguous()
    _need_cuda(pts)
    B, _, N = pts.shape
    H, W = int(intHeight), int(intWidth)
    if N != H * W:
        raise RuntimeError(f"generate_mask: the cloud must be the H*W grid ({H * W} points), got {N}")   # :829 views it as an image
    L = nat.lib()
    mask = torch.empty(B, N, device=pts.device, dtype=torch.float32)
    ws = torch.empty(L.kb_mask_workspace_bytes(B, N, H, W), device=pts.device, dtype=torch.uint8)
    nat.check(L.kb_generate_mask(_ptr(pts), B, N, float(dblFocal), float(dblBaseline), H, W, _ptr(mask), _ptr(ws), _stream()),
              "kb_generate_mask")
    return spatial_filter(mask.view(-1, 1, H, W), 'median-5')


def fill_disocclusion(tensorInput, tensorDepth):
    """utils/common.py:833-937.  [B,C,H,W], [B,1,H,W] -> filled copy of the input."""
    _need_cuda(tensorInput, tensorDepth)
    x = tensorInput.contiguous()
    d = tensorDepth.contiguous()
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    nat.check(nat.lib().kb_fill(_ptr(x), _ptr(d), _ptr(out), B, C, H, W, _stream()), "kb_fill")
    return out


# ---------------------------------------------------------------------------------------------------------
# camera path (reference: utils/common.py:83-112, :172-263)
# ---------------------------------------------------------------------------------------------------------

def _shift_scalars(objectSettings, objectCommon, dblFocal):
    """Python-double camera shift of process_shift, utils/common.py:88-100 (incl. the crop-relative argmin
    location quirk: objectDepthrange[2] is relative to the [128:-128] crop and stays so)."""
    rng = objectCommon['objectDepthrange']
    closest = rng[0] + (objectSettings['dblDepthTo'] - objectSettings['dblDepthFrom'])
    from_u, from_v = rng[2][0], rng[2][1]
    to_u = from_u + objectSettings['dblShiftU']
    to_v = from_v + objectSettings['dblShiftV']
    half_w = objectCommon['intWidth'] / 2.0
    half_h = objectCommon['intHeight'] / 2.0
    from_x = ((from_u - half_w) * closest) / dblFocal
    from_y = ((from_v - half_h) * closest) / dblFocal
    to_x = ((to_u - half_w) * closest) / dblFocal
    to_y = ((to_v - half_h) * closest) / dblFocal
    return from_x - to_x, from_y - to_y, objectSettings['dblDepthTo'] - objectSettings['dblDepthFrom']


def process_shift(objectSettings, objectCommon, dblFocal=None):
    """utils/common.py:83-112 -> (shifted points [B,3,N], tensorShift [1,3,1])."""
    if dblFocal is None:
        dblFocal = objectCommon['dblFocal']
    sx, sy, sz = _shift_scalars(objectSettings, objectCommon, dblFocal)
    pts = objectSettings['tensorPoints']
    _need_cuda(pts)
    tensorShift = torch.FloatTensor([sx, sy, sz]).view(1, 3, 1).to(pts.device)
    src = pts.contiguous()
    B, _, N = src.shape
    out = torch.empty_like(src)
    sh = tensorShift.view(1, 3).expand(B, 3).contiguous()
    nat.check(nat.lib().kb_shift_points(_ptr(src), _ptr(sh), _ptr(out), B, N, _stream()), "kb_shift_points")
    return out, tensorShift


def _pose_settings(objectSettings, objectCommon, dblStep):
    """Per-step camera scalars of process_kenburns, utils/common.py:182-198 and :223-236."""
    dblFrom = 1.0 - dblStep
    dblTo = 1.0 - dblFrom
    f_w = objectSettings['objectFrom']['intCropWidth']
    t_w = objectSettings['objectTo']['intCropWidth']
    if objectSettings['dolly']:
        focalScaling = t_w / f_w
        currentFocal = objectCommon['dblFocal'] * (1 - dblStep) + dblStep * objectCommon['dblFocal'] * focalScaling
    else:
        currentFocal = objectCommon['dblFocal']
    dblShiftU = ((dblFrom * objectSettings['objectFrom']['dblCenterU']) + (dblTo * objectSettings['objectTo']['dblCenterU'])) - (objectCommon['intWidth'] / 2.0)
    dblShiftV = ((dblFrom * objectSettings['objectFrom']['dblCenterV']) + (dblTo * objectSettings['objectTo']['dblCenterV'])) - (objectCommon['intHeight'] / 2.0)
    dblCropWidth = (dblFrom * f_w) + (dblTo * t_w)
    dblDepthFrom = objectCommon['objectDepthrange'][0]
    dblDepthTo = objectCommon['objectDepthrange'][0] * (dblCropWidth / max(f_w, t_w))
    return {'dblShiftU': dblShiftU, 'dblShiftV': dblShiftV, 'dblDepthFrom': dblDepthFrom,
            'dblDepthTo': dblDepthTo}, currentFocal


def kenburns_poses(objectSettings, objectCommon):
    """[(shift_xyz as fp32 triple, focal double)] for every step of objectSettings['dblSteps']."""
    poses = []
    for dblStep in objectSettings['dblSteps']:
        st, focal = _pose_settings(objectSettings, objectCommon, dblStep)
        sx, sy, sz = _shift_scalars(st, objectCommon, focal)
        sh = np.array([sx, sy, sz], dtype=np.float64).astype(np.float32)   # torch.FloatTensor([...]) rounding
        poses.append((sh, float(focal)))
    return poses


def process_autozoom(objectSettings, objectCommon):
    """utils/common.py:114-170: pick the end window of an automatic zoom -- among a 16 x 16 grid of centre shifts within
    +-objectSettings['dblShift'], the one whose view of the RAW cloud covers the most pixels (existing > 0).  The reference's
    version cannot run (its process_shift call, :146-152, lacks the objectCommon argument); this is that function with the call
    completed.  The (up to) 256 candidate views are rendered through the multi-pose z-buffer kernels in groups of KB_MAX_POSES
    (kb_coverage) instead of 256 x (clone + 3 kernels + 2 reductions + .item())."""
    n = 16
    shift_u = np.linspace(-objectSettings['dblShift'], objectSettings['dblShift'], n)[None, :].repeat(n, 0)
    shift_v = np.linspace(-objectSettings['dblShift'], objectSettings['dblShift'], n)[:, None].repeat(n, 1)
    frm = objectSettings['objectFrom']
    crop_w = frm['intCropWidth'] / objectSettings['dblZoom']
    crop_h = frm['intCropHeight'] / objectSettings['dblZoom']
    depth_from = objectCommon['objectDepthrange'][0]
    depth_to = objectCommon['objectDepthrange'][0] * (crop_w / frm['intCropWidth'])
    cands = []
    for iu in range(n):
        for iv in range(n):
            su, sv = shift_u[iu, iv].item(), shift_v[iu, iv].item()
            if frm['dblCenterU'] + su < crop_w / 2.0 or frm['dblCenterU'] + su > objectCommon['intWidth'] - (crop_w / 2.0):
                continue
            if frm['dblCenterV'] + sv < crop_h / 2.0 or frm['dblCenterV'] + sv > objectCommon['intHeight'] - (crop_h / 2.0):
                continue
            sx, sy, sz = _shift_scalars({'dblShiftU': su, 'dblShiftV': sv, 'dblDepthFrom': depth_from, 'dblDepthTo': depth_to},
                                        objectCommon, objectCommon['dblFocal'])
            cands.append((su, sv, np.array([sx, sy, sz], dtype=np.float64).astype(np.float32)))
    if not cands:
        raise ValueError("process_autozoom: no candidate window fits the image")
    cover = coverage_counts(objectCommon['tensorRawPoints'], [c[2] for c in cands], objectCommon['intWidth'], objectCommon['intHeight'],
                            objectCommon['dblFocal'], objectCommon['dblBaseline'])
    best, best_u, best_v = 0.0, None, None
    for (su, sv, _), cnt in zip(cands, cover):            # the reference's scan order and strict '<' (:160-164)
        if best < cnt:
            best, best_u, best_v = cnt, su, sv
    if best_u is None:
        raise ValueError("process_autozoom: every candidate view is empty")
    return {'dblCenterU': frm['dblCenterU'] + best_u, 'dblCenterV': frm['dblCenterV'] + best_v,
            'intCropWidth': int(round(frm['intCropWidth'] / objectSettings['dblZoom'])),
            'intCropHeight': int(round(frm['intCropHeight'] / objectSettings['dblZoom']))}


def coverage_counts(tensorPoints, shifts, intWidth, intHeight, dblFocal, dblBaseline):
    """For each camera shift (fp32 triple): how many pixels of render_pointcloud(process_shift(points), ...)'s `existing` map are
    > 0 (utils/common.py:154-160) -> list of ints.  existing > 0 at a pixel iff some point's gated bilinear weight reaches it, which
    does not depend on the data channels: the z-buffer passes and a weight-only accumulation decide it."""
    _need_cuda(tensorPoints)
    L = nat.lib()
    xyz = tensorPoints.reshape(3, -1).contiguous()
    N = xyz.shape[1]
    H, W = int(intHeight), int(intWidth)
    out = []
    K = nat.KB_MAX_POSES
    ws = torch.empty(L.kb_coverage_workspace_bytes(H, W, K) + 256, device=xyz.device, dtype=torch.uint8)
    ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    counts = torch.empty(K, device=xyz.device, dtype=torch.int32)
    for a in range(0, len(shifts), K):
        grp = shifts[a:a + K]
        arr = (nat.KBPose * len(grp))()
        for i, sh in enumerate(grp):
            arr[i].shift[0], arr[i].shift[1], arr[i].shift[2] = float(sh[0]), float(sh[1]), float(sh[2])
            arr[i].focal = float(dblFocal)
        nat.check(L.kb_coverage(_ptr(xyz), N, arr, len(grp), H, W, float(dblBaseline), ctypes.c_void_p(ws_ptr), _ptr(counts), _stream()),
                  "kb_coverage")
        out += counts[:len(grp)].tolist()
    return out


def process_inpaint(tensorShift, objectCommon, moduleInpaint, dblFocal):
    """utils/common.py:47-81.  A list [colour network, disparity network] (`kbe.py --inpaint-depth`) takes colour and the
    existing-mask from the first and the disparity from the second, which is what the reference's list branch (:50-62) sets out
    to do before it trips over an undefined name and a wrong key."""
    if isinstance(moduleInpaint, (list, tuple)):
        if len(moduleInpaint) != 2:
            raise ValueError("process_inpaint: a list of inpainting networks must be [colour, disparity]")
        col = moduleInpaint[0].pointcloud_inpainting(objectCommon['tensorRawImage'], objectCommon['tensorRawDisparity'],
                                                     tensorShift, objectCommon, dblFocal)
        dep = moduleInpaint[1].pointcloud_inpainting(objectCommon['tensorRawImage'], objectCommon['tensorRawDisparity'],
                                                     tensorShift, objectCommon, dblFocal)
        obj = dict(col)
        obj['tensorDisparity'] = dep['tensorDisparity']
    else:
        obj = moduleInpaint.pointcloud_inpainting(objectCommon['tensorRawImage'], objectCommon['tensorRawDisparity'],
                                                  tensorShift, objectCommon, dblFocal)
    disp = obj['tensorDisparity']
    depth = (dblFocal * objectCommon['dblBaseline']) / (disp + 0.0000001)
    valid = (spatial_filter(disp / disp.max(), 'laplacian').abs() < 0.03).float()
    points = depth_to_points(depth * valid, dblFocal).view(1, 3, -1) - tensorShift
    # points that were missing in the shifted view get appended to the cloud (:75-80).  The reference tests
    # 'tensorExisting' of the network output; modules may expose the input mask separately (PartialInpaint).
    existing = obj.get('tensorExistingInput', obj['tensorExisting'])
    mask = (existing[:, 0:1] == 0.0).view(-1)
    idx = mask.nonzero(as_tuple=True)[0]
    objectCommon['tensorInpaImage'] = torch.cat([objectCommon['tensorInpaImage'], obj['tensorImage'].view(1, 3, -1)[:, :, idx]], 2)
    objectCommon['tensorInpaDisparity'] = torch.cat([objectCommon['tensorInpaDisparity'], disp.view(1, 1, -1)[:, :, idx]], 2)
    objectCommon['tensorInpaDepth'] = torch.cat([objectCommon['tensorInpaDepth'], depth.view(1, 1, -1)[:, :, idx]], 2)
    objectCommon['tensorInpaPoints'] = torch.cat([objectCommon['tensorInpaPoints'], points[:, :, idx]], 2)


class FrameRenderer:
    """Fused per-frame loop (utils/common.py:238-257) over one point cloud: K poses per call of
    kb_render_frames, frames land in pinned host memory through an async copy."""

    def __init__(self, tensorPoints, tensorImage, tensorDepth, intWidth, intHeight, dblBaseline, crop_w, crop_h,
                 batch=FRAME_BATCH):
        _need_cuda(tensorPoints, tensorImage, tensorDepth)
        self.device = tensorPoints.device
        self.set_cloud(tensorPoints, tensorImage, tensorDepth)
        self.H, self.W = int(intHeight), int(intWidth)
        self.batch = max(1, min(int(batch), nat.KB_MAX_POSES))
        self.params = nat.KBFrameParams(self.H, self.W, int(crop_w), int(crop_h), float(dblBaseline))
        L = nat.lib()
        nbytes = L.kb_frames_workspace_bytes(ctypes.byref(self.params), self.batch)
        self.ws = torch.empty(nbytes + 256, device=self.device, dtype=torch.uint8)
        off = (-self.ws.data_ptr()) % 256
        self.ws_ptr = self.ws.data_ptr() + off
        # host destinations: two device staging buffers + a copy stream, so the D2H of batch i overlaps the
        # kernels of batch i+1 (2.36 MB per 1024x768 frame: PCIe is the end-to-end bound)
        self.dev_frames = [torch.empty(self.batch, self.H, self.W, 3, device=self.device, dtype=torch.uint8) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.ev_rendered = [torch.cuda.Event() for _ in range(2)]
        self.ev_copied = [torch.cuda.Event() for _ in range(2)]

    def set_cloud(self, tensorPoints, tensorImage, tensorDepth):
        """Point the renderer at a (new) cloud; frame geometry and workspace are kept.  When image and depth are
        adjacent rows of one packed [7,N] buffer (shard.unpack_cloud) nothing is copied."""
        _need_cuda(tensorPoints, tensorImage, tensorDepth)
        self.N = tensorPoints.shape[-1]
        self.xyz = tensorPoints.reshape(3, self.N).contiguous()
        img, dep = tensorImage.reshape(3, self.N), tensorDepth.reshape(1, self.N)
        if (img.is_contiguous() and dep.is_contiguous() and dep.data_ptr() == img.data_ptr() + 12 * self.N
                and img.untyped_storage().data_ptr() == dep.untyped_storage().data_ptr()):
            self.rgbd = torch.as_strided(img, (4, self.N), (self.N, 1))
        else:
            self.rgbd = torch.cat([img, dep], 0).contiguous()

    def render_into(self, poses, out_frames, on_batch=None):
        """poses: list of (shift fp32[3], focal); out_frames: uint8 tensor [len(poses),H,W,3] (device or
        pinned host).  Enqueue only -- the caller synchronises the current stream.
        on_batch(start, count, event): called after every batch has been enqueued, `event` fires when out_frames[start:start+count]
        is complete (after the device-to-host copy for host destinations) -- utils/sink.py consumes frames this way while the
        next batch renders."""
        L = nat.lib()
        n = len(poses)
        done = 0
        main = torch.cuda.current_stream(self.device)
        to_host = not out_frames.is_cuda
        it = 0
        while done < n:
            k = min(self.batch if not to_host else min(self.batch, FRAME_BATCH_TO_HOST), n - done)
            arr = (nat.KBPose * k)()
            for i in range(k):
                sh, focal = poses[done + i]
                arr[i].shift[0], arr[i].shift[1], arr[i].shift[2] = float(sh[0]), float(sh[1]), float(sh[2])
                arr[i].focal = float(focal)
            dst = out_frames[done:done + k]
            slot = it & 1
            if to_host:
                if it >= 2:
                    main.wait_event(self.ev_copied[slot])        # staging buffer free again
                target = self.dev_frames[slot][:k]
            else:
                target = dst
            nat.check(L.kb_render_frames(_ptr(self.xyz), _ptr(self.rgbd), self.N, arr, k, ctypes.byref(self.params),
                                         ctypes.c_void_p(self.ws_ptr), _ptr(target), _stream()), "kb_render_frames")
            if to_host:
                self.ev_rendered[slot].record(main)
                self.copy_stream.wait_event(self.ev_rendered[slot])
                with torch.cuda.stream(self.copy_stream):
                    dst.copy_(target, non_blocking=True)
                    self.ev_copied[slot].record(self.copy_stream)
                    if on_batch is not None:
                        ev = torch.cuda.Event()
                        ev.record(self.copy_stream)
                        on_batch(done, k, ev)
            elif on_batch is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                on_batch(done, k, ev)
            done += k
            it += 1
        if to_host:
            main.wait_stream(self.copy_stream)                   # the caller's sync on the current stream covers the copies
        return out_frames


def prepare_cloud(objectSettings, objectCommon, moduleInpaint):
    """Stage A of process_kenburns, utils/common.py:175-220: reset the working cloud to the raw one and, unless in
    dolly mode, append the points the inpainting network hallucinates at both extremes of the camera path."""
    dev = objectCommon['tensorRawPoints'].device
    objectCommon['tensorInpaImage'] = objectCommon['tensorRawImage'].view(1, 3, -1)
    objectCommon['tensorInpaDisparity'] = objectCommon['tensorRawDisparity'].view(1, 1, -1)
    objectCommon['tensorInpaDepth'] = objectCommon['tensorRawDepth'].view(1, 1, -1)
    objectCommon['tensorInpaPoints'] = objectCommon['tensorRawPoints'].view(1, 3, -1)
    for dblStep in [0.0, 1.0]:
        st, focal = _pose_settings(objectSettings, objectCommon, dblStep)
        sx, sy, sz = _shift_scalars(st, objectCommon, focal)
        tensorShift = torch.FloatTensor([sx, sy, sz]).view(1, 3, 1).to(dev)
        # (the reference also renders the current cloud here, :208-215, and discards the result)
        if not objectSettings['dolly']:
            process_inpaint(1.1 * tensorShift, objectCommon, moduleInpaint, focal)


def crop_size(objectSettings):
    """Patch size of the per-frame getRectSubPix, utils/common.py:256."""
    f, t = objectSettings['objectFrom'], objectSettings['objectTo']
    return max(f['intCropWidth'], t['intCropWidth']), max(f['intCropHeight'], t['intCropHeight'])


_PINNED = {}      # shape -> [pinned uint8 tensors]; cudaHostAlloc of 354 MB costs more than rendering 150 frames
_PINNED_ORDER = []              # shapes, least recently used first
PINNED_POOL_BYTES = 4 << 30     # idle buffers beyond this are released, oldest shape first
_RENDERERS = {}   # (device, H, W, crop, baseline) -> FrameRenderer (workspace, staging buffers, copy stream)


def _storage_users(t):
    """How many tensors / numpy arrays share t's storage.  `tensor.numpy()` hangs a NEW tensor wrapper on the array (its .base),
    so the Python refcount of `t` itself says nothing about frames a caller still holds; the storage's use count does.
    None when this torch build does not expose the counter: such a buffer is then never recycled."""
    try:
        return int(torch._C._storage_Use_Count(t.untyped_storage()._cdata))
    except Exception:
        return None


def pinned_frames(shape):
    """A pinned uint8 host buffer of `shape`, recycled from earlier calls once nobody else holds its storage (frames handed out
    as numpy views keep the storage in use, so they are never overwritten)."""
    import sys
    shape = tuple(int(v) for v in shape)
    pool = _PINNED.setdefault(shape, [])
    if shape in _PINNED_ORDER:
        _PINNED_ORDER.remove(shape)
    _PINNED_ORDER.append(shape)

    def idle(entry):
        # nobody holds the tensor object itself (refs here: the pool's tuple, `t`, getrefcount's argument) and nobody holds
        # another view of its storage (numpy arrays made by .numpy(), slices)
        t = entry[0]
        users = _storage_users(t)
        return sys.getrefcount(t) <= 3 and users is not None and users <= entry[1]

    for entry in pool:
        if idle(entry):
            return entry[0]
    t = torch.empty(*shape, dtype=torch.uint8).pin_memory()
    base = _storage_users(t)
    if base is not None and len(pool) < 4:
        pool.append((t, base))
    # bound the pool: varying image sizes / frame counts must not pile up pinned host memory
    total = sum(e[0].numel() for bufs in _PINNED.values() for e in bufs)
    for old in list(_PINNED_ORDER[:-1]):
        if total <= PINNED_POOL_BYTES:
            break
        bufs = _PINNED.get(old, [])
        for e in [e for e in bufs if idle(e)]:
            bufs.remove(e)
            total -= e[0].numel()
        if not bufs:
            _PINNED.pop(old, None)
            _PINNED_ORDER.remove(old)
    return t


def render_poses(objectSettings, objectCommon, poses, to_host=True, sink=None, out=None):
    """Stage B, utils/common.py:222-260, for the given poses of the path -> uint8 [len(poses),H,W,3]
    (pinned host memory when to_host, else on the cloud's device).  sink: a utils.sink.FrameSink that receives the frames batch
    by batch while later batches still render (host destinations only); out: destination to use instead of a pooled buffer."""
    crop_w, crop_h = crop_size(objectSettings)
    pts = objectCommon['tensorInpaPoints']
    key = (pts.device, int(objectCommon['intHeight']), int(objectCommon['intWidth']), crop_w, crop_h, float(objectCommon['dblBaseline']))
    renderer = _RENDERERS.get(key)
    if renderer is None:
        _RENDERERS.clear()                    # one geometry at a time: the workspace is ~0.45 GB at 1024x768
        renderer = _RENDERERS[key] = FrameRenderer(pts, objectCommon['tensorInpaImage'], objectCommon['tensorInpaDepth'],
                                                   objectCommon['intWidth'], objectCommon['intHeight'],
                                                   objectCommon['dblBaseline'], crop_w, crop_h)
    else:
        renderer.set_cloud(pts, objectCommon['tensorInpaImage'], objectCommon['tensorInpaDepth'])
    if out is None:
        if to_host:
            out = pinned_frames((len(poses), renderer.H, renderer.W, 3))
        else:
            out = torch.empty(len(poses), renderer.H, renderer.W, 3, dtype=torch.uint8, device=renderer.device)
    feeder = None
    if sink is not None:
        if out.is_cuda:
            raise RuntimeError("render_poses: a frame sink consumes host frames (to_host=True)")
        from .sink import BatchFeeder
        feeder = BatchFeeder(sink, out)
    if len(poses):
        renderer.render_into(poses, out, on_batch=feeder)
    torch.cuda.current_stream(renderer.device).synchronize()
    if feeder is not None:
        feeder.finish()
    return out


def process_kenburns(objectSettings, objectCommon, moduleInpaint, sink=None):
    """utils/common.py:172-263 -> list of uint8 [H,W,3] frames (RGB order of the input tensor's channels).
    sink (additive): a utils.sink.FrameSink fed while the frames render."""
    if 'boolInpaint' not in objectSettings or objectSettings['boolInpaint'] == True:  # noqa: E712
        prepare_cloud(objectSettings, objectCommon, moduleInpaint)
    poses = kenburns_poses(objectSettings, objectCommon)
    frames = render_poses(objectSettings, objectCommon, poses, sink=sink).numpy()
    return [frames[i] for i in range(len(poses))]
