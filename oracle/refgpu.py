"""Run the REFERENCE's own CUDA kernels (cubins compiled by build_ref.py from /root/reference/utils/common.py)
on the GPU box.  TEST INFRASTRUCTURE: this is the ground truth that pins both the CPU oracle and the product.

Launch geometry is the reference's: grid = ceil(n/512), block = 512, args = [int n, device pointers...]
(utils/common.py:516-521, :577-582, :679-684, :929-934).
"""
import ctypes
import json
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_cache = {}


def available():
    return os.path.exists(os.path.join(_REF, "manifest.json"))


def manifest():
    with open(os.path.join(_REF, "manifest.json")) as f:
        return json.load(f)


def _drv():
    from cuda.bindings import driver
    return driver


def _chk(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _function(row):
    key = (row["file"], row["entry"])
    if key not in _cache:
        drv = _drv()
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make sure the primary context is current
        with open(os.path.join(_REF, row["file"]), "rb") as f:
            image = f.read()
        mod = _chk(drv.cuModuleLoadData(image))
        fn = _chk(drv.cuModuleGetFunction(mod, row["entry"].encode()))
        _cache[key] = (mod, fn)
    return _cache[key][1]


def find(role, **shape):
    for row in manifest():
        if row["role"] == role and all(row[k] == v for k, v in shape.items()):
            return row
    raise KeyError(f"no reference cubin for {role} {shape}; add the shape to oracle/build_ref.py:SHAPES")


def _launch(row, n, tensors):
    drv = _drv()
    fn = _function(row)
    vals = [ctypes.c_int(n)] + [ctypes.c_void_p(t.data_ptr()) for t in tensors]
    ptrs = (ctypes.c_void_p * len(vals))(*[ctypes.addressof(v) for v in vals])
    stream = torch.cuda.current_stream().cuda_stream
    _chk(drv.cuLaunchKernel(fn, (n + 511) // 512, 1, 1, 512, 1, 1, 0, stream, ctypes.addressof(ptrs), 0))
    torch.cuda.synchronize()


def render_pointcloud(points, data, W, H, focal, baseline, stages=False):
    """The reference's render_pointcloud (utils/common.py:428-686) with its own three kernels.
    Returns (render, existing) or, with stages=True, also zee after updateZee and after updateDegrid."""
    B, C, N = data.shape
    shape = dict(H=H, W=W, N=N, C=C, focal=focal, baseline=baseline, B=B)
    data1 = torch.cat([data, data.new_ones(B, 1, N)], 1).contiguous()          # :429
    zee = points.new_zeros(B, 1, H, W).fill_(1000000.0)                        # :430
    out = points.new_zeros(B, C + 1, H, W)                                     # :431
    points = points.contiguous()
    _launch(find("updateZee", **shape), B * N, [points, data1, zee])
    zee_raw = zee.clone()
    _launch(find("updateDegrid", **shape), zee.numel(), [points, data1, zee])
    zee_deg = zee.clone()
    _launch(find("updateOutput", **shape), B * N, [points, data1, zee, out])
    render = out[:, :-1] / (out[:, -1:] + 0.0000001)                           # :686
    existing = out[:, -1:].clone()
    if stages:
        return render, existing, zee_raw, zee_deg, out
    return render, existing


def fill_disocclusion(inp, depth, N=None, focal=None, baseline=None):
    """The reference's fill_disocclusion kernel (utils/common.py:833-937)."""
    B, C, H, W = inp.shape
    row = None
    for r in manifest():
        if r["role"] == "discfill" and r["H"] == H and r["W"] == W and r["C"] == C and r["B"] == B:
            row = r
            break
    if row is None:
        raise KeyError(f"no reference discfill cubin for {tuple(inp.shape)}")
    out = inp.clone()
    _launch(row, B * H * W, [inp.contiguous(), depth.contiguous(), out])
    return out


def mask_zee(points_shifted, W, H, focal, baseline):
    """The kernel of the reference's generate_mask (utils/common.py:696-827) on already shifted points [B,3,N] -> the raw
    masks [B,N] before the median filter (:829).  Racy by construction: two runs may differ."""
    B, _, N = points_shifted.shape
    row = None
    for r in manifest():
        if r["role"] == "maskZee" and r["H"] == H and r["W"] == W and r["N"] == N and r["B"] == B and abs(r["focal"] - focal) < 1e-9:
            row = r
            break
    if row is None:
        raise KeyError(f"no reference maskZee cubin for B={B} N={N} {W}x{H} f={focal}")
    zee = points_shifted.new_zeros(B, 1, H, W).fill_(1000000.0)      # :692
    masks = points_shifted.new_zeros(B, 1, N)                        # :693
    ids = points_shifted.new_ones(B, H, W) * -1                      # :694 (a FLOAT tensor the kernel reads as int*)
    _launch(row, B * N, [points_shifted.contiguous(), masks, zee, ids])
    return masks.view(B, N)
