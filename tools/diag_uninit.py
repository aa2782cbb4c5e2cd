#!/usr/bin/env python3
"""Do the network forwards read memory nobody wrote?  Every torch.empty / empty_like the product makes is filled with NaN (then with
a large finite value) and the outputs are compared with a normal run: a read of uninitialised memory shows up as NaN / a changed
result.  Usage (GPU box): python tools/diag_uninit.py"""
import os
import sys

os.environ.setdefault("KB200_RANDOM_VGG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import kb_helpers  # noqa: E402
from ken_burns_effect_b200.utils import synthetic  # noqa: E402

_empty, _empty_like = torch.empty, torch.empty_like
FILL = [None]


def empty(*a, **k):
    t = _empty(*a, **k)
    if FILL[0] is not None and t.is_cuda and t.is_floating_point():
        t.fill_(FILL[0])
    return t


def empty_like(*a, **k):
    t = _empty_like(*a, **k)
    if FILL[0] is not None and t.is_cuda and t.is_floating_point():
        t.fill_(FILL[0])
    return t


torch.empty, torch.empty_like = empty, empty_like


def run(name, fn):
    outs = {}
    for tag, fill in (("plain", None), ("ctrl", None), ("nan", float("nan")), ("big", 1.0e30)):
        FILL[0] = fill
        with torch.no_grad():
            o = fn()
        torch.cuda.synchronize()
        outs[tag] = [t.float().clone() for t in o]
    FILL[0] = None
    for tag in ("ctrl", "nan", "big"):
        for i, (a, b) in enumerate(zip(outs["plain"], outs[tag])):
            nans = int(torch.isnan(b).sum())
            diff = float((a - torch.nan_to_num(b)).abs().max())
            print(f"{name:28s} output {i} fill={tag:4s} NaNs {nans:9d}  max|diff| vs plain {diff:.3e}", flush=True)


def main():
    from ken_burns_effect_b200.models.partial_inpainting import Inpaint as PartialInpaint
    from ken_burns_effect_b200.models.pointcloud_inpainting import Inpaint
    from ken_burns_effect_b200.models.disparity_refinement import Refine
    from ken_burns_effect_b200.models.disparity_estimation import Disparity, Semantics
    from ken_burns_effect_b200.utils import convstack as cs
    cs.GRAPHS_ENABLED = False
    W, H = 1024, 768
    img, disp = synthetic.synthetic_scene(W, H, 1234)
    image = torch.from_numpy(img).permute(2, 0, 1).contiguous().float().div(255).view(1, 3, H, W).cuda()
    disparity = torch.from_numpy(disp).view(1, 1, H, W).cuda()
    oc = {'dblFocal': 512.0, 'dblBaseline': 120, 'intWidth': W, 'intHeight': H}
    shift = torch.tensor([14.0, -9.0, -30.0], device='cuda').view(1, 3, 1)
    for cls, name in ((PartialInpaint, "PartialInpaint.pc_inpainting"), (Inpaint, "Inpaint.pc_inpainting")):
        net = kb_helpers.deterministic_state(cls()).cuda().eval()

        def fn(net=net):
            o = net.pointcloud_inpainting(image.clone(), disparity.clone(), shift, oc)
            return [o['tensorImage'], o['tensorDisparity']]
        run(name, fn)
    ref = kb_helpers.deterministic_state(Refine()).cuda().eval()
    small = torch.rand(1, 1, H // 4, W // 4, device='cuda')
    run("Refine", lambda: [ref(image, small)])
    sem = kb_helpers.deterministic_state(Semantics()).cuda().eval()
    dis = kb_helpers.deterministic_state(Disparity()).cuda().eval()
    half = torch.nn.functional.interpolate(image, scale_factor=0.5, mode='bilinear')
    run("Semantics+Disparity", lambda: [dis(half, sem(half))])


if __name__ == "__main__":
    main()
