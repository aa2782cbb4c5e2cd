#!/usr/bin/env python3
"""Why does PartialInpaint.pointcloud_inpainting at 1024x768 sit at 2.4e-3 rel. L2 from the reference in a fresh process and at
1e-2 after other product forwards ran in the same process?  Runs the comparison fresh, then after a 384x320 partial pipeline,
and prints WHERE the error lives.  Usage (GPU box): python tools/diag_partial.py"""
import contextlib
import io
import os
import sys

os.environ.setdefault("KB200_RANDOM_VGG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import kb_helpers  # noqa: E402
from oracle import refshim  # noqa: E402
from ken_burns_effect_b200.utils import synthetic  # noqa: E402

torch.set_grad_enabled(False)
refshim.fp32_convs()
ref = refshim.load()
from ken_burns_effect_b200.models.partial_inpainting import Inpaint as PartialInpaint  # noqa: E402
from ken_burns_effect_b200.utils.pipeline import Pipeline  # noqa: E402

W, H = 1024, 768
img, disp = synthetic.synthetic_scene(W, H, 1234)
image = torch.from_numpy(img).permute(2, 0, 1).contiguous().float().div(255).view(1, 3, H, W).cuda()
disparity = torch.from_numpy(disp).view(1, 1, H, W).cuda()
oc = {'dblFocal': 512.0, 'dblBaseline': 120, 'intWidth': W, 'intHeight': H}
shift = torch.tensor([14.0, -9.0, -30.0], device='cuda').view(1, 3, 1)
rnet = kb_helpers.deterministic_state(ref.PartialInpaint()).cuda().eval()
with contextlib.redirect_stdout(io.StringIO()):
    theirs = rnet.pointcloud_inpainting(image.clone(), disparity.clone(), shift, oc)


def compare(tag):
    net = PartialInpaint().cuda().eval()
    net.load_state_dict(rnet.state_dict())
    mine = net.pointcloud_inpainting(image.clone(), disparity.clone(), shift, oc)
    for key in ('tensorImage', 'tensorDisparity'):
        a, b = mine[key], theirs[key]
        d = (a - b).abs().amax(1)[0]
        big = d > 0.05 * float(b.abs().max())
        ys, xs = torch.nonzero(big, as_tuple=True)
        box = (int(ys.min()), int(ys.max()), int(xs.min()), int(xs.max())) if ys.numel() else None
        print(f"{tag:10s} {key:16s} rel_l2 {kb_helpers.rel_l2(a.cpu().numpy(), b.cpu().numpy()):.3e}  pixels off by >5%: {int(big.sum()):7d}"
              f"  bbox(y0,y1,x0,x1) {box}  rows>{H - 8}: {int(big[H - 8:].sum())} cols>{W - 8}: {int(big[:, W - 8:].sum())}", flush=True)
    return mine


m1 = compare("fresh")
m2 = compare("fresh2")
print("product run-to-run image rel_l2", kb_helpers.rel_l2(m1['tensorImage'].cpu().numpy(), m2['tensorImage'].cpu().numpy()))
torch.manual_seed(1)
img2, _ = synthetic.synthetic_scene(384, 320, seed=6)
t = torch.from_numpy(img2).permute(2, 0, 1).contiguous().float().div(255).view(1, 3, 320, 384)
Pipeline(model_paths=None, partial_inpainting=True, dolly=False, frames=3)(t, synthetic.default_zoom(384, 320))
m3 = compare("after384")
Pipeline(model_paths=None, partial_inpainting=False, dolly=False, frames=3)(t, synthetic.default_zoom(384, 320))
m4 = compare("after384b")
