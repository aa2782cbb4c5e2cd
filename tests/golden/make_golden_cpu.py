#!/usr/bin/env python3
"""Generate tests/golden/ref_torch_cpu.npz by running the REFERENCE's own Python code on CPU.

The reference (/root/reference) cannot travel to the GPU box and has no tests of its own, so its torch-level
functions and nn.Modules are executed here, on seeded inputs with name-seeded weights
(kb_helpers.deterministic_state), and their outputs are committed as small fixtures.  Shims: a stub `cupy`
(the CUDA kernels are not exercised here), stub kornia/matplotlib/imageio/moviepy, torchvision's vgg19_bn
without the pretrained download, and Tensor.cuda() as identity (utils/common.py:102 calls it).

Run from the repo root:  python tests/golden/make_golden_cpu.py
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kb_helpers  # noqa: E402

warnings.filterwarnings("ignore")
for name in ("cupy", "cupy.util", "cupy.cuda", "kornia", "matplotlib", "matplotlib.pyplot", "imageio", "moviepy",
             "moviepy.editor"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["cupy"].util = sys.modules["cupy.util"]
sys.modules["cupy"].cuda = sys.modules["cupy.cuda"]
sys.modules["cupy.util"].memoize = lambda for_each_device=False: (lambda f: f)


class _S:
    cuda_stream = 0


torch.cuda.current_stream = lambda *a, **k: _S()
torch.Tensor.cuda = lambda self, *a, **k: self
import torchvision  # noqa: E402

_vgg = torchvision.models.vgg19_bn
torchvision.models.vgg19_bn = lambda pretrained=False, **kw: _vgg(weights=None)
sys.path.insert(0, REF)
import utils.common as rc  # noqa: E402
from models.disparity_estimation import Disparity, Semantics  # noqa: E402
from models.disparity_refinement import Refine  # noqa: E402
from models.disparity_refinement_pretrained import Refine as RefineP  # noqa: E402
from models.partial_inpainting import Inpaint as PartialInpaint  # noqa: E402
from models.pointcloud_inpainting import Inpaint  # noqa: E402
from utils.partial_conv import PartialConv2d  # noqa: E402

torch.set_grad_enabled(False)
out = {}
g = torch.Generator().manual_seed(2024)

# ---- utils/common.py torch-level helpers ---------------------------------------------------------------
depth = torch.rand(2, 1, 12, 16, generator=g) * 100 + 1
out["d2p_depth"], out["d2p_points"] = depth.numpy(), rc.depth_to_points(depth, 8.0).numpy()
x = torch.rand(1, 1, 20, 24, generator=g)
out["sf_in"], out["sf_laplacian"] = x.numpy(), rc.spatial_filter(x, "laplacian").numpy()
out["sf_median3"], out["sf_median5"] = rc.spatial_filter(x, "median-3").numpy(), rc.spatial_filter(x, "median-5").numpy()
mb = (torch.rand(1, 1, 20, 24, generator=g) > 0.4).float()
out["sfb_in"], out["sfb_median5"] = mb.numpy(), rc.spatial_filter(mb, "median-5").numpy()

# ---- camera path scalars: process_shift (:83-112) driven like process_kenburns (:222-244) ------------------
W, H = 1024, 768
common = {"dblFocal": 512.0, "dblBaseline": 120, "intWidth": W, "intHeight": H,
          "objectDepthrange": (523.25, 3056.1, (264, 189), (0, 0))}
for dolly in (False, True):
    import math
    if not dolly:
        frm = dict(dblCenterU=W / 2.15, dblCenterV=H / 2.15, intCropWidth=int(math.floor(0.90 * W)), intCropHeight=int(math.floor(0.90 * H)))
        to = dict(dblCenterU=W / 1.85, dblCenterV=H / 1.85, intCropWidth=int(math.floor(0.85 * W)), intCropHeight=int(math.floor(0.85 * H)))
    else:
        frm = dict(dblCenterU=W / 2, dblCenterV=H / 2, intCropWidth=int(math.floor(0.8 * W)), intCropHeight=int(math.floor(0.8 * H)))
        to = dict(dblCenterU=W / 2, dblCenterV=H / 2, intCropWidth=int(math.floor(0.3 * W)), intCropHeight=int(math.floor(0.3 * H)))
    st = {"objectFrom": frm, "objectTo": to, "dolly": dolly}
    pts = torch.rand(1, 3, 50, generator=g) * 1000
    pts[0, 2, :5] = 0.0
    shifts, focals, moved = [], [], []
    for dblStep in np.linspace(0.0, 1.0, 7).tolist():
        # the per-step scalars of utils/common.py:223-236, evaluated by the reference's own process_shift
        dblFrom = 1.0 - dblStep
        dblTo = 1.0 - dblFrom
        if dolly:
            focalScaling = to["intCropWidth"] / frm["intCropWidth"]
            currentFocal = common["dblFocal"] * (1 - dblStep) + dblStep * common["dblFocal"] * focalScaling
        else:
            currentFocal = common["dblFocal"]
        dblShiftU = ((dblFrom * frm["dblCenterU"]) + (dblTo * to["dblCenterU"])) - (W / 2.0)
        dblShiftV = ((dblFrom * frm["dblCenterV"]) + (dblTo * to["dblCenterV"])) - (H / 2.0)
        dblCropWidth = (dblFrom * frm["intCropWidth"]) + (dblTo * to["intCropWidth"])
        dblDepthFrom = common["objectDepthrange"][0]
        dblDepthTo = common["objectDepthrange"][0] * (dblCropWidth / max(frm["intCropWidth"], to["intCropWidth"]))
        p, s = rc.process_shift({"tensorPoints": pts, "dblShiftU": dblShiftU, "dblShiftV": dblShiftV,
                                 "dblDepthFrom": dblDepthFrom, "dblDepthTo": dblDepthTo}, common, currentFocal)
        shifts.append(s.view(3).numpy())
        focals.append(currentFocal)
        moved.append(p.numpy())
    tag = "dolly" if dolly else "kbe"
    out[f"shift_{tag}_points"] = pts.numpy()
    out[f"shift_{tag}_shifts"] = np.stack(shifts)
    out[f"shift_{tag}_focals"] = np.array(focals, np.float64)
    out[f"shift_{tag}_moved"] = np.stack(moved)

# ---- networks ------------------------------------------------------------------------------------------------
img = torch.rand(1, 3, 72, 104, generator=g)                      # odd feature sizes on the way down: exercises the crop
sem = kb_helpers.deterministic_state(Semantics().eval())
dis = kb_helpers.deterministic_state(Disparity().eval())
s = sem(img.clone())
out["net_img"], out["net_semantics"], out["net_disparity"] = img.numpy(), s.numpy(), dis(img, s).numpy()

img2 = torch.rand(1, 3, 48, 64, generator=g)
disp_lo = torch.rand(1, 1, 12, 16, generator=g) * 30
out["ref_img"], out["ref_disp_lo"] = img2.numpy(), disp_lo.numpy()
out["ref_refine"] = kb_helpers.deterministic_state(Refine().eval())(img2, disp_lo).numpy()
out["ref_refine_pretrained"] = kb_helpers.deterministic_state(RefineP().eval())(img2, disp_lo).numpy()

disp = torch.rand(1, 1, 48, 64, generator=g) * 100 + 5
mask = (torch.rand(1, 1, 48, 64, generator=g) > 0.3).float()
out["inp_disp"], out["inp_mask"] = disp.numpy(), mask.numpy()
for tag, cls in (("inpaint", Inpaint), ("partial", PartialInpaint)):
    net = kb_helpers.deterministic_state(cls().eval())
    if tag == "partial":
        import io, contextlib
        with contextlib.redirect_stdout(io.StringIO()):           # the reference prints 'nomask' from its heads
            o = net(mask, tensorImage=img2 * mask, tensorDisparity=disp * mask)
    else:
        o = net(mask, tensorImage=img2 * mask, tensorDisparity=disp * mask)
    for k, v in o.items():
        out[f"{tag}_{k}"] = v.numpy()

pc = kb_helpers.deterministic_state(PartialConv2d(6, 5, kernel_size=3, stride=2, padding=1, multi_channel=True, return_mask=True))
xin = torch.randn(2, 6, 15, 17, generator=g)
min_ = (torch.rand(2, 6, 15, 17, generator=g) > 0.5).float()
y, m = pc(xin, mask_in=min_)
out["pconv_x"], out["pconv_mask"], out["pconv_y"], out["pconv_mask_out"] = xin.numpy(), min_.numpy(), y.numpy(), m.numpy()
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    pc2 = kb_helpers.deterministic_state(PartialConv2d(6, 5, kernel_size=3, stride=1, padding=1, multi_channel=True, return_mask=False))
    out["pconv_nomask_y"] = pc2(xin).numpy()

np.savez_compressed(os.path.join(HERE, "ref_torch_cpu.npz"), **out)
print("wrote", os.path.join(HERE, "ref_torch_cpu.npz"), {k: v.shape for k, v in out.items()})
