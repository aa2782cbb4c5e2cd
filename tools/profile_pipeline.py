#!/usr/bin/env python3
"""Where the time of one full KBE goes on the GPU: torch.profiler (CUPTI) kernel table of Pipeline.estimate_depth +
prepare_cloud + render_poses at 1024x768, random-init weights.  Usage (GPU box): python tools/profile_pipeline.py > out.md"""
import os

os.environ.setdefault("KB200_RANDOM_VGG", "1")   # synthetic weights: there are no checkpoints offline (explicit opt-in)
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ken_burns_effect_b200.utils import common as kb   # noqa: E402
from ken_burns_effect_b200.utils import synthetic   # noqa: E402
from ken_burns_effect_b200.utils.pipeline import Pipeline   # noqa: E402

W, H, frames = 1024, 768, 150
torch.manual_seed(1234)
img, _ = synthetic.synthetic_scene(W, H, seed=1234)
t = torch.from_numpy(img).permute(2, 0, 1).float().div(255).view(1, 3, H, W).pin_memory()
pipe = Pipeline(model_paths=None, dolly=False, frames=frames)
zoom = synthetic.default_zoom(W, H)
settings = {'dblSteps': np.linspace(0.0, 1.0, frames).tolist(), 'objectFrom': zoom['objectFrom'], 'objectTo': zoom['objectTo'],
            'boolInpaint': True, 'dolly': False}


def one():
    pipe.estimate_depth(t)
    kb.prepare_cloud(settings, pipe.objectCommon, pipe.moduleInpaint)
    poses = kb.kenburns_poses(settings, pipe.objectCommon)
    return kb.render_poses(settings, pipe.objectCommon, poses)


with torch.no_grad():
    for _ in range(2):
        one()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        one()
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
