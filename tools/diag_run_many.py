#!/usr/bin/env python3
"""Host-side timeline of Pipeline.run_many (throughput mode): when does the CNN stage of image i+1 run relative to the frame loop
of image i?  Usage (GPU box): python tools/diag_run_many.py"""
import os
import sys
import time

os.environ.setdefault("KB200_RANDOM_VGG", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ken_burns_effect_b200.utils import common as kb  # noqa: E402
from ken_burns_effect_b200.utils import synthetic  # noqa: E402
from ken_burns_effect_b200.utils.pipeline import Pipeline  # noqa: E402

W, H, frames = 1024, 768, 150
imgs = []
for i in range(8):
    img, _ = synthetic.synthetic_scene(W, H, seed=1234 + i)
    imgs.append(torch.from_numpy(img).permute(2, 0, 1).contiguous().float().div(255).view(1, 3, H, W).pin_memory())
pipe = Pipeline(model_paths=None, dolly=False, frames=frames)
zoom = synthetic.default_zoom(W, H)
for t in imgs[:4]:
    pipe(t, zoom)
pipe.run_many(imgs[:2], zoom, keep=False)
torch.cuda.synchronize()
trace = []
t0 = time.perf_counter()
pipe.run_many(imgs, zoom, keep=False, trace=trace)
torch.cuda.synchronize()
total = time.perf_counter() - t0
print(f"run_many: {1e3 * total / len(imgs):.2f} ms per image")
for i, stage, a, b in sorted(trace, key=lambda r: r[2]):
    print(f"image {i} {stage:6s} {1e3 * (a - t0):8.2f} -> {1e3 * (b - t0):8.2f}  ({1e3 * (b - a):6.2f} ms)")
