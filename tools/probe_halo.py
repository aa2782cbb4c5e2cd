#!/usr/bin/env python3
"""GPU probe: which shared-memory descriptor conventions make the persistent halo convolution exact?
Runs integer-valued convolutions (exactly representable in TF32) under the four combinations of
KB_CONV_PITCH16 x KB_CONV_DESC_MODE and prints the max abs error against a float64 CPU reference."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ken_burns_effect_b200.utils import convstack as cs  # noqa: E402

torch.set_grad_enabled(False)
CASES = [(32, 32, 3, 29, 41), (64, 48, 3, 29, 41), (69, 32, 1, 33, 47), (128, 128, 3, 40, 24), (256, 256, 3, 24, 32),
         (512, 512, 3, 12, 16), (32, 3, 3, 40, 40)]


def run(algo):
    out = {}
    for (Cin, Cout, k, H, W) in CASES:
        g = torch.Generator().manual_seed(Cin + Cout + k)
        conv = torch.nn.Conv2d(Cin, Cout, k, 1, k // 2).cuda()
        conv.weight.copy_(torch.randint(-3, 4, conv.weight.shape, generator=g).float())
        conv.bias.copy_(torch.randint(-3, 4, (Cout,), generator=g).float())
        x = torch.randint(-4, 5, (2, Cin, H, W), generator=g).float().cuda()
        want = F.conv2d(x.double().cpu(), conv.weight.double().cpu(), conv.bias.double().cpu(), 1, k // 2)
        got, = cs.conv2d(cs.to_nhwc(x), cs.packed(conv), [(None, False, None)], algo=algo)
        torch.cuda.synchronize()
        out[f"{Cin}->{Cout}k{k}"] = float((got.permute(0, 3, 1, 2).double().cpu() - want).abs().max())
    return out


print(json.dumps({"algo": 1, "err": run(1)}), flush=True)
for pitch16 in (0, 1):
    for mode in (1, 0):
        os.environ["KB_CONV_PITCH16"] = str(pitch16)
        os.environ["KB_CONV_DESC_MODE"] = str(mode)
        print(json.dumps({"algo": 2, "pitch16": pitch16, "desc_mode": mode, "err": run(2)}), flush=True)
