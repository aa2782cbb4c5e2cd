#!/usr/bin/env python3
"""GPU micro-benchmark of the tcgen05 convolution path: every conv class of Inpaint at 1024x768 (SURVEY.md 8(a) conv
mix) and the whole networks, CUDA-event timed, next to cuDNN (TF32 allowed, channels_last) on the same shapes.
Prints one JSON line per case.  Usage (GPU box): python tools/bench_conv.py [--nets] [--iters 20]"""
import argparse
import json
import os

os.environ.setdefault("KB200_RANDOM_VGG", "1")   # synthetic weights: there are no checkpoints offline (explicit opt-in)
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from ken_burns_effect_b200.utils import convstack as cs  # noqa: E402

TF32_PEAK = 1590.0 / 2   # TFLOP/s: half of the bf16 fallback peak in B200_PROFILING.md (MEASURED_PEAKS.json absent)
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(peaks):
    TF32_PEAK = json.load(open(peaks))["bf16_tflops"] / 2


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# (Cin, Cout, k, stride, H, W, count in Inpaint.forward)
SHAPES = [
    (32, 32, 3, 1, 768, 1024, 11), (64, 64, 3, 1, 384, 512, 10), (128, 128, 3, 1, 192, 256, 10),
    (256, 256, 3, 1, 96, 128, 8), (64, 64, 3, 1, 768, 1024, 1), (256, 128, 3, 1, 192, 256, 2),
    (128, 64, 3, 1, 384, 512, 2), (64, 32, 3, 1, 768, 1024, 2), (69, 32, 3, 1, 768, 1024, 1),
    (32, 64, 3, 2, 768, 1024, 2), (64, 128, 3, 2, 384, 512, 2), (128, 256, 3, 2, 192, 256, 2),
    (4, 64, 3, 1, 768, 1024, 1), (512, 512, 3, 1, 24, 32, 13),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--nets", action="store_true")
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--tile_w", type=int, default=0)
    ap.add_argument("--n_block", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--no-cudnn", action="store_true")
    ap.add_argument("--f16", action="store_true", help="per-layer timings with fp16 operands (kind::f16): fp16 input and output")
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    dev = "cuda"
    for i, (Cin, Cout, k, stride, H, W, cnt) in enumerate(SHAPES):
        if args.only >= 0 and i != args.only:
            continue
        conv = torch.nn.Conv2d(Cin, Cout, k, stride, k // 2).to(dev)
        x = torch.randn(1, Cin, H, W, device=dev)
        xh = cs.to_nhwc(x)
        slope = torch.full((Cout,), 0.25, device=dev)
        pc = cs.packed(conv)
        Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
        dst = cs.new_act(1, Ho, Wo, Cout, dev, torch.float16 if args.f16 else torch.float32)
        if args.f16:
            x16 = cs.new_act(1, H, W, Cin, dev, torch.float16)
            x16.copy_(xh)
            xh = x16
        ms = timed(lambda: cs.conv2d(xh, pc, [(slope, True, dst)], tile_w=args.tile_w, n_block=args.n_block,
                                     stages=args.stages), args.iters)
        flops = 2.0 * Cout * Cin * k * k * Ho * Wo
        out = {"case": f"{Cin}->{Cout} k{k} s{stride} @{H}x{W}" + (" f16" if args.f16 else ""), "count_in_inpaint": cnt, "ms": ms,
               "tflops": flops / ms / 1e9, "frac_tf32_peak": flops / ms / 1e9 / TF32_PEAK,
               "algorithmic_GBps": 4.0 * (Cin * H * W + Cout * Ho * Wo) / ms / 1e6}
        if not args.no_cudnn:
            xc = x.contiguous(memory_format=torch.channels_last)
            cc = conv.to(memory_format=torch.channels_last)
            ms_c = timed(lambda: F.prelu(cc(xc), slope), args.iters)
            out["cudnn_tf32_ms"] = ms_c
        print(json.dumps(out), flush=True)
    if args.nets:
        import kb_helpers
        from ken_burns_effect_b200.models.pointcloud_inpainting import Inpaint
        from ken_burns_effect_b200.models.disparity_refinement import Refine
        from ken_burns_effect_b200.models.disparity_estimation import Disparity, Semantics
        H, W = 768, 1024
        net = kb_helpers.deterministic_state(Inpaint().eval()).to(dev)
        data = torch.randn(1, 68, H, W, device=dev)
        mask = (torch.rand(1, 1, H, W, device=dev) > 0.2).float()
        net.normalize_images_disp(torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H, W, device=dev), True)
        ms = timed(lambda: net(mask, tensorData=data), 5, 4)
        print(json.dumps({"net": "Inpaint.forward 1024x768", "ms": ms, "tflops": 819.5 / ms, "frac_tf32_peak": 819.5 / ms / TF32_PEAK}))
        from ken_burns_effect_b200.models.partial_inpainting import Inpaint as PartialInpaint
        pnet = kb_helpers.deterministic_state(PartialInpaint().eval()).to(dev)
        pnet.normalize_images_disp(torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H, W, device=dev), True)
        ms = timed(lambda: pnet(mask, tensorData=data), 5, 4)
        # FLOPs of the data convolutions only: the mask "convolutions" of the reference are a box filter here
        print(json.dumps({"net": "PartialInpaint.forward 1024x768 (kbe.py --partial-conv)", "ms": ms, "tflops": 819.5 / ms,
                          "frac_tf32_peak": 819.5 / ms / TF32_PEAK}))
        ref = kb_helpers.deterministic_state(Refine().eval()).to(dev)
        img, dlo = torch.rand(1, 3, H, W, device=dev), torch.rand(1, 1, H // 4, W // 4, device=dev)
        ms = timed(lambda: ref(img, dlo), 5, 4)
        print(json.dumps({"net": "Refine.forward 1024x768", "ms": ms, "tflops": 311.3 / ms, "frac_tf32_peak": 311.3 / ms / TF32_PEAK}))
        sem = kb_helpers.deterministic_state(Semantics().eval()).to(dev)
        dis = kb_helpers.deterministic_state(Disparity().eval()).to(dev)
        small = torch.rand(1, 3, H // 2, W // 2, device=dev)
        s = sem(small)
        ms = timed(lambda: sem(small), 5, 4)
        print(json.dumps({"net": "Semantics.forward 512x384", "ms": ms, "tflops": 138.4 / ms, "frac_tf32_peak": 138.4 / ms / TF32_PEAK}))
        ms = timed(lambda: dis(small, s), 5, 4)
        print(json.dumps({"net": "Disparity.forward 512x384", "ms": ms, "tflops": 87.7 / ms, "frac_tf32_peak": 87.7 / ms / TF32_PEAK}))


if __name__ == "__main__":
    main()
