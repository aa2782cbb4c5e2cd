"""Refine -- mirror of models/disparity_refinement.py:65-113 (Basic blocks WITHOUT shortcut)."""
import torch
import torch.nn as nn

from ..utils import convstack as cs
from .gridnet import Basic, Downsample, Upsample, sample_norm


class Refine(nn.Module):
    SHORTCUT = 'none'

    def __init__(self):
        super().__init__()
        self.spectral_norm = False
        s = self.SHORTCUT
        self.moduleImageOne = Basic('conv-relu-conv', [3, 24, 24], s)
        self.moduleImageTwo = Downsample([24, 48, 48])
        self.moduleImageThr = Downsample([48, 96, 96])
        self.moduleDisparityOne = Basic('conv-relu-conv', [1, 96, 96], s)
        self.moduleDisparityTwo = Upsample([192, 96, 96])
        self.moduleDisparityThr = Upsample([144, 48, 48])
        self.moduleDisparityFou = Basic('conv-relu-conv', [72, 24, 24], s)
        self.moduleRefine = Basic('conv-relu-conv', [24, 24, 1], s)

    def forward(self, tensorImage, tensorDisparity):
        mean, std = sample_norm(tensorImage, tensorDisparity)
        img = (tensorImage - mean[0]) / (std[0] + 0.0000001)
        disp = (tensorDisparity - mean[1]) / (std[1] + 0.0000001)
        if img.is_cuda:
            return cs.graphed(self, 'forward', self._forward_b200, img.contiguous(), disp.contiguous()) * (std[1] + 0.0000001) + mean[1]
        one = self.moduleImageOne(img)
        two = self.moduleImageTwo(one)
        thr = self.moduleImageThr(two)
        up = self.moduleDisparityOne(disp)
        up = self.moduleDisparityTwo(torch.cat([thr, up], 1))
        up = self.moduleDisparityThr(torch.cat([two, up], 1))
        up = self.moduleDisparityFou(torch.cat([one, up], 1))
        out = self.moduleRefine(up)
        return out * (std[1] + 0.0000001) + mean[1]

    def _forward_b200(self, img, disp):
        """The same network on libkb200 convolutions.  torch.cat of the reference (:100-104) = channel slices of three
        NHWC concat buffers that the producing convolutions write directly."""
        with cs.f16_operands():       # every `round` output / up-sampled tensor of this stack is read by convolutions only
            return self._forward_b200_body(img, disp)

    def _forward_b200_body(self, img, disp):
        N, _, H, W = img.shape
        dev = img.device
        h2, w2, h4, w4 = (H + 1) // 2, (W + 1) // 2, ((H + 1) // 2 + 1) // 2, ((W + 1) // 2 + 1) // 2
        cat1 = torch.empty(N, h4, w4, 192, device=dev)      # [thr(96) | DisparityOne(96)]
        cat2 = torch.empty(N, h2, w2, 144, device=dev)      # [two(48) | DisparityTwo(96)]
        cat3 = torch.empty(N, H, W, 72, device=dev)         # [one(24) | DisparityThr(48)]
        x = cs.to_nhwc(img)
        d = cs.to_nhwc(disp)
        raw = (None, False, None)
        _, one_act = cs.run_block(self.moduleImageOne, x, [(None, False, cat3[..., :24]),
                                                           (cs.pre_slope(self.moduleImageTwo), True, None)], x_raw=x)
        _, two_act = cs.run_block(self.moduleImageTwo, one_act, [(None, False, cat2[..., :48]),
                                                                 (cs.pre_slope(self.moduleImageThr), True, None)])
        cs.run_block(self.moduleImageThr, two_act, [(None, False, cat1[..., :96])])
        cs.run_block(self.moduleDisparityOne, d, [(None, False, cat1[..., 96:192])], x_raw=d)
        cs.run_block(self.moduleDisparityTwo, cat1, [(None, False, cat2[..., 48:144])], crop=(h2, w2))
        cs.run_block(self.moduleDisparityThr, cat2, [(None, False, cat3[..., 24:72])], crop=(H, W))
        fou, = cs.run_block(self.moduleDisparityFou, cat3, [raw], x_raw=cat3)
        out, = cs.run_block(self.moduleRefine, fou, [raw], x_raw=fou)
        return cs.to_nchw(out)
