"""Refine -- mirror of models/disparity_refinement.py:65-113 (Basic blocks WITHOUT shortcut)."""
import torch
import torch.nn as nn

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
        one = self.moduleImageOne(img)
        two = self.moduleImageTwo(one)
        thr = self.moduleImageThr(two)
        up = self.moduleDisparityOne(disp)
        up = self.moduleDisparityTwo(torch.cat([thr, up], 1))
        up = self.moduleDisparityThr(torch.cat([two, up], 1))
        up = self.moduleDisparityFou(torch.cat([one, up], 1))
        out = self.moduleRefine(up)
        return out * (std[1] + 0.0000001) + mean[1]
