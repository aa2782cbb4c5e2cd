"""PartialInpaint -- mirror of models/partial_inpainting.py:99-279: the Inpaint topology built from
PartialConv2d, masks travelling with the activations (up-sampled masks thresholded at 0.5, skip merges by
min).  Parameter names follow the reference: p_relu_1 / conv1 / p_relu_2 / conv2 / moduleShortcut."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..utils.partial_conv import PartialConv2d
from .gridnet import grid_name
from .pointcloud_inpainting import Inpaint as _DenseInpaint


def _pconv(cin, cout, k=3, stride=1, return_mask=True):
    return PartialConv2d(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=k // 2,
                         multi_channel=True, return_mask=return_mask)


class _Chain(nn.Module):
    """[up x2] -> [PReLU] -> pconv -> PReLU -> pconv, the mask threading through both partial convs."""
    pre_relu = True
    upsample_first = False
    stride = 1

    def __init__(self, intChannels):
        super().__init__()
        c0, c1, c2 = intChannels
        if self.upsample_first:
            self.upsample = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False)
        if self.pre_relu:
            self.p_relu_1 = nn.PReLU(num_parameters=c0, init=0.25)
        self.conv1 = _pconv(c0, c1, stride=self.stride)
        self.p_relu_2 = nn.PReLU(num_parameters=c1, init=0.25)
        self.conv2 = _pconv(c1, c2)

    def chain(self, x, mask_in):
        if self.upsample_first:
            x = self.upsample(x)
            mask_in = (self.upsample(mask_in) > 0.5).float()           # partial_inpainting.py:90
        if self.pre_relu:
            x = self.p_relu_1(x)
        x, mask_in = self.conv1(x, mask_in=mask_in)
        x = self.p_relu_2(x)
        return self.conv2(x, mask_in=mask_in)

    def forward(self, tensorInput, mask_in=None):
        return self.chain(tensorInput, mask_in)


class Basic(_Chain):
    def __init__(self, strType, intChannels):
        self.pre_relu = strType == 'relu-conv-relu-conv'
        super().__init__(intChannels)
        self.strType = strType
        c0, _, c2 = intChannels
        self.moduleShortcut = None if c0 == c2 else _pconv(c0, c2, k=1, return_mask=False)

    def forward(self, tensorInput, mask_in=None):
        # the 1x1 shortcut is a partial conv called WITHOUT a mask (partial_inpainting.py:47)
        shortcut = tensorInput if self.moduleShortcut is None else self.moduleShortcut(tensorInput)
        out, mask = self.chain(tensorInput, mask_in)
        return out + shortcut, mask


class Downsample(_Chain):
    stride = 2


class Upsample(_Chain):
    upsample_first = True


class Inpaint(_DenseInpaint):
    def __init__(self):
        nn.Module.__init__(self)
        self.spectral_norm = False
        self.moduleContext = nn.Sequential(
            nn.Conv2d(in_channels=4, out_channels=64, kernel_size=3, stride=1, padding=1, bias=True),
            nn.PReLU(num_parameters=64, init=0.25),
            nn.Conv2d(in_channels=64, out_channels=64, kernel_size=3, stride=1, padding=1, bias=True),
            nn.PReLU(num_parameters=64, init=0.25))
        self.moduleInput = Basic('conv-relu-conv', [3 + 1 + 64, 32, 32])
        F_ = self.FEATURES
        R = len(F_)
        for r, f in enumerate(F_):
            for c in range(3):
                self.add_module(grid_name(r, c, r, c + 1), Basic('relu-conv-relu-conv', [f, f, f]))
        for c in (0, 1):
            for r in range(R - 1):
                self.add_module(grid_name(r, c, r + 1, c), Downsample([F_[r], F_[r + 1], F_[r + 1]]))
        for c in (2, 3):
            for r in range(R - 1, 0, -1):
                self.add_module(grid_name(r, c, r - 1, c), Upsample([F_[r], F_[r - 1], F_[r - 1]]))
        self.moduleImage = Basic('conv-relu-conv', [32, 32, 3])
        self.moduleDisparity = Basic('conv-relu-conv', [32, 32, 1])

    def forward(self, tensorMasks, tensorImage=None, tensorDisparity=None, tensorData=None, tensorContext=None):
        if tensorImage is not None and tensorContext is None:
            tensorImage, tensorDisparity = self.normalize_images_disp(tensorImage, tensorDisparity, not_normed=True)
        if tensorData is None and tensorContext is not None:
            tensorData = torch.cat([tensorImage, tensorDisparity, tensorContext], 1)
        elif tensorData is None:
            tensorContext = self.moduleContext(torch.cat([tensorImage, tensorDisparity], 1))
            tensorData = torch.cat([tensorImage, tensorDisparity, tensorContext], 1)

        m = self._modules
        R = len(self.FEATURES)
        x = [None] * R
        k = [None] * R
        x[0], k[0] = self.moduleInput(tensorData, mask_in=tensorMasks.expand_as(tensorData))
        for r in range(1, R):
            x[r], k[r] = m[grid_name(r - 1, 0, r, 0)](x[r - 1], k[r - 1])
        for r in range(R):                                              # column 1, top -> bottom
            x[r], k[r] = m[grid_name(r, 0, r, 1)](x[r], k[r])
            if r > 0:
                d, dk = m[grid_name(r - 1, 1, r, 1)](x[r - 1], k[r - 1])
                x[r] = x[r] + d
                k[r] = torch.min(k[r], dk)                              # partial_inpainting.py:167
        for c in (2, 3):                                                # columns 2, 3, bottom -> top
            for r in range(R - 1, -1, -1):
                x[r], k[r] = m[grid_name(r, c - 1, r, c)](x[r], k[r])
                if r < R - 1:
                    u, uk = m[grid_name(r + 1, c, r, c)](x[r + 1], k[r + 1])
                    u = u[:, :, :x[r].size(2), :x[r].size(3)]
                    uk = uk[:, :, :x[r].size(2), :x[r].size(3)]
                    x[r] = x[r] + u
                    k[r] = torch.min(k[r], uk)
        img, _ = self.moduleImage(x[0])                                  # heads run without a mask (:212-213)
        disp, _ = self.moduleDisparity(x[0])
        img, disp = self.normalize_images_disp(img, disp, not_normed=False)
        return {
            'tensorExisting': k[0],
            'tensorExistingInput': tensorMasks,   # what process_inpaint needs (SURVEY.md discrepancy table)
            'tensorImage': img.clamp(0.0, 1.0) if self.training == False else img,  # noqa: E712
            'tensorDisparity': F.threshold(input=disp, threshold=0.0, value=0.0),
        }
