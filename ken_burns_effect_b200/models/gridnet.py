"""Building blocks shared by the reference's four conv networks, with the reference's parameter names.

Reference: models/disparity_estimation.py:6-80, models/disparity_refinement.py:6-63,
models/disparity_refinement_pretrained.py:6-78, models/pointcloud_inpainting.py:7-81 (the same three
blocks are pasted into each file there).  state_dict contract kept: `moduleMain.<i>.weight|bias` with
PReLU at the even slots of 'relu-conv-relu-conv', `moduleShortcut.weight|bias` for the 1x1 shortcut.

Every block is a chain of  [bilinear x2] -> [PReLU] -> conv3x3(s1|s2) -> PReLU -> conv3x3 (+ residual), which is
also the fusion unit of the sm_100a conv path; `conv_chain()` exposes that structure so an
execution backend can run a block without re-deriving it from module internals.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _prelu(ch):
    return nn.PReLU(num_parameters=ch, init=0.25)


def _conv3(cin, cout, stride=1):
    return nn.Conv2d(in_channels=cin, out_channels=cout, kernel_size=3, stride=stride, padding=1)


class Basic(nn.Module):
    """'relu-conv-relu-conv' or 'conv-relu-conv', optionally with the residual / 1x1-shortcut sum.

    shortcut: 'auto'  -> identity when Cin == Cout else a 1x1 conv   (disparity_estimation.py:25-44)
              'none'  -> plain chain, no sum                          (disparity_refinement.py:24-27)
    """

    def __init__(self, strType, intChannels, shortcut='auto'):
        super().__init__()
        c0, c1, c2 = intChannels
        layers = []
        if strType == 'relu-conv-relu-conv':
            layers.append(_prelu(c0))
        elif strType != 'conv-relu-conv':
            raise ValueError(strType)
        layers += [_conv3(c0, c1), _prelu(c1), _conv3(c1, c2)]
        self.moduleMain = nn.Sequential(*layers)
        self.residual = shortcut != 'none'
        if self.residual:
            # attribute must exist (as None) so that state_dict keys match the reference exactly
            self.moduleShortcut = None if c0 == c2 else nn.Conv2d(c0, c2, kernel_size=1, stride=1, padding=0)

    def forward(self, tensorInput):
        out = self.moduleMain(tensorInput)
        if not self.residual:
            return out
        if self.moduleShortcut is None:
            return out + tensorInput
        return out + self.moduleShortcut(tensorInput)


class Downsample(nn.Module):
    """PReLU -> conv3x3 stride 2 -> PReLU -> conv3x3 (disparity_estimation.py:47-62)."""

    def __init__(self, intChannels):
        super().__init__()
        c0, c1, c2 = intChannels
        self.moduleMain = nn.Sequential(_prelu(c0), _conv3(c0, c1, stride=2), _prelu(c1), _conv3(c1, c2))

    def forward(self, tensorInput):
        return self.moduleMain(tensorInput)


class Upsample(nn.Module):
    """bilinear x2 (align_corners=False) -> PReLU -> conv3x3 -> PReLU -> conv3x3 (disparity_estimation.py:64-80)."""

    def __init__(self, intChannels):
        super().__init__()
        c0, c1, c2 = intChannels
        self.moduleMain = nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False),
                                        _prelu(c0), _conv3(c0, c1), _prelu(c1), _conv3(c1, c2))

    def forward(self, tensorInput):
        return self.moduleMain(tensorInput)


def grid_name(r0, c0, r1, c1):
    """Registered name of the block from cell (r0,c0) to cell (r1,c1): '0x0 - 1x0' etc. (spaces included)."""
    return f"{r0}x{c0} - {r1}x{c1}"


def add_grid(module, features):
    """Register the GridNet blocks of a `len(features)`-row, 4-column grid on `module` under the reference's
    names and in the reference's order (models/pointcloud_inpainting.py:100-116, disparity_estimation.py:128-148)."""
    R = len(features)
    for r, f in enumerate(features):
        for c in range(3):
            module.add_module(grid_name(r, c, r, c + 1), Basic('relu-conv-relu-conv', [f, f, f]))
    for c in (0, 1):
        for r in range(R - 1):
            module.add_module(grid_name(r, c, r + 1, c), Downsample([features[r], features[r + 1], features[r + 1]]))
    for c in (2, 3):
        for r in range(R - 1, 0, -1):
            module.add_module(grid_name(r, c, r - 1, c), Upsample([features[r], features[r - 1], features[r - 1]]))


def _crop_like(up, ref):
    # the x2 upsample of an odd-sized map is one row/column too large (reference: F.pad(..., -1))
    if up.size(2) != ref.size(2):
        up = up[:, :, :ref.size(2), :]
    if up.size(3) != ref.size(3):
        up = up[:, :, :, :ref.size(3)]
    return up


def grid_forward(module, rows):
    """Columns 1..3 of the GridNet given the column-0 activations `rows` (list, one tensor per row).
    Order of evaluation as in the reference (models/pointcloud_inpainting.py:141-172): column 1 top->bottom
    with the down-sampled, already updated row above added in; columns 2 and 3 bottom->top with the
    up-sampled, already updated row below added in."""
    m = module._modules
    R = len(rows)
    rows = list(rows)
    for r in range(R):
        rows[r] = m[grid_name(r, 0, r, 1)](rows[r])
        if r > 0:
            rows[r] = rows[r] + m[grid_name(r - 1, 1, r, 1)](rows[r - 1])
    for c in (2, 3):
        for r in range(R - 1, -1, -1):
            rows[r] = m[grid_name(r, c - 1, r, c)](rows[r])
            if r < R - 1:
                rows[r] = rows[r] + _crop_like(m[grid_name(r + 1, c, r, c)](rows[r + 1]), rows[r])
    return rows


def sample_norm(tensorImage, tensorDisparity):
    """Per-sample mean / unbiased std of image and disparity as [B,1,1,1] tensors
    (models/disparity_refinement.py:84-85, models/pointcloud_inpainting.py:219-220)."""
    B = tensorImage.size(0)
    img, disp = tensorImage.reshape(B, -1), tensorDisparity.reshape(B, -1)   # (.view in the reference: contiguous inputs there)
    mean = [img.mean(1, True).view(B, 1, 1, 1), disp.mean(1, True).view(B, 1, 1, 1)]
    std = [img.std(1, True).view(B, 1, 1, 1), disp.std(1, True).view(B, 1, 1, 1)]
    return mean, std
