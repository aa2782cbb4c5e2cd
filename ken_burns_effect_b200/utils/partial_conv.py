"""PartialConv2d -- mirror of utils/partial_conv.py:14-84 (NVIDIA partial convolution, Liu et al. 2018) with the
same constructor keywords (multi_channel, return_mask), the same cached-mask behaviour and the same
arithmetic:  out = ((conv(x*mask) - b) * ratio + b) * update_mask,  ratio = slide_winsize / (sum(mask) + 1e-8)
clamped through update_mask = clamp(sum(mask), 0, 1)."""
import torch
import torch.nn.functional as F
from torch import nn


class PartialConv2d(nn.Conv2d):
    def __init__(self, *args, **kwargs):
        self.multi_channel = kwargs.pop('multi_channel', False)
        self.return_mask = kwargs.pop('return_mask', False)
        super().__init__(*args, **kwargs)
        kh, kw = self.kernel_size
        if self.multi_channel:
            self.weight_maskUpdater = torch.ones(self.out_channels, self.in_channels, kh, kw)
        else:
            self.weight_maskUpdater = torch.ones(1, 1, kh, kw)
        # plain attribute, not a buffer: it must stay out of the state_dict (partial_conv.py:33)
        shp = self.weight_maskUpdater.shape
        self.slide_winsize = shp[1] * shp[2] * shp[3]
        self.last_size = (None, None, None, None)
        self.update_mask = None
        self.mask_ratio = None

    def _refresh_mask(self, input, mask_in):
        with torch.no_grad():
            if self.weight_maskUpdater.type() != input.type():
                self.weight_maskUpdater = self.weight_maskUpdater.to(input)
            if mask_in is None:
                if self.multi_channel:
                    mask = torch.ones(input.shape[0], input.shape[1], input.shape[2], input.shape[3]).to(input)
                else:
                    mask = torch.ones(1, 1, input.shape[2], input.shape[3]).to(input)
            else:
                mask = mask_in
            self.update_mask = F.conv2d(mask, self.weight_maskUpdater, bias=None, stride=self.stride,
                                        padding=self.padding, dilation=self.dilation, groups=1)
            self.mask_ratio = self.slide_winsize / (self.update_mask + 1e-8)
            self.update_mask = torch.clamp(self.update_mask, 0, 1)
            self.mask_ratio = torch.mul(self.mask_ratio, self.update_mask)

    def forward(self, input, mask_in=None):
        assert len(input.shape) == 4
        if mask_in is not None or self.last_size != tuple(input.shape):
            self.last_size = tuple(input.shape)
            self._refresh_mask(input, mask_in)
        raw_out = super().forward(torch.mul(input, mask_in) if mask_in is not None else input)
        if self.bias is not None:
            bias_view = self.bias.view(1, self.out_channels, 1, 1)
            output = torch.mul(raw_out - bias_view, self.mask_ratio) + bias_view
            output = torch.mul(output, self.update_mask)
        else:
            output = torch.mul(raw_out, self.mask_ratio)
        if self.return_mask:
            return output, self.update_mask
        return output
