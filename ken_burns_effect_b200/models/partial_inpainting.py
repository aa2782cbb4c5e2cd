"""PartialInpaint -- mirror of models/partial_inpainting.py:99-279: the Inpaint topology built from
PartialConv2d, masks travelling with the activations (up-sampled masks thresholded at 0.5, skip merges by
min).  Parameter names follow the reference: p_relu_1 / conv1 / p_relu_2 / conv2 / moduleShortcut."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..utils import convstack as cs
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
            tensorContext = (self._context(tensorImage, tensorDisparity) if tensorImage.is_cuda
                             else self.moduleContext(torch.cat([tensorImage, tensorDisparity], 1)))
            tensorData = torch.cat([tensorImage, tensorDisparity, tensorContext], 1)

        if tensorData.is_cuda:
            img, disp, k0 = cs.graphed(self, 'grid', self._forward_b200, tensorData.contiguous(), tensorMasks.contiguous())
            k0 = k0[:, None].expand(tensorData.shape[0], self.FEATURES[0], tensorData.shape[2], tensorData.shape[3])
            img, disp = self.normalize_images_disp(img, disp, not_normed=False)
            return {
                'tensorExisting': k0,
                'tensorExistingInput': tensorMasks,
                'tensorImage': img.clamp(0.0, 1.0) if self.training == False else img,  # noqa: E712
                'tensorDisparity': F.threshold(input=disp, threshold=0.0, value=0.0),
            }

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

    # -- the same forward on libkb200 (NHWC activations, tcgen05 convolutions with the partial-conv renormalisation, the
    #    next layer's PReLU and the consumer's mask folded into the epilogue; utils/convstack.py) --------------------------
    # Every mask of this network has identical channels (it starts as tensorMasks.expand_as(data), :152, and every update
    # is channel-independent), so masks travel as ONE channel [N,H,W]; kb_pconv_mask turns a mask into the per-pixel
    # mask_ratio / update_mask of a layer exactly (sums of 0/1 values).
    @staticmethod
    def _pblock(block, x_in, m_in, make_outs, x_raw=None, extra_res=None, merge_mask=None, crop=None):
        """One Basic / Downsample / Upsample chain of PartialConv2d layers.
        x_in : conv1's input with the block's leading PReLU and mask already applied by its producer (Upsample: the RAW
               coarse input -- up-sampling, PReLU and masking happen here, partial_inpainting.py:88-93);
        m_in : the mask travelling with x_in, [N,H,W], or None (heads: called without a mask, :212-213);
        merge_mask: mask of the other branch arriving at the same grid cell (skip merge: torch.min, :167);
        make_outs(k) -> output specs of cs.conv2d for the cell mask k.  Returns (outputs, k)."""
        if getattr(block, 'upsample_first', False):
            m_in = (F.interpolate(m_in[:, None], scale_factor=2, mode='bilinear', align_corners=False)[:, 0] > 0.5).float().contiguous()
            x_in = cs.upsample2x_prelu(x_in, block.p_relu_1.weight, mul=m_in)
        c1, c2 = block.conv1, block.conv2
        N, H, W, _ = x_in.shape
        r1, u1 = cs.pconv_mask(m_in, (N, H, W), c1.in_channels, c1.kernel_size[0], c1.stride[0], c1.padding[0])
        t, = cs.conv2d(x_in, cs.packed(c1), [(block.p_relu_2.weight, True, None)], partial=(r1, u1))
        r2, u2 = cs.pconv_mask(u1, tuple(u1.shape), c2.in_channels, c2.kernel_size[0], 1, c2.padding[0])
        if crop is not None:
            r2, u2 = r2[:, :crop[0], :crop[1]].contiguous(), u2[:, :crop[0], :crop[1]].contiguous()
        k = u2 if merge_mask is None else torch.min(merge_mask, u2)
        res = extra_res
        if isinstance(block, Basic):
            assert extra_res is None and x_raw is not None
            sc = block.moduleShortcut
            # the 1x1 shortcut is a PartialConv2d called without a mask (:47): ratio = Cin / (Cin + 1e-8) = 1, update_mask = 1
            res = x_raw if sc is None else cs.conv2d(x_raw, cs.packed(sc), [(None, False, None)])[0]
        return cs.conv2d(t, cs.packed(c2), make_outs(k), res=res, crop=crop, partial=(r2, u2)), k

    def _forward_b200(self, tensorData, tensorMasks):
        with cs.f16_operands():       # every `round` output / up-sampled tensor of this stack is read by convolutions only
            return self._forward_b200_body(tensorData, tensorMasks)

    def _forward_b200_body(self, tensorData, tensorMasks):
        m = self._modules
        F_ = self.FEATURES
        R = len(F_)
        N, C, H, W = tensorData.shape
        m0 = tensorMasks[:, 0].contiguous()
        x_raw = cs.to_nhwc(tensorData)
        x_msk = cs.to_nhwc(tensorData * tensorMasks)                    # conv(input * mask), partial_conv.py:71

        def pre(block):
            return block.p_relu_1.weight if block.pre_relu else None

        def cell_specs(r, c):
            keys = []
            if c < 3 or r > 0 or (r == 0 and c == 3):
                keys.append(('raw', None))
            if c < 3:
                keys.append(('basic', pre(m[grid_name(r, c, r, c + 1)])))
            if c in (0, 1) and r < R - 1:
                keys.append(('down', pre(m[grid_name(r, c, r + 1, c)])))
            return keys

        def run(block, x_in, m_in, r, c, **kw):
            keys = cell_specs(r, c)
            outs, k = self._pblock(block, x_in, m_in,
                                   lambda k: [(sl, sl is not None, None, k if sl is not None else None) for _, sl in keys], **kw)
            cell = {name: t for (name, _), t in zip(keys, outs)}
            cell['mask'] = k
            return cell

        def raw_only(block, x_in, m_in, **kw):
            outs, k = self._pblock(block, x_in, m_in, lambda k: [(None, False, None)], **kw)
            return outs[0], k

        V = [None] * R
        V[0] = run(self.moduleInput, x_msk, m0, 0, 0, x_raw=x_raw)
        for r in range(1, R):                                            # column 0
            V[r] = run(m[grid_name(r - 1, 0, r, 0)], V[r - 1]['down'], V[r - 1]['mask'], r, 0)
        for r in range(R):                                               # column 1, top -> bottom
            basic = m[grid_name(r, 0, r, 1)]
            if r == 0:
                V[0] = run(basic, V[0]['basic'], V[0]['mask'], 0, 1, x_raw=V[0]['raw'])
            else:
                t1, kb = raw_only(basic, V[r]['basic'], V[r]['mask'], x_raw=V[r]['raw'])
                V[r] = run(m[grid_name(r - 1, 1, r, 1)], V[r - 1]['down'], V[r - 1]['mask'], r, 1, extra_res=t1, merge_mask=kb)
        for c in (2, 3):                                                 # columns 2, 3, bottom -> top
            for r in range(R - 1, -1, -1):
                basic = m[grid_name(r, c - 1, r, c)]
                if r == R - 1:
                    V[r] = run(basic, V[r]['basic'], V[r]['mask'], r, c, x_raw=V[r]['raw'])
                else:
                    t1, kb = raw_only(basic, V[r]['basic'], V[r]['mask'], x_raw=V[r]['raw'])
                    hw = (t1.size(1), t1.size(2))
                    V[r] = run(m[grid_name(r + 1, c, r, c)], V[r + 1]['raw'], V[r + 1]['mask'], r, c, extra_res=t1,
                               merge_mask=kb, crop=hw)
        row0 = V[0]['raw']
        img, _ = raw_only(self.moduleImage, row0, None, x_raw=row0)      # heads run without a mask (:212-213)
        disp, _ = raw_only(self.moduleDisparity, row0, None, x_raw=row0)
        return cs.to_nchw(img), cs.to_nchw(disp), V[0]['mask']
