"""Semantics (VGG19-bn trunk) and Disparity (6-row GridNet) -- mirrors of models/disparity_estimation.py:82-198
with identical constructors, forwards and state_dict keys."""
import torch
import torch.nn as nn
import torchvision

from ..utils import convstack as cs
from .gridnet import Basic, add_grid, grid_forward, grid_name


# Opt-in for a default-initialised VGG trunk (tests, benchmarks, `kbe.py --random-weights`, Pipeline(model_paths=None)):
# set by those callers or by KB200_RANDOM_VGG=1.  Without it a missing ImageNet checkpoint is an error, never a silent
# random network: Semantics' weights are in none of the released .tar files, so garbage here is garbage frames.
ALLOW_RANDOM_VGG = False


def _vgg19_bn_features():
    """torchvision's VGG19-bn feature stack with the ImageNet weights the reference asks for
    (disparity_estimation.py:86: vgg19_bn(pretrained=True)).  Order: an explicit state_dict file named by
    KB200_VGG19_BN_WEIGHTS; torchvision's own loader (local hub cache, else download); and only with the explicit opt-in above a
    default-initialised trunk, announced on stderr."""
    import os
    import sys
    explicit = os.environ.get('KB200_VGG19_BN_WEIGHTS')
    if explicit:
        vgg = torchvision.models.vgg19_bn(weights=None)
        vgg.load_state_dict(torch.load(explicit, map_location='cpu'))
        return vgg.features.eval()
    weights = torchvision.models.VGG19_BN_Weights.IMAGENET1K_V1
    cached = os.path.join(torch.hub.get_dir(), 'checkpoints', os.path.basename(weights.url))
    random_ok = ALLOW_RANDOM_VGG or os.environ.get('KB200_RANDOM_VGG') == '1'
    if os.path.exists(cached) or not random_ok:
        try:
            return torchvision.models.vgg19_bn(weights=weights).features.eval()
        except Exception as exc:                                   # no network / unwritable cache
            if not random_ok:
                raise RuntimeError(
                    "Semantics needs torchvision's ImageNet VGG19-bn weights (the reference loads vgg19_bn(pretrained=True)) and "
                    f"they could not be loaded ({exc}).  Put {os.path.basename(weights.url)} into {os.path.dirname(cached)}, or "
                    "point KB200_VGG19_BN_WEIGHTS at a state_dict file, or opt in to a randomly initialised trunk with "
                    "--random-weights / KB200_RANDOM_VGG=1.") from exc
    print("ken_burns_effect_b200: WARNING -- VGG19-bn trunk is RANDOMLY initialised (opt-in); disparities are meaningless",
          file=sys.stderr)
    return torchvision.models.vgg19_bn(weights=None).features.eval()


class Semantics(nn.Module):
    def __init__(self):
        super().__init__()
        vgg = _vgg19_bn_features()
        pool = lambda: nn.MaxPool2d(kernel_size=2, stride=2, ceil_mode=True)  # noqa: E731
        # slices keep torchvision's indices as keys ("moduleVgg.3.7.weight"), disparity_estimation.py:88-105
        self.moduleVgg = nn.Sequential(
            vgg[0:3], vgg[3:6], pool(),
            vgg[7:10], vgg[10:13], pool(),
            vgg[14:17], vgg[17:20], vgg[20:23], vgg[23:26], pool(),
            vgg[27:30], vgg[30:33], vgg[33:36], vgg[36:39], pool())

    def forward(self, tensorInput):
        # BGR -> RGB, ImageNet normalisation (disparity_estimation.py:108-116); done out of place
        x = tensorInput[:, [2, 1, 0], :, :]
        mean = x.new_tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
        std = x.new_tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
        x = (x - mean) / std
        return cs.graphed(self, 'vgg', self._vgg_b200, x.contiguous()) if x.is_cuda else self.moduleVgg(x)

    def _vgg_b200(self, x):
        """The VGG19-bn trunk on libkb200 convolutions: eval-mode BatchNorm folded into filter and bias, ReLU in the
        epilogue (PReLU with slope 0), ceil-mode 2x2 max-pool as its own NHWC kernel."""
        h = cs.to_nhwc(x)
        for layer in self.moduleVgg:
            if isinstance(layer, nn.MaxPool2d):
                h = cs.maxpool2_ceil(h)
                continue
            conv, bn, _relu = list(layer)
            zero = getattr(self, '_relu_slopes', {}).get(conv.out_channels)
            if zero is None or zero.device != x.device:
                if not hasattr(self, '_relu_slopes'):
                    object.__setattr__(self, '_relu_slopes', {})
                zero = self._relu_slopes[conv.out_channels] = torch.zeros(conv.out_channels, device=x.device)
            h, = cs.conv2d(h, cs.packed(conv, bn), [(zero, True, None)])
        return cs.to_nchw(h)


class Disparity(nn.Module):
    FEATURES = (32, 48, 64, 512, 512, 512)

    def __init__(self):
        super().__init__()
        self.spectral_norm = False
        self.moduleImage = nn.Conv2d(in_channels=3, out_channels=32, kernel_size=7, stride=2, padding=3)
        self.moduleSemantics = nn.Conv2d(in_channels=512, out_channels=512, kernel_size=3, stride=1, padding=1)
        add_grid(self, self.FEATURES)
        self.moduleDisparity = Basic('conv-relu-conv', [32, 32, 1])

    def _forward_b200(self, tensorImage, tensorSemantics):
        x = cs.to_nhwc(tensorImage)
        sem, = cs.conv2d(cs.to_nhwc(tensorSemantics), cs.packed(self.moduleSemantics), [(None, False, None)])
        with cs.f16_operands():       # every `round` output of the GridNet is read by convolutions only
            row0 = cs.grid_forward_nhwc(self, self.FEATURES, lambda outs: cs.conv2d(x, cs.packed(self.moduleImage), outs),
                                        semantics_res=sem)
            return cs.to_nchw(cs.head_nhwc(self.moduleDisparity, row0))

    def forward(self, tensorImage, tensorSemantics):
        if tensorImage.is_cuda:
            return cs.graphed(self, 'forward', self._forward_b200, tensorImage.contiguous(), tensorSemantics.contiguous())
        m = self._modules
        rows = [self.moduleImage(tensorImage)]
        for r in range(1, len(self.FEATURES)):
            nxt = m[grid_name(r - 1, 0, r, 0)](rows[r - 1])
            if r == 3:
                nxt = nxt + self.moduleSemantics(tensorSemantics)      # disparity_estimation.py:159
            rows.append(nxt)
        rows = grid_forward(self, rows)
        return self.moduleDisparity(rows[0])
