"""Execution of the reference's conv stacks on libkb200's tcgen05 convolution (include/kb200.h: kb_conv2d & co).

The nn.Module mirrors in ken_burns_effect_b200/models keep the reference's parameters and state_dict keys; when
their forward() is called with CUDA tensors they hand the work to the functions below, which walk the same
blocks (models/pointcloud_inpainting.py:7-81: Basic / Downsample / Upsample) but on NHWC activations and with
every bias, PReLU, residual and GridNet skip sum folded into a convolution epilogue:

    PReLU -> conv -> PReLU -> conv (+x)         becomes        conv[epilogue: +b, PReLU] -> conv[epilogue: +b, +x, outputs]

where "outputs" are the raw sum plus one copy per consumer that starts with its own PReLU.

An activation is a torch view [N,H,W,C] of a channels-last buffer whose pixel stride (stride(2)) is a multiple of 4
floats; a channel slice of a wider buffer is how torch.cat of the reference is expressed (no copy).
There is no fallback: a missing libkb200.so raises from _native.lib().
"""
import ctypes

import torch
import torch.nn as nn

from .. import _native as nat


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def round4(c):
    return (c + 3) & ~3


# Operand precision of the convolutions.  False: TF32 operands read from fp32 activations (the reference's own arithmetic under
# PyTorch's default cudnn.allow_tf32).  True: every tensor that is ONLY ever the input of a convolution -- the PReLU'd, already
# TF32-rounded copies the epilogues write (`round` outputs) and the up-sampled tensors -- is stored as fp16 (the same 10 mantissa
# bits, half the bytes) and those convolutions run tcgen05 kind::f16: twice the MACs per MMA at the same operand-fetch cost
# (DESIGN.md section 5).  Sums, residual streams, biases and network outputs stay fp32.  fp16's RANGE (|x| <= 65504, saturated)
# is the one thing TF32 does not share: KB200_CONV_F16=0 selects the TF32 path.
import contextlib as _contextlib
import os as _os0
F16_ENABLED = _os0.environ.get("KB200_CONV_F16", "0") == "1"
F16 = False          # what conv2d / upsample2x_prelu read; raised inside f16_operands() by the stacks that are fp16-clean


@_contextlib.contextmanager
def f16_operands():
    """Inside this scope `round` outputs are fp16 and their consumers run kind::f16 -- entered by the network forwards whose
    `round` outputs are read by convolutions only (the GridNets); a no-op unless KB200_CONV_F16=1."""
    global F16
    old = F16
    F16 = F16_ENABLED
    try:
        yield
    finally:
        F16 = old


def new_act(N, H, W, C, device, dtype=torch.float32):
    """Fresh NHWC activation with C logical channels (allocation padded so that a pixel is a multiple of 16 bytes)."""
    pad = 8 if dtype == torch.float16 else 4
    return torch.empty(N, H, W, (C + pad - 1) // pad * pad, device=device, dtype=dtype)[..., :C]


def _check_act(t):
    unit = 8 if t.dtype == torch.float16 else 4
    if not (t.is_cuda and t.dtype in (torch.float32, torch.float16) and t.dim() == 4 and t.stride(3) == 1 and t.stride(2) % unit == 0
            and t.stride(1) == t.stride(2) * t.size(2) and t.stride(0) == t.stride(1) * t.size(1)):
        raise RuntimeError(f"not an NHWC activation view: shape {tuple(t.shape)} strides {t.stride()} dtype {t.dtype}")
    return t


class PackedConv:
    """A convolution's parameters in the layout kb_conv2d reads (weights per tap and 32-channel slice, TF32-rounded)."""

    def __init__(self, conv, bn=None):
        w = conv.weight.detach()
        assert w.is_cuda and w.dtype == torch.float32
        self.Cout, self.Cin, kh, kw = w.shape
        assert kh == kw and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1] and conv.groups == 1
        self.k, self.stride, self.pad = kh, conv.stride[0], conv.padding[0]
        bias = conv.bias.detach() if conv.bias is not None else torch.zeros(self.Cout, device=w.device)
        scale = None
        if bn is not None:   # eval-mode BatchNorm folded into the filter and bias (VGG trunk, disparity_estimation.py:86)
            scale = (bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)).contiguous()
            bias = (bias - bn.running_mean) * scale + bn.bias.detach()
        self.bias = bias.contiguous().float()
        L = nat.lib()
        n = L.kb_conv_packed_floats(self.Cout, self.Cin, self.k)
        self.w = torch.empty(n, device=w.device, dtype=torch.float32)
        nat.check(L.kb_conv_pack_weights(_ptr(w.contiguous()), self.Cout, self.Cin, self.k, _ptr(scale), _ptr(self.w),
                                         _stream()), "kb_conv_pack_weights")
        self._w16 = None
        self._src = (w, scale)

    @property
    def w16(self):
        """The same filters as fp16 panels (kind::f16), packed on first use."""
        if self._w16 is None:
            L = nat.lib()
            w, scale = self._src
            self._w16 = torch.empty(L.kb_conv_packed_halves(self.Cout, self.Cin, self.k), device=w.device, dtype=torch.float16)
            nat.check(L.kb_conv_pack_weights_f16(_ptr(w.contiguous()), self.Cout, self.Cin, self.k, _ptr(scale), _ptr(self._w16),
                                                 _stream()), "kb_conv_pack_weights_f16")
        return self._w16


def packed(conv, bn=None):
    """PackedConv of an nn.Conv2d, cached on the module and rebuilt when its parameters change (load_state_dict)."""
    key = (conv.weight.data_ptr(), conv.weight._version, None if conv.bias is None else conv.bias._version,
           None if bn is None else (bn.weight._version, bn.running_var._version, bn.running_mean._version))
    cached = getattr(conv, '_kb_packed', None)
    if cached is None or cached[0] != key:
        cached = (key, PackedConv(conv, bn))
        object.__setattr__(conv, '_kb_packed', cached)
    return cached[1]


def conv2d(x, pc, outs, res=None, crop=None, tile_w=0, n_block=0, stages=0, algo=0, partial=None):
    """outs: list of (slope tensor | None, round_tf32 bool, destination view | None[, mul mask [N,Ho,Wo] | None]).
    Returns the output views.  out_o = prelu_o(pc(conv(x) + bias) + res) * mul_o; crop=(Ho,Wo) keeps only the top-left
    part (reference: F.pad(..., -1)); partial=(ratio, update_mask), both [N,Ho,Wo] from pconv_mask(), turns the layer into
    a PartialConv2d (utils/partial_conv.py:62-77)."""
    _check_act(x)
    N, H, W, Cin = x.shape
    if Cin != pc.Cin:
        raise RuntimeError(f"conv2d: input has {Cin} channels, filter expects {pc.Cin}")
    Ho = (H + 2 * pc.pad - pc.k) // pc.stride + 1
    Wo = (W + 2 * pc.pad - pc.k) // pc.stride + 1
    if crop is not None:
        Ho, Wo = min(Ho, crop[0]), min(Wo, crop[1])
    a = nat.KBConvArgs()
    a.x, a.N, a.H, a.W, a.Cin, a.x_stride = x.data_ptr(), N, H, W, Cin, x.stride(2)
    a.x_f16 = 1 if x.dtype == torch.float16 else 0
    a.w_packed, a.bias = (pc.w16 if a.x_f16 else pc.w).data_ptr(), pc.bias.data_ptr()
    a.Cout, a.ksize, a.stride, a.pad = pc.Cout, pc.k, pc.stride, pc.pad
    if res is not None:
        _check_act(res)
        if tuple(res.shape) != (N, Ho, Wo, pc.Cout):
            raise RuntimeError(f"conv2d: residual {tuple(res.shape)} does not match output {(N, Ho, Wo, pc.Cout)}")
        a.res, a.res_stride = res.data_ptr(), res.stride(2)
    keep = []
    if partial is not None:
        ratio, um = partial
        for t in (ratio, um):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (N, Ho, Wo)):
                raise RuntimeError(f"conv2d: partial-conv maps must be contiguous fp32 {(N, Ho, Wo)}, got {tuple(t.shape)}")
        a.pc_ratio, a.pc_um = ratio.data_ptr(), um.data_ptr()
        keep += [ratio, um]
    a.n_out = len(outs)
    result = []
    for i, spec in enumerate(outs):
        slope, rnd, dst = spec[:3]
        mul = spec[3] if len(spec) > 3 else None
        if mul is not None:
            if not (mul.is_cuda and mul.dtype == torch.float32 and mul.is_contiguous() and tuple(mul.shape) == (N, Ho, Wo)):
                raise RuntimeError(f"conv2d: mul mask must be contiguous fp32 {(N, Ho, Wo)}, got {tuple(mul.shape)}")
            a.out[i].mul = mul.data_ptr()
            keep.append(mul)
        if dst is None:     # a `round` output only ever feeds convolutions: fp16 when the fp16 operand path is on
            dst = new_act(N, Ho, Wo, pc.Cout, x.device, torch.float16 if (rnd and F16) else torch.float32)
        _check_act(dst)
        if tuple(dst.shape) != (N, Ho, Wo, pc.Cout):
            raise RuntimeError(f"conv2d: destination {tuple(dst.shape)} does not match output {(N, Ho, Wo, pc.Cout)}")
        a.out[i].ptr, a.out[i].pixel_stride = dst.data_ptr(), dst.stride(2)
        if slope is not None:
            s = slope.detach()
            keep.append(s)
            a.out[i].slope = s.data_ptr()
        a.out[i].round_tf32 = 1 if rnd else 0
        a.out[i].store_f16 = 1 if dst.dtype == torch.float16 else 0
        result.append(dst)
    a.out_H, a.out_W = (Ho, Wo) if crop is not None else (0, 0)
    a.tile_w, a.n_block, a.stages, a.algo = tile_w, n_block, stages, algo
    nat.check(nat.lib().kb_conv2d(ctypes.byref(a), _stream()), "kb_conv2d")
    return result


def upsample2x_prelu(x, slope, out_hw=None, rnd=True, mul=None):
    """nn.Upsample(x2, bilinear) -> PReLU [-> * mul, a [N,Ho,Wo] mask] on an NHWC activation."""
    _check_act(x)
    N, H, W, C = x.shape
    Ho, Wo = (2 * H, 2 * W) if out_hw is None else out_hw
    if x.dtype != torch.float32:
        raise RuntimeError("upsample2x_prelu reads the raw fp32 value of the coarser row")
    f16 = bool(rnd and F16)
    y = new_act(N, Ho, Wo, C, x.device, torch.float16 if f16 else torch.float32)
    if mul is not None and not (mul.is_cuda and mul.dtype == torch.float32 and mul.is_contiguous() and tuple(mul.shape) == (N, Ho, Wo)):
        raise RuntimeError(f"upsample2x_prelu: mul mask must be contiguous fp32 {(N, Ho, Wo)}")
    nat.check(nat.lib().kb_upsample2x_prelu(_ptr(x), x.stride(2), N, H, W, C, _ptr(slope.detach()) if slope is not None else None,
                                            _ptr(y), y.stride(2), Ho, Wo, 2 if f16 else (1 if rnd else 0), _ptr(mul), _stream()),
              "kb_upsample2x_prelu")
    return y


def pconv_mask(mask, shape_nhw, Cin, ksize, stride, pad):
    """Mask bookkeeping of a PartialConv2d(multi_channel=True) layer whose mask channels are identical
    (utils/partial_conv.py:43-69): mask [N,H,W] of {0,1} or None (= no mask) -> (mask_ratio, update_mask), [N,Ho,Wo] each."""
    N, H, W = shape_nhw
    Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    dev = mask.device if mask is not None else torch.device('cuda', torch.cuda.current_device())
    if mask is not None and not (mask.is_cuda and mask.dtype == torch.float32 and mask.is_contiguous() and tuple(mask.shape) == (N, H, W)):
        raise RuntimeError(f"pconv_mask: mask must be contiguous fp32 {(N, H, W)}, got {tuple(mask.shape)}")
    ratio = torch.empty(N, Ho, Wo, device=dev, dtype=torch.float32)
    um = torch.empty(N, Ho, Wo, device=dev, dtype=torch.float32)
    nat.check(nat.lib().kb_pconv_mask(_ptr(mask), N, H, W, Cin, ksize, stride, pad, _ptr(ratio), _ptr(um), _stream()), "kb_pconv_mask")
    return ratio, um


def prelu(x, slope, rnd=True, dst=None):
    _check_act(x)
    N, H, W, C = x.shape
    y = new_act(N, H, W, C, x.device) if dst is None else _check_act(dst)
    nat.check(nat.lib().kb_prelu_nhwc(_ptr(x), x.stride(2), N * H * W, C, _ptr(slope.detach()) if slope is not None else None,
                                      _ptr(y), y.stride(2), 1 if rnd else 0, _stream()), "kb_prelu_nhwc")
    return y


def maxpool2_ceil(x):
    _check_act(x)
    N, H, W, C = x.shape
    y = new_act(N, (H + 1) // 2, (W + 1) // 2, C, x.device)
    nat.check(nat.lib().kb_maxpool2_ceil(_ptr(x), x.stride(2), N, H, W, C, _ptr(y), y.stride(2), _stream()), "kb_maxpool2_ceil")
    return y


def to_nhwc(x_nchw, dst=None, sub=0.0, mul=1.0):
    """[N,C,H,W] contiguous -> NHWC activation, y = (x - sub) * mul (dst: view to fill, e.g. a slice of a concat buffer)."""
    x = x_nchw.contiguous()
    N, C, H, W = x.shape
    y = new_act(N, H, W, C, x.device) if dst is None else _check_act(dst)
    nat.check(nat.lib().kb_nchw_to_nhwc(_ptr(x), N, C, H, W, _ptr(y), y.stride(2), float(sub), float(mul), _stream()),
              "kb_nchw_to_nhwc")
    return y


def to_nchw(x, mul=1.0, add=0.0):
    _check_act(x)
    N, H, W, C = x.shape
    y = torch.empty(N, C, H, W, device=x.device, dtype=torch.float32)
    nat.check(nat.lib().kb_nhwc_to_nchw(_ptr(x), x.stride(2), N, C, H, W, _ptr(y), float(mul), float(add), _stream()),
              "kb_nhwc_to_nchw")
    return y


# -----------------------------------------------------------------------------------------------------------------
# blocks (models/pointcloud_inpainting.py:7-81 and the copies in the other model files)
# -----------------------------------------------------------------------------------------------------------------

def block_parts(block):
    """-> (has_upsample, pre PReLU | None, conv1, PReLU, conv2) of a Basic / Downsample / Upsample block."""
    layers = list(block.moduleMain)
    up = isinstance(layers[0], nn.Upsample)
    if up:
        layers = layers[1:]
    pre = None
    if isinstance(layers[0], nn.PReLU):
        pre, layers = layers[0], layers[1:]
    conv1, act, conv2 = layers
    return up, pre, conv1, act, conv2


def pre_slope(block):
    """Weight of the PReLU a block applies to its input before its first convolution (None for 'conv-relu-conv')."""
    pre = block_parts(block)[1]
    return None if pre is None else pre.weight


def run_block(block, x_in, outs, x_raw=None, extra_res=None, crop=None):
    """One Basic / Downsample / Upsample block.
    x_in : the block input with the block's own leading PReLU already applied by its producer
           (for an Upsample block: the RAW low-resolution input; up-sampling + PReLU happen here);
    x_raw: the un-activated input, needed when the block has a residual / 1x1 shortcut (Basic);
    extra_res: a tensor added to the block output (GridNet: the other branch arriving at the same cell);
    outs : output specs of conv2d()."""
    up, pre, conv1, act, conv2 = block_parts(block)
    if up:
        x_in = upsample2x_prelu(x_in, pre.weight)
    t, = conv2d(x_in, packed(conv1), [(act.weight, True, None)])
    res = extra_res
    if getattr(block, 'residual', False):
        assert extra_res is None and x_raw is not None
        sc = block.moduleShortcut
        res = x_raw if sc is None else conv2d(x_raw, packed(sc), [(None, False, None)])[0]
    return conv2d(t, packed(conv2), outs, res=res, crop=crop)


def grid_forward_nhwc(module, features, stem, semantics_res=None):
    """The GridNet of models/pointcloud_inpainting.py:136-172 / disparity_estimation.py:157-195 on NHWC activations.
    stem(specs) must produce the value of cell (0,0) for the given output specs; returns the raw value of cell (0,3).
    Cell values are dicts {'raw', 'basic', 'down'}: the sum itself and the copies pre-activated for the Basic block of
    the next column and the Downsample block to the next row."""
    from ..models.gridnet import grid_name
    m = module._modules
    R = len(features)

    def specs(r, c):
        keys, out = [], []
        need_raw = c < 3 or r > 0 or (r == 0 and c == 3)
        if need_raw:
            keys.append('raw'); out.append((None, False, None))
        if r == 0 and c == 3 and F16:
            keys.append('head'); out.append((None, True, None))      # what the heads' first convolutions read (no PReLU before them)
        if c < 3:
            keys.append('basic'); out.append((pre_slope(m[grid_name(r, c, r, c + 1)]), True, None))
        if c in (0, 1) and r < R - 1:
            keys.append('down'); out.append((pre_slope(m[grid_name(r, c, r + 1, c)]), True, None))
        return keys, out

    def cell(keys, tensors):
        return dict(zip(keys, tensors))

    V = [None] * R
    keys, out = specs(0, 0)
    V[0] = cell(keys, stem(out))
    for r in range(1, R):                                   # column 0
        keys, out = specs(r, 0)
        extra = semantics_res if (semantics_res is not None and r == 3) else None
        V[r] = cell(keys, run_block(m[grid_name(r - 1, 0, r, 0)], V[r - 1]['down'], out, extra_res=extra))
    for r in range(R):                                      # column 1, top -> bottom
        keys, out = specs(r, 1)
        basic = m[grid_name(r, 0, r, 1)]
        if r == 0:
            V[0] = cell(keys, run_block(basic, V[0]['basic'], out, x_raw=V[0]['raw']))
        else:
            t1, = run_block(basic, V[r]['basic'], [(None, False, None)], x_raw=V[r]['raw'])
            V[r] = cell(keys, run_block(m[grid_name(r - 1, 1, r, 1)], V[r - 1]['down'], out, extra_res=t1))
    for c in (2, 3):                                        # columns 2, 3, bottom -> top
        for r in range(R - 1, -1, -1):
            keys, out = specs(r, c)
            basic = m[grid_name(r, c - 1, r, c)]
            if r == R - 1:
                V[r] = cell(keys, run_block(basic, V[r]['basic'], out, x_raw=V[r]['raw']))
            else:
                t1, = run_block(basic, V[r]['basic'], [(None, False, None)], x_raw=V[r]['raw'])
                hw = (t1.size(1), t1.size(2))
                V[r] = cell(keys, run_block(m[grid_name(r + 1, c, r, c)], V[r + 1]['raw'], out, extra_res=t1, crop=hw))
    return V[0]['raw'] if 'head' not in V[0] else (V[0]['raw'], V[0]['head'])


# -----------------------------------------------------------------------------------------------------------------
# CUDA-graph replay of a whole network forward
# -----------------------------------------------------------------------------------------------------------------
# A network forward is 60-130 launches issued from Python through ctypes (~30 us of host time each).  In isolation the host
# stays ahead of the GPU (profiles/bench_nets_r01h.md: Inpaint 5.33 ms eager vs 5.36 ms replayed), but inside the pipeline every
# host sync (.item(), nonzero, minMaxLoc) drains the queue and the GPU then waits for the host to refill it: the CNN stage of
# a KBE takes 22.5 ms eager and 20.8 ms with the forwards replayed from CUDA graphs.  The forwards are pure functions of their
# input tensors with static shapes, so from the THIRD call with the same shapes and weights on they are captured and replayed
# (a single image -- two inpainting passes -- never pays the capture; a server processing a stream of images does once).
# KB_GRAPHS=0 disables it.
import os as _os

GRAPHS_ENABLED = _os.environ.get("KB_GRAPHS", "1") != "0"
GRAPH_EAGER_CALLS = 2


def _fingerprint(owner):
    return tuple((p.data_ptr(), p._version) for p in owner.parameters()) + tuple((b.data_ptr(), b._version) for b in owner.buffers())


def graphed(owner, tag, fn, *tensors):
    """fn(*tensors) -> tensor or tuple of tensors; eager for the first GRAPH_EAGER_CALLS calls with a given signature (they
    also pack the weights and warm the allocator), captured into a CUDA graph on the next one and replayed afterwards.
    Outputs are fresh tensors (copies of the graph's static outputs).  The graph is dropped when the input shapes or any
    parameter of `owner` change (load_state_dict, .to())."""
    if not GRAPHS_ENABLED or not tensors[0].is_cuda or torch.cuda.is_current_stream_capturing():
        return fn(*tensors)
    store = owner.__dict__.get('_kb_graphs')          # lives and dies with the module (not a registered attribute)
    if store is None:
        store = {}
        object.__setattr__(owner, '_kb_graphs', store)
    shapes = tuple((tuple(t.shape), t.dtype, t.device.index) for t in tensors)
    sig = (shapes, _fingerprint(owner))
    # one graph per (forward, input shapes): a pipeline that alternates batch sizes (estimate_depth_batch next to single images)
    # keeps both; a change of the owner's parameters drops every graph of that forward
    ent = store.get((tag, shapes))
    if ent is None or ent['sig'] != sig:
        if ent is not None:
            for k in [k for k in store if k[0] == tag]:
                del store[k]
        ent = store[(tag, shapes)] = {'sig': sig, 'graph': None, 'calls': 0}
    if ent['graph'] is None:
        ent['calls'] += 1
        if ent['calls'] <= GRAPH_EAGER_CALLS:
            return fn(*tensors)
        static_in = [t.clone() for t in tensors]
        torch.cuda.current_stream().synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):    # a render thread may be launching on its own stream
            out = fn(*static_in)
        ent.update(graph=g, static_in=static_in, static_out=out)
    else:
        for s, t in zip(ent['static_in'], tensors):
            s.copy_(t)
    ent['graph'].replay()
    out = ent['static_out']
    return out.clone() if torch.is_tensor(out) else tuple(o.clone() for o in out)


def head_nhwc(block, x_raw):
    """Basic('conv-relu-conv') head with its 1x1 shortcut (moduleImage / moduleDisparity).  x_raw: the raw row-0 value, or
    (raw fp32, fp16 copy) from grid_forward_nhwc under f16_operands()."""
    x_in, x_raw = (x_raw[1], x_raw[0]) if isinstance(x_raw, tuple) else (x_raw, x_raw)
    return run_block(block, x_in, [(None, False, None)], x_raw=x_raw)[0]
