"""Import and RUN the reference's own Python (staged by oracle/stage_ref.py into baseline/_ref/) on the GPU box.
TEST INFRASTRUCTURE: the ground truth of tests/test_gpu_reference_e2e.py -- never imported by the product.

The reference needs `cupy` (absent from this image) only as an NVRTC front end (utils/common.py:377-380:
`cupy.cuda.compile_with_cache(src, opts).get_function(name)(grid=, block=, args=, stream=)`), and a few packages it
imports but never uses on the inference path (imageio, moviepy, kornia, matplotlib).  This module provides
  * a `cupy` stand-in backed by cuda-python: NVRTC -> cubin for sm_100a, cuModuleLoadData, cuLaunchKernel with the
    reference's own grid/block/args (first arg a 32-bit int n, the rest device pointers: common.py:516-521);
  * empty stand-ins for the unused imports; `moviepy.editor.ImageSequenceClip` records the frame sequence
    Pipeline.__call__ hands to it (utils/pipeline.py:130-134) in `captured_clips` instead of encoding it;
  * torchvision's `vgg19_bn(pretrained=True)` / `maskrcnn_resnet50_fpn(pretrained=True)` without the download
    (no network): default-initialised VGG19-bn (the caller seeds / overwrites the weights), an empty module for the
    Mask R-CNN the reference builds and never calls (utils/pipeline.py:36).
Nothing in the staged files is edited; `utils.common.path_to_math_helper` (a private home-directory path,
common.py:14) is pointed at the staged helper_math.h after import, as any user of the reference has to.
"""
import ctypes
import hashlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")

captured_clips = []          # [(frames list, fps)] handed to moviepy by the reference's Pipeline
_loaded = None
_modules = {}                # sha1(src) -> CUmodule
nvrtc_compiles = 0


def available():
    return os.path.exists(os.path.join(STAGED, "utils", "common.py"))


def _chk(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA error {err}")
    return res[1] if len(res) == 2 else res[1:]


def _nvrtc_cubin(src, name="ref_kernel"):
    from cuda.bindings import nvrtc
    global nvrtc_compiles
    prog = _chk(nvrtc.nvrtcCreateProgram(src.encode(), (name + ".cu").encode(), 0, [], []))
    opts = [b"--gpu-architecture=sm_100a", b"-I/usr/local/cuda/include", b"-I" + os.path.join(STAGED, "utils").encode()]
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        n = _chk(nvrtc.nvrtcGetProgramLogSize(prog))
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC failed:\n" + log.decode(errors="replace"))
    n = _chk(nvrtc.nvrtcGetCUBINSize(prog))
    cubin = b" " * n
    _chk(nvrtc.nvrtcGetCUBIN(prog, cubin))
    nvrtc_compiles += 1
    return cubin


class _Function:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, grid=None, block=None, args=None, stream=None):
        from cuda.bindings import driver
        vals = [ctypes.c_int(int(args[0]))] + [ctypes.c_void_p(int(a)) for a in args[1:]]
        ptrs = (ctypes.c_void_p * len(vals))(*[ctypes.addressof(v) for v in vals])
        s = getattr(stream, "ptr", 0) or 0
        _chk(driver.cuLaunchKernel(self.fn, int(grid[0]), int(grid[1]), int(grid[2]), int(block[0]), int(block[1]), int(block[2]),
                                   0, s, ctypes.addressof(ptrs), 0))


class _Module:
    def __init__(self, src):
        import torch
        from cuda.bindings import driver
        key = hashlib.sha1(src.encode()).hexdigest()
        if key not in _modules:
            torch.zeros(1, device="cuda")                      # primary context current
            _modules[key] = _chk(driver.cuModuleLoadData(_nvrtc_cubin(src)))
        self.mod = _modules[key]

    def get_function(self, name):
        from cuda.bindings import driver
        return _Function(_chk(driver.cuModuleGetFunction(self.mod, name.encode())))


def _memoize(for_each_device=False):
    def deco(f):
        cache = {}

        def wrapped(*a):
            if a not in cache:
                cache[a] = f(*a)
            return cache[a]
        return wrapped
    return deco


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Clip:
    def __init__(self, sequence=None, fps=None, **kw):
        self.sequence, self.fps = list(sequence), fps

    def write_videofile(self, path, codec=None, **kw):
        captured_clips.append((self.sequence, self.fps, path))


def load():
    """-> namespace(common, pipeline, utils, Inpaint, PartialInpaint, Semantics, Disparity, Refine, RefineP, PartialConv2d):
    the reference's own modules, imported from the staged copy."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("baseline/_ref is empty: run oracle/stage_ref.py where /root/reference exists")
    import torch
    import torchvision

    cupy = _stub("cupy")
    cupy.util = _stub("cupy.util", memoize=_memoize)
    cupy.cuda = _stub("cupy.cuda", compile_with_cache=lambda src, options=(): _Module(src))
    for name in ("imageio", "moviepy", "kornia", "matplotlib"):
        if name not in sys.modules:
            _stub(name)
    sys.modules["moviepy"].editor = _stub("moviepy.editor", ImageSequenceClip=_Clip)
    sys.modules["matplotlib"].pyplot = _stub("matplotlib.pyplot")
    os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")

    _vgg = torchvision.models.vgg19_bn
    if not getattr(_vgg, "_kb_offline", False):
        def vgg19_bn(pretrained=False, **kw):
            return _vgg(weights=None)
        vgg19_bn._kb_offline = True
        torchvision.models.vgg19_bn = vgg19_bn
        torchvision.models.detection.maskrcnn_resnet50_fpn = lambda pretrained=False, **kw: torch.nn.Identity()

    if not torch.cuda.is_available():                      # CPU container: common.py:268 runs at import time
        class _S:
            cuda_stream = 0
        torch.cuda.current_stream = lambda *a, **k: _S()

    # the reference's top-level package names are `utils` and `models`
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k == "models" or k.startswith("models.")]:
        raise RuntimeError(f"a module named {k} is already imported; the reference needs that name")
    sys.path.insert(0, STAGED)
    try:
        import utils.common as common
        common.path_to_math_helper = os.path.join(STAGED, "utils", "helper_math.h")
        import utils.utils as rutils
        import utils.pipeline as pipeline
        from models.disparity_estimation import Disparity, Semantics
        from models.disparity_refinement import Refine
        from models.disparity_refinement_pretrained import Refine as RefineP
        from models.partial_inpainting import Inpaint as PartialInpaint
        from models.pointcloud_inpainting import Inpaint
        from utils.partial_conv import PartialConv2d
    finally:
        sys.path.remove(STAGED)
    _loaded = types.SimpleNamespace(common=common, pipeline=pipeline, utils=rutils, Inpaint=Inpaint, PartialInpaint=PartialInpaint,
                                    Semantics=Semantics, Disparity=Disparity, Refine=Refine, RefineP=RefineP,
                                    PartialConv2d=PartialConv2d)
    return _loaded


def fp32_convs():
    """The reference's arithmetic is fp32 (README pins PyTorch 1.3.1: no TF32 anywhere): keep cuDNN / cuBLAS from using TF32
    while the reference modules run."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
