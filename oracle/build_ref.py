#!/usr/bin/env python3
"""Compile the REFERENCE's own CUDA kernels into cubins under oracle/_ref/  (test infrastructure).

The reference (pierlj/ken-burns-effect) keeps its five CUDA kernels as Python strings inside
utils/common.py and specialises them per tensor shape with a regex pre-pass (common.py:271-375) before
handing them to cupy/NVRTC (common.py:377-380).  This script imports that file *where it lies* under
/root/reference behind a tiny `cupy` stand-in, lets the reference's own preprocess_kernel() produce the
specialised source for each shape in SHAPES, and compiles that source with NVRTC for sm_100a.

Only binaries (+ a manifest) are written, only into oracle/_ref/ (git-ignored, shipped to the GPU box by
gpurun).  No reference source text is copied into the repository.  On the GPU box the parity tests load
these cubins with the CUDA driver API (oracle/refgpu.py) and use them as the ground truth that pins the
CPU restatement (oracle/kb_oracle.c) and the product kernels.

Run:  python oracle/build_ref.py        (needs /root/reference; no GPU needed)
"""
import hashlib
import json
import os
import sys
import types
import warnings

REF = os.environ.get("KB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

# (H, W, N, C, focal, baseline, B) tuples the GPU parity tests use.  focal/baseline keep the Python types
# the reference pastes into the source (float 512.0, int 120 -- pipeline.py:26-27).
SHAPES = [
    dict(H=48, W=64, N=48 * 64, C=4, focal=32.0, baseline=120, B=1),
    dict(H=48, W=64, N=48 * 64 + 517, C=4, focal=32.0, baseline=120, B=1),
    dict(H=96, W=128, N=96 * 128, C=4, focal=64.0, baseline=120, B=2),
    dict(H=192, W=256, N=192 * 256 + 4099, C=4, focal=128.0, baseline=120, B=1),
    dict(H=192, W=256, N=192 * 256, C=68, focal=128.0, baseline=120, B=1),
    dict(H=192, W=256, N=192 * 256, C=4, focal=101.37, baseline=120, B=1),   # dolly-style focal
    dict(H=768, W=1024, N=768 * 1024, C=4, focal=512.0, baseline=120, B=1),
    dict(H=768, W=1024, N=768 * 1024 + 70001, C=4, focal=512.0, baseline=120, B=1),
    dict(H=768, W=1024, N=768 * 1024, C=68, focal=512.0, baseline=120, B=1),
    dict(H=2160, W=3840, N=2160 * 3840, C=4, focal=1920.0, baseline=120, B=1),      # configs[3]
]


def _install_shims(captured):
    import torch

    class _Fn:
        def __init__(self, name):
            self.name = name

        def __call__(self, grid=None, block=None, args=None, stream=None):
            return None  # no GPU here: the launch is a no-op, we only want the specialised source

    class _Mod:
        def __init__(self, src):
            self.src = src

        def get_function(self, name):
            captured.append((name, self.src))
            return _Fn(name)

    cupy = types.ModuleType("cupy")
    cupy.util = types.ModuleType("cupy.util")
    cupy.cuda = types.ModuleType("cupy.cuda")
    cupy.util.memoize = lambda for_each_device=False: (lambda f: f)
    cupy.cuda.compile_with_cache = lambda src, options=(): _Mod(src)
    sys.modules["cupy"] = cupy
    sys.modules["cupy.util"] = cupy.util
    sys.modules["cupy.cuda"] = cupy.cuda

    class _S:
        cuda_stream = 0

    torch.cuda.current_stream = lambda *a, **k: _S()  # common.py:268 runs at import time
    import torchvision

    _orig = torchvision.models.vgg19_bn
    torchvision.models.vgg19_bn = lambda pretrained=False, **kw: _orig(weights=None)
    os.environ.setdefault("CUDA_HOME", "/usr/local/cuda")


def _nvrtc_cubin(src, name):
    from cuda.bindings import nvrtc

    def chk(res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError(f"nvrtc error {err}")
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    prog = chk(nvrtc.nvrtcCreateProgram(src.encode(), (name + ".cu").encode(), 0, [], []))
    opts = [b"--gpu-architecture=sm_100a", b"-I/usr/local/cuda/include", b"-I" + os.path.join(REF, "utils").encode()]
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        n = chk(nvrtc.nvrtcGetProgramLogSize(prog))
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC failed:\n" + log.decode(errors="replace"))
    n = chk(nvrtc.nvrtcGetCUBINSize(prog))
    cubin = b" " * n
    chk(nvrtc.nvrtcGetCUBIN(prog, cubin))
    return cubin


def main():
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent: keeping whatever oracle/_ref/ already holds")
        return 0
    warnings.filterwarnings("ignore")
    import torch

    captured = []
    _install_shims(captured)
    sys.path.insert(0, REF)
    import utils.common as rc  # the reference file itself

    rc.path_to_math_helper = os.path.join(REF, "utils", "helper_math.h")
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    seen = {}
    for shp in SHAPES:
        H, W, N, C, B = shp["H"], shp["W"], shp["N"], shp["C"], shp["B"]
        focal, baseline = shp["focal"], shp["baseline"]
        del captured[:]
        pts = torch.zeros(B, 3, N)
        dat = torch.zeros(B, C, N)
        rc.render_pointcloud(pts, dat, W, H, focal, baseline)          # common.py:428 -> 3 kernels
        rc.fill_disocclusion(torch.zeros(B, C, H, W), torch.zeros(B, 1, H, W))  # common.py:833
        if N == H * W and C == 4:
            rc.device = "cpu"
            rc.generate_mask(torch.zeros(B, 3, N), torch.zeros(B, 3, 1), W, H, focal, baseline)  # common.py:689
        roles = ["updateZee", "updateDegrid", "updateOutput", "discfill", "maskZee"]
        for role, (name, src) in zip(roles, captured):
            key = hashlib.sha1(src.encode()).hexdigest()[:16]
            fname = f"{name}_{key}.cubin"
            if key not in seen:
                cubin = _nvrtc_cubin(src, name)
                with open(os.path.join(OUT, fname), "wb") as f:
                    f.write(cubin)
                seen[key] = fname
            manifest.append(dict(shp, role=role, entry=name, file=fname))
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(f"[build_ref] wrote {len(seen)} cubins, {len(manifest)} manifest rows into {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
