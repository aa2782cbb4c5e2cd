"""ctypes binding of libkb200.so (the C ABI in include/kb200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.  The
library is built in-tree by `make -C ken_burns_effect_b200/csrc` (see __graft_entry__.build()).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libkb200.so")

KB_MAX_POSES = 32

c_void = ctypes.c_void_p
c_int = ctypes.c_int
c_long = ctypes.c_long
c_double = ctypes.c_double
c_size_t = ctypes.c_size_t
c_f32p = ctypes.POINTER(ctypes.c_float)


class KBPose(ctypes.Structure):
    _fields_ = [("shift", ctypes.c_float * 3), ("_pad", ctypes.c_float), ("focal", ctypes.c_double)]


class KBFrameParams(ctypes.Structure):
    _fields_ = [("H", c_int), ("W", c_int), ("crop_w", c_int), ("crop_h", c_int), ("baseline", c_double)]


class KBConvOut(ctypes.Structure):
    _fields_ = [("ptr", c_void), ("pixel_stride", c_long), ("slope", c_void), ("mul", c_void), ("round_tf32", c_int), ("store_f16", c_int)]


class KBConvArgs(ctypes.Structure):
    _fields_ = [("x", c_void), ("N", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("x_stride", c_long),
                ("w_packed", c_void), ("bias", c_void), ("Cout", c_int), ("ksize", c_int), ("stride", c_int),
                ("pad", c_int), ("res", c_void), ("res_stride", c_long), ("pc_ratio", c_void), ("pc_um", c_void), ("n_out", c_int),
                ("out", KBConvOut * 3),
                ("out_H", c_int), ("out_W", c_int), ("tile_w", c_int), ("n_block", c_int), ("stages", c_int), ("algo", c_int),
                ("x_f16", c_int)]


# name -> (restype, argtypes); must list every symbol include/kb200.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "kb_version": (c_int, []),
    "kb_last_error": (ctypes.c_char_p, []),
    "kb_launch_count": (ctypes.c_longlong, []),
    "kb_shift_points": (c_int, [c_void, c_void, c_void, c_int, c_long, c_void]),
    "kb_splat_min": (c_int, [c_void, c_int, c_long, c_f32p, c_double, c_double, c_void, c_int, c_int, c_void, c_void]),
    "kb_degrid": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void]),
    "kb_accum_channels": (c_int, [c_int]),
    "kb_splat_accum": (c_int, [c_void, c_void, c_int, c_long, c_int, c_f32p, c_double, c_double, c_void, c_void,
                               c_int, c_int, c_void]),
    "kb_splat_accum_rows": (c_int, [c_void, c_void, c_long, c_int, c_long, c_int, c_f32p, c_double, c_double, c_void, c_void,
                                    c_int, c_int, c_void]),
    "kb_accum_weight": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_void]),
    "kb_normalize_rows": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_void]),
    "kb_normalize": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_void, c_void]),
    "kb_render_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "kb_render_pointcloud": (c_int, [c_void, c_void, c_int, c_long, c_int, c_int, c_int, c_double, c_double,
                                     c_void, c_void, c_void, c_void]),
    "kb_fill": (c_int, [c_void, c_void, c_void, c_int, c_int, c_int, c_int, c_void]),
    "kb_median5_binary": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void]),
    "kb_mask_workspace_bytes": (c_size_t, [c_int, c_long, c_int, c_int]),
    "kb_generate_mask": (c_int, [c_void, c_int, c_long, c_double, c_double, c_int, c_int, c_void, c_void, c_void]),
    "kb_laplacian5": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void]),
    "kb_frames_workspace_bytes": (c_size_t, [ctypes.POINTER(KBFrameParams), c_int]),
    "kb_render_frames": (c_int, [c_void, c_void, c_long, ctypes.POINTER(KBPose), c_int,
                                 ctypes.POINTER(KBFrameParams), c_void, c_void, c_void]),
    "kb_image_front_end": (c_int, [c_void, c_int, c_int, c_int, c_void, c_void]),
    "kb_coverage_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "kb_coverage": (c_int, [c_void, c_long, ctypes.POINTER(KBPose), c_int, c_int, c_int, c_double, c_void, c_void, c_void]),
    "kb_conv_packed_floats": (c_long, [c_int, c_int, c_int]),
    "kb_conv_pack_weights": (c_int, [c_void, c_int, c_int, c_int, c_void, c_void, c_void]),
    "kb_conv_packed_halves": (c_long, [c_int, c_int, c_int]),
    "kb_conv_pack_weights_f16": (c_int, [c_void, c_int, c_int, c_int, c_void, c_void, c_void]),
    "kb_conv2d": (c_int, [ctypes.POINTER(KBConvArgs), c_void]),
    "kb_upsample2x_prelu": (c_int, [c_void, c_long, c_int, c_int, c_int, c_int, c_void, c_void, c_long, c_int, c_int,
                                    c_int, c_void, c_void]),
    "kb_pconv_mask": (c_int, [c_void, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void, c_void, c_void]),
    "kb_prelu_nhwc": (c_int, [c_void, c_long, c_long, c_int, c_void, c_void, c_long, c_int, c_void]),
    "kb_maxpool2_ceil": (c_int, [c_void, c_long, c_int, c_int, c_int, c_int, c_void, c_long, c_void]),
    "kb_nchw_to_nhwc": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_long, ctypes.c_float, ctypes.c_float, c_void]),
    "kb_nhwc_to_nchw": (c_int, [c_void, c_long, c_int, c_int, c_int, c_int, c_void, ctypes.c_float, ctypes.c_float, c_void]),
    "kb_selftest_arith": (c_int, [c_int, c_void, c_void, c_long, c_int, c_void, c_void]),
    "kb_profile_enable": (c_int, [c_int]),
    "kb_profile_read": (c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]),
}

FRAME_STAGES = ("memset_accum", "init_zbuf_tables", "splat_min", "degrid", "splat_accum", "resolve", "fill", "crop_resize")

_lib = None


def lib():
    """Load libkb200.so once; raise loudly when it is absent (no CPU or eager fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C ken_burns_effect_b200/csrc` "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no fallback path.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().kb_last_error()
        raise RuntimeError(f"{what} failed (rc={rc}): {msg.decode(errors='replace') if msg else ''}")


def launch_count():
    return int(lib().kb_launch_count())


def profile_enable(on):
    lib().kb_profile_enable(1 if on else 0)


def profile_read():
    """-> ({stage: summed ms}, calls) for all kb_render_frames calls since the last read."""
    ms = (ctypes.c_double * len(FRAME_STAGES))()
    calls = ctypes.c_longlong(0)
    lib().kb_profile_read(ms, ctypes.byref(calls))
    return {k: float(ms[i]) for i, k in enumerate(FRAME_STAGES)}, int(calls.value)
