"""CPU: the C-ABI shared library loads and exports exactly what include/kb200.h declares; the ctypes
binding lists the same symbols.  No compute call is made (no GPU here)."""
import os
import re
import subprocess

from ken_burns_effect_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "kb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_are_exported_and_bound():
    declared = _header_symbols()
    assert {"kb_render_frames", "kb_splat_min", "kb_degrid", "kb_splat_accum", "kb_fill"} <= declared
    nm = subprocess.check_output(["nm", "-D", "--defined-only", _native.LIB_PATH], text=True)
    exported = set(re.findall(r" T (kb_[a-z0-9_]+)", nm))
    assert declared <= exported, f"declared but not exported: {declared - exported}"
    assert exported <= declared, f"exported but not declared in kb200.h: {exported - declared}"
    assert declared == set(_native.SIGNATURES), set(_native.SIGNATURES) ^ declared


def test_library_loads_without_gpu_and_reports_version():
    L = _native.lib()
    assert L.kb_version() >= 100
    assert L.kb_accum_channels(4) == 8 and L.kb_accum_channels(68) == 72
    assert L.kb_render_workspace_bytes(1, 4, 768, 1024) == 4 * 768 * 1024 * (2 + 8)


def test_argument_errors_are_reported_not_raised_in_c():
    L = _native.lib()
    rc = L.kb_degrid(None, None, 1, 8, 8, None)
    assert rc == -1 and b"kb_degrid" in L.kb_last_error()


def test_sass_is_sm100_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _native.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_never_touches_the_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ken_burns_effect_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
                assert "kb_oracle" not in text and "libkb_oracle" not in text, f"{f} references the oracle library"
