"""The C-ABI shared library loads and exports every symbol include/sr_b200.h declares.
No compute calls (no GPU needed)."""
import ctypes
import os
import re

import pytest

from stereoreconstruction_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported():
    if capi.needs_build():
        capi.build()
    L = ctypes.CDLL(capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/sr_b200.h but not exported"
    assert sorted(capi.EXPORTS) == syms


def test_struct_layouts_match_header():
    from stereoreconstruction_b200.types import SrCamera, SrParams
    # sr_camera: 9*4 + 3 + 3 + 5 + 3 + 1 + 1 + 3 doubles + 2 int32
    assert ctypes.sizeof(SrCamera) == (36 + 3 + 3 + 5 + 3 + 1 + 1 + 3) * 8 + 8
    assert SrParams.image_scale.offset == 24 and SrParams.second_best_factor.offset == 56
    from oracle.oracle_api import OrcCamera, OrcParams
    assert ctypes.sizeof(OrcCamera) == ctypes.sizeof(SrCamera) and ctypes.sizeof(OrcParams) == ctypes.sizeof(SrParams)


def test_no_device_fails_loudly():
    """Without a CUDA device the product refuses to run; it never falls back to a CPU path."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.SrError):
        capi.Context(0)


def test_product_never_references_the_oracle():
    """oracle/ is test infrastructure: nothing of the shipped path (package, CUDA sources, C++
    headers) may import, include, link or load it."""
    offenders = []
    for top in ("stereoreconstruction_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                    continue
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"^\s*(from|import)\s+oracle", r'#include\s*[<"].*oracle', r"liboracle", r"oracle_api", r"_ref/libref"):
                    if re.search(pat, text, flags=re.M):
                        offenders.append((os.path.join(dirpath, f), pat))
    assert not offenders, offenders
