"""The C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import videorenderer_b200 as V
from videorenderer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpcvr_[a-z0-9_]+)\s*\(", text)) - {"mpcvr_sub_render_fn", "mpcvr_sub_set_device_fn"})


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(V.lib_path())
    names = _declared("mpcvr_b200.h") + _declared("mpcvr_b200_debug.h")
    assert len(names) >= 45
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in _abi.SIGNATURES, "python binding missing: " + n
    assert set(_abi.SIGNATURES) <= set(names)
    assert V.lib().mpcvr_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the C side was compiled with (any drift between include/mpcvr_b200.h and _abi.py shows up here)
    assert ctypes.sizeof(_abi.Settings) == 33 * 4
    assert ctypes.sizeof(_abi.Rect) == 16
    assert ctypes.sizeof(_abi.MediaType) == 5 * 4 + 16 + 4 * 4 + 4 + 8 + 2 * 4
    assert ctypes.sizeof(_abi.SideDataHDR) == 80 and ctypes.sizeof(_abi.SideDataHDRCLL) == 8   # Include/IMediaSideData.h:35-60, pack(1)
    assert ctypes.sizeof(_abi.PlanInfo) == 15 * 4
    assert ctypes.sizeof(_abi.DoviMetadata) == 5736 and ctypes.sizeof(_abi.Sample) == 120
    s = _abi.Sample()
    assert ctypes.sizeof(s) % 8 == 0


def test_product_does_not_link_or_import_the_oracle():
    """the oracle is test infrastructure: nothing under videorenderer_b200/ may reference it"""
    pkg = os.path.join(ROOT, "videorenderer_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "orclib" not in text and "liborc" not in text and "orc_" not in text.replace("force_", ""), os.path.join(dirpath, f)
