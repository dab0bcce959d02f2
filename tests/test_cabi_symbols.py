"""The C-ABI library must load on a CPU-only box and export every symbol include/xmaps_b200.h
declares (no compute calls here)."""
import ctypes
import os
import re

from xm_helpers import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "xmaps_b200.h")) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    lib = ctypes.CDLL(os.path.join(ROOT, "x-maps_b200", "libxmaps_b200.so"))
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_covers_header():
    from xmaps_b200 import _native

    assert sorted(_native.SIGNATURES) == declared_symbols()
    assert _native.lib.xm_abi_version() == 1
    assert _native.launch_count() >= 0


def test_struct_layout_matches_header():
    """ctypes mirrors of the header structs: sizes that a C compiler would produce."""
    from xmaps_b200 import _native as N

    assert ctypes.sizeof(N.XmTables) == 10 * 4 + 8 + 6 * 8
    assert ctypes.sizeof(N.XmFrameArgs) == 8 + 8 + 4 * 4 + 8 + 8 + 8 + 4 + 4
    assert ctypes.sizeof(N.XmFrameStatus) == 5 * 8 + 4 * 4


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device."""
    from xmaps_b200 import _native as N

    out = ctypes.c_void_p()
    assert N.lib.xm_ctx_create(None, 0, ctypes.byref(out)) == N.ERR_INVALID_ARG
    assert b"null" in N.lib.xm_last_error()
    t = N.XmTables()
    assert N.lib.xm_ctx_create(ctypes.byref(t), 0, ctypes.byref(out)) == N.ERR_INVALID_ARG
    assert N.lib.xm_frame(None, None, None) == N.ERR_INVALID_ARG


def test_engine_refuses_cpu():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from xm_helpers import load_golden_tables
    from xmaps_b200.engine import DepthEngine, TableSet

    t, _ = load_golden_tables("small")
    with pytest.raises(RuntimeError):
        DepthEngine(TableSet(t.lut_x, t.lut_y, t.x_map, t.remap_xy, t.rect_w, t.rect_h, t.t_px_scale, t.x_offset, t.depth_scale))
