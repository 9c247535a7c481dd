"""C-ABI library: loads without a GPU, exports every symbol include/*.h declares, struct layouts agree."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from xsdeepfwfm_deprecated_b200 import _lib
    return _lib.load()


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(dfw_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound(lib):
    from xsdeepfwfm_deprecated_b200 import _lib
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert decl == set(_lib.SYMBOLS), (decl ^ set(_lib.SYMBOLS))


def test_struct_layout_matches(lib):
    from xsdeepfwfm_deprecated_b200 import _lib
    assert lib.dfw_version() == _lib.DFW_ABI_VERSION
    assert lib.dfw_struct_bytes(0) == ctypes.sizeof(_lib.Model)
    assert lib.dfw_struct_bytes(1) == ctypes.sizeof(_lib.FieldDesc)
    assert lib.dfw_struct_bytes(2) == ctypes.sizeof(_lib.Csr)


def test_argument_errors_do_not_need_a_gpu(lib):
    from xsdeepfwfm_deprecated_b200 import _lib
    # NULL model -> DFW_E_ARG, wrong struct size -> DFW_E_ABI; both checked before any CUDA call
    assert lib.dfw_forward(None, None, 0, 0, None, 0, 0, 1, 0, None, 0, None, None, None, None) == -1
    m = _lib.Model()
    m.struct_bytes = 8
    assert lib.dfw_forward(ctypes.byref(m), None, 0, 0, None, 0, 0, 1, 0, None, 0, None, None, None, None) == -3
    assert b"ABI mismatch" in lib.dfw_last_error_string()


def test_prune_entry_points_check_arguments_before_touching_the_device(lib):
    from xsdeepfwfm_deprecated_b200 import _lib
    assert lib.dfw_prune_workspace_bytes() >= 64
    span = (_lib.PruneSpan * 1)()
    span[0].ptr, span[0].count = None, 10
    assert lib.dfw_prune_threshold(None, 1, 0, 0.5, 10, None, 0, None, None, None) == -1          # NULL spans
    assert lib.dfw_prune_threshold(span, 1, 0, 0.5, 10, None, 0, None, None, None) == -1          # NULL tensor with a count
    assert lib.dfw_prune_threshold(span, 500, 0, 0.5, 10, None, 0, None, None, None) == -1        # too many spans
    span[0].count = 0
    assert lib.dfw_prune_threshold(span, 1, 7, 0.5, 49, None, 0, None, None, None) == -1          # sym needs F*F elements
    assert lib.dfw_prune_apply(span, 1, 0, None, None, None) == -1                               # NULL threshold
    assert b"threshold" in lib.dfw_last_error_string()


def test_sass_is_sm100a_only():
    import subprocess
    from xsdeepfwfm_deprecated_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
