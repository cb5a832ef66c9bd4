"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/moolib_b200.h
declares (no compute calls -- there is no GPU here)."""
import os
import re

from moolib_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "moolib_b200.h")).read()
    return sorted(set(re.findall(r"MB_API\s+[\w\s\*]+?\b(mb_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for s in declared_symbols():
        assert hasattr(L, s), f"libmoolib_b200.so does not export {s}"
    assert L.mb_version() == 1


def test_struct_layouts_match_header():
    import ctypes
    assert ctypes.sizeof(_lib.CopyJob) == 48
    assert ctypes.sizeof(_lib.ArHdr) == 32
    assert ctypes.sizeof(_lib.ArHandle) == 192


def test_flat_layout_rule():
    # tensors start on 4-float (16 B) boundaries
    assert _lib.flat_numel([5, 8, 3]) == 8 + 8 + 4
    assert _lib.flat_numel([992, 31]) == 992 + 32
    import oracle
    assert oracle.flat_layout([5, 8, 3]) == ([0, 8, 16], 20)


def test_argument_errors_are_reported_without_a_gpu():
    import ctypes
    L = _lib.load()
    ctx = ctypes.c_void_p()
    rc = L.mb_ar_ctx_create(0, 99, 0, 1024, 1, ctypes.byref(ctx))
    assert rc == _lib.MB_EINVAL and b"world" in L.mb_last_error()
    rc = L.mb_stack_slot(None, 1, 4, 9, 16, None, None)
    assert rc == _lib.MB_EINVAL and b"slot" in L.mb_last_error()
