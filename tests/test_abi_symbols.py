"""The C-ABI library loads and exports every symbol include/gpw.h declares (no compute calls)."""
import ctypes
import os
import re

import gpw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gpw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpw_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(gpw.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libgpw.so does not export %s" % n
        assert n in gpw.SYMBOLS, "python binding lacks %s" % n
    assert sorted(gpw.SYMBOLS) == names


def test_no_cpu_fallback_without_device():
    if gpw.device_count() > 0:
        return
    try:
        gpw.Context(0)
    except gpw.GpwError as e:
        assert e.code == -2 and "no CPU fallback" in str(e)
    else:
        raise AssertionError("Context creation must fail loudly without a GPU")
