"""The C-ABI boundary, without a GPU: the library loads, exports exactly what
include/sepfilt.h declares, validates arguments before touching CUDA, and the product
package never reaches into oracle/."""
import ctypes
import os
import re

import pytest

from cupyimg_b200 import _build, _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "sepfilt.h")).read()
    return sorted(set(re.findall(r"SEPFILT_API\s+[\w\s\*]+?\b(sepfilt_\w+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = _declared()
    assert len(names) >= 7
    L = ctypes.CDLL(_build.LIB_PATH)
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_ffi.EXPORTS) == names


def test_version_and_struct_layout():
    L = _ffi.lib()
    assert L.sepfilt_version() == 100
    # sepfilt_tensor: ptr, 2 x i32, 2 x 8 x i64, 2 x i32
    assert ctypes.sizeof(_ffi.Tensor) == 8 + 8 + 64 + 64 + 8
    assert ctypes.sizeof(_ffi.Pass) == 4 + 4 + 8 + 16


def test_invalid_arguments_fail_before_cuda():
    L = _ffi.lib()
    t = _ffi.Tensor()
    t.ndim, t.dtype = 1, 8
    t.shape[0], t.stride_bytes[0] = 4, 4
    p, keep = _ffi.make_pass(0, [1.0, 2.0, 1.0], 0, 0)
    assert L.sepfilt_correlate1d(None, t, p, 0, 0.0, 0, None, 0, None) == _ffi.ERR_INVALID
    assert "NULL" in _ffi.last_error()
    p.origin = 2
    rc = L.sepfilt_correlate1d(t, t, p, 0, 0.0, 0, None, 0, None)
    assert rc == _ffi.ERR_VALUE                       # the exception class travels in the status code
    assert "origin" in _ffi.last_error()
    with pytest.raises(ValueError):
        _ffi.check(rc)
    p.origin, p.mode = 0, 9
    assert L.sepfilt_correlate1d(t, t, p, 0, 0.0, 0, None, 0, None) == _ffi.ERR_INVALID
    with pytest.raises(RuntimeError):
        _ffi.check(_ffi.ERR_INVALID)
    p.mode, p.axis = 0, 3
    rc = L.sepfilt_correlate1d(t, t, p, 0, 0.0, 0, None, 0, None)
    assert rc == _ffi.ERR_VALUE
    with pytest.raises(ValueError):
        _ffi.check(rc)
    t.dtype = 42
    p.axis = 0
    assert L.sepfilt_correlate1d(t, t, p, 0, 0.0, 0, None, 0, None) == _ffi.ERR_INVALID
    assert L.sepfilt_gradmag_step(None, None, 4, 8, 7, None) == _ffi.ERR_INVALID
    # empty output: success without any launch
    t.dtype = 8
    t.shape[0] = 0
    assert L.sepfilt_correlate1d(t, t, p, 0, 0.0, 0, None, 0, None) == _ffi.OK


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_ffi, "_lib", None)
    monkeypatch.setenv("SEPFILT_LIB", "/nonexistent/libsepfilt_b200.so")
    with pytest.raises(RuntimeError, match="not built"):
        _ffi.lib()


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cupyimg_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"\boracle\b", text) or re.search(r"^\s*(?:import\s+scipy|from\s+scipy)\b", text, re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_sass_is_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
