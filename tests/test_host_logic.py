"""Host-side logic that needs no GPU: argument validation, tap generation, aliasing
analysis, descriptor construction."""
import itertools

import numpy as np
import pytest

from cupyimg_b200 import _array, _ffi
from cupyimg_b200.scipy.ndimage import filters as F


def test_gaussian_kernel_matches_scipy_bitwise():
    from scipy.ndimage._filters import _gaussian_kernel1d
    for sigma, order in itertools.product([0.5, 1.0, 1.5, 2.0, 4.0, 7.3], range(4)):
        lw = int(4 * sigma + 0.5)
        np.testing.assert_array_equal(F._gaussian_kernel1d(sigma, order, lw), _gaussian_kernel1d(sigma, order, lw))
    with pytest.raises(ValueError):
        F._gaussian_kernel1d(1.0, -1, 4)


def test_origin_mode_axis_validation():
    for width in range(1, 9):
        lo, hi = -(width // 2), (width - 1) // 2
        for o in range(lo, hi + 1):
            assert F._check_origin(o, width) == o
        for o in (lo - 1, hi + 1):
            with pytest.raises(ValueError):
                F._check_origin(o, width)
    for m in ("reflect", "constant", "nearest", "mirror", "wrap", "grid-mirror", "grid-wrap", "grid-constant"):
        assert 0 <= F._check_mode(m) <= 4
    assert F._check_mode("grid-constant") == F._check_mode("constant")
    assert F._check_mode("grid-wrap") == F._check_mode("wrap")
    assert F._check_mode("grid-mirror") == F._check_mode("reflect")
    for bad in ("", "foo", None, 3):
        with pytest.raises(RuntimeError):
            F._check_mode(bad)
    assert F._normalize_axis_index(-1, 3) == 2
    for bad in (3, -4):
        with pytest.raises(ValueError):
            F._normalize_axis_index(bad, 3)
    with pytest.raises(RuntimeError):
        F._normalize_sequence([1, 2], 3)
    assert F._normalize_sequence("wrap", 2) == ["wrap", "wrap"]
    with pytest.raises(ValueError):
        F._check_dtype_mode("numpy")


def test_pass_spec_radius():
    s = F._PassSpec(0, np.ones(17), 0, 0)
    assert s.radius() == 8
    s = F._PassSpec(0, np.ones(4), -2, 0)       # taps cover offsets 0..3
    assert s.radius() == 3
    s = F._PassSpec(0, None, 1, 0, uniform=True, size=5)
    assert s.radius() == 3
    st, keep = s.struct()
    assert st.uniform == 1 and st.ntaps == 5 and not st.taps


def _fake(ptr, shape, strides, dtype="float32"):
    return _array.DevArray(ptr, shape, strides, dtype, 0, None)


def test_overlap_and_contiguity_analysis():
    a = _fake(1000, (4, 5), (20, 4))
    assert a.c_contiguous() and a.byte_bounds() == (1000, 1080)
    b = _fake(1080, (4, 5), (20, 4))
    assert not a.may_overlap(b)
    c = _fake(1076, (2,), (4,))
    assert a.may_overlap(c)
    neg = _fake(1076, (4, 5), (-20, 4))          # flipped view of the same block
    assert neg.byte_bounds() == (1016, 1096)
    assert not neg.c_contiguous()
    assert _fake(0, (0, 3), (12, 4)).byte_bounds() == (0, 0)
    assert not _fake(1000, (0, 5), (20, 4)).may_overlap(a)
    w = a.view_axis_window(0, 1, 2)
    assert w.shape == (2, 5) and w.ptr == 1020
    t = a.tensor()
    assert (t.ndim, t.dtype, t.shape[1], t.stride_bytes[0]) == (2, 8, 5, 20)
    with pytest.raises(RuntimeError):
        _fake(0, (2,), (2,), "float16").tensor()


def test_dtype_helpers():
    import torch
    assert _array.to_numpy_dtype(torch.uint16) == np.dtype("uint16")
    assert _array.to_numpy_dtype("f4") == np.dtype("float32")
    assert _array.to_numpy_dtype(np.int8) == np.dtype("int8")
    assert issubclass(_array.OutputShapeError, ValueError) and issubclass(_array.OutputShapeError, RuntimeError)
    w = _array.host_weights([1, 2, 3])
    assert w.dtype == np.float64
    with pytest.raises(TypeError):
        _array.ingest(np.zeros(3))
    with pytest.raises(TypeError):
        _array.ingest(torch.zeros(3))


def test_complex_views_and_output_rules():
    """Complex arrays are filtered by component: strided float views of the same memory; a complex input or
    complex weights need a complex output (_util.py:52-75)."""
    c = _fake(4096, (3, 5), (40, 8), "complex64")
    re, im = c.component(0), c.component(1)
    assert re.dtype == np.dtype("float32") and im.dtype == np.dtype("float32")
    assert (re.ptr, im.ptr) == (4096, 4100) and re.strides == im.strides == (40, 8)
    assert not re.c_contiguous() and re.may_overlap(im)         # interleaved: byte ranges overlap
    z = _fake(0, (2,), (16,), "complex128")
    assert z.component(1).ptr == 8 and z.component(1).dtype == np.dtype("float64")
    with pytest.raises(TypeError):
        _fake(0, (2,), (4,), "float32").component(0)
    with pytest.raises(RuntimeError):
        F._get_output(np.float32, c)                            # complex input, real output dtype
    with pytest.raises(RuntimeError):
        F._get_output(np.float64, _fake(0, (2, 2), (8, 4)), complex_output=True)   # complex weights
    assert F._split_cval(1.5, False) == (1.5, 0.0)
    assert F._split_cval(1.5 - 2j, True) == (1.5, -2.0)
    with pytest.raises(ValueError):
        F._split_cval(1j, False)


def test_window_filter_specs():
    """minimum / maximum passes never take the float32 tiled or fused paths (exact per-axis kernels)."""
    a = _fake(1024, (8, 16, 32), (2048, 128, 4))
    b = _fake(1 << 20, (8, 16, 32), (2048, 128, 4))
    mx = F._PassSpec(1, None, 0, 0, uniform=F._MAX, size=5)
    mean = F._PassSpec(1, None, 0, 0, uniform=True, size=5)
    assert mx.radius() == 2 and mx.struct()[0].uniform == 3 and mean.struct()[0].uniform == 1
    assert F._f32_tiled_ok(a, b, mean) and not F._f32_tiled_ok(a, b, mx)
    assert F._fused_candidate(a, b, [mean, F._PassSpec(2, None, 0, 0, uniform=True, size=3)], False)
    assert not F._fused_candidate(a, b, [mx, F._PassSpec(2, None, 0, 0, uniform=F._MAX, size=3)], False)
    with pytest.raises(NotImplementedError):
        F._check_minmax_cval(float("nan"))


def test_gaussian_pass_specs_are_cached_and_immutable():
    """One pass spec (and one ctypes struct) per (axis, sigma, order, radius, mode): a loop that calls the same filter
    spends no host time on taps; the cached taps cannot be written through."""
    a = F._gaussian_spec(1, 2.0, 0, F._check_mode("reflect"), 4.0)
    b = F._gaussian_spec(1, np.float64(2.0), 0, F._check_mode("reflect"), 4.0)
    assert a is b and a.struct()[0] is b.struct()[0]
    assert F._gaussian_spec(1, 2.0, 1, F._check_mode("reflect"), 4.0) is not a          # another order
    assert F._gaussian_spec(0, 2.0, 0, F._check_mode("reflect"), 4.0) is not a          # another axis
    assert F._gaussian_spec(1, 2.0, 0, F._check_mode("mirror"), 4.0) is not a           # another mode
    assert F._gaussian_spec(1, 2.0, 0, F._check_mode("reflect"), 4.0, radius=5).size == 11
    assert a.size == 17 and a.radius() == 8
    with pytest.raises(ValueError):
        a.taps[0] = 1.0
    with pytest.raises(ValueError):
        F._gaussian_spec(1, 2.0, 0, F._check_mode("reflect"), 4.0, radius=-1)


def test_status_codes_carry_the_exception_class():
    """include/sepfilt.h: SEPFILT_ERR_VALUE -> ValueError, SEPFILT_ERR_UNSUPPORTED -> NotImplementedError, the rest
    RuntimeError (the class no longer depends on the message text)."""
    assert (_ffi.OK, _ffi.ERR_INVALID, _ffi.ERR_UNSUPPORTED, _ffi.ERR_SCRATCH, _ffi.ERR_CUDA, _ffi.ERR_VALUE) == (0, -1, -2, -3, -4, -5)
    _ffi.check(_ffi.OK)
    for rc, exc in [(_ffi.ERR_VALUE, ValueError), (_ffi.ERR_UNSUPPORTED, NotImplementedError),
                    (_ffi.ERR_INVALID, RuntimeError), (_ffi.ERR_SCRATCH, RuntimeError), (_ffi.ERR_CUDA, RuntimeError)]:
        with pytest.raises(exc):
            _ffi.check(rc)
    header = open(__import__("os").path.join(__import__("os").path.dirname(__file__), "..", "include", "sepfilt.h")).read()
    assert "SEPFILT_ERR_VALUE = -5" in header


def test_component_view_detection():
    class Owner:
        dtype = np.dtype("complex64")
    c = _array.DevArray(0, (1,), (8,), np.complex64, 0, Owner())
    assert F._is_component_view(c.component(0)) and F._is_component_view(c.component(1))
    assert not F._is_component_view(c)
    assert c.component(0).c_contiguous()            # why contiguity alone is not enough for a size-1 array
    assert not F._is_component_view(_array.DevArray(0, (4,), (4,), np.float32, 0, None))
