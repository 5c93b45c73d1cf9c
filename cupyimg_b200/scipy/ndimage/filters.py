"""Separable n-d correlation filters with the call signatures of
``cupyimg.scipy.ndimage.filters`` (themselves ``scipy.ndimage``'s), executed by the
hand-written sm_100a kernels of libsepfilt_b200 through its C ABI.

Drop-in for the functions on the reference's separable hot path
(/root/reference/cupyimg/scipy/ndimage/filters.py):

    correlate1d :213    convolve1d :286    uniform_filter1d :549    uniform_filter :602
    gaussian_filter1d :668    gaussian_filter :725    prewitt :828    sobel :889
    generic_laplace :963    laplace :1041    gaussian_laplace :1077
    generic_gradient_magnitude :1125    gaussian_gradient_magnitude :1207
    minimum_filter / maximum_filter :1296-1396 (separable sizes)    minimum_filter1d / maximum_filter1d :1422-1508
    correlate :65    convolve :136 (dense N-d weights, exact arithmetic)

Inputs are CUDA arrays (``torch.Tensor``, ``cupy.ndarray`` or anything exposing
``__cuda_array_interface__`` / ``__dlpack__``); the result is the same kind of object.
There is no CPU path: a missing / unbuilt shared library raises ``RuntimeError``.

Arithmetic policy (keyword ``dtype_mode``, cf. _util.py:28-40):
  * ``None`` (default): float32 -> float32 runs the float32 tiled / fused kernels
    (float32 FMA accumulate, within rtol 1e-5 of scipy); every other dtype pair runs the
    exact kernels (float64, scipy's summation order, no FMA contraction — bit-exact
    integer outputs, bit-identical float64).
  * ``"ndimage"``: always the exact float64 kernels.
  * ``"float"``: like ``None`` (the reference's float32 accumulation for <= 16-bit
    integers is served by the exact kernels, which are more accurate).
Where the reference deviates from scipy this module follows scipy:
``uniform_filter`` uses an exact window sum / size, ``convolve1d`` honours ``cval``,
``grid-constant`` == ``constant``, Gaussian taps stay float64.
"""
import functools
import numbers

import numpy as np
from numpy.exceptions import AxisError

from ... import _array, _ffi
from ..._array import DevArray

__all__ = [
    "correlate1d", "convolve1d", "uniform_filter1d", "uniform_filter",
    "gaussian_filter1d", "gaussian_filter", "prewitt", "sobel",
    "generic_laplace", "laplace", "gaussian_laplace",
    "generic_gradient_magnitude", "gaussian_gradient_magnitude",
    "minimum_filter1d", "maximum_filter1d", "minimum_filter", "maximum_filter",
    "correlate", "convolve",
]

_F32 = np.dtype("float32")


# ----------------------------------------------------------------------------
# argument handling (mirrors _util.py / _filters_core.py of the reference)
# ----------------------------------------------------------------------------
def _check_mode(mode):
    """_util._check_mode (_util.py:105-119)."""
    if not isinstance(mode, str) or mode not in _ffi.MODE_CODES:
        raise RuntimeError("boundary mode not supported (actual: {})".format(mode))
    return _ffi.MODE_CODES[mode]


def _check_origin(origin, width):
    """_util._check_origin (_util.py:98-102)."""
    origin = int(origin)
    if (width // 2 + origin < 0) or (width // 2 + origin >= width):
        raise ValueError("invalid origin")
    return origin


def _normalize_axis_index(axis, ndim):
    """_misc._normalize_axis_index (_misc.py:134-157); AxisError is a ValueError."""
    axis = int(axis)
    if not -ndim <= axis < ndim:
        raise AxisError("axis {} is out of bounds for array of dimension {}".format(axis, ndim))
    return axis % ndim


def _normalize_sequence(arg, rank):
    """_util._normalize_sequence (_util.py:137-151)."""
    if hasattr(arg, "__iter__") and not isinstance(arg, str):
        if _array.is_device_array(arg):
            arg = _array.host_weights(arg)
        normalized = list(arg)
        if len(normalized) != rank:
            raise RuntimeError("sequence argument must have length equal to input rank")
    else:
        normalized = [arg] * rank
    return normalized


def _check_dtype_mode(dtype_mode):
    if dtype_mode not in (None, "auto", "ndimage", "float"):
        raise ValueError("dtype_mode={!r} not supported".format(dtype_mode))
    return dtype_mode


_COMPLEX = (np.dtype("complex64"), np.dtype("complex128"))


def _get_output(output, inp, shape=None, complex_output=None):
    """_util._get_output (_util.py:43-81) without the memset: returns (DevArray, user_owned).
    A complex input or complex weights need a complex output (_util.py:52-58, :69-75); with
    ``output=None`` it is promote_types(input, complex64) (_util.py:66-67)."""
    shape = inp.shape if shape is None else tuple(shape)
    if complex_output is None:
        complex_output = inp.dtype.kind == "c"

    def check(dt):
        if complex_output:
            if dt.kind != "c":
                raise RuntimeError("output must have complex dtype if either the input or "
                                   "weights are complex-valued.")
            if dt not in _COMPLEX:
                raise RuntimeError("array type {} not supported".format(dt))
        elif dt not in _ffi.DTYPE_CODES or dt == np.dtype("bool"):
            raise RuntimeError("array type {} not supported".format(dt))

    if output is None:
        dt = np.promote_types(inp.dtype, np.complex64) if complex_output else inp.dtype
        return _array.empty(shape, dt, inp.device), False
    if _array.is_device_array(output):
        out = _array.ingest(output, "output")
        if out.shape != shape:
            raise _array.OutputShapeError("output shape not correct")
        if out.device != inp.device:
            raise RuntimeError("output must live on the same device as the input")
        check(out.dtype)
        return out, True
    dt = _array.to_numpy_dtype(output)
    check(dt)
    return _array.empty(shape, dt, inp.device), False


def _ingest_input(input):
    inp = _array.ingest(input, "input")
    if inp.dtype not in _ffi.DTYPE_CODES and inp.dtype not in _COMPLEX:
        raise RuntimeError("array type {} not supported".format(inp.dtype))
    return inp


def _host_taps(weights):
    return _array.host_weights(weights)


def _split_cval(cval, complex_input):
    """Complex arrays are filtered by real and imaginary component, each with its part of cval."""
    if isinstance(cval, complex) or np.iscomplexobj(cval):
        if not complex_input:
            raise ValueError("Cannot provide a complex-valued cval when the input is real.")
        return float(np.real(cval)), float(np.imag(cval))
    return float(cval), 0.0


# ----------------------------------------------------------------------------
# pass execution
# ----------------------------------------------------------------------------
class _PassSpec:
    """One 1-D pass: taps (host float64) or a uniform window, in correlation orientation."""

    __slots__ = ("axis", "taps", "size", "origin", "mode", "uniform", "_struct")

    def __init__(self, axis, taps, origin, mode, uniform=False, size=0):
        self.axis = axis
        self.taps = taps
        self.size = int(size) if uniform else int(len(taps))
        self.origin = int(origin)
        self.mode = mode                # integer mode code
        self.uniform = uniform
        self._struct = None

    def radius(self):
        before = self.size // 2 + self.origin
        return max(before, self.size - 1 - before)

    def struct(self):
        """(ctypes Pass, keepalive) — built once per spec."""
        if self._struct is None:
            self._struct = _ffi.make_pass(self.axis, self.taps, self.origin, self.mode, self.uniform, self.size)
        return self._struct


def _f32_tiled_ok(src, dst, spec):
    return (int(spec.uniform) <= 1 and src.dtype == _F32 and dst.dtype == _F32 and src.c_contiguous() and dst.c_contiguous()
            and spec.radius() <= _ffi.FAST_MAX_RADIUS)


def _launch_pass(src, dst, spec, cval, exact, in_offset=0):
    """One sepfilt_correlate1d call: src -> dst (must not overlap)."""
    L = _ffi.lib()
    p, keep = spec.struct()
    acc = _ffi.ACC_F64_EXACT
    if not exact and _f32_tiled_ok(src, dst, spec):
        acc = _ffi.ACC_F32
    scratch = None
    sptr, sbytes = None, 0
    if acc == _ffi.ACC_F64_EXACT and not spec.uniform and spec.size > _ffi.PARAM_TAPS:
        scratch = _array.empty((spec.size,), np.float64, src.device)
        sptr, sbytes = scratch.ptr, spec.size * 8
    ti, to = src.tensor(), dst.tensor()
    rc = L.sepfilt_correlate1d(ti, to, p, int(in_offset), float(cval), acc, sptr, sbytes,
                               _array.current_stream(src.device))
    if rc == _ffi.ERR_UNSUPPORTED and acc == _ffi.ACC_F32:
        rc = L.sepfilt_correlate1d(ti, to, p, int(in_offset), float(cval), _ffi.ACC_F64_EXACT,
                                   sptr, sbytes, _array.current_stream(src.device))
    _ffi.check(rc)
    _ffi.count_launch()
    del keep, scratch


def _copy_cast(src, dst):
    """output[...] = input[...] under the filter's store rules (filters.py:663-664, :790-791)."""
    if dst.size == 0:
        return
    _ffi.check(_ffi.lib().sepfilt_copy_cast(src.tensor(), dst.tensor(), _array.current_stream(src.device)))
    _ffi.count_launch()


def _fused_candidate(inp, out, specs, exact, gradmag=False):
    """Cheap host-side screen for the fused float32 kernel; the library has the last word
    (sepfilt_separable_f32 answers SEPFILT_ERR_UNSUPPORTED and the caller falls back per axis)."""
    if exact or inp.dtype != _F32 or out.dtype != _F32 or inp.ndim not in (2, 3):
        return False
    if any(int(s.uniform) > 1 for s in specs):
        return False                    # minimum / maximum windows: exact per-axis passes
    if not (inp.c_contiguous() and out.c_contiguous()) or inp.size == 0 or inp.may_overlap(out):
        return False
    if len(specs) < 2 and not gradmag:
        return False                    # a single axis is served by the tiled 1-D pass
    return all(s.radius() <= _ffi.FAST_MAX_RADIUS and s.radius() <= inp.shape[s.axis] for s in specs)


def _try_fused(inp, out, specs, cval, dspecs=None, in_offset0=0):
    """One sepfilt_separable_f32 call; False when the library declines the request."""
    structs = [s.struct() for s in specs]
    arr = (_ffi.Pass * len(structs))(*[s[0] for s in structs])
    darr = None
    if dspecs is not None:
        dstructs = [s.struct() for s in dspecs]
        darr = (_ffi.Pass * len(dstructs))(*[s[0] for s in dstructs])
    rc = _ffi.lib().sepfilt_separable_f32(inp.tensor(), out.tensor(), arr, len(structs), darr,
                                          1 if dspecs is not None else 0, int(in_offset0), float(cval),
                                          _array.current_stream(inp.device))
    if rc == _ffi.ERR_UNSUPPORTED:
        return False
    _ffi.check(rc)
    _ffi.count_launch(_ffi.lib().sepfilt_last_launch_count())
    return True


def _run_passes(inp, out, specs, cval, dtype_mode):
    """Apply ``specs`` in order, writing the OUTPUT dtype after every pass like the
    per-axis loops of the reference (filters.py:651-662, :777-789) and scipy do.  One
    launch when the fused float32 kernel applies; otherwise one tiled / exact launch per
    axis, ping-ponging between ``out`` and one temporary instead of the reference's
    temp + copy-back per in-place pass (_filters_core.py:148-155)."""
    if out.size == 0:
        return out
    if inp.dtype.kind == "c" or out.dtype.kind == "c":
        # real taps act on the real and the imaginary component independently (the reference's kernel
        # does the complex multiply-add with a real weight, _filters_core.py:239-312): two real runs
        # over strided component views, every pass rounded to the component dtype like the complex store
        if inp.dtype.kind != "c" or out.dtype.kind != "c":
            raise RuntimeError("output must have complex dtype if either the input or "
                               "weights are complex-valued.")
        cre, cim = _split_cval(cval, True)
        _run_passes(inp.component(0), out.component(0), specs, cre, dtype_mode)
        _run_passes(inp.component(1), out.component(1), specs, cim, dtype_mode)
        return out
    cval, _ = _split_cval(cval, False)
    exact = dtype_mode == "ndimage"
    if not specs:
        if inp.may_overlap(out):
            if inp.ptr == out.ptr and inp.strides == out.strides and inp.dtype == out.dtype:
                return out
            tmp = _array.empty(out.shape, out.dtype, out.device)
            _copy_cast(inp, tmp)
            _copy_cast(tmp, out)
        else:
            _copy_cast(inp, out)
        return out
    if _fused_candidate(inp, out, specs, exact) and _try_fused(inp, out, specs, cval):
        return out
    n = len(specs)
    aliased = inp.may_overlap(out)
    if n == 1 and not aliased:
        _launch_pass(inp, out, specs[0], cval, exact)
        return out
    # buffers for the intermediates: [out, tmp] when out is free, two temporaries otherwise
    tmp_a = _array.empty(out.shape, out.dtype, out.device)
    if aliased:
        tmp_b = _array.empty(out.shape, out.dtype, out.device) if n > 2 else None
        ring = [tmp_a, tmp_b]
    else:
        ring = [tmp_a, out] if n % 2 == 0 else [out, tmp_a]
    src = inp
    for i, spec in enumerate(specs):
        last = i == n - 1
        if last and aliased and n == 1:
            _launch_pass(src, tmp_a, spec, cval, exact)
            _copy_cast(tmp_a, out)
            break
        dst = out if last else ring[i % 2]
        _launch_pass(src, dst, spec, cval, exact)
        src = dst
    return out


def _run_passes_window(inp, out, specs, cval, dtype_mode, in_offset0):
    """Like :func:`_run_passes` for an output that is a WINDOW of the input along axis 0
    (out plane z <-> in plane z + in_offset0; the other extents match).  This is what the
    z-slab sharding calls: halo planes are read but never written.  ``specs`` must be in
    increasing axis order so that the axis-0 pass — the only one that sees the window — runs
    first, which is also the reference's pass order (filters.py:777-789)."""
    if out.size == 0:
        return out
    exact = dtype_mode == "ndimage"
    if inp.may_overlap(out):
        raise RuntimeError("windowed filtering cannot run in place")
    if _fused_candidate(inp, out, specs, exact) and _try_fused(inp, out, specs, cval, in_offset0=in_offset0):
        return out
    specs = list(specs)
    if not specs or specs[0].axis != 0:
        # identity along axis 0: the window is a plain sub-view
        inp = inp.view_axis_window(0, in_offset0, out.shape[0])
        return _run_passes(inp, out, specs, cval, dtype_mode)
    if any(b.axis <= a.axis for a, b in zip(specs, specs[1:])):
        raise RuntimeError("passes must be in increasing axis order")
    n = len(specs)
    tmp = _array.empty(out.shape, out.dtype, out.device) if n > 1 else None
    ring = [tmp, out] if n % 2 == 0 else [out, tmp]
    src = inp
    for i, spec in enumerate(specs):
        dst = out if i == n - 1 else ring[i % 2]
        _launch_pass(src, dst, spec, cval, exact, in_offset=in_offset0 if i == 0 else 0)
        src = dst
    return out


def _gradient_magnitude_window(inp, out, smooth, deriv, cval, dtype_mode, in_offset0):
    """sqrt(sum_a (D_a prod_{b != a} S_b inp)^2) for an output that windows the input along axis 0
    (the z-slab call of the sharded path).  ``smooth`` / ``deriv`` hold one pass per axis in
    increasing axis order.  One fused launch per axis for float32 volumes; otherwise the
    reference's staging (filters.py:1187-1201): one windowed separable filter per axis into a
    temporary, squares / sum / sqrt in the output dtype."""
    if out.size == 0:
        return out
    exact = dtype_mode == "ndimage"
    if _fused_candidate(inp, out, smooth, exact, gradmag=True) and len(smooth) == inp.ndim \
            and _try_fused(inp, out, smooth, cval, dspecs=deriv, in_offset0=in_offset0):
        return out
    acc = out if out.c_contiguous() else _array.empty(out.shape, out.dtype, out.device)
    for i, d in enumerate(deriv):
        specs = [d if s.axis == d.axis else s for s in smooth]
        if i == 0:
            _run_passes_window(inp, acc, specs, cval, dtype_mode, in_offset0)
            _accumulate(acc, acc, 0)
        else:
            tmp = _array.empty(out.shape, out.dtype, out.device)
            _run_passes_window(inp, tmp, specs, cval, dtype_mode, in_offset0)
            _accumulate(acc, tmp, 1)
    _accumulate(acc, acc, 2)
    if acc is not out:
        _copy_cast(acc, out)
    return out


# ----------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------
def correlate1d(input, weights, axis=-1, output=None, mode="reflect", cval=0.0, origin=0, *,
                backend="ndimage", dtype_mode=None):
    """One-dimensional correlation along ``axis`` (reference filters.py:213-283)."""
    if backend != "ndimage":
        raise NotImplementedError("backend={!r} is not available; only 'ndimage'".format(backend))
    _check_dtype_mode(dtype_mode)
    inp = _ingest_input(input)
    w = _host_taps(weights)
    if w.ndim != 1 or w.size < 1:
        raise RuntimeError("incorrect filter size")      # _filters_core.py:52-53
    axis = _normalize_axis_index(axis, inp.ndim)
    origin = _check_origin(origin, w.size)
    mode_code = _check_mode(mode)
    if w.dtype.kind == "c":
        return _correlate1d_complex_taps(inp, w.conj(), axis, output, mode_code, cval, origin, dtype_mode)
    out, _ = _get_output(output, inp)
    _run_passes(inp, out, [_PassSpec(axis, w, origin, mode_code)], cval, dtype_mode)
    return _array.export(out, inp)


def _correlate1d_complex_taps(inp, w, axis, output, mode_code, cval, origin, dtype_mode):
    """Complex weights (already conjugated: correlation conjugates the weights, not the input —
    filters.py:467-469): a linear combination of real passes, the way scipy evaluates it
    (scipy/ndimage/_filters.py ``_complex_via_real_components``)."""
    out, _ = _get_output(output, inp, complex_output=True)
    if out.size == 0:
        return _array.export(out, inp)
    w_re = _PassSpec(axis, np.ascontiguousarray(w.real), origin, mode_code)
    w_im = _PassSpec(axis, np.ascontiguousarray(w.imag), origin, mode_code)
    o_re, o_im = out.component(0), out.component(1)
    if inp.dtype.kind != "c":
        cval, _ = _split_cval(cval, False)
        if inp.may_overlap(out):
            raise RuntimeError("in-place filtering of a real array with complex weights is not possible")
        _run_passes(inp, o_re, [w_re], cval, dtype_mode)
        _run_passes(inp, o_im, [w_im], cval, dtype_mode)
        return _array.export(out, inp)
    cre, cim = _split_cval(cval, True)
    i_re, i_im = inp.component(0), inp.component(1)
    part = o_re.dtype

    def one(src, spec, c):
        tmp = _array.empty(out.shape, part, out.device)
        _run_passes(src, tmp, [spec], c, dtype_mode)
        return tmp

    re = one(i_re, w_re, cre)
    _accumulate(re, one(i_im, w_im, cim), 4)          # real part: re*re - im*im
    im = one(i_re, w_im, cre)
    _accumulate(im, one(i_im, w_re, cim), 3)          # imaginary part: re*im + im*re
    _copy_cast(re, o_re)
    _copy_cast(im, o_im)
    return _array.export(out, inp)


def convolve1d(input, weights, axis=-1, output=None, mode="reflect", cval=0.0, origin=0, *,
               crop=True, backend="ndimage", dtype_mode=None):
    """One-dimensional convolution along ``axis`` (reference filters.py:286-438; the
    flip / origin rule is filters.py:459-466).  ``cval`` is honoured."""
    if backend != "ndimage":
        raise NotImplementedError("backend={!r} is not available; only 'ndimage'".format(backend))
    if not crop:
        raise ValueError("crop=False requires backend='fast_upfirdn'")
    w = _host_taps(weights)
    if w.ndim != 1:
        raise ValueError("expected a 1d weights array")
    if w.size < 1:
        raise RuntimeError("incorrect filter size")
    origin = _check_origin(origin, w.size)
    w = w[::-1]
    origin = -origin
    if not w.size & 1:
        origin -= 1
    if w.dtype.kind == "c":
        w = w.conj()                  # convolution does not conjugate: undo correlate1d's conjugation
    return correlate1d(input, w, axis, output, mode, cval, origin, dtype_mode=dtype_mode)


def uniform_filter1d(input, size, axis=-1, output=None, mode="reflect", cval=0.0, origin=0, *,
                     dtype_mode=None):
    """One-dimensional uniform filter (reference filters.py:549-599, scipy semantics:
    exact window sum divided by ``size``)."""
    _check_dtype_mode(dtype_mode)
    inp = _ingest_input(input)
    size = int(size)
    if size < 1:
        raise RuntimeError("incorrect filter size")
    axis = _normalize_axis_index(axis, inp.ndim)
    out, _ = _get_output(output, inp)
    origin = _check_origin(origin, size)
    mode_code = _check_mode(mode)
    _run_passes(inp, out, [_PassSpec(axis, None, origin, mode_code, uniform=True, size=size)],
                cval, dtype_mode)
    return _array.export(out, inp)


def _filter_axes(ndim, axes):
    if axes is None:
        return list(range(ndim))
    if isinstance(axes, numbers.Integral):
        axes = (axes,)
    axes = [_normalize_axis_index(a, ndim) for a in axes]
    if len(set(axes)) != len(axes):
        raise ValueError("axes must be unique")
    return axes


def uniform_filter(input, size=3, output=None, mode="reflect", cval=0.0, origin=0, *,
                   axes=None, dtype_mode=None):
    """Multi-dimensional uniform filter (reference filters.py:602-665)."""
    _check_dtype_mode(dtype_mode)
    inp = _ingest_input(input)
    out, _ = _get_output(output, inp)
    axes = _filter_axes(inp.ndim, axes)
    sizes = _normalize_sequence(size, len(axes))
    origins = _normalize_sequence(origin, len(axes))
    modes = _normalize_sequence(mode, len(axes))
    specs = []
    for a, sz, og, md in zip(axes, sizes, origins, modes):
        if sz > 1:
            sz = int(sz)
            specs.append(_uniform_pass_cached(int(a), sz, int(_check_origin(og, sz)), int(_check_mode(md))))
    _run_passes(inp, out, specs, cval, dtype_mode)
    return _array.export(out, inp)


@functools.lru_cache(maxsize=256)
def _gaussian_taps_cached(sigma, order, radius):
    """Correlation-orientation taps (reversed kernel, filters.py:717-718), cached: a scale-space loop
    or a benchmark calls the same (sigma, order, radius) thousands of times."""
    taps = _gaussian_kernel1d(sigma, order, radius)[::-1].copy()
    taps.setflags(write=False)
    return taps


def _gaussian_kernel1d(sigma, order, radius):
    """Taps of a Gaussian (or its ``order``-th derivative), convolution orientation
    (reference filters.py:795-825 == scipy's; float64 throughout)."""
    if order < 0:
        raise ValueError("order must be non-negative")
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    if order == 0:
        return phi_x
    # (q phi)' = (q' - x q / sigma^2) phi : advance the polynomial coefficients `order` times
    q = np.zeros(order + 1)
    q[0] = 1
    inv = 1.0 / -sigma2
    for _ in range(order):
        nxt = np.zeros(order + 1)
        for i in range(order + 1):
            low = q[i - 1] * inv if i >= 1 else 0.0
            high = (i + 1) * q[i + 1] if i < order else 0.0
            nxt[i] = low + high
        q = nxt
    return (x[:, None] ** np.arange(order + 1)).dot(q) * phi_x


def _gaussian_spec(axis, sigma, order, mode_code, truncate, radius=None):
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    if radius is not None:
        lw = radius
    if not isinstance(lw, numbers.Integral) or lw < 0:
        raise ValueError("Radius must be a nonnegative integer.")
    return _gaussian_pass_cached(int(axis), sd, int(order), int(lw), int(mode_code))


@functools.lru_cache(maxsize=512)
def _uniform_pass_cached(axis, size, origin, mode_code):
    """Like :func:`_gaussian_pass_cached` for the mean window of ``uniform_filter``."""
    return _PassSpec(axis, None, origin, mode_code, uniform=True, size=size)


@functools.lru_cache(maxsize=512)
def _gaussian_pass_cached(axis, sigma, order, radius, mode_code):
    """One immutable pass spec per (axis, sigma, order, radius, mode): its ctypes struct is built once, so a loop that
    calls the same filter (a benchmark, a sharded step) spends no host time on taps."""
    return _PassSpec(axis, _gaussian_taps_cached(sigma, order, radius), 0, mode_code)


def gaussian_filter1d(input, sigma, axis=-1, order=0, output=None, mode="reflect", cval=0.0,
                      truncate=4.0, *, radius=None, dtype_mode=None):
    """One-dimensional Gaussian filter (reference filters.py:668-722)."""
    _check_dtype_mode(dtype_mode)
    inp = _ingest_input(input)
    if order < 0:
        raise ValueError("order must be non-negative")
    axis = _normalize_axis_index(axis, inp.ndim)
    spec = _gaussian_spec(axis, sigma, order, _check_mode(mode), truncate, radius)
    out, _ = _get_output(output, inp)
    _run_passes(inp, out, [spec], cval, dtype_mode)
    return _array.export(out, inp)


def _gaussian_specs(inp, sigma, order, mode, truncate, radius=None, axes=None):
    axes = _filter_axes(inp.ndim, axes)
    orders = _normalize_sequence(order, len(axes))
    sigmas = _normalize_sequence(sigma, len(axes))
    modes = _normalize_sequence(mode, len(axes))
    radiuses = _normalize_sequence(radius, len(axes))
    specs = []
    for a, sg, od, md, rd in zip(axes, sigmas, orders, modes, radiuses):
        if od < 0:
            raise ValueError("order must be non-negative")
        if sg > 1e-15:
            specs.append(_gaussian_spec(a, sg, od, _check_mode(md), truncate, rd))
    return specs


def gaussian_filter(input, sigma, order=0, output=None, mode="reflect", cval=0.0, truncate=4.0, *,
                    radius=None, axes=None, dtype_mode=None):
    """Multi-dimensional Gaussian filter (reference filters.py:725-792)."""
    _check_dtype_mode(dtype_mode)
    inp = _ingest_input(input)
    out, _ = _get_output(output, inp)
    specs = _gaussian_specs(inp, sigma, order, mode, truncate, radius, axes)
    _run_passes(inp, out, specs, cval, dtype_mode)
    return _array.export(out, inp)


def _derivative_then_smooth(input, axis, output, mode, cval, smooth, dtype_mode):
    _check_dtype_mode(dtype_mode)
    inp = _ingest_input(input)
    ndim = inp.ndim
    if axis < -ndim or axis >= ndim:
        raise ValueError("invalid axis")          # filters.py:871-872, :930-931
    axis = axis % ndim
    out, _ = _get_output(output, inp)
    modes = [_check_mode(m) for m in _normalize_sequence(mode, ndim)]
    specs = [_PassSpec(axis, np.array([-1.0, 0.0, 1.0]), 0, modes[axis])]
    specs += [_PassSpec(a, np.asarray(smooth, np.float64), 0, modes[a]) for a in range(ndim) if a != axis]
    _run_passes(inp, out, specs, cval, dtype_mode)
    return _array.export(out, inp)


def prewitt(input, axis=-1, output=None, mode="reflect", cval=0.0, *, dtype_mode=None):
    """Prewitt filter: [-1, 0, 1] along ``axis`` then [1, 1, 1] along every other axis
    (reference filters.py:828-883)."""
    return _derivative_then_smooth(input, axis, output, mode, cval, [1.0, 1.0, 1.0], dtype_mode)


def sobel(input, axis=-1, output=None, mode="reflect", cval=0.0, *, dtype_mode=None):
    """Sobel filter: [-1, 0, 1] along ``axis`` then [1, 2, 1] along every other axis
    (reference filters.py:889-941)."""
    return _derivative_then_smooth(input, axis, output, mode, cval, [1.0, 2.0, 1.0], dtype_mode)


def _contiguous_same_dtype(arr, dtype):
    if arr.dtype == dtype and arr.c_contiguous():
        return arr
    tmp = _array.empty(arr.shape, dtype, arr.device)
    _copy_cast(arr, tmp)
    return tmp


def _accumulate(acc, a, op):
    _ffi.check(_ffi.lib().sepfilt_gradmag_step(acc.ptr, a.ptr, acc.size, _ffi.DTYPE_CODES[acc.dtype], op,
                                               _array.current_stream(acc.device)))
    _ffi.count_launch()


def _is_component_view(a):
    """True for a real DevArray carved out of a complex owner (DevArray.component)."""
    try:
        return a.dtype.kind != "c" and _array.to_numpy_dtype(a.obj.dtype).kind == "c"
    except Exception:
        return False


def _generic_axis_reduce(input, derivative, output, mode, cval, extra_arguments, extra_keywords, magnitude):
    """Shared body of generic_laplace (sum of per-axis results, filters.py:1011-1038) and
    generic_gradient_magnitude (sqrt of the sum of squares, filters.py:1173-1204); all
    elementwise steps run in the OUTPUT dtype like the reference's cupy ufunc calls."""
    if extra_keywords is None:
        extra_keywords = {}
    inp = _ingest_input(input)
    out, _ = _get_output(output, inp)
    if inp.dtype.kind == "c":
        if magnitude:
            raise NotImplementedError("gradient magnitude of a complex array is not supported")
        # a sum of per-axis real-tap filters is linear: filter the components
        cre, cim = _split_cval(cval, True)
        for k, c in ((0, cre), (1, cim)):
            _generic_axis_reduce(inp.component(k), derivative, out.component(k), mode, c,
                                 extra_arguments, extra_keywords, magnitude)
        return _array.export(out, inp)
    ndim = inp.ndim
    if ndim == 0 or out.size == 0:
        if out.size:
            _copy_cast(inp, out)
        return _array.export(out, inp)
    modes = _normalize_sequence(mode, ndim)
    # the derivative callable is handed an array OBJECT: a real / imaginary component view has none of its own (its
    # owner is the complex array, and a size-1 view even counts as contiguous), so it accumulates in a temporary
    acc = out if out.c_contiguous() and not _is_component_view(out) else _array.empty(out.shape, out.dtype, out.device)
    derivative(input, 0, acc.obj, modes[0], cval, *extra_arguments, **extra_keywords)
    if magnitude:
        _accumulate(acc, acc, 0)
    for ax in range(1, ndim):
        tmp = derivative(input, ax, out.dtype, modes[ax], cval, *extra_arguments, **extra_keywords)
        tmp = _contiguous_same_dtype(_array.ingest(tmp, "derivative result"), out.dtype)
        _accumulate(acc, tmp, 1 if magnitude else 3)
    if magnitude:
        _accumulate(acc, acc, 2)
    if acc is not out:
        _copy_cast(acc, out)
    return _array.export(out, inp)


def generic_laplace(input, derivative2, output=None, mode="reflect", cval=0.0,
                    extra_arguments=(), extra_keywords=None):
    """N-d Laplace filter from a user second-derivative callable (reference filters.py:963-1038)."""
    return _generic_axis_reduce(input, derivative2, output, mode, cval, extra_arguments,
                                extra_keywords, magnitude=False)


def laplace(input, output=None, mode="reflect", cval=0.0, *, dtype_mode=None):
    """N-d Laplace filter from [1, -2, 1] second differences (reference filters.py:1041-1074)."""

    def derivative2(input, axis, output, mode, cval):
        return correlate1d(input, [1.0, -2.0, 1.0], axis, output, mode, cval, 0, dtype_mode=dtype_mode)

    return generic_laplace(input, derivative2, output, mode, cval)


def gaussian_laplace(input, sigma, output=None, mode="reflect", cval=0.0, **kwargs):
    """N-d Laplace filter from Gaussian second derivatives (reference filters.py:1077-1122)."""
    ndim = _ingest_input(input).ndim

    def derivative2(input, axis, output, mode, cval, sigma, **kwargs):
        order = [0] * ndim
        order[axis] = 2
        return gaussian_filter(input, sigma, order, output, mode, cval, **kwargs)

    return generic_laplace(input, derivative2, output, mode, cval,
                           extra_arguments=(sigma,), extra_keywords=kwargs)


def generic_gradient_magnitude(input, derivative, output=None, mode="reflect", cval=0.0,
                               extra_arguments=(), extra_keywords=None):
    """Gradient magnitude from a user derivative callable (reference filters.py:1125-1204)."""
    return _generic_axis_reduce(input, derivative, output, mode, cval, extra_arguments,
                                extra_keywords, magnitude=True)


def gaussian_gradient_magnitude(input, sigma, output=None, mode="reflect", cval=0.0, **kwargs):
    """Gradient magnitude from Gaussian first derivatives (reference filters.py:1207-1252).

    float32 volumes take one fused launch (smoothing + derivative taps per axis, squares,
    sum and sqrt in the kernel epilogue); everything else follows the reference staging
    through :func:`generic_gradient_magnitude`."""
    inp = _ingest_input(input)
    ndim = inp.ndim
    dtype_mode = _check_dtype_mode(kwargs.get("dtype_mode"))
    truncate = kwargs.get("truncate", 4.0)
    # a mode SEQUENCE means "mode[a] for every pass of the derivative along axis a" (filters.py:1175-1201:
    # derivative(input, axis, output, modes[axis], ...)), not one mode per filtered axis: only a single mode
    # maps onto the fused launches, whose smoothing and derivative passes share one mode per filtered axis
    modes = _normalize_sequence(mode, ndim)
    single_mode = all(m == modes[0] for m in modes)
    if set(kwargs) <= {"dtype_mode", "truncate"} and dtype_mode != "ndimage" and inp.dtype == _F32 \
            and ndim in (2, 3) and inp.size and single_mode:
        mode = modes[0]
        out, _ = _get_output(output, inp)
        if out.dtype == _F32:
            smooth = _gaussian_specs(inp, sigma, 0, mode, truncate)
            deriv = _gaussian_specs(inp, sigma, 1, mode, truncate)
            if len(smooth) == ndim and len(deriv) == ndim and _fused_candidate(inp, out, smooth, False, gradmag=True) \
                    and _try_fused(inp, out, smooth, cval, dspecs=deriv):
                return _array.export(out, inp)
        output = out.obj

    def derivative(input, axis, output, mode, cval, sigma, **kwargs):
        order = [0] * ndim
        order[axis] = 1
        return gaussian_filter(input, sigma, order, output, mode, cval, **kwargs)

    return generic_gradient_magnitude(input, derivative, output, mode, cval,
                                      extra_arguments=(sigma,), extra_keywords=kwargs)


# ----------------------------------------------------------------------------
# minimum / maximum filters: the same per-axis skeleton with (min, max)
# instead of (x, +)
# ----------------------------------------------------------------------------
_MIN, _MAX = 2, 3     # sepfilt_pass.uniform codes


def _min_or_max_1d(input, size, axis, output, mode, cval, origin, kind):
    """reference filters.py:1475-1508 (one generated kernel launch with a 1-D all-ones footprint)."""
    inp = _ingest_input(input)
    if inp.dtype.kind == "c":
        raise TypeError("Complex type not supported")
    size = int(size)
    if size < 1:
        raise RuntimeError("incorrect filter size")
    axis = _normalize_axis_index(axis, inp.ndim)
    origin = _check_origin(origin, size)
    mode_code = _check_mode(mode)
    _check_minmax_cval(cval)
    out, _ = _get_output(output, inp)
    _run_passes(inp, out, [_PassSpec(axis, None, origin, mode_code, uniform=kind, size=size)], cval, "ndimage")
    return _array.export(out, inp)


def _check_minmax_cval(cval):
    if isinstance(cval, float) and cval != cval:
        raise NotImplementedError("NaN cval is unsupported")         # filters.py:1382-1383


def minimum_filter1d(input, size, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    """Minimum filter along one axis (reference filters.py:1422-1446)."""
    return _min_or_max_1d(input, size, axis, output, mode, cval, origin, _MIN)


def maximum_filter1d(input, size, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    """Maximum filter along one axis (reference filters.py:1449-1472)."""
    return _min_or_max_1d(input, size, axis, output, mode, cval, origin, _MAX)


def _min_or_max_filter(input, size, footprint, output, mode, cval, origin, kind, axes=None):
    """reference filters.py:1373-1419: a size (or an all-True footprint, as in scipy) is separable and
    runs as one 1-D pass per axis with size > 1 (_filters_core._run_1d_filters, _filters_core.py:79-109);
    a general footprint needs the N-d kernel, which is outside this path."""
    inp = _ingest_input(input)
    if inp.dtype.kind == "c":
        raise TypeError("Complex type not supported")
    axes = _filter_axes(inp.ndim, axes)
    if size is None and footprint is None:
        raise RuntimeError("no footprint or filter size provided")
    if footprint is not None:
        fp = np.asarray(_array.host_weights(footprint) if _array.is_device_array(footprint) else footprint).astype(bool)
        if fp.ndim != len(axes):
            raise RuntimeError("footprint array has incorrect shape.")
        if not fp.any():
            raise ValueError("All-zero footprint is not supported.")
        if not fp.all():
            raise NotImplementedError("only separable minimum / maximum filters (a size or an all-True "
                                      "footprint) are on this path")
        sizes = list(fp.shape)
    else:
        sizes = [int(s) for s in _normalize_sequence(size, len(axes))]
    _check_minmax_cval(cval)
    origins = _normalize_sequence(origin, len(axes))
    modes = _normalize_sequence(mode, len(axes))
    out, _ = _get_output(output, inp)
    specs = []
    for a, sz, og, md in zip(axes, sizes, origins, modes):
        if sz < 1:
            raise RuntimeError("incorrect filter size")
        mode_code = _check_mode(md)
        og = _check_origin(og, sz)
        if sz > 1:
            specs.append(_PassSpec(a, None, og, mode_code, uniform=kind, size=sz))
    _run_passes(inp, out, specs, cval, "ndimage")
    return _array.export(out, inp)


def minimum_filter(input, size=None, footprint=None, output=None, mode="reflect", cval=0.0, origin=0, *,
                   axes=None):
    """Multi-dimensional minimum filter (reference filters.py:1296-1332), separable case."""
    return _min_or_max_filter(input, size, footprint, output, mode, cval, origin, _MIN, axes)


def maximum_filter(input, size=None, footprint=None, output=None, mode="reflect", cval=0.0, origin=0, *,
                   axes=None):
    """Multi-dimensional maximum filter (reference filters.py:1335-1370), separable case."""
    return _min_or_max_filter(input, size, footprint, output, mode, cval, origin, _MAX, axes)


# ----------------------------------------------------------------------------
# dense N-d correlation for small kernels
# ----------------------------------------------------------------------------
def _correlate_or_convolve(input, weights, output, mode, cval, origin, convolution):
    """reference filters.py:441-495 for N-d weights: validation (_filters_core._check_nd_args, :63-76),
    the convolution flip / origin rule (:459-466), one launch (sepfilt_correlate_nd) in scipy's arithmetic."""
    import ctypes
    inp = _ingest_input(input)
    w = _host_taps(weights)
    if inp.dtype.kind == "c" or w.dtype.kind == "c":
        raise NotImplementedError("complex-valued N-d correlation is not on this path (1-D weights are: correlate1d)")
    wshape = [s for s in w.shape if s != 0]
    if w.ndim != inp.ndim or len(wshape) != inp.ndim:
        raise RuntimeError("filter weights array has incorrect shape.")
    if not isinstance(mode, str) and hasattr(mode, "__iter__"):
        raise RuntimeError("A sequence of modes is not supported")
    mode_code = _check_mode(mode)
    origins = [int(o) for o in _normalize_sequence(origin, inp.ndim)]
    if convolution:
        w = w[tuple([slice(None, None, -1)] * w.ndim)]
        for i in range(len(origins)):
            origins[i] = -origins[i]
            if not w.shape[i] & 1:
                origins[i] -= 1
    for o, width in zip(origins, wshape):
        _check_origin(o, width)
    cval, _ = _split_cval(cval, False)
    out, _ = _get_output(output, inp)
    if out.size == 0:
        return _array.export(out, inp)
    w = np.ascontiguousarray(w, dtype=np.float64)
    dst = out
    if inp.may_overlap(out):
        dst = _array.empty(out.shape, out.dtype, out.device)     # _filters_core.py:148-155
    scratch, sptr, sbytes = None, None, 0
    if w.size > _ffi.PARAM_TAPS:
        if w.size > _ffi.MAX_TAPS:
            raise NotImplementedError("more than %d filter weights: use an FFT-based convolution" % _ffi.MAX_TAPS)
        scratch = _array.empty((w.size,), np.float64, inp.device)
        sptr, sbytes = scratch.ptr, w.size * 8
    i32 = ctypes.c_int32 * inp.ndim
    rc = _ffi.lib().sepfilt_correlate_nd(inp.tensor(), dst.tensor(), w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                         i32(*wshape), i32(*origins), mode_code, float(cval), sptr, sbytes,
                                         _array.current_stream(inp.device))
    _ffi.check(rc)
    _ffi.count_launch()
    if dst is not out:
        _copy_cast(dst, out)
    del scratch
    return _array.export(out, inp)


def correlate(input, weights, output=None, mode="reflect", cval=0.0, origin=0, *, dtype_mode=None):
    """Multi-dimensional correlation with a small dense kernel (reference filters.py:65-133)."""
    _check_dtype_mode(dtype_mode)
    return _correlate_or_convolve(input, weights, output, mode, cval, origin, False)


def convolve(input, weights, output=None, mode="reflect", cval=0.0, origin=0, *, dtype_mode=None):
    """Multi-dimensional convolution with a small dense kernel (reference filters.py:136-210)."""
    _check_dtype_mode(dtype_mode)
    return _correlate_or_convolve(input, weights, output, mode, cval, origin, True)
