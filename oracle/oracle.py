"""CPU oracle for the separable-correlation hot path (numpy in, numpy out).

TEST INFRASTRUCTURE ONLY — the header of ``sepfilt_oracle.c`` says why.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; nothing under ``cupyimg_b200/``
does (``tests/test_boundary.py`` greps for it).

Each function restates one reference function (paths relative to
/root/reference/cupyimg/scipy/ndimage/) on top of the C line filter in
``sepfilt_oracle.c``:

    correlate1d                  filters.py:213-283 + _filters_core.py:51-60
    convolve1d                   filters.py:286-438, flip/origin rule :459-466
    uniform_filter1d / _filter   filters.py:549-665  (scipy's running-sum semantics)
    gaussian_kernel1d            filters.py:795-825
    gaussian_filter1d / _filter  filters.py:668-792
    prewitt / sobel              filters.py:828-941
    generic_laplace / laplace / gaussian_laplace   filters.py:963-1122
    generic_gradient_magnitude   filters.py:1125-1204
    gaussian_gradient_magnitude  filters.py:1207-1252
    convolve_separable           ../../_misc.py:39-77

Pinned (tests/test_oracle.py) against the reference test-suite's literal known
answers (tests/golden/reference_kats.json), against golden vectors produced by
executing the reference's own generated kernel source on the CPU
(tests/golden/make_reference_kernel_vectors.py) and against scipy.ndimage, the
oracle the reference's tests themselves use.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsepfilt_oracle.so")

MODES = {"reflect": 0, "grid-mirror": 0, "constant": 1, "grid-constant": 1,
         "nearest": 2, "mirror": 3, "wrap": 4, "grid-wrap": 4}
_DT = {np.dtype(k): v for k, v in {
    "int8": 0, "uint8": 1, "int16": 2, "uint16": 3, "int32": 4, "uint32": 5,
    "int64": 6, "uint64": 7, "float32": 8, "float64": 9, "bool": 10}.items()}


def build(force=False):
    """Compile sepfilt_oracle.c with gcc (no-op when the .so is newer than the source)."""
    src = os.path.join(_HERE, "sepfilt_oracle.c")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(src)):
        return _SO
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64, dbl, vp, ci = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
        L.oracle_correlate1d.argtypes = [vp, ci, vp, ci, i64, i64, i64, i64,
                                         ctypes.POINTER(dbl), ci, ci, ci, dbl, i64, ci]
        L.oracle_correlate1d.restype = ci
        L.oracle_correlate1d_lines.argtypes = L.oracle_correlate1d.argtypes + [i64, i64]
        L.oracle_correlate1d_lines.restype = ci
        L.oracle_remap.argtypes = [ci, i64, i64]
        L.oracle_remap.restype = i64
        L.oracle_gradmag_step.argtypes = [vp, vp, i64, ci, ci]
        L.oracle_gradmag_step.restype = ci
        L.oracle_copy_cast.argtypes = [vp, ci, vp, ci, i64]
        L.oracle_copy_cast.restype = ci
        _lib = L
    return _lib


def remap(mode, ix, n):
    return int(lib().oracle_remap(MODES[mode], int(ix), int(n)))


def _axis(axis, ndim):
    if not -ndim <= axis < ndim:
        raise ValueError("invalid axis")
    return axis % ndim


def _get_output(output, input):
    if output is None:
        return np.zeros(input.shape, input.dtype)
    if isinstance(output, np.ndarray):
        if output.shape != input.shape:
            raise RuntimeError("output shape not correct")
        return output
    return np.zeros(input.shape, np.dtype(output))


def _seq(v, n):
    if hasattr(v, "__iter__") and not isinstance(v, str):
        v = list(v)
        if len(v) != n:
            raise RuntimeError("sequence argument must have length equal to input rank")
        return v
    return [v] * n


THREADS = 1   # host threads per pass (bench.py's cpu_baseline raises it; results are identical)


def _line_pass(input, output, axis, w, K, origin, mode, cval, uniform=0, in_offset=0):
    """One C call; ``output`` may alias ``input`` (the C code works line by line but a
    strided alias would still be unsafe, so alias -> temporary, like _filters_core.py:148-155)."""
    if mode not in MODES:
        raise RuntimeError("boundary mode not supported")
    x = np.ascontiguousarray(input)
    if x.dtype not in _DT or output.dtype not in _DT:
        raise RuntimeError("data type not supported")
    y = output
    tmp = None
    if (not y.flags.c_contiguous) or np.shares_memory(x, y):
        tmp = np.empty(y.shape, y.dtype)
        y = tmp
    n_in, n_out = x.shape[axis], y.shape[axis]
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, np.float64)
        wp = w.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    args = (x.ctypes.data, _DT[x.dtype], y.ctypes.data, _DT[y.dtype], outer, n_in, n_out, inner, wp,
            int(K), int(origin), MODES[mode], float(cval), int(in_offset), int(uniform))
    nlines = outer * inner
    nthreads = max(1, min(THREADS, nlines // 64))
    if nthreads == 1:
        rc = lib().oracle_correlate1d(*args)
    else:
        import threading
        rcs = [0] * nthreads

        def work(i):   # ctypes releases the GIL during the C call
            rcs[i] = lib().oracle_correlate1d_lines(*args, i * nlines // nthreads, (i + 1) * nlines // nthreads)

        ts = [threading.Thread(target=work, args=(i,)) for i in range(nthreads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        rc = min(rcs)
    if rc != 0:
        raise ValueError("invalid origin")
    if tmp is not None:
        output[...] = tmp
    return output


def _complex_output(output, input):
    """_util._get_output's complex rule (_util.py:52-58, :66-75)."""
    if output is None:
        return np.zeros(input.shape, np.promote_types(input.dtype, np.complex64))
    if isinstance(output, np.ndarray):
        if output.shape != input.shape:
            raise RuntimeError("output shape not correct")
        if output.dtype.kind != "c":
            raise RuntimeError("output must have complex dtype")
        return output
    if np.dtype(output).kind != "c":
        raise RuntimeError("output must have complex dtype")
    return np.zeros(input.shape, np.dtype(output))


def _correlate1d_complex(input, weights, axis, output, mode, cval, origin):
    """Complex input and / or weights.  The reference conjugates the weights (filters.py:467-469) and
    lets its kernel do the complex multiply-add; restated here the way scipy evaluates the same sum,
    as a linear combination of real passes (scipy/ndimage/_filters.py, _complex_via_real_components)."""
    if weights.dtype.kind == "c":
        weights = weights.conj().astype(np.complex128)
    out = _complex_output(output, input)
    kw = dict(axis=axis, mode=mode, origin=origin)
    if input.dtype.kind == "c" and weights.dtype.kind == "c":
        correlate1d(input.real, weights.real, output=out.real, cval=np.real(cval), **kw)
        out.real -= correlate1d(input.imag, weights.imag, cval=np.imag(cval), **kw)
        correlate1d(input.real, weights.imag, output=out.imag, cval=np.real(cval), **kw)
        out.imag += correlate1d(input.imag, weights.real, cval=np.imag(cval), **kw)
    elif input.dtype.kind == "c":
        correlate1d(input.real, weights, output=out.real, cval=np.real(cval), **kw)
        correlate1d(input.imag, weights, output=out.imag, cval=np.imag(cval), **kw)
    else:
        if np.iscomplexobj(cval):
            raise ValueError("Cannot provide a complex-valued cval when the input is real.")
        correlate1d(input, weights.real, output=out.real, cval=cval, **kw)
        correlate1d(input, weights.imag, output=out.imag, cval=cval, **kw)
    return out


def correlate1d(input, weights, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    input = np.asarray(input)
    weights = np.asarray(weights)
    if input.dtype.kind == "c" or weights.dtype.kind == "c":
        return _correlate1d_complex(input, weights, axis, output, mode, cval, origin)
    weights = np.asarray(weights, np.float64)
    if weights.ndim != 1 or weights.size < 1:
        raise RuntimeError("no filter weights given")
    axis = _axis(axis, input.ndim)
    output = _get_output(output, input)
    if input.size == 0:
        return output
    return _line_pass(input, output, axis, weights, weights.size, origin, mode, cval)


def convolve1d(input, weights, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    weights = np.asarray(weights)
    weights = (weights.conj() if weights.dtype.kind == "c" else weights.astype(np.float64))[::-1]
    origin = -origin
    if weights.size and not weights.size & 1:
        origin -= 1
    return correlate1d(input, weights, axis, output, mode, cval, origin)


def uniform_filter1d(input, size, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    input = np.asarray(input)
    if size < 1:
        raise RuntimeError("incorrect filter size")
    axis = _axis(axis, input.ndim)
    output = _get_output(output, input)
    if size // 2 + origin < 0 or size // 2 + origin >= size:
        raise ValueError("invalid origin")
    if input.size == 0:
        return output
    return _line_pass(input, output, axis, None, size, origin, mode, cval, uniform=1)


def uniform_filter(input, size=3, output=None, mode="reflect", cval=0.0, origin=0):
    input = np.asarray(input)
    output = _get_output(output, input)
    sizes, origins, modes = (_seq(v, input.ndim) for v in (size, origin, mode))
    axes = [a for a in range(input.ndim) if sizes[a] > 1]
    if not axes:
        lib().oracle_copy_cast(np.ascontiguousarray(input).ctypes.data, _DT[input.dtype],
                               output.ctypes.data, _DT[output.dtype], input.size)
        return output
    for a in axes:
        uniform_filter1d(input, int(sizes[a]), a, output, modes[a], cval, origins[a])
        input = output
    return output


def gaussian_kernel1d(sigma, order, radius):
    """Taps of the (derivative of a) Gaussian, convolution orientation (filters.py:795-825)."""
    if order < 0:
        raise ValueError("order must be non-negative")
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    if order == 0:
        return phi
    # d/dx [q(x) phi(x)] = (q'(x) - x q(x) / sigma^2) phi(x), iterated on the coefficients of q
    q = np.zeros(order + 1)
    q[0] = 1
    for _ in range(order):
        nq = np.zeros(order + 1)
        for i in range(order + 1):
            lo = q[i - 1] * (1.0 / -sigma2) if i >= 1 else 0.0
            hi = (i + 1) * q[i + 1] if i + 1 <= order else 0.0
            nq[i] = lo + hi
        q = nq
    powers = np.arange(order + 1)
    return (x[:, None] ** powers).dot(q) * phi


def gaussian_filter1d(input, sigma, axis=-1, order=0, output=None, mode="reflect",
                      cval=0.0, truncate=4.0):
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    weights = gaussian_kernel1d(sd, order, lw)[::-1]
    return correlate1d(input, weights, axis, output, mode, cval, 0)


def gaussian_filter(input, sigma, order=0, output=None, mode="reflect", cval=0.0, truncate=4.0):
    input = np.asarray(input)
    output = _get_output(output, input)
    orders, sigmas, modes = (_seq(v, input.ndim) for v in (order, sigma, mode))
    axes = [a for a in range(input.ndim) if sigmas[a] > 1e-15]
    if not axes:
        lib().oracle_copy_cast(np.ascontiguousarray(input).ctypes.data, _DT[input.dtype],
                               output.ctypes.data, _DT[output.dtype], input.size)
        return output
    for a in axes:
        gaussian_filter1d(input, sigmas[a], a, orders[a], output, modes[a], cval, truncate)
        input = output
    return output


def _edge(input, axis, output, mode, cval, smooth):
    input = np.asarray(input)
    axis = _axis(axis, input.ndim)
    output = _get_output(output, input)
    modes = _seq(mode, input.ndim)
    correlate1d(input, [-1, 0, 1], axis, output, modes[axis], cval, 0)
    for a in range(input.ndim):
        if a != axis:
            correlate1d(output, smooth, a, output, modes[a], cval, 0)
    return output


def prewitt(input, axis=-1, output=None, mode="reflect", cval=0.0):
    return _edge(input, axis, output, mode, cval, [1, 1, 1])


def sobel(input, axis=-1, output=None, mode="reflect", cval=0.0):
    return _edge(input, axis, output, mode, cval, [1, 2, 1])


def generic_gradient_magnitude(input, derivative, output=None, mode="reflect", cval=0.0,
                               extra_arguments=(), extra_keywords=None):
    extra_keywords = extra_keywords or {}
    input = np.asarray(input)
    output = _get_output(output, input)
    if input.ndim == 0:
        output[...] = input
        return output
    modes = _seq(mode, input.ndim)
    t = _DT[output.dtype]
    derivative(input, 0, output, modes[0], cval, *extra_arguments, **extra_keywords)
    acc = np.ascontiguousarray(output)
    lib().oracle_gradmag_step(acc.ctypes.data, acc.ctypes.data, acc.size, t, 0)
    for a in range(1, input.ndim):
        tmp = derivative(input, a, output.dtype, modes[a], cval, *extra_arguments, **extra_keywords)
        tmp = np.ascontiguousarray(tmp)
        lib().oracle_gradmag_step(acc.ctypes.data, tmp.ctypes.data, acc.size, t, 1)
    lib().oracle_gradmag_step(acc.ctypes.data, acc.ctypes.data, acc.size, t, 2)
    if acc is not output:
        output[...] = acc
    return output


def gaussian_gradient_magnitude(input, sigma, output=None, mode="reflect", cval=0.0, **kwargs):
    input = np.asarray(input)

    def derivative(input, axis, output, mode, cval, sigma, **kw):
        order = [0] * input.ndim
        order[axis] = 1
        return gaussian_filter(input, sigma, order, output, mode, cval, **kw)

    return generic_gradient_magnitude(input, derivative, output, mode, cval,
                                      extra_arguments=(sigma,), extra_keywords=kwargs)


def convolve_separable(x, w, axes=None, **kwargs):
    x = np.asarray(x)
    axes = tuple(range(x.ndim)) if axes is None else tuple(axes)
    if any(a < -x.ndim or a > x.ndim - 1 for a in axes):
        raise ValueError("axis out of range")
    if isinstance(w, np.ndarray) and w.ndim == 1:
        w = [w] * len(axes)
    elif len(w) != len(axes):
        raise ValueError("user should supply one filter per axis")
    for a, w0 in zip(axes, w):
        x = convolve1d(x, w0, axis=a, **kwargs)
    return x


def generic_laplace(input, derivative2, output=None, mode="reflect", cval=0.0,
                    extra_arguments=(), extra_keywords=None):
    """filters.py:963-1038: sum over axes of a second-derivative pass, added in the output dtype."""
    extra_keywords = extra_keywords or {}
    input = np.asarray(input)
    output = _get_output(output, input)
    if input.ndim == 0:
        output[...] = input
        return output
    modes = _seq(mode, input.ndim)
    t = _DT[output.dtype]
    derivative2(input, 0, output, modes[0], cval, *extra_arguments, **extra_keywords)
    acc = np.ascontiguousarray(output)
    for a in range(1, input.ndim):
        tmp = derivative2(input, a, output.dtype, modes[a], cval, *extra_arguments, **extra_keywords)
        tmp = np.ascontiguousarray(tmp)
        lib().oracle_gradmag_step(acc.ctypes.data, tmp.ctypes.data, acc.size, t, 3)
    if acc is not output:
        output[...] = acc
    return output


def laplace(input, output=None, mode="reflect", cval=0.0):
    def derivative2(input, axis, output, mode, cval):
        return correlate1d(input, [1, -2, 1], axis, output, mode, cval, 0)

    return generic_laplace(input, derivative2, output, mode, cval)


def gaussian_laplace(input, sigma, output=None, mode="reflect", cval=0.0, **kwargs):
    input = np.asarray(input)

    def derivative2(input, axis, output, mode, cval, sigma, **kw):
        order = [0] * input.ndim
        order[axis] = 2
        return gaussian_filter(input, sigma, order, output, mode, cval, **kw)

    return generic_laplace(input, derivative2, output, mode, cval,
                           extra_arguments=(sigma,), extra_keywords=kwargs)


# ---- minimum / maximum filters ---------------------------------------
def _min_or_max_1d(input, size, axis, output, mode, cval, origin, func):
    """The reference's generated min / max kernel with a 1-D all-ones footprint (filters.py:1475-1557):
    the window [i - size//2 - origin, ... + size) through the boundary rule (_util.py:170-228), compared
    as double ("value = min(cast<double>(x), value)"), stored with the C cast."""
    input = np.asarray(input)
    size = int(size)
    if size < 1:
        raise RuntimeError("incorrect filter size")
    axis = _axis(axis, input.ndim)
    if size // 2 + origin < 0 or size // 2 + origin >= size:
        raise ValueError("invalid origin")
    if mode not in MODES:
        raise RuntimeError("boundary mode not supported")
    output = _get_output(output, input)
    if input.size == 0:
        return output
    n = input.shape[axis]
    x = np.moveaxis(input.astype(np.float64), axis, -1)
    ext = np.concatenate([x, np.full(x.shape[:-1] + (1,), float(cval))], axis=-1)   # index n = cval
    idx = np.empty((n, size), np.int64)
    for i in range(n):
        for k in range(size):
            m = remap(mode, i - size // 2 - origin + k, n)
            idx[i, k] = n if m < 0 else m
    win = ext[..., idx]                                                           # (..., n, size)
    red = win.min(axis=-1) if func == "min" else win.max(axis=-1)
    red = np.ascontiguousarray(np.moveaxis(red, -1, axis))
    res = np.empty(input.shape, output.dtype)
    lib().oracle_copy_cast(red.ctypes.data, _DT[np.dtype(np.float64)], res.ctypes.data, _DT[res.dtype], red.size)
    output[...] = res
    return output


def minimum_filter1d(input, size, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    return _min_or_max_1d(input, size, axis, output, mode, cval, origin, "min")


def maximum_filter1d(input, size, axis=-1, output=None, mode="reflect", cval=0.0, origin=0):
    return _min_or_max_1d(input, size, axis, output, mode, cval, origin, "max")


def _min_or_max_filter(input, size, output, mode, cval, origin, func):
    """Separable case of filters.py:1373-1419 through _filters_core._run_1d_filters (:79-109): one 1-D
    pass per axis with size > 1, every pass stored in the output dtype."""
    input = np.asarray(input)
    output = _get_output(output, input)
    sizes, origins, modes = (_seq(v, input.ndim) for v in (size, origin, mode))
    axes = [a for a in range(input.ndim) if sizes[a] > 1]
    if not axes:
        lib().oracle_copy_cast(np.ascontiguousarray(input).ctypes.data, _DT[input.dtype],
                               output.ctypes.data, _DT[output.dtype], input.size)
        return output
    for a in axes:
        _min_or_max_1d(input, int(sizes[a]), a, output, modes[a], cval, origins[a], func)
        input = output.copy()
    return output


def minimum_filter(input, size=3, output=None, mode="reflect", cval=0.0, origin=0):
    return _min_or_max_filter(input, size, output, mode, cval, origin, "min")


def maximum_filter(input, size=3, output=None, mode="reflect", cval=0.0, origin=0):
    return _min_or_max_filter(input, size, output, mode, cval, origin, "max")


# ---- dense N-d correlation -------------------------------------------
def _correlate_nd(input, weights, output, mode, cval, origin, convolution):
    """The reference's generated N-d correlate kernel (filters.py:441-495, body _filters_core.py:239-312)
    in scipy's NI_Correlate arithmetic: float64 `tmp += x * w` over the taps with |w| > DBL_EPSILON in C order
    of the weights, boundary extension per axis (_util.py:170-228), C-cast store."""
    input = np.asarray(input)
    w = np.asarray(weights, np.float64)
    if w.ndim != input.ndim or any(s == 0 for s in w.shape):
        raise RuntimeError("filter weights array has incorrect shape.")
    if mode not in MODES:
        raise RuntimeError("boundary mode not supported")
    origins = [int(o) for o in _seq(origin, input.ndim)]
    if convolution:
        w = w[tuple([slice(None, None, -1)] * w.ndim)]
        for i in range(len(origins)):
            origins[i] = -origins[i]
            if not w.shape[i] & 1:
                origins[i] -= 1
    for o, width in zip(origins, w.shape):
        if width // 2 + o < 0 or width // 2 + o >= width:
            raise ValueError("invalid origin")
    output = _get_output(output, input)
    if input.size == 0:
        return output
    x = input.astype(np.float64)
    acc = np.zeros(input.shape, np.float64)
    # per axis and tap offset: index vector through the boundary rule (-1 -> cval)
    maps = []
    for d, n in enumerate(input.shape):
        before = w.shape[d] // 2 + origins[d]
        maps.append([np.array([remap(mode, i + k - before, n) for i in range(n)], np.int64) for k in range(w.shape[d])])
    for kk in np.ndindex(*w.shape):
        wk = w[kk]
        if not abs(wk) > np.finfo(np.float64).eps:
            continue
        v = x
        outside = np.zeros(input.shape, bool)
        for d, k in enumerate(kk):
            m = maps[d][k]
            v = np.take(v, np.where(m < 0, 0, m), axis=d)
            shape = [1] * input.ndim
            shape[d] = -1
            outside |= (m < 0).reshape(shape)
        v = np.where(outside, float(cval), v)
        acc = acc + v * wk
    res = np.empty(input.shape, output.dtype)
    acc = np.ascontiguousarray(acc)
    lib().oracle_copy_cast(acc.ctypes.data, _DT[np.dtype(np.float64)], res.ctypes.data, _DT[res.dtype], acc.size)
    output[...] = res
    return output


def correlate(input, weights, output=None, mode="reflect", cval=0.0, origin=0):
    return _correlate_nd(input, weights, output, mode, cval, origin, False)


def convolve(input, weights, output=None, mode="reflect", cval=0.0, origin=0):
    return _correlate_nd(input, weights, output, mode, cval, origin, True)
