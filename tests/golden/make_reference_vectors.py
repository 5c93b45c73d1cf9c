#!/usr/bin/env python
"""Generate golden vectors by EXECUTING the reference (mritools/cupyimg) on the CPU.

The reference is pure Python on top of CuPy and cannot run without CuPy + a GPU.  Its GPU
kernels, however, are plain C++ strings produced by its own generator
(cupyimg/scipy/ndimage/_filters_core.py:190-348).  This script

  1. installs a shim `cupy` module backed by numpy (arrays, dtype helpers, memoize, ...);
  2. loads the reference's own `_util.py`, `_filters_core.py` and `filters.py` from
     /root/reference UNMODIFIED (nothing is copied into this repository);
  3. implements `cupy.ElementwiseKernel` by compiling the kernel source the reference generates
     (its preamble + operation strings, taken at run time) with g++ into a small shared library
     whose loop body is exactly the generated per-element code, with CuPy's CArray / `i` /
     `_raw_y` conventions;
  4. calls the reference's public functions (correlate1d, convolve1d, gaussian_filter, ...) on
     seeded numpy inputs and stores inputs + outputs in tests/golden/reference_vectors.npz.

Run it in the build container (needs /root/reference and g++):
    python tests/golden/make_reference_vectors.py
tests/test_oracle.py::test_reference_executed_vectors checks the oracle against the file; the
GPU tests check the CUDA path against the oracle on the same cases.
"""
import ctypes
import hashlib
import importlib.util
import os
import re
import subprocess
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "reference_vectors.npz")

CTYPES = {"int8": "signed char", "uint8": "unsigned char", "int16": "short", "uint16": "unsigned short",
          "int32": "int", "uint32": "unsigned int", "int64": "long long", "uint64": "unsigned long long",
          "float32": "float", "float64": "double", "bool": "bool"}

HARNESS = r"""
#include <algorithm>
#include <cstddef>
#include <cmath>
#include <type_traits>
using std::min; using std::max;
#define __device__
#define __forceinline__ inline
typedef ptrdiff_t ssize_t_;
template <class T> struct CArray {
    const T* data; const ptrdiff_t* shp; const ptrdiff_t* str;
    const ptrdiff_t* shape() const { return shp; }
    const ptrdiff_t* strides() const { return str; }
    const T& operator[](ptrdiff_t i) const { return data[i]; }
};
%(preamble)s
typedef %(X)s X; typedef %(W)s W; typedef %(Y)s Y;
extern "C" void run(const X* xd, const ptrdiff_t* xshape, const ptrdiff_t* xstrides,
                    const W* wd, const ptrdiff_t* wshape, const ptrdiff_t* wstrides,
                    Y* yd, const ptrdiff_t* yshape, const ptrdiff_t* ystrides, ptrdiff_t n)
{
    CArray<X> x{xd, xshape, xstrides};
    CArray<W> w{wd, wshape, wstrides};
    CArray<Y> _raw_y{yd, yshape, ystrides};
    for (ptrdiff_t i = 0; i < n; ++i) {
        Y y;
        { %(operation)s }
        yd[i] = y;
    }
}
"""


class ElementwiseKernel:
    """CPU stand-in for cupy.ElementwiseKernel restricted to what the correlate generator emits:
    in_params 'raw X x, raw W w', out_params 'Y y', reduce_dims=False."""
    _cache = {}

    def __init__(self, in_params, out_params, operation, name="kernel", reduce_dims=True, preamble="",
                 options=(), **kw):
        assert in_params.replace(" ", "") == "rawXx,rawWw" and out_params.strip() == "Y y", (in_params, out_params)
        self.operation, self.name = operation, name
        # the reference's preamble, minus CuPy-only headers and the float16 / complex specialisations
        keep = []
        for line in preamble.splitlines():
            if "#include" in line or "float16" in line or "complex<" in line:
                continue
            keep.append(line)
        self.preamble = "\n".join(keep)

    def _lib(self, xt, wt, yt):
        src = HARNESS % dict(preamble=self.preamble, operation=self.operation,
                             X=CTYPES[xt.name], W=CTYPES[wt.name], Y=CTYPES[yt.name])
        key = hashlib.sha1(src.encode()).hexdigest()
        if key not in self._cache:
            d = tempfile.mkdtemp(prefix="refkern_")
            cpp, so = os.path.join(d, "k.cpp"), os.path.join(d, "k.so")
            with open(cpp, "w") as f:
                f.write(src)
            subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++14", "-shared", "-fPIC", "-w",
                                   "-o", so, cpp])
            self._cache[key] = ctypes.CDLL(so)
        return self._cache[key]

    def __call__(self, x, w, y):
        lib = self._lib(x.dtype, w.dtype, y.dtype)
        assert y.flags.c_contiguous
        P = ctypes.c_ssize_t

        def arr(v):
            return (P * len(v))(*v)
        lib.run(ctypes.c_void_p(x.ctypes.data), arr(x.shape), arr(x.strides),
                ctypes.c_void_p(w.ctypes.data), arr(w.shape), arr(w.strides),
                ctypes.c_void_p(y.ctypes.data), arr(y.shape), arr(y.strides), P(y.size))
        return y


def install_shim():
    cupy = types.ModuleType("cupy")

    def memoize(for_each_device=False):
        def deco(f):
            cache = {}

            def wrapper(*a):
                if a not in cache:
                    cache[a] = f(*a)
                return cache[a]
            return wrapper
        return deco

    cupy.ndarray = np.ndarray
    cupy.ElementwiseKernel = ElementwiseKernel
    cupy.memoize = memoize
    cupy.asnumpy = np.asarray
    cupy.shares_memory = lambda a, b, max_work=None: np.may_share_memory(a, b)
    cupy.sqrt = lambda a, out=None, casting="same_kind": np.sqrt(a, out, casting=casting)
    cupy.multiply = np.multiply
    for sub in ("util", "_util"):
        m = types.ModuleType("cupy." + sub)
        m.memoize = memoize
        setattr(cupy, sub, m)
        sys.modules["cupy." + sub] = m
    cupy.__getattr__ = lambda name: getattr(np, name)     # everything else is numpy
    sys.modules["cupy"] = cupy

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    base = os.path.join(REF, "cupyimg")
    root = pkg("cupyimg", base)
    root.memoize = memoize
    pkg("cupyimg.scipy", os.path.join(base, "scipy"))
    nd = pkg("cupyimg.scipy.ndimage", os.path.join(base, "scipy", "ndimage"))
    pkg("cupyimg.scipy.ndimage._kernels", os.path.join(base, "scipy", "ndimage", "_kernels"))
    # cupyimg/_misc.py imports the whole package back; only two helpers are needed by filters.py
    misc = types.ModuleType("cupyimg._misc")

    def _normalize_axis_index(axis, ndim):
        if axis < 0:
            axis += ndim
        if not 0 <= axis < ndim:
            raise np.exceptions.AxisError("axis out of bounds")
        return axis

    def _reshape_nd(arr, ndim, axis):
        axis %= ndim
        return arr.reshape((1,) * axis + (arr.size,) + (1,) * (ndim - axis - 1))
    misc._normalize_axis_index, misc._reshape_nd = _normalize_axis_index, _reshape_nd
    sys.modules["cupyimg._misc"] = misc
    root._misc = misc
    nddir = os.path.join(base, "scipy", "ndimage")
    nd._util = load("cupyimg.scipy.ndimage._util", os.path.join(nddir, "_util.py"))
    load("cupyimg.scipy.ndimage._kernels.support", os.path.join(nddir, "_kernels", "support.py"))
    load("cupyimg.scipy.ndimage._kernels.filters_v2", os.path.join(nddir, "_kernels", "filters_v2.py"))
    nd._filters_core = load("cupyimg.scipy.ndimage._filters_core", os.path.join(nddir, "_filters_core.py"))
    nd._filters_optimal_medians = load("cupyimg.scipy.ndimage._filters_optimal_medians",
                                       os.path.join(nddir, "_filters_optimal_medians.py"))
    return load("cupyimg.scipy.ndimage.filters", os.path.join(nddir, "filters.py"))


def main():
    ref = install_shim()
    rng = np.random.default_rng(20261017)
    # compact storage: every input / output is appended to one byte blob; the JSON index records
    # (offset, shape, dtype) so that thousands of tiny arrays cost no per-entry zip overhead
    index, xblob, yblob = [], bytearray(), bytearray()
    xseen = {}

    def add(func, x, kwargs, out):
        xb = np.ascontiguousarray(x).tobytes()
        xkey = (xb, x.dtype.str, x.shape)
        if xkey not in xseen:
            xseen[xkey] = len(xblob)
            xblob.extend(xb)
        yb = np.ascontiguousarray(out).tobytes()
        index.append({"func": func, "kwargs": kwargs, "x": [xseen[xkey], list(x.shape), x.dtype.str],
                      "y": [len(yblob), list(out.shape), out.dtype.str]})
        yblob.extend(yb)

    # --- correlate1d / convolve1d: dtype x mode x K x origin x axis, integer and float taps ---
    shapes = [(9,), (3, 4, 6)]
    for dt in ["uint8", "uint16", "int32", "float32", "float64"]:
        for shape in shapes:
            x = (rng.random(shape) * 50).astype(dt)
            for mode in ["reflect", "constant", "nearest", "mirror", "wrap"]:
                for K in (2, 3, 8):
                    wi = rng.integers(-3, 4, K).astype(np.float64)
                    wf = rng.standard_normal(K)
                    for axis in range(len(shape)):
                        for origin in sorted({-(K // 2), 0, (K - 1) // 2}):
                            for fn in ("correlate1d", "convolve1d"):
                                for wname, w in (("int", wi), ("float", wf)):
                                    if wname == "float" and fn == "convolve1d":
                                        continue
                                    kw = dict(weights=w.tolist(), axis=axis, mode=mode, cval=0.0, origin=origin)
                                    y = getattr(ref, fn)(x.copy(), np.asarray(w), axis=axis, mode=mode, cval=0.0,
                                                         origin=origin, dtype_mode="ndimage")
                                    add(fn, x, kw, y)
    # --- composite filters (the reference's own per-axis staging, taps, roundings) ---
    for dt in ["float32", "float64", "uint16", "int32"]:
        for shape in [(12, 14), (6, 9, 10)]:
            x = (rng.random(shape) * 100).astype(dt)
            for mode in ["reflect", "mirror", "nearest", "wrap", "constant"]:
                add("gaussian_filter", x, dict(sigma=1.5, mode=mode), ref.gaussian_filter(x.copy(), 1.5, mode=mode))
                add("gaussian_filter", x, dict(sigma=1.0, order=1, mode=mode),
                    ref.gaussian_filter(x.copy(), 1.0, order=1, mode=mode))
                add("uniform_filter", x, dict(size=3, mode=mode), ref.uniform_filter(x.copy(), 3, mode=mode))
                add("sobel", x, dict(axis=0, mode=mode), ref.sobel(x.copy(), 0, mode=mode))
                add("prewitt", x, dict(axis=-1, mode=mode), ref.prewitt(x.copy(), -1, mode=mode))
                add("laplace", x, dict(mode=mode), ref.laplace(x.copy(), mode=mode))
                add("gaussian_gradient_magnitude", x, dict(sigma=1.0, mode=mode),
                    ref.gaussian_gradient_magnitude(x.copy(), 1.0, mode=mode))
    import json
    np.savez_compressed(OUT, index=np.array(json.dumps(index)), x=np.frombuffer(bytes(xblob), np.uint8),
                        y=np.frombuffer(bytes(yblob), np.uint8))
    print("wrote %d cases to %s (%.0f KiB)" % (len(index), OUT, os.path.getsize(OUT) / 1024))


if __name__ == "__main__":
    main()
