"""Device-array ingestion: torch tensors, ``__cuda_array_interface__`` (CuPy, Numba,
...) and DLPack producers become a plain (pointer, shape, byte strides, dtype, device)
descriptor — the ``sepfilt_tensor`` of the C ABI.  PyTorch is used only to allocate
outputs / scratch and to name the current stream.

Replaces the implicit ``cupy.ndarray`` contract of the reference (every function in
cupyimg/scipy/ndimage/filters.py takes and returns cupy arrays; outputs come from
``_util._get_output``, _util.py:43-81).
"""
import importlib

import numpy as np
import torch

from . import _ffi

_TORCH_DTYPES = {
    np.dtype("int8"): torch.int8, np.dtype("uint8"): torch.uint8,
    np.dtype("int16"): torch.int16, np.dtype("uint16"): torch.uint16,
    np.dtype("int32"): torch.int32, np.dtype("uint32"): torch.uint32,
    np.dtype("int64"): torch.int64, np.dtype("uint64"): torch.uint64,
    np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
    np.dtype("bool"): torch.bool,
    np.dtype("complex64"): torch.complex64, np.dtype("complex128"): torch.complex128,
    np.dtype("float16"): torch.float16,
}
_NUMPY_DTYPES = {v: k for k, v in _TORCH_DTYPES.items()}


class OutputShapeError(ValueError, RuntimeError):
    """Wrong ``output`` shape: the reference raises ValueError (_util.py:49-50), scipy
    RuntimeError (scipy/ndimage/_ni_support.py:103-104); this is both."""


def to_numpy_dtype(dt):
    """Accept numpy / torch dtypes, type objects and dtype strings."""
    if isinstance(dt, torch.dtype):
        return _NUMPY_DTYPES[dt]
    return np.dtype(dt)


class DevArray:
    """A strided view of device memory. ``obj`` keeps the owner alive; ``foreign`` is the
    module name to hand results back to (None for torch)."""

    __slots__ = ("ptr", "shape", "strides", "dtype", "device", "obj", "foreign")

    def __init__(self, ptr, shape, strides, dtype, device, obj, foreign=None):
        self.ptr = int(ptr)
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self.dtype = np.dtype(dtype)
        self.device = int(device)
        self.obj = obj
        self.foreign = foreign

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def itemsize(self):
        return self.dtype.itemsize

    def c_contiguous(self):
        expect = self.itemsize
        for s, st in zip(reversed(self.shape), reversed(self.strides)):
            if s != 1 and st != expect:
                return False
            expect *= s
        return True

    def byte_bounds(self):
        """[lo, hi) of the bytes this view can touch."""
        if self.size == 0:
            return self.ptr, self.ptr
        lo = hi = self.ptr
        for s, st in zip(self.shape, self.strides):
            if st >= 0:
                hi += (s - 1) * st
            else:
                lo += (s - 1) * st
        return lo, hi + self.itemsize

    def may_overlap(self, other):
        """Conservative aliasing test (the reference uses shares_memory(.., 'MAY_SHARE_BOUNDS'),
        _filters_core.py:147)."""
        if self.device != other.device or self.size == 0 or other.size == 0:
            return False
        a0, a1 = self.byte_bounds()
        b0, b1 = other.byte_bounds()
        return a0 < b1 and b0 < a1

    def tensor(self):
        """ctypes ``sepfilt_tensor`` for this view."""
        if self.ndim > _ffi.MAX_NDIM:
            raise RuntimeError("arrays of rank > %d are not supported" % _ffi.MAX_NDIM)
        if self.dtype not in _ffi.DTYPE_CODES:
            raise RuntimeError("array type %s not supported" % self.dtype)
        t = _ffi.Tensor()
        t.ptr = self.ptr if self.ptr else None
        t.dtype = _ffi.DTYPE_CODES[self.dtype]
        t.ndim = self.ndim
        for i, (s, st) in enumerate(zip(self.shape, self.strides)):
            t.shape[i] = s
            t.stride_bytes[i] = st
        t.device = self.device
        return t

    def component(self, which):
        """Real (0) / imaginary (1) part of a complex array as a strided real view (no copy)."""
        if self.dtype.kind != "c":
            raise TypeError("component() needs a complex array")
        part = np.dtype("float32") if self.dtype.itemsize == 8 else np.dtype("float64")
        return DevArray(self.ptr + which * part.itemsize, self.shape, self.strides, part, self.device,
                        self.obj, self.foreign)

    def view_axis_window(self, axis, start, length):
        """Sub-view [start, start+length) along ``axis`` (no copy)."""
        shape = list(self.shape)
        shape[axis] = length
        return DevArray(self.ptr + start * self.strides[axis], shape, self.strides, self.dtype,
                        self.device, self.obj, self.foreign)


def _from_torch(t, foreign=None):
    if not t.is_cuda:
        raise TypeError("expected an array in CUDA device memory, got a CPU tensor")
    if t.dtype not in _NUMPY_DTYPES:
        raise RuntimeError("array type %s not supported" % t.dtype)
    dt = _NUMPY_DTYPES[t.dtype]
    if t.requires_grad:
        t = t.detach()
    strides = [s * dt.itemsize for s in t.stride()]
    return DevArray(t.data_ptr(), t.shape, strides, dt, t.device.index, t, foreign)


def is_device_array(x):
    return isinstance(x, (torch.Tensor, DevArray)) or hasattr(x, "__cuda_array_interface__") \
        or (hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray))


def ingest(x, name="input"):
    """Any supported device array -> DevArray."""
    if isinstance(x, DevArray):
        return x
    if isinstance(x, torch.Tensor):
        return _from_torch(x)
    cai = getattr(x, "__cuda_array_interface__", None)
    if cai is not None:
        shape = tuple(cai["shape"])
        dt = np.dtype(cai["typestr"])
        strides = cai.get("strides")
        if strides is None:
            strides, acc = [], dt.itemsize
            for s in reversed(shape):
                strides.append(acc)
                acc *= s
            strides = tuple(reversed(strides))
        dev = getattr(getattr(x, "device", None), "id", None)
        if dev is None:
            dev = torch.cuda.current_device()
        # CAI v3: work the producer queued on its stream must be complete before the array is read.  1 is the
        # legacy default stream (torch's default stream: already ordered), 2 the per-thread default stream.
        stream = cai.get("stream")
        if isinstance(stream, int) and stream > 2:
            ext = torch.cuda.ExternalStream(stream, device=torch.device("cuda", dev))
            torch.cuda.current_stream(dev).wait_stream(ext)
        elif stream == 2:
            torch.cuda.synchronize(dev)
        foreign = type(x).__module__.split(".")[0]
        return DevArray(cai["data"][0] or 0, shape, strides, dt, dev, x, foreign)
    if hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        foreign = type(x).__module__.split(".")[0]
        return _from_torch(torch.from_dlpack(x), foreign)
    raise TypeError("%s must be a CUDA device array (torch.Tensor, cupy.ndarray or any object with "
                    "__cuda_array_interface__ / __dlpack__), got %s" % (name, type(x).__name__))


def empty(shape, dtype, device):
    """Uninitialised device array from torch's caching allocator (the reference memsets
    every output with cupy.zeros, _util.py:80; every element is written exactly once here)."""
    dt = np.dtype(dtype)
    if dt not in _TORCH_DTYPES:
        raise RuntimeError("array type %s not supported" % dt)
    t = torch.empty(tuple(shape), dtype=_TORCH_DTYPES[dt], device=torch.device("cuda", device))
    return _from_torch(t)


def export(arr, like):
    """Hand a result back as the same kind of object the caller passed in."""
    if arr.obj is not None and not isinstance(arr.obj, torch.Tensor):
        return arr.obj               # caller-provided foreign output array
    t = arr.obj
    foreign = like.foreign if isinstance(like, DevArray) else None
    if foreign and foreign != "torch":
        try:
            mod = importlib.import_module(foreign)
            if hasattr(mod, "from_dlpack"):
                return mod.from_dlpack(t)
            if hasattr(mod, "asarray"):
                return mod.asarray(t)
        except Exception:            # pragma: no cover - foreign library missing at return time
            pass
    return t


def host_weights(w):
    """Filter weights -> host float64 (or complex128) numpy array.  Weights are a few
    hundred bytes; the C ABI takes them as host doubles."""
    if isinstance(w, torch.Tensor):
        w = w.detach().cpu().numpy()
    elif hasattr(w, "__cuda_array_interface__") or (hasattr(w, "__dlpack__") and not isinstance(w, np.ndarray)):
        if hasattr(w, "get"):
            w = w.get()
        else:
            w = torch.as_tensor(w, device="cuda").cpu().numpy()
    w = np.asarray(w)
    if w.dtype.kind == "c":
        return w.astype(np.complex128)
    return w.astype(np.float64)


def current_stream(device):
    return torch.cuda.current_stream(device).cuda_stream
