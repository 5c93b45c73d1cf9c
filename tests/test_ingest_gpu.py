"""Array ingestion through ``__cuda_array_interface__`` and ``__dlpack__`` (north_star: "accepts cupy.ndarray or
torch tensors through DLPack / __cuda_array_interface__"; reference callers pass cupy.ndarray everywhere,
e.g. filters.py:286-438).  CuPy is not installed here, so the producers are minimal stand-ins that expose ONLY
the protocol under test: positive, negative and sliced strides, a DLPack-only object, and the return-type
round trip (a foreign output array comes back as the same object; a foreign input picks the producer's
``from_dlpack``)."""
import sys
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


class CAIArray:
    """Exposes nothing but __cuda_array_interface__ (v3) over memory owned by a torch tensor."""

    def __init__(self, owner, ptr, shape, strides, typestr):
        self._owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "strides": None if strides is None else tuple(strides), "version": 3}

    @classmethod
    def of(cls, t, with_strides=True):
        strides = tuple(s * t.element_size() for s in t.stride()) if with_strides else None
        return cls(t, t.data_ptr(), t.shape, strides, np.dtype(str(t.dtype).replace("torch.", "")).str)


class DLPackOnly:
    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None, **kw):
        return self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def _ndi():
    from cupyimg_b200.scipy import ndimage as ndi
    return ndi


def test_cai_contiguous_and_implicit_strides():
    ndi = _ndi()
    x = torch.rand((20, 33, 40), device="cuda")
    want = ndi.gaussian_filter(x, 1.5)
    for with_strides in (True, False):
        got = ndi.gaussian_filter(CAIArray.of(x, with_strides), 1.5)
        assert isinstance(got, torch.Tensor)            # the stand-in's module has no from_dlpack / asarray
        assert torch.equal(got, want)


def test_cai_sliced_strides():
    ndi = _ndi()
    base = torch.rand((24, 40, 64), device="cuda")
    view = base[2:20:2, 5:35, 3:60:3]
    # the same strided view as a torch tensor takes the same (strided, float64-accumulate) kernel: equal bits;
    # the contiguous copy takes the float32 kernel: equal within the float32 contract
    got = ndi.correlate1d(CAIArray.of(view), [1.0, 2.0, 3.0], axis=1, mode="mirror")
    assert torch.equal(got, ndi.correlate1d(view, [1.0, 2.0, 3.0], axis=1, mode="mirror"))
    assert torch.allclose(got, ndi.correlate1d(view.contiguous(), [1.0, 2.0, 3.0], axis=1, mode="mirror"), rtol=1e-5, atol=1e-6)
    got = ndi.uniform_filter(CAIArray.of(view), 3)
    assert torch.equal(got, ndi.uniform_filter(view, 3))
    assert torch.allclose(got, ndi.uniform_filter(view.contiguous(), 3), rtol=1e-5, atol=1e-6)


def test_cai_negative_strides():
    """cupy allows negative strides (x[::-1]); torch does not, so the view is described by hand: the data pointer is
    the address of the LAST row and the row stride is negative."""
    ndi = _ndi()
    x = torch.rand((17, 48), device="cuda")
    es = x.element_size()
    flipped = CAIArray(x, x.data_ptr() + (x.shape[0] - 1) * x.stride(0) * es, x.shape, (-x.stride(0) * es, es), "<f4")
    # same arithmetic on both sides (float64 accumulate): equal bits; the float32 kernel of the contiguous copy
    # agrees within the float32 contract
    kw = dict(axis=0, mode="reflect", origin=-1)
    got = ndi.correlate1d(flipped, [1.0, 0.0, -1.0, 0.5], dtype_mode="ndimage", **kw)
    assert torch.equal(got, ndi.correlate1d(torch.flip(x, [0]).contiguous(), [1.0, 0.0, -1.0, 0.5], dtype_mode="ndimage", **kw))
    got = ndi.correlate1d(flipped, [1.0, 0.0, -1.0, 0.5], **kw)
    assert torch.allclose(got, ndi.correlate1d(torch.flip(x, [0]).contiguous(), [1.0, 0.0, -1.0, 0.5], **kw), rtol=1e-5, atol=1e-6)
    # both axes reversed, integer data on the exact path
    xi = torch.randint(0, 60000, (31, 50), device="cuda", dtype=torch.int32).to(torch.uint16)
    both = CAIArray(xi, xi.data_ptr() + (xi.numel() - 1) * 2, xi.shape, (-xi.stride(0) * 2, -2), "<u2")
    want = ndi.uniform_filter(torch.flip(xi.to(torch.int32), [0, 1]).to(torch.uint16).contiguous(), 5)
    assert torch.equal(ndi.uniform_filter(both, 5), want)


def test_dlpack_only_object():
    ndi = _ndi()
    x = torch.rand((12, 64, 64), device="cuda")
    got = ndi.sobel(DLPackOnly(x), axis=1)
    assert torch.equal(got, ndi.sobel(x, axis=1))


def test_foreign_output_array_is_returned_and_filled():
    ndi = _ndi()
    x = torch.rand((16, 40, 48), device="cuda")
    out_t = torch.zeros_like(x)
    out = CAIArray.of(out_t)
    res = ndi.gaussian_filter(CAIArray.of(x), 1.0, output=out)
    assert res is out                                   # like the reference: the caller's array comes back
    assert torch.equal(out_t, ndi.gaussian_filter(x, 1.0))


def test_return_type_follows_the_producer_module():
    """A producer whose top-level module offers from_dlpack (cupy does) gets its own array type back."""
    ndi = _ndi()
    mod = types.ModuleType("fakegpuarr")

    class Arr(DLPackOnly):
        pass

    Arr.__module__ = "fakegpuarr"
    mod.Arr = Arr
    mod.from_dlpack = lambda t: Arr(torch.from_dlpack(t))
    sys.modules["fakegpuarr"] = mod
    try:
        x = torch.rand((10, 32, 32), device="cuda")
        res = ndi.uniform_filter(Arr(x), 3)
        assert isinstance(res, Arr)
        assert torch.equal(res._t, ndi.uniform_filter(x, 3))
    finally:
        del sys.modules["fakegpuarr"]


def test_cai_producer_stream_is_honoured():
    """CAI v3 'stream': work queued on the producer's stream must be visible to the filter."""
    ndi = _ndi()
    side = torch.cuda.Stream()
    x = torch.empty((64, 256, 256), device="cuda")
    with torch.cuda.stream(side):
        torch.cuda._sleep(20_000_000)                   # ~10 ms of delay before the producer writes
        x.fill_(3.0)
        arr = CAIArray.of(x)
        arr.__cuda_array_interface__["stream"] = side.cuda_stream
    got = ndi.uniform_filter(arr, 3)                   # default stream: must wait for `side`
    torch.cuda.synchronize()
    assert float(got.min()) == 3.0 and float(got.max()) == 3.0
