"""minimum / maximum filters on the GPU path against the oracle (pinned to the
reference's known answers and to scipy in tests/test_oracle.py): bit-exact for every dtype."""
import itertools

import numpy as np
import pytest

from helpers import TYPES, to_device, to_host
from oracle import oracle
from test_oracle import MINMAX_KATS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ndi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cupyimg_b200.scipy import ndimage
    return ndimage


def test_reference_known_answers(ndi):
    for fn, x, size, want in MINMAX_KATS:
        got = to_host(getattr(ndi, fn)(to_device(np.asarray(x)), size))
        np.testing.assert_array_equal(got, np.asarray(want), err_msg=fn)
    # an all-True footprint is the separable case (tests/test_ndimage.py:789-795, :865-873)
    x = np.asarray([[3, 2, 5, 1, 4], [7, 6, 9, 3, 5], [5, 8, 3, 7, 1]])
    got = to_host(ndi.minimum_filter(to_device(x), footprint=np.ones((2, 3), bool)))
    np.testing.assert_array_equal(got, [[2, 2, 1, 1, 1], [2, 2, 1, 1, 1], [5, 3, 3, 1, 1]])
    with pytest.raises(NotImplementedError):
        ndi.minimum_filter(to_device(x), footprint=np.asarray([[1, 0, 1], [1, 1, 0]]))
    with pytest.raises(RuntimeError):
        ndi.maximum_filter(to_device(x))
    with pytest.raises(ValueError):
        ndi.maximum_filter1d(to_device(x), 3, origin=2)
    with pytest.raises(NotImplementedError):
        ndi.minimum_filter(to_device(x), 3, mode="constant", cval=float("nan"))


@pytest.mark.parametrize("dtype", TYPES)
def test_minmax_1d_bit_exact(dtype, ndi):
    rng = np.random.default_rng(TYPES.index(dtype))
    x = (rng.random((6, 13, 10)) * 200 - (0 if dtype[0] == "u" else 60)).astype(dtype)
    xd = to_device(x)
    for mode, size, fn in itertools.product(["reflect", "constant", "nearest", "mirror", "wrap"],
                                            [1, 2, 3, 8, 15], ["minimum_filter1d", "maximum_filter1d"]):
        for origin in sorted({-(size // 2), 0, (size - 1) // 2}):
            for axis in (0, 1, 2):
                want = getattr(oracle, fn)(x, size, axis=axis, mode=mode, cval=4.0, origin=origin)
                got = to_host(getattr(ndi, fn)(xd, size, axis=axis, mode=mode, cval=4.0, origin=origin))
                assert got.dtype == want.dtype
                np.testing.assert_array_equal(got, want, err_msg="%s %s size %d origin %d axis %d" % (fn, mode, size, origin, axis))


def test_minmax_nd_outputs_views_inplace(ndi):
    import torch
    rng = np.random.default_rng(5)
    x = (rng.random((20, 33, 17)) * 1000 - 300).astype(np.float32)
    xd = to_device(x)
    for fn in ("minimum_filter", "maximum_filter"):
        for size, mode, origin in [(3, "reflect", 0), ([5, 1, 2], ["wrap", "nearest", "mirror"], [1, 0, -1]),
                                   ((2, 7, 4), "constant", 0)]:
            want = getattr(oracle, fn)(x, size, mode=mode, cval=-7.5, origin=origin)
            got = to_host(getattr(ndi, fn)(xd, size, mode=mode, cval=-7.5, origin=origin))
            np.testing.assert_array_equal(got, want)
        # dtype conversion at every pass, like _run_1d_filters (_filters_core.py:79-109)
        want = getattr(oracle, fn)(x, 3, output=np.int16)
        got = to_host(getattr(ndi, fn)(xd, 3, output=np.int16))
        np.testing.assert_array_equal(got, want)
        # strided input view, in place
        v = xd[::2, :, 1:]
        np.testing.assert_array_equal(to_host(getattr(ndi, fn)(v, 3)), getattr(oracle, fn)(x[::2, :, 1:], 3))
        y = xd.clone()
        assert getattr(ndi, fn)(y, 3, output=y) is y
        np.testing.assert_array_equal(to_host(y), getattr(oracle, fn)(x, 3))
        # axes= (scipy >= 1.11 keyword)
        got = to_host(getattr(ndi, fn)(xd, 5, axes=(0, 2)))
        np.testing.assert_array_equal(got, getattr(oracle, fn)(x, [5, 1, 5]))
    e = torch.empty((0, 4), device="cuda")
    assert ndi.minimum_filter(e, 3).shape == (0, 4)
    # idempotence-style property at a BASELINE-sized image: max of max over the same window grows monotonically
    big = torch.rand((2048, 2048), device="cuda")
    m1 = ndi.maximum_filter(big, 5)
    assert bool((m1 >= big).all()) and bool((ndi.minimum_filter(big, 5) <= big).all())
    assert bool((ndi.maximum_filter(m1, 5) >= m1).all())


@pytest.mark.parametrize("dtype", ["uint8", "int16", "uint16", "float32", "float64"])
def test_minmax_streaming_kernels_bit_exact(dtype, ndi):
    """csrc/minmax_stream.cu (window sizes 2..9, dtype-preserving, vector-aligned shapes): column kernel
    with several segments, row kernel with long rows, origins, every mode, cvals that fold to T exactly and
    cvals that do not (fractional, negative for unsigned types, out of range: the general kernel takes over)."""
    rng = np.random.default_rng(17)
    for shape, axes in [((3, 150, 64), (0, 1, 2)), ((2, 9, 2064), (1, 2)), ((40, 48), (0, 1))]:
        if np.dtype(dtype).kind == "f":
            x = (rng.standard_normal(shape) * 50).astype(dtype)
        else:
            info = np.iinfo(dtype)
            x = rng.integers(info.min, info.max, shape, endpoint=True).astype(dtype)
        xd = to_device(x)
        for axis, size, fn in itertools.product(axes, [2, 3, 4, 5, 7, 9], ["minimum_filter1d", "maximum_filter1d"]):
            for mode, cval in [("reflect", 0.0), ("constant", 3.7), ("constant", -2.5), ("constant", 70000.0),
                               ("nearest", 0.0), ("mirror", 0.0), ("wrap", 0.0)]:
                for origin in sorted({0, -(size // 2), (size - 1) // 2}):
                    want = getattr(oracle, fn)(x, size, axis=axis, mode=mode, cval=cval, origin=origin)
                    got = to_host(getattr(ndi, fn)(xd, size, axis=axis, mode=mode, cval=cval, origin=origin))
                    np.testing.assert_array_equal(got, want, err_msg="%s size %d axis %d %s cval %g origin %d" % (
                        fn, size, axis, mode, cval, origin))
