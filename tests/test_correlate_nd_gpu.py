"""Dense N-d correlate / convolve on the GPU path against the oracle (pinned to the
reference's known answers and to scipy in tests/test_oracle.py): bit-exact for every dtype pair."""
import itertools

import numpy as np
import pytest

from helpers import TYPES, to_device, to_host
from oracle import oracle
from test_oracle import run_correlate_nd_kats

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ndi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cupyimg_b200.scipy import ndimage
    return ndimage


def test_reference_known_answers(ndi):
    run_correlate_nd_kats(ndi, to_device, to_host)
    import torch
    e = torch.empty((1, 0), device="cuda")
    assert ndi.correlate(e, np.ones((1, 2))).shape == (1, 0)           # correlate10


@pytest.mark.parametrize("dtype", TYPES)
def test_correlate_nd_bit_exact(dtype, ndi):
    rng = np.random.default_rng(TYPES.index(dtype) + 100)
    for shape, wshapes in [((13, 17), [(3, 3), (2, 4), (5, 1), (7, 7)]), ((6, 7, 9), [(3, 3, 3), (2, 1, 4)]), ((20,), [(5,), (4,)]),
                           ((3, 4, 5, 6), [(2, 3, 1, 2)])]:
        x = (rng.random(shape) * 100 - (0 if dtype[0] == "u" else 30)).astype(dtype)
        xd = to_device(x)
        for ws in wshapes:
            w = rng.standard_normal(ws)
            w[tuple(0 for _ in ws)] = 0.0                         # a zero tap: scipy's footprint skips it
            for mode in ["reflect", "constant", "nearest", "mirror", "wrap"]:
                for origin in (0, [-(s // 2) for s in ws], [(s - 1) // 2 for s in ws]):
                    for fn in ("correlate", "convolve"):
                        want = getattr(oracle, fn)(x, w, mode=mode, cval=2.5, origin=origin)
                        got = to_host(getattr(ndi, fn)(xd, w, mode=mode, cval=2.5, origin=origin))
                        assert got.dtype == want.dtype
                        np.testing.assert_array_equal(got, want, err_msg="%s %s %s %s" % (fn, ws, mode, origin))


def test_correlate_nd_outputs_views_large_kernels_errors(ndi):
    import torch
    rng = np.random.default_rng(9)
    x = (rng.random((40, 50)) * 255).astype(np.uint8)
    xd = to_device(x)
    w = rng.standard_normal((5, 5))
    for t_out in TYPES:
        want = oracle.correlate(x, w, output=np.dtype(t_out), mode="mirror")
        got = to_host(ndi.correlate(xd, w, output=np.dtype(t_out), mode="mirror"))
        np.testing.assert_array_equal(got, want, err_msg=t_out)
    # strided view in, in place, more taps than fit in kernel parameters (13 x 13 = 169 > 129)
    v = xd[::2, 3:]
    np.testing.assert_array_equal(to_host(ndi.convolve(v, w)), oracle.convolve(x[::2, 3:], w))
    y = xd.clone()
    assert ndi.correlate(y, w, output=y) is y
    np.testing.assert_array_equal(to_host(y), oracle.correlate(x, w))
    big = rng.standard_normal((13, 13))
    xf = rng.random((64, 70))
    np.testing.assert_array_equal(to_host(ndi.correlate(to_device(xf), big, mode="wrap")),
                                  oracle.correlate(xf, big, mode="wrap"))
    with pytest.raises(RuntimeError):
        ndi.correlate(xd, np.ones(3))                              # weights rank != input rank
    with pytest.raises(ValueError):
        ndi.correlate(xd, np.ones((3, 3)), origin=2)
    with pytest.raises(RuntimeError):
        ndi.correlate(xd, np.ones((3, 3)), mode="foo")
    with pytest.raises(RuntimeError):
        ndi.correlate(xd, np.ones((3, 3)), mode=["reflect", "wrap"])
    # separable kernel == the separable path (integer taps: exact either way)
    xi = to_device((rng.random((30, 31, 32)) * 100).astype(np.int32))
    k = np.array([1.0, 2.0, 1.0])
    k3 = k[:, None, None] * k[None, :, None] * k[None, None, :]
    a = ndi.correlate(xi, k3)
    b = xi
    for ax in range(3):
        b = ndi.correlate1d(b, k, axis=ax)
    assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", ["float32", "uint8", "uint16", "int32", "float64"])
def test_correlate_2d_tile_kernel_multi_tile(dtype, ndi):
    """Weights on the last two axes up to 7 x 7 take the shared-memory tile kernel (correlate_nd.cu): stacks and
    images spanning several 128 x 32 tiles, ragged edges, every boundary mode, shifted origins, non-finite data."""
    rng = np.random.default_rng(41)
    for shape, ws in [((3, 70, 300), (1, 3, 3)), ((2, 45, 260), (1, 5, 5)), ((97, 131), (7, 7)), ((40, 516), (3, 6)),
                      ((33, 129), (6, 2)), ((5, 4), (7, 7))]:
        x = (rng.random(shape) * 200 - (0 if dtype[0] == "u" else 60)).astype(dtype)
        if dtype[0] == "f":
            x[(0,) * x.ndim] = np.inf
            x[tuple(s // 2 for s in shape)] = np.nan
        w = rng.standard_normal(ws)
        w[(0,) * len(ws)] = 0.0
        if len(ws) == 2 and ws[1] > 2:
            w[ws[0] // 2, ws[1] - 1] = 0.0
        xd = to_device(x)
        for mode in ["reflect", "constant", "nearest", "mirror", "wrap"]:
            for origin in (0, [-(s // 2) for s in ws], [(s - 1) // 2 for s in ws]):
                want = oracle.correlate(x, w, mode=mode, cval=-1.5, origin=origin)
                got = to_host(ndi.correlate(xd, w, mode=mode, cval=-1.5, origin=origin))
                np.testing.assert_array_equal(got, want, err_msg="%s %s %s %s" % (shape, ws, mode, origin))
    # another output dtype than the input's (scipy's cast rules in the store)
    x = (rng.random((50, 300)) * 300).astype(dtype)
    w = rng.standard_normal((3, 3))
    for t_out in ("uint8", "float32", "float64", "int64"):
        np.testing.assert_array_equal(to_host(ndi.correlate(to_device(x), w, output=np.dtype(t_out))),
                                      oracle.correlate(x, w, output=np.dtype(t_out)), err_msg=t_out)
