"""Complex-valued arrays and weights on the GPU path (X, Y in {..., c64, c128};
complex weights are conjugated for correlation, filters.py:467-469) against the oracle, which is
pinned bit-for-bit to scipy in tests/test_oracle.py::test_complex_oracle_matches_scipy."""
import itertools

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ndi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from cupyimg_b200.scipy import ndimage
    return ndimage


def dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def host(t):
    return t.cpu().numpy()


TYPES = [np.float32, np.float64, np.complex64, np.complex128]


@pytest.mark.parametrize("dx,dh", [(a, b) for a, b in itertools.product(TYPES, TYPES)
                                   if np.dtype(a).kind == "c" or np.dtype(b).kind == "c"])
def test_correlate_convolve_complex_matrix(dx, dh, ndi):
    """The dtype matrix of the reference's test_correlate1d_complex (len_h up to 2 len_x + 1, every
    mode), with non-zero imaginary parts, complex cval and both call directions."""
    rng = np.random.default_rng(1)
    cx, ch = np.dtype(dx).kind == "c", np.dtype(dh).kind == "c"
    x = (rng.standard_normal((4, 6)) + (1j * rng.standard_normal((4, 6)) if cx else 0)).astype(dx)
    xd = dev(x)
    for len_h in range(1, 14):
        h = (rng.standard_normal(len_h) + (1j * rng.standard_normal(len_h) if ch else 0)).astype(dh)
        for mode in ("constant", "mirror", "nearest", "reflect", "wrap"):
            cval = (0.25 + 1.5j) if cx else 0.25
            for fn in ("correlate1d", "convolve1d"):
                want = getattr(oracle, fn)(x, h, axis=1, mode=mode, cval=cval)
                got = host(getattr(ndi, fn)(xd, h, axis=1, mode=mode, cval=cval, dtype_mode="ndimage"))
                assert got.dtype == want.dtype, (fn, got.dtype, want.dtype)
                np.testing.assert_array_equal(got, want, err_msg="%s %s len_h=%d" % (fn, mode, len_h))
                # default policy (float32 components may use float32 accumulation): rtol 1e-5
                got = host(getattr(ndi, fn)(xd, h, axis=1, mode=mode, cval=cval))
                np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_composite_filters_on_complex_arrays(dtype, ndi):
    """Real-tap filters act on the real and the imaginary component independently."""
    rng = np.random.default_rng(2)
    x = (rng.random((9, 12, 10)) + 1j * rng.random((9, 12, 10))).astype(dtype)
    xd = dev(x)
    part = np.float32 if dtype == np.complex64 else np.float64
    tol = dict(rtol=1e-5, atol=1e-6) if dtype == np.complex64 else dict(rtol=1e-12, atol=1e-13)
    for name, args, kw in [
        ("gaussian_filter", (1.2,), {}),
        ("gaussian_filter", ([1.0, 0.0, 2.0],), {"order": [0, 0, 1], "mode": "mirror"}),
        ("uniform_filter", (3,), {"mode": "wrap"}),
        ("uniform_filter1d", (4,), {"axis": 1, "origin": -1}),
        ("gaussian_filter1d", (1.5,), {"axis": 0, "order": 2, "mode": "constant", "cval": 1.0 - 2.0j}),
        ("sobel", (1,), {}), ("prewitt", (0,), {"mode": "nearest"}),
        ("laplace", (), {}), ("gaussian_laplace", (1.1,), {}),
    ]:
        kre, kim = dict(kw), dict(kw)
        if "cval" in kw:
            kre["cval"], kim["cval"] = kw["cval"].real, kw["cval"].imag
        want = (getattr(oracle, name)(x.real.astype(part), *args, **kre)
                + 1j * getattr(oracle, name)(x.imag.astype(part), *args, **kim))
        got = host(getattr(ndi, name)(xd, *args, **kw))
        assert got.dtype == np.dtype(dtype), name
        np.testing.assert_allclose(got, want, err_msg=name, **tol)


def test_complex_output_rules_and_errors(ndi):
    import torch
    x = dev((np.arange(12) + 1j * np.arange(12)[::-1]).reshape(3, 4).astype(np.complex64))
    w = np.array([1.0, 2.0, 1.0])
    assert ndi.correlate1d(x, w).dtype == torch.complex64
    assert ndi.correlate1d(x, w, output=np.complex128).dtype == torch.complex128
    out = torch.empty((3, 4), dtype=torch.complex128, device="cuda")
    assert ndi.correlate1d(x, w, output=out) is out
    np.testing.assert_array_equal(host(out), oracle.correlate1d(host(x), w, output=np.complex128))
    for bad in (np.float64, torch.empty((3, 4), dtype=torch.float32, device="cuda")):
        with pytest.raises(RuntimeError):
            ndi.correlate1d(x, w, output=bad)
    xr = dev(np.arange(12.0).reshape(3, 4))
    assert ndi.correlate1d(xr, w * 1j).dtype == torch.complex128      # promote_types(float64, complex64)
    assert ndi.correlate1d(xr.float(), w * 1j).dtype == torch.complex64
    with pytest.raises(RuntimeError):
        ndi.correlate1d(xr, w * 1j, output=np.float64)
    with pytest.raises(ValueError):
        ndi.correlate1d(xr, w, mode="constant", cval=1j)
    with pytest.raises(NotImplementedError):
        ndi.gaussian_gradient_magnitude(x, 1.0)
    # in place on a complex array
    y = x.clone()
    ndi.gaussian_filter(y, 1.0, output=y)
    np.testing.assert_allclose(host(y), host(ndi.gaussian_filter(x, 1.0)), rtol=1e-6)
    # empty
    e = torch.empty((0, 4), dtype=torch.complex64, device="cuda")
    assert ndi.correlate1d(e, w).shape == (0, 4)


def test_laplace_of_a_size_one_complex_array(ndi):
    """A size-1 component view counts as contiguous; the per-axis accumulation must still not write through the
    complex owner."""
    for shape in [(1,), (1, 1), (1, 1, 1)]:
        x = np.full(shape, 1.5 - 2.5j, dtype=np.complex64)
        got = host(ndi.laplace(dev(x)))
        want = oracle.laplace(x.real.copy()) + 1j * oracle.laplace(x.imag.copy())
        assert got.dtype == np.complex64 and got.shape == shape
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
