"""NaN / Inf in the DATA: a non-finite sample must poison exactly the outputs whose true filter footprint
contains it — as in scipy.ndimage and in the reference's per-tap kernel (_filters_core.py:239-312) — and never a
wider, zero-padded footprint (0 * NaN = NaN).  Every kernel family is instantiated per radius for that reason;
these tests sweep the radii that used to share a wider bucket (5, 7, 9-11, 13-15) and anisotropic filters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _poisoned(shape, dtype, seed):
    rng = np.random.default_rng(seed)
    a = rng.random(shape).astype(dtype)
    flat = a.reshape(-1)
    idx = rng.choice(flat.size, size=6, replace=False)
    flat[idx[:3]] = np.nan
    flat[idx[3:5]] = np.inf
    flat[idx[5]] = -np.inf
    return a


def _same_nonfinite(got, want):
    return (np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isposinf(got), np.isposinf(want))
            and np.array_equal(np.isneginf(got), np.isneginf(want)))


def _close_finite(got, want, rtol):
    m = np.isfinite(want)
    return np.allclose(got[m], want[m], rtol=rtol, atol=rtol * np.abs(want[m]).max())


@pytest.mark.parametrize("radius", [1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 13, 16])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_correlate1d_f32_footprint(radius, axis):
    from cupyimg_b200.scipy import ndimage as ndi
    from oracle import oracle
    a = _poisoned((24, 40, 64), np.float32, 100 * radius + axis)
    w = np.linspace(1.0, 2.0, 2 * radius + 1)
    w /= w.sum()
    want = oracle.correlate1d(a, w, axis=axis, mode="reflect")
    got = ndi.correlate1d(torch.from_numpy(a).cuda(), w, axis=axis, mode="reflect").cpu().numpy()
    assert _same_nonfinite(got, want)
    assert _close_finite(got, want, 1e-5)


@pytest.mark.parametrize("sigma", [0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 2.0, (2.0, 2.0, 0.0), (0.0, 1.0, 2.0), (1.25, 0.5, 1.0), 2.5, 4.0])
def test_gaussian_filter_f32_footprint(sigma):
    from cupyimg_b200.scipy import ndimage as ndi
    from oracle import oracle
    a = _poisoned((36, 48, 80), np.float32, 7)
    want = oracle.gaussian_filter(a, sigma, mode="nearest")
    got = ndi.gaussian_filter(torch.from_numpy(a).cuda(), sigma, mode="nearest").cpu().numpy()
    assert _same_nonfinite(got, want)
    assert _close_finite(got, want, 1e-5)


@pytest.mark.parametrize("sigma", [0.75, 1.0, 1.25, 1.5, 2.0])
def test_gradient_magnitude_f32_footprint(sigma):
    from cupyimg_b200.scipy import ndimage as ndi
    from oracle import oracle
    a = _poisoned((36, 48, 80), np.float32, 11)
    want = oracle.gaussian_gradient_magnitude(a, sigma, mode="reflect")
    got = ndi.gaussian_gradient_magnitude(torch.from_numpy(a).cuda(), sigma, mode="reflect").cpu().numpy()
    assert _same_nonfinite(got, want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("radius", [1, 3, 5, 6, 7, 8, 12])
def test_exact_path_footprint(dtype, radius):
    """dtype_mode='ndimage' (float64 accumulate, scipy order): finite outputs stay bit-identical, non-finite ones
    sit exactly where the oracle has them."""
    from cupyimg_b200.scipy import ndimage as ndi
    from oracle import oracle
    a = _poisoned((12, 64, 96), dtype, 3 * radius)
    d = np.arange(-radius, radius + 1, dtype=np.float64)
    w = np.exp(-0.5 * (d / max(radius / 2.0, 0.5)) ** 2)
    w /= w.sum()
    for axis in (1, 2):
        want = oracle.correlate1d(a, w, axis=axis, mode="mirror")
        got = ndi.correlate1d(torch.from_numpy(a).cuda(), w, axis=axis, mode="mirror", dtype_mode="ndimage").cpu().numpy()
        assert _same_nonfinite(got, want)
        m = np.isfinite(want)
        assert np.array_equal(got[m], want[m])
