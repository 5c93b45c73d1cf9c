"""skimage-level consumers of the separable filters: cupyimg_b200.skimage.filters.gaussian /
difference_of_gaussians, .feature.structure_tensor, .metrics.structural_similarity.

CPU tier: the numpy restatement (oracle/consumers.py) against the literal examples of the reference's docstrings
and against the same formulas on scipy.ndimage's filters.  GPU tier: the CUDA path against that restatement."""
import warnings

import numpy as np
import pytest

from oracle import consumers as oc
from oracle import oracle


# ---------------------------------------------------------------- oracle pinned (CPU)
def test_oracle_gaussian_reference_docstring_examples():
    """/root/reference/cupyimg/skimage/filters/_gaussian.py:86-101 (values as printed there)."""
    a = np.zeros((3, 3))
    a[1, 1] = 1
    np.testing.assert_allclose(oc.gaussian(a, sigma=0.4), [[0.00163116, 0.03712502, 0.00163116],
                                                           [0.03712502, 0.84496158, 0.03712502],
                                                           [0.00163116, 0.03712502, 0.00163116]], atol=5e-9)
    np.testing.assert_allclose(oc.gaussian(a, sigma=1), [[0.05855018, 0.09653293, 0.05855018],
                                                         [0.09653293, 0.15915589, 0.09653293],
                                                         [0.05855018, 0.09653293, 0.05855018]], atol=5e-9)
    np.testing.assert_allclose(oc.gaussian(a, sigma=1, mode="reflect"), [[0.08767308, 0.12075024, 0.08767308],
                                                                         [0.12075024, 0.16630671, 0.12075024],
                                                                         [0.08767308, 0.12075024, 0.08767308]], atol=5e-9)


def test_oracle_structure_tensor_reference_docstring_example():
    """/root/reference/cupyimg/skimage/feature/corner.py:89-98: square image, sigma 0.1 -> Acc."""
    square = np.zeros((5, 5))
    square[2, 2] = 1
    Arr, Arc, Acc = oc.structure_tensor(square, sigma=0.1, order="rc")
    np.testing.assert_allclose(Acc, [[0, 0, 0, 0, 0], [0, 1, 0, 1, 0], [0, 4, 0, 4, 0], [0, 1, 0, 1, 0],
                                     [0, 0, 0, 0, 0]], atol=1e-12)


def test_oracle_ssim_against_scipy_filters_and_identity():
    sndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(0)
    x = rng.random((40, 48))
    y = np.clip(x + 0.1 * rng.standard_normal(x.shape), 0, 1)
    for gw in (False, True):
        m0, S0 = oc.structural_similarity(x, y, data_range=1.0, gaussian_weights=gw, full=True)
        m1, S1 = oc.structural_similarity(x, y, data_range=1.0, gaussian_weights=gw, full=True, filters=sndi)
        np.testing.assert_allclose(S0, S1, rtol=1e-10, atol=1e-12)
        assert abs(m0 - m1) < 1e-12 and 0 < m0 < 1
    assert oc.structural_similarity(x, x, data_range=1.0) == pytest.approx(1.0, abs=1e-12)
    # Wang et al. 2004 settings
    m = oc.structural_similarity(x, y, data_range=1.0, gaussian_weights=True, sigma=1.5, use_sample_covariance=False)
    assert 0 < m < 1


# ---------------------------------------------------------------- CUDA path vs the restatement (GPU)
gpu = pytest.mark.gpu


def _dev(a):
    import torch
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.uint16)
    return torch.from_numpy(a).cuda()


def _host(t):
    return t.cpu().numpy()


@gpu
def test_gaussian_docstring_examples_and_dtypes():
    from cupyimg_b200.skimage import filters as skf
    a = np.zeros((3, 3))
    a[1, 1] = 1
    np.testing.assert_array_equal(_host(skf.gaussian(_dev(a), sigma=0.4)), oc.gaussian(a, sigma=0.4))
    rng = np.random.default_rng(1)
    img8 = rng.integers(0, 256, (40, 52), dtype=np.uint8)
    np.testing.assert_array_equal(_host(skf.gaussian(_dev(img8), 1.5)), oc.gaussian(img8, 1.5))          # float64, bit-exact
    np.testing.assert_array_equal(_host(skf.gaussian(_dev(img8), 2, preserve_range=True, mode="reflect")),
                                  oc.gaussian(img8, 2, preserve_range=True, mode="reflect"))
    img16 = rng.integers(0, 65536, (12, 30, 34), dtype=np.uint16)
    np.testing.assert_array_equal(_host(skf.gaussian(_dev(img16), (1, 2, 0.5))), oc.gaussian(img16, (1, 2, 0.5)))
    f32 = rng.random((24, 40, 64), dtype=np.float32)
    got, want = _host(skf.gaussian(_dev(f32), 2.0)), oc.gaussian(f32, 2.0)
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)
    # colour image: the channel axis is not filtered; the ambiguous (M, N, 3) case warns like the reference
    rgb = rng.random((32, 36, 3))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        got = _host(skf.gaussian(_dev(rgb), 1.0))
        assert any(issubclass(x.category, RuntimeWarning) for x in w)
    np.testing.assert_array_equal(got, oc.gaussian(rgb, 1.0, multichannel=True))
    with pytest.raises(ValueError):
        skf.gaussian(_dev(f32), -1.0)
    import torch
    with pytest.raises(ValueError):
        skf.gaussian(_dev(f32), 1.0, output=torch.empty((24, 40, 64), dtype=torch.int32, device="cuda"))
    out = torch.empty((24, 40, 64), dtype=torch.float32, device="cuda")
    assert skf.gaussian(_dev(f32), 1.0, output=out) is out


@gpu
def test_difference_of_gaussians():
    from cupyimg_b200.skimage import filters as skf
    rng = np.random.default_rng(2)
    img = rng.random((48, 64))
    np.testing.assert_array_equal(_host(skf.difference_of_gaussians(_dev(img), 1.5)), oc.difference_of_gaussians(img, 1.5))
    np.testing.assert_array_equal(_host(skf.difference_of_gaussians(_dev(img), (1, 2), (2, 3.5), mode="reflect")),
                                  oc.difference_of_gaussians(img, (1, 2), (2, 3.5), mode="reflect"))
    with pytest.raises(ValueError):
        skf.difference_of_gaussians(_dev(img), 2.0, 1.0)


@gpu
@pytest.mark.parametrize("shape", [(5, 5), (40, 56), (12, 24, 32)])
def test_structure_tensor(shape):
    from cupyimg_b200.skimage import feature as skfe
    rng = np.random.default_rng(3)
    if shape == (5, 5):
        img = np.zeros(shape)
        img[2, 2] = 1
        sigma = 0.1
    else:
        img = rng.random(shape)
        sigma = 1.5
    got = skfe.structure_tensor(_dev(img), sigma=sigma, order="rc")
    want = oc.structure_tensor(img, sigma=sigma, order="rc")
    assert len(got) == len(want)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(_host(g), w)                  # float64 path: bit-exact
    if len(shape) == 2:
        with pytest.warns(FutureWarning):
            xy = skfe.structure_tensor(_dev(img), sigma=sigma)
        np.testing.assert_array_equal(_host(xy[0]), want[2])         # legacy default order is "xy"
    f32 = img.astype(np.float32)
    for g, w in zip(skfe.structure_tensor(_dev(f32), sigma=sigma, order="rc", mode="reflect"),
                    oc.structure_tensor(f32, sigma=sigma, order="rc", mode="reflect")):
        np.testing.assert_allclose(_host(g), w, rtol=1e-5, atol=1e-5)


@gpu
@pytest.mark.parametrize("shape", [(64, 80), (20, 40, 48)])
@pytest.mark.parametrize("gaussian_weights", [False, True])
def test_structural_similarity(shape, gaussian_weights):
    from cupyimg_b200.skimage import metrics as skm
    rng = np.random.default_rng(4)
    x = rng.random(shape)
    y = np.clip(x + 0.1 * rng.standard_normal(shape), 0, 1)
    kw = dict(data_range=1.0, gaussian_weights=gaussian_weights)
    m, S = skm.structural_similarity(_dev(x), _dev(y), full=True, **kw)
    m0, S0 = oc.structural_similarity(x, y, full=True, **kw)
    if gaussian_weights:
        np.testing.assert_array_equal(_host(S), S0)                 # float64 everywhere: the map is bit-exact
    else:
        np.testing.assert_allclose(_host(S), S0, rtol=1e-10, atol=1e-12)   # uniform f64: running sum vs window sum (1e-12)
    assert abs(float(m) - m0) < 1e-12
    assert float(skm.structural_similarity(_dev(x), _dev(x), **kw)) == pytest.approx(1.0, abs=1e-12)
    # float32 computation
    m32 = skm.structural_similarity(_dev(x.astype(np.float32)), _dev(y.astype(np.float32)), data_dtype=np.float32, **kw)
    assert abs(float(m32) - m0) < 2e-5
    # uint8 images: data_range from the dtype
    a8 = (x * 255).astype(np.uint8)
    b8 = (y * 255).astype(np.uint8)
    assert abs(float(skm.structural_similarity(_dev(a8), _dev(b8), gaussian_weights=gaussian_weights)) -
               oc.structural_similarity(a8, b8, gaussian_weights=gaussian_weights)) < 1e-10
    with pytest.raises(ValueError):
        skm.structural_similarity(_dev(x), _dev(y), win_size=6, data_range=1.0)
    with pytest.raises(ValueError):
        skm.structural_similarity(_dev(x), _dev(y), win_size=101, data_range=1.0)
    with pytest.raises(NotImplementedError):
        skm.structural_similarity(_dev(x), _dev(y), gradient=True, data_range=1.0)


@gpu
def test_ssim_multichannel():
    from cupyimg_b200.skimage import metrics as skm
    rng = np.random.default_rng(5)
    x = rng.random((48, 56, 3))
    y = np.clip(x + 0.05 * rng.standard_normal(x.shape), 0, 1)
    m = float(skm.structural_similarity(_dev(x), _dev(y), multichannel=True, data_range=1.0))
    want = np.mean([oc.structural_similarity(x[..., c], y[..., c], data_range=1.0) for c in range(3)])
    assert abs(m - want) < 1e-12
