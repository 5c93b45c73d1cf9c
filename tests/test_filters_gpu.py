"""GPU parity tests: the public API (-> ctypes -> C ABI -> sm_100a kernels) against the
CPU oracle on the same seeded inputs.  Integer / float64 outputs must be bit-exact;
float32 through the float32 kernels must satisfy |a-b| <= atol + 1e-5 |b| (north_star:
rtol 1e-5; atol 1e-6 * max|b| for near-zero derivative outputs)."""
import itertools
import threading

import numpy as np
import pytest

from helpers import (TYPES, call_filter, check_against_reference, gpu_call, load_kats, load_reference_vectors,
                     run_kat, to_device, to_host)
from oracle import oracle

pytestmark = pytest.mark.gpu

MODES = ["reflect", "constant", "nearest", "mirror", "wrap"]
KATS = load_kats()


@pytest.fixture(scope="module")
def ndi():
    import torch
    assert torch.cuda.is_available()
    from cupyimg_b200.scipy import ndimage
    return ndimage


def assert_f32_close(got, want, rtol=1e-5, atol_scale=1e-6, atol=None):
    want64 = want.astype(np.float64)
    if atol is None:
        atol = atol_scale * max(float(np.abs(want64).max()) if want64.size else 0.0, 1e-30)
    err = np.abs(got.astype(np.float64) - want64)
    bound = atol + rtol * np.abs(want64)
    assert (err <= bound).all(), "max abs err %.3e (bound %.3e)" % (err.max(), bound.flat[err.argmax()])


# ---------------------------------------------------------------- known answers
def _gpu_kat_call(func, x, **kw):
    if "weights" in kw:
        w = kw.pop("weights")
        import cupyimg_b200.scipy.ndimage as ndi
        return to_host(getattr(ndi, func)(to_device(x), np.asarray(w), **kw))
    if func == "uniform_filter1d":
        return gpu_call(func, x, **kw)
    return gpu_call(func, x, **kw)


@pytest.mark.parametrize("case", KATS, ids=[c["id"] for c in KATS])
def test_reference_known_answers(case, ndi):
    run_kat(case, _gpu_kat_call)


def test_reference_executed_vectors(ndi):
    """The CUDA path (exact kernels, dtype_mode='ndimage') against outputs of the reference itself,
    executed on the CPU from its own generated kernel source (tests/golden/make_reference_vectors.py)."""
    class GpuNs:
        def __getattr__(self, name):
            fn = getattr(ndi, name)
            return lambda x, *a, **kw: to_host(fn(to_device(x), *a, dtype_mode="ndimage", **kw))
    n = {"checked": 0, "skipped": 0}
    for func, kwargs, x, ref in load_reference_vectors():
        if func in ("laplace", "gaussian_gradient_magnitude"):       # no dtype_mode keyword on these two
            kw = dict(kwargs)
            first = [kw.pop("sigma")] if "sigma" in kw else []
            got = to_host(getattr(ndi, func)(to_device(x), *first, **kw))
            if ref.dtype == np.float32:                               # float32 kernels: tolerance, not bits
                np.testing.assert_allclose(got, ref, rtol=1e-5, atol=2e-6 * float(np.abs(ref).max()))
                n["checked"] += 1
                continue
        else:
            got = call_filter(GpuNs(), func, x, kwargs)
        n[check_against_reference(func, x, got, ref)] += 1
    assert n["checked"] > 1000


# ---------------------------------------------------------------- exact path
@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("len_x", [1, 2, 3, 6, 7])
def test_length_origin_sweep(mode, len_x, ndi):
    """tests/test_ndimage_vs_scipy.py:24-111 — every length up to 2n+1 x every origin; f64 bit-exact."""
    x = np.arange(1, 1 + len_x, dtype=np.float64)
    xd = to_device(x)
    for len_h in range(1, 2 * len_x + 2):
        h = np.arange(1, 1 + len_h, dtype=np.float64)
        lo, hi = -(len_h // 2), (len_h - 1) // 2
        for origin in range(lo, hi + 1):
            for fn in ("correlate1d", "convolve1d"):
                want = getattr(oracle, fn)(x, h, mode=mode, cval=0.25, origin=origin)
                got = to_host(getattr(ndi, fn)(xd, h, mode=mode, cval=0.25, origin=origin))
                np.testing.assert_array_equal(got, want, err_msg="%s K=%d origin=%d" % (fn, len_h, origin))
        for origin in (lo - 1, hi + 1):
            with pytest.raises(ValueError):
                ndi.correlate1d(xd, h, mode=mode, origin=origin)
            with pytest.raises(ValueError):
                ndi.convolve1d(xd, h, mode=mode, origin=origin)


@pytest.mark.parametrize("t_in", TYPES)
def test_dtype_matrix_bit_exact(t_in, ndi):
    rng = np.random.default_rng(TYPES.index(t_in))
    x = (rng.random((5, 7, 6)) * 100).astype(t_in)
    xd = to_device(x)
    taps = [oracle.gaussian_kernel1d(1.0, 0, 4), oracle.gaussian_kernel1d(1.0, 1, 4),
            rng.standard_normal(4), np.array([1.0, 2.0, 1.0])]
    for t_out in TYPES:
        for w, axis in itertools.product(taps, range(3)):
            want = oracle.correlate1d(x, w, axis=axis, output=np.dtype(t_out), mode="mirror")
            got = to_host(ndi.correlate1d(xd, w, axis=axis, output=np.dtype(t_out), mode="mirror",
                                          dtype_mode="ndimage"))
            assert got.dtype == want.dtype
            np.testing.assert_array_equal(got, want, err_msg="%s->%s axis %d" % (t_in, t_out, axis))


@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int16", "int32", "int64", "float64"])
def test_composite_filters_exact(dtype, ndi):
    rng = np.random.default_rng(11)
    x = (rng.random((12, 17, 9)) * 200).astype(dtype)
    xd = to_device(x)
    for name, args, kw in [
        ("gaussian_filter", (1.5,), {}),
        ("gaussian_filter", ([1.0, 0.0, 2.0],), {"order": [0, 0, 1]}),
        ("uniform_filter", (5,), {}),
        ("uniform_filter", ([3, 1, 4],), {"origin": [0, 0, -1], "mode": "wrap"}),
        ("sobel", (0,), {}), ("prewitt", (-1,), {"mode": ["reflect", "wrap", "mirror"]}),
        ("gaussian_gradient_magnitude", (1.5,), {}),
        ("laplace", (), {}), ("gaussian_laplace", (1.2,), {"mode": "nearest"}),
        ("gaussian_filter1d", (2.0,), {"axis": 1, "order": 2, "mode": "constant", "cval": 3.0}),
        ("uniform_filter1d", (4,), {"axis": 0, "origin": 1, "mode": "mirror"}),
    ]:
        want = getattr(oracle, name)(x, *args, **kw)
        got = to_host(getattr(ndi, name)(xd, *args, **kw))
        assert got.dtype == want.dtype
        if dtype == "float64" and name.startswith("uniform"):
            # scipy's running sum vs an exact window sum: float64 contract is rtol 1e-12
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, err_msg=name)
        else:
            np.testing.assert_array_equal(got, want, err_msg=name)


def test_long_filters_and_multireflection(ndi):
    """K > SEPFILT_PARAM_TAPS goes through device scratch; K >> n wraps many times."""
    rng = np.random.default_rng(5)
    x = rng.random((9, 40)).astype(np.float64)
    xd = to_device(x)
    for K, mode in itertools.product([129, 130, 301], MODES):
        w = rng.standard_normal(K)
        for axis in (0, 1):
            want = oracle.correlate1d(x, w, axis=axis, mode=mode, cval=-1.5)
            got = to_host(ndi.correlate1d(xd, w, axis=axis, mode=mode, cval=-1.5))
            np.testing.assert_array_equal(got, want)
    want = oracle.gaussian_filter(x, 25.0)           # 201 taps, symmetric, f64
    got = to_host(ndi.gaussian_filter(xd, 25.0))
    np.testing.assert_array_equal(got, want)


def test_strided_and_inplace(ndi):
    import torch
    rng = np.random.default_rng(9)
    base = (rng.random((6, 10, 8)) * 50).astype(np.int32)
    bd = to_device(base)
    w = np.array([1.0, -2.0, 3.0, 0.5])
    # non-contiguous input views (tests/test_filters_from_cupy.py:147-150)
    for view_np, view_t in [(base[..., :4], bd[..., :4]), (base[:, ::2], bd[:, ::2]),
                            (base.transpose(2, 0, 1), bd.permute(2, 0, 1))]:
        for axis in range(3):
            want = oracle.correlate1d(view_np, w, axis=axis, mode="reflect")
            got = to_host(ndi.correlate1d(view_t, w, axis=axis, mode="reflect"))
            np.testing.assert_array_equal(got, want)
    # strided output
    out = torch.zeros((6, 10, 16), dtype=torch.float64, device="cuda")
    res = ndi.correlate1d(bd, w, axis=1, output=out[..., ::2])
    want = oracle.correlate1d(base, w, axis=1, output=np.float64)
    np.testing.assert_array_equal(to_host(out[..., ::2]), want)
    assert res.data_ptr() == out.data_ptr()
    # in place: output is input
    x = (rng.random((7, 9)) * 10).astype(np.float64)
    for fn, args in [("correlate1d", (w,)), ("gaussian_filter", (1.0,)), ("uniform_filter", (3,)), ("sobel", ())]:
        xd = to_device(x)
        getattr(ndi, fn)(xd, *args, output=xd)
        np.testing.assert_allclose(to_host(xd), getattr(oracle, fn)(x, *args), rtol=1e-12 if fn == "uniform_filter" else 0,
                                   atol=0, err_msg=fn)


def test_degenerate_inputs_and_errors(ndi):
    import torch
    e = torch.zeros((0,), dtype=torch.float64, device="cuda")
    assert ndi.correlate1d(e, [1.0, 1.0]).shape == (0,)
    assert ndi.uniform_filter(torch.zeros((3, 0), device="cuda"), 3).shape == (3, 0)
    x = to_device(np.arange(12, dtype=np.float32).reshape(3, 4))
    np.testing.assert_array_equal(to_host(ndi.gaussian_filter(x, 0)), to_host(x))
    np.testing.assert_array_equal(to_host(ndi.uniform_filter(x, 1)), to_host(x))
    got = ndi.uniform_filter(x, [1, 0], output=np.int16)
    assert got.dtype == torch.int16
    np.testing.assert_array_equal(to_host(got), np.arange(12).reshape(3, 4))
    one = torch.zeros((1,), dtype=torch.float64, device="cuda")
    assert float(ndi.gaussian_filter(one, 1, order=3)) == 0.0      # tests/test_filters.py:80-88
    with pytest.raises(ValueError):
        ndi.gaussian_filter(one, 1, -1)
    with pytest.raises(ValueError):
        ndi.gaussian_filter1d(one, 1, -1, -1)
    with pytest.raises(RuntimeError):
        ndi.correlate1d(x, np.ones((2, 2)))
    with pytest.raises(RuntimeError):
        ndi.correlate1d(x, [])
    with pytest.raises(RuntimeError):
        ndi.correlate1d(x, [1.0], mode="unknown")
    with pytest.raises(ValueError):
        ndi.correlate1d(x, [1.0], axis=2)
    with pytest.raises(ValueError):
        ndi.sobel(x, axis=2)
    with pytest.raises(RuntimeError):
        ndi.uniform_filter1d(x, 0)
    with pytest.raises(ValueError):
        ndi.uniform_filter(x, 3, origin=2)                          # tests/test_filters.py:91-159
    with pytest.raises((ValueError, RuntimeError)):
        ndi.correlate1d(x, [1.0], output=torch.zeros((4, 3), device="cuda"))
    with pytest.raises(RuntimeError):
        ndi.uniform_filter(x, [3, 3, 3])
    with pytest.raises(RuntimeError):
        ndi.correlate1d(x.half(), [1.0])
    with pytest.raises(TypeError):
        ndi.correlate1d(np.zeros(3), [1.0])


def test_modes_sequentially_bit_equal(ndi):
    """tests/test_filters.py:203-225: n-d call with per-axis modes == sequential 1-D calls, bit-equal."""
    arr = to_device(np.array([[1.0, 0.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 0.0]]))
    modes = ["reflect", "wrap"]
    exp = ndi.gaussian_filter1d(arr, 1, axis=0, mode=modes[0])
    exp = ndi.gaussian_filter1d(exp, 1, axis=1, mode=modes[1])
    np.testing.assert_array_equal(to_host(exp), to_host(ndi.gaussian_filter(arr, 1, mode=modes)))
    exp = ndi.uniform_filter1d(arr, 5, axis=0, mode=modes[0])
    exp = ndi.uniform_filter1d(exp, 5, axis=1, mode=modes[1])
    np.testing.assert_array_equal(to_host(exp), to_host(ndi.uniform_filter(arr, 5, mode=modes)))
    for fn, args in [("gaussian_filter", (1,)), ("prewitt", ()), ("sobel", ()), ("laplace", ()),
                     ("gaussian_laplace", (1,)), ("gaussian_gradient_magnitude", (1,)), ("uniform_filter", (5,))]:
        a = getattr(ndi, fn)(arr, *args, mode="reflect")
        b = getattr(ndi, fn)(arr, *args, mode=["reflect", "reflect"])
        np.testing.assert_array_equal(to_host(a), to_host(b), err_msg=fn)


def test_gaussian_truncate(ndi):
    """tests/test_filters.py:313-351: truncate -> support size."""
    import torch
    arr = torch.zeros((100, 100), dtype=torch.float64, device="cuda")
    arr[50, 50] = 1
    assert int((ndi.gaussian_filter(arr, 5, truncate=2) > 0).sum()) == 21 ** 2
    assert int((ndi.gaussian_filter(arr, 5, truncate=5) > 0).sum()) == 51 ** 2
    f = to_host(ndi.gaussian_filter(arr, [0.5, 2.5], truncate=3.5)) > 0
    assert f.any(axis=0).sum() == 19 and f.any(axis=1).sum() == 5
    x = torch.zeros(51, dtype=torch.float64, device="cuda")
    x[25] = 1
    assert int((ndi.gaussian_filter1d(x, sigma=2, truncate=3.5) > 0).sum()) == 15
    for fn in (ndi.gaussian_laplace, ndi.gaussian_gradient_magnitude):
        nz = np.nonzero(to_host(fn(x, sigma=2, truncate=3.5)) != 0)[0]
        assert np.ptp(nz) + 1 == 15


def test_threading(ndi):
    """tests/test_filters.py:354-412: 4 Python threads give the serial answers."""
    rng = np.random.default_rng(2)
    xs = [to_device(rng.random((40, 50)).astype(np.float64)) for _ in range(4)]
    serial = [to_host(ndi.gaussian_filter(x, 1.5)) for x in xs]
    results = [None] * 4

    def work(i):
        import torch
        with torch.cuda.stream(torch.cuda.Stream()):
            results[i] = ndi.gaussian_filter(xs[i], 1.5)
            torch.cuda.current_stream().synchronize()

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for a, b in zip(serial, results):
        np.testing.assert_array_equal(a, to_host(b))


def test_convolve_separable(ndi):
    import cupyimg_b200
    rng = np.random.default_rng(4)
    x = rng.random((8, 9, 10)).astype(np.float64)
    w1, w2 = rng.standard_normal(3), rng.standard_normal(4)
    got = to_host(cupyimg_b200.convolve_separable(to_device(x), w1))
    np.testing.assert_array_equal(got, oracle.convolve_separable(x, w1))
    got = to_host(cupyimg_b200.convolve_separable(to_device(x), [w1, w2], axes=(0, 2), mode="mirror"))
    np.testing.assert_array_equal(got, oracle.convolve_separable(x, [w1, w2], axes=(0, 2), mode="mirror"))
    with pytest.raises(ValueError):
        cupyimg_b200.convolve_separable(to_device(x), [w1, w2], axes=(0,))
    with pytest.raises(ValueError):
        cupyimg_b200.convolve_separable(to_device(x), w1, axes=(3,))


# ---------------------------------------------------------------- float32 kernels
F32_SHAPES = [(33, 47, 70), (64, 64, 64), (5, 300), (130, 9), (1, 1, 517), (3, 128, 256), (700,), (2, 3, 4, 40)]


@pytest.mark.parametrize("shape", F32_SHAPES, ids=[str(s) for s in F32_SHAPES])
@pytest.mark.parametrize("mode", MODES)
def test_f32_correlate1d_all_axes(shape, mode, ndi):
    import zlib
    rng = np.random.default_rng(zlib.crc32(repr((shape, mode)).encode()))
    x = rng.random(shape).astype(np.float32)
    xd = to_device(x)
    for K, origin in [(3, 0), (5, 1), (4, -2), (17, 0), (9, -4), (33, 0), (2, 0)]:
        w = rng.standard_normal(K)
        for axis in range(len(shape)):
            want = oracle.correlate1d(x, w, axis=axis, mode=mode, cval=0.75, origin=origin)
            got = to_host(ndi.correlate1d(xd, w, axis=axis, mode=mode, cval=0.75, origin=origin))
            assert got.dtype == np.float32
            # random-sign taps cancel: float32 accumulation error scales with sum|w| * max|x|, not |result|
            assert_f32_close(got, want, atol=1e-6 * float(np.abs(w).sum()))


@pytest.mark.parametrize("mode", MODES)
def test_f32_filters_match_oracle(mode, ndi):
    rng = np.random.default_rng(21)
    for shape in [(40, 52, 64), (37, 41, 43), (96, 200), (20, 16, 8)]:
        x = rng.random(shape).astype(np.float32)
        xd = to_device(x)
        for name, args, kw in [
            ("gaussian_filter", (2.0,), {}), ("gaussian_filter", (1.0,), {"truncate": 3.0}),
            ("gaussian_filter", (4.0,), {}), ("gaussian_gradient_magnitude", (3.0,), {}),
            ("gaussian_filter", ([1.5, 0.0, 2.5][:len(shape)],), {}),
            ("gaussian_filter", (1.5,), {"order": ([0, 1, 0][:len(shape)])}),
            ("uniform_filter", (5,), {}), ("uniform_filter", (4,), {"origin": -1}),
            ("sobel", (0,), {}), ("prewitt", (-1,), {}),
            ("gaussian_gradient_magnitude", (1.5,), {}),
            ("laplace", (), {}), ("gaussian_laplace", (1.5,), {}),
        ]:
            want = getattr(oracle, name)(x, *args, mode=mode, cval=0.5, **kw)
            got = to_host(getattr(ndi, name)(xd, *args, mode=mode, cval=0.5, **kw))
            assert got.dtype == np.float32 and got.shape == want.shape
            assert_f32_close(got, want, atol_scale=2e-6)


def test_f32_dtype_mode_ndimage_is_exact(ndi):
    """dtype_mode='ndimage' forces scipy's float64 arithmetic: float32 results bit-equal."""
    rng = np.random.default_rng(8)
    x = rng.random((30, 40, 20)).astype(np.float32)
    want = oracle.gaussian_filter(x, 1.5)
    got = to_host(ndi.gaussian_filter(to_device(x), 1.5, dtype_mode="ndimage"))
    np.testing.assert_array_equal(got, want)
    fast = to_host(ndi.gaussian_filter(to_device(x), 1.5))
    np.testing.assert_allclose(fast, got, rtol=1e-4)                  # tests/test_filters_new.py:41-79


def test_f32_unaligned_and_window(ndi):
    """Misaligned base pointers / odd row lengths take the scalar staging path; windows
    (out[p] <-> in[p+offset]) are what the z-slab sharding uses."""
    import torch
    from cupyimg_b200 import _array
    from cupyimg_b200.scipy.ndimage import filters as F
    rng = np.random.default_rng(12)
    big = rng.random(40 * 51 + 3).astype(np.float32)
    bd = to_device(big)
    x, xd = big[1:1 + 40 * 51].reshape(40, 51), bd[1:1 + 40 * 51].view(40, 51)
    w = oracle.gaussian_kernel1d(2.0, 0, 8)
    for axis in (0, 1):
        assert_f32_close(to_host(ndi.correlate1d(xd, w, axis=axis)), oracle.correlate1d(x, w, axis=axis))
    vol = rng.random((30, 24, 36)).astype(np.float32)
    vd = to_device(vol)
    full = oracle.correlate1d(vol, w, axis=0, mode="reflect")
    out = torch.empty((10, 24, 36), dtype=torch.float32, device="cuda")
    spec = F._PassSpec(0, w, 0, F._check_mode("reflect"))
    F._launch_pass(_array.ingest(vd[2:28]), _array.ingest(out), spec, 0.0, False, in_offset=8)
    assert_f32_close(to_host(out), full[10:20])
    # a window at the array start still reflects at index 0
    F._launch_pass(_array.ingest(vd[0:18]), _array.ingest(out), spec, 0.0, False, in_offset=0)
    assert_f32_close(to_host(out), full[0:10])


def test_large_volume_properties(ndi):
    """Size-independent checks at a BASELINE-sized volume (256^3 here; 512^3 in bench):
    DC gain 1 on a constant field, linearity, and sub-brick parity with the oracle."""
    import torch
    n = 256
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand((n, n, n), device="cuda", generator=g)
    y = ndi.gaussian_filter(x, 2.0)
    c = ndi.gaussian_filter(torch.full((n, n, n), 3.25, device="cuda"), 2.0)
    assert float((c - 3.25).abs().max()) < 1e-5
    y2 = ndi.gaussian_filter(2.0 * x + 1.0, 2.0)
    assert float((y2 - (2.0 * y + 1.0)).abs().max()) < 2e-5
    assert abs(float(y.double().mean()) - float(x.double().mean())) < 1e-6   # reflect preserves the mean
    xh = x.cpu().numpy()
    for z0, y0, x0 in [(0, 0, 0), (n - 40, n - 40, n - 40), (100, 0, n - 40), (60, 90, 120)]:
        # brick = the 40^3 region plus an 8-voxel halo where the volume has one; where the
        # region touches the array edge the brick edge IS the array edge, so reflect agrees
        lo = [max(a - 8, 0) for a in (z0, y0, x0)]
        hi = [min(a + 48, n) for a in (z0, y0, x0)]
        brick = xh[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        want = oracle.gaussian_filter(brick, 2.0)
        oz, oy, ox = z0 - lo[0], y0 - lo[1], x0 - lo[2]
        w_sub = want[oz:oz + 40, oy:oy + 40, ox:ox + 40]
        g_sub = y[z0:z0 + 40, y0:y0 + 40, x0:x0 + 40].cpu().numpy()
        assert_f32_close(g_sub, w_sub)


def test_host_pipeline_matches_resident_filter(ndi):
    """Chunked H2D / filter / D2H streaming of a host volume == filtering the resident volume, bit for bit."""
    import torch
    from cupyimg_b200 import host
    g = torch.Generator().manual_seed(5)
    x = torch.rand((150, 48, 64), generator=g).pin_memory()
    for mode in ["reflect", "nearest", "mirror", "constant", "wrap"]:
        want = ndi.gaussian_filter(x.cuda(), 2.0, mode=mode).cpu()
        got = host.gaussian_filter_host(x, 2.0, mode=mode, chunk_planes=32)
        torch.cuda.synchronize()
        assert torch.equal(got, want), mode
    want = ndi.uniform_filter(x.cuda(), [5, 3, 4]).cpu()
    got = host.uniform_filter_host(x, [5, 3, 4], chunk_planes=40)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    xi = (x * 1000).to(torch.int32)
    want = ndi.gaussian_filter(xi.cuda(), 1.5).cpu()
    got = host.gaussian_filter_host(xi, 1.5, chunk_planes=32)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    # chunk edge cases: a last chunk thinner than the halo (150 = 4 x 36 + 6, radius 8: merged into its
    # predecessor), chunks smaller than two halos (raised to 2 r), more chunks than chunk buffers
    want = ndi.gaussian_filter(x.cuda(), 2.0).cpu()
    for c in (36, 8, 16, 20, 149, 150, 400):
        got = host.gaussian_filter_host(x, 2.0, chunk_planes=c)
        torch.cuda.synchronize()
        assert torch.equal(got, want), c


@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror", "wrap"])
def test_f32_radius_9_to_16(mode, ndi):
    """sigma 2.25 .. 4 (radius 9 .. 16).  3-D volumes: ONE launch of the warp-specialised fused kernel on 8-row tiles
    (fused_ws.cuh, round 2; before: three single-axis passes), except `wrap` along y / x (far-side sources: per-axis
    passes).  Stacks of 2-D images (no z pass): two single-axis passes (the fused 2-D kernel for radius 12 / 16 was
    dropped when the single-axis passes overtook it: 0.54 against 0.71 ms for sigma (0, 4, 4) on 512^3)."""
    from cupyimg_b200 import _ffi
    rng = np.random.default_rng(33)
    fused = 1 if mode != "wrap" else 3
    for shape, sig, launches in [((40, 52, 64), lambda s: s, fused), ((33, 70, 132), lambda s: s, fused),
                                 ((12, 70, 200), lambda s: (0, s, s), 2)]:
        x = rng.random(shape).astype(np.float32)
        xd = to_device(x)
        for sigma in (2.25, 2.5, 2.75, 3.0, 3.25, 3.5, 3.75, 4.0):
            want = oracle.gaussian_filter(x, sig(sigma), mode=mode)
            _ffi.LAUNCHES = 0
            got = to_host(ndi.gaussian_filter(xd, sig(sigma), mode=mode))
            assert _ffi.LAUNCHES == launches, (shape, sigma, _ffi.LAUNCHES)
            assert_f32_close(got, want, atol_scale=2e-6)


@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror"])
def test_fused_ws_every_radius(mode, ndi):
    """Every radius 1 .. 8 of the warp-specialised fused kernel (14- / 16-row tiles, fifth Y warp) on shapes with ragged
    tile edges, against the oracle; uniform (box) taps as well as gaussian ones."""
    from cupyimg_b200 import _ffi
    rng = np.random.default_rng(77)
    for shape in [(20, 45, 132), (9, 30, 64), (35, 14, 260)]:
        x = rng.random(shape).astype(np.float32)
        xd = to_device(x)
        for radius in range(1, 9):
            sigma = radius / 4.0
            want = oracle.gaussian_filter(x, sigma, mode=mode)
            _ffi.LAUNCHES = 0
            got = to_host(ndi.gaussian_filter(xd, sigma, mode=mode))
            assert _ffi.LAUNCHES == 1, (shape, radius, _ffi.LAUNCHES)
            assert_f32_close(got, want, atol_scale=2e-6)
            size = 2 * radius + 1
            want = oracle.uniform_filter(x, size, mode=mode)
            got = to_host(ndi.uniform_filter(xd, size, mode=mode))
            assert_f32_close(got, want, atol_scale=2e-6)


@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror", "wrap"])
def test_f32_streaming_single_axis_passes(mode, ndi):
    """csrc/f32_stream.cu: column pass (radius <= 8) and row pass (radius <= 4) without shared memory, on
    aligned shapes with several segments / long rows, origins, short lines, and the windowed call."""
    from cupyimg_b200 import _array
    from cupyimg_b200.scipy.ndimage import filters as F
    import torch
    rng = np.random.default_rng(44)
    for shape, axes in [((3, 700, 64), (0, 1, 2)), ((2, 37, 2056), (1, 2)), ((5, 6, 8), (0, 1, 2)), ((40, 12), (0, 1))]:
        x = rng.random(shape).astype(np.float32)
        xd = to_device(x)
        for axis in axes:
            for radius, order in [(1, 0), (2, 0), (3, 1), (4, 0), (6, 0), (8, 0), (8, 1)]:
                w = oracle.gaussian_kernel1d(max(radius / 3.0, 0.6), order, radius)[::-1].copy()
                for origin in (0, -radius, radius // 2):
                    want = oracle.correlate1d(x, w, axis=axis, mode=mode, cval=0.75, origin=origin)
                    got = to_host(ndi.correlate1d(xd, w, axis=axis, mode=mode, cval=0.75, origin=origin))
                    if order:
                        # derivative taps cancel: the float32 rounding error scales with the INPUT, not with
                        # the (near-zero) output — the reference's own absolute setting
                        assert_f32_close(got, want, atol=1e-6)
                    else:
                        assert_f32_close(got, want, atol_scale=2e-6)
    x = rng.random((64, 24, 64)).astype(np.float32)
    xd = to_device(x)
    w = oracle.gaussian_kernel1d(2.0, 0, 8)[::-1].copy()
    full = to_host(ndi.correlate1d(xd, w, axis=0, mode=mode, cval=0.75))
    spec = F._PassSpec(0, w, 0, F._check_mode(mode))
    for off, n in ((0, 64), (9, 30), (40, 24), (63, 1)):
        out = torch.empty((n, 24, 64), dtype=torch.float32, device="cuda")
        F._launch_pass(_array.ingest(xd), _array.ingest(out), spec, 0.75, False, in_offset=off)
        np.testing.assert_array_equal(to_host(out), full[off:off + n])


def test_gradient_magnitude_mode_sequence(ndi):
    """mode=[m0, m1, m2] means: the derivative along axis a is a full Gaussian filter with mode m_a on EVERY
    axis (reference filters.py:1175-1201), not one mode per filtered axis."""
    rng = np.random.default_rng(51)
    for shape in [(26, 60, 100), (40, 64)]:
        x = rng.random(shape).astype(np.float32)
        xd = to_device(x)
        for mode in (["constant", "mirror", "nearest"][:len(shape)], ["wrap", "reflect", "constant"][:len(shape)],
                     ["nearest"] * len(shape), "mirror"):
            want = oracle.gaussian_gradient_magnitude(x, 1.5, mode=mode)
            got = to_host(ndi.gaussian_gradient_magnitude(xd, 1.5, mode=mode))
            assert_f32_close(got, want, atol=2e-6)
            want = oracle.gaussian_laplace(x, 1.5, mode=mode)
            got = to_host(ndi.gaussian_laplace(xd, 1.5, mode=mode))
            assert_f32_close(got, want, atol=2e-6)


def test_uniform_integer_output_inexact_sums(ndi):
    """Integer outputs stay bit-exact where the window sums round (constant mode with a fractional cval, float
    input): those calls follow scipy's running sum along the line, found by the fuzz test."""
    rng = np.random.default_rng(77)
    x = rng.integers(0, 30000, (32, 53, 256)).astype(np.uint32)
    xd = to_device(x)
    for size, origin in [(3, 0), (6, -1), ([2, 5, 4], 0)]:
        want = oracle.uniform_filter(x, size, mode="constant", cval=7.7, origin=origin)
        got = to_host(ndi.uniform_filter(xd, size, mode="constant", cval=7.7, origin=origin))
        np.testing.assert_array_equal(got, want)
    xf = (rng.standard_normal((40, 300)) * 1000).astype(np.float32)
    for t_out in ("int16", "int32", "uint8"):
        want = oracle.uniform_filter(xf, 5, output=np.dtype(t_out), mode="mirror")
        got = to_host(ndi.uniform_filter(to_device(xf), 5, output=np.dtype(t_out), mode="mirror", dtype_mode="ndimage"))
        np.testing.assert_array_equal(got, want)
    want = oracle.uniform_filter1d(x, 7, axis=1, mode="constant", cval=-0.3)
    got = to_host(ndi.uniform_filter1d(xd, 7, axis=1, mode="constant", cval=-0.3))
    np.testing.assert_array_equal(got, want)
