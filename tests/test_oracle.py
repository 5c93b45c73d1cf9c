"""Pins the CPU oracle: reference known-answer tables, the reference's exhaustive
length x origin x mode sweep (tests/test_ndimage_vs_scipy.py:24-111), and scipy.ndimage
bit-for-bit (the oracle the reference's own tests use)."""
import itertools

import numpy as np
import pytest
from scipy import ndimage as sndi

from oracle import oracle
from helpers import TYPES, call_filter, check_against_reference, load_kats, load_reference_vectors, run_kat

KATS = load_kats()


def _oracle_call(func, x, **kw):
    if "weights" in kw:
        w = kw.pop("weights")
        return getattr(oracle, func)(x, w, **kw)
    if func == "uniform_filter1d":
        return oracle.uniform_filter1d(x, kw.pop("size"), **kw)
    return getattr(oracle, func)(x, **kw)


@pytest.mark.parametrize("case", KATS, ids=[c["id"] for c in KATS])
def test_reference_known_answers(case):
    run_kat(case, _oracle_call)


@pytest.mark.parametrize("mode", ["constant", "mirror", "nearest", "reflect", "wrap"])
@pytest.mark.parametrize("len_x", [1, 2, 3, 6, 7])
def test_length_origin_sweep_vs_scipy(mode, len_x):
    """Every filter length up to 2*len_x+1 (multi-reflection) x every valid origin."""
    x = np.arange(1, 1 + len_x, dtype=np.float64)
    for len_h in range(1, 2 * len_x + 2):
        h = np.arange(1, 1 + len_h, dtype=np.float64)
        lo, hi = -(len_h // 2), (len_h - 1) // 2
        for origin in range(lo, hi + 1):
            for fn in ("correlate1d", "convolve1d"):
                want = getattr(sndi, fn)(x, h, mode=mode, cval=0.25, origin=origin)
                got = getattr(oracle, fn)(x, h, mode=mode, cval=0.25, origin=origin)
                np.testing.assert_array_equal(got, want)
        for origin in (lo - 1, hi + 1):
            with pytest.raises(ValueError):
                oracle.correlate1d(x, h, mode=mode, origin=origin)


def test_remap_tables():
    """_util.py:170-228 index rules, spelled out (d c b a | a b c d | d c b a etc.)."""
    n = 4
    want = {
        "reflect": [3, 2, 1, 0, 0, 1, 2, 3, 3, 2, 1, 0],
        "mirror": [2, 3, 2, 1, 0, 1, 2, 3, 2, 1, 0, 1],
        "nearest": [0, 0, 0, 0, 0, 1, 2, 3, 3, 3, 3, 3],
        "wrap": [0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3],
        "constant": [-1, -1, -1, -1, 0, 1, 2, 3, -1, -1, -1, -1],
    }
    for mode, row in want.items():
        assert [oracle.remap(mode, i, n) for i in range(-4, 8)] == row
    assert oracle.remap("mirror", -5, 1) == 0


@pytest.mark.parametrize("t_in,t_out", list(itertools.product(TYPES, TYPES)))
def test_dtype_matrix_bit_exact_vs_scipy(t_in, t_out):
    rng = np.random.default_rng(hash((t_in, t_out)) % 2**32)
    x = (rng.random((5, 7, 6)) * 100).astype(t_in)
    taps = [oracle.gaussian_kernel1d(1.0, 0, 4), oracle.gaussian_kernel1d(1.0, 1, 4),
            rng.standard_normal(4), np.array([1.0, 2.0, 1.0])]
    for w, axis in itertools.product(taps, range(3)):
        want = sndi.correlate1d(x, w, axis=axis, output=t_out, mode="mirror")
        got = oracle.correlate1d(x, w, axis=axis, output=t_out, mode="mirror")
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("dtype", ["float32", "float64", "uint16", "int32", "uint8"])
def test_composite_filters_bit_exact_vs_scipy(dtype):
    rng = np.random.default_rng(7)
    x = (rng.random((12, 17, 9)) * 200).astype(dtype)
    for name, args, kw in [
        ("gaussian_filter", (1.5,), {}), ("gaussian_filter", ([1.0, 0.0, 2.0],), {"order": [0, 0, 1]}),
        ("uniform_filter", (5,), {}), ("uniform_filter", ([3, 1, 4],), {"origin": [0, 0, -1], "mode": "wrap"}),
        ("sobel", (0,), {}), ("prewitt", (-1,), {"mode": ["reflect", "wrap", "mirror"]}),
        ("gaussian_gradient_magnitude", (1.5,), {}),
        ("laplace", (), {}), ("gaussian_laplace", (1.2,), {"mode": "nearest"}),
        ("gaussian_filter1d", (2.0,), {"axis": 1, "order": 2, "mode": "constant", "cval": 3.0}),
    ]:
        want = getattr(sndi, name)(x, *args, **kw)
        got = getattr(oracle, name)(x, *args, **kw)
        np.testing.assert_array_equal(got, want, err_msg=name)


def test_gaussian_taps_equal_scipy_and_closed_forms():
    from scipy.ndimage._filters import _gaussian_kernel1d
    for sigma, order in itertools.product([0.5, 1.0, 1.5, 2.0, 4.0, 7.3], range(4)):
        lw = int(4 * sigma + 0.5)
        np.testing.assert_array_equal(oracle.gaussian_kernel1d(sigma, order, lw),
                                      _gaussian_kernel1d(sigma, order, lw))
    # closed forms, reference tests/test_filters.py:61-77
    radius, sigma = 10, 2
    s2 = sigma * sigma
    x = np.arange(-radius, radius + 1, dtype=np.double)
    phi = np.exp(-0.5 * x * x / s2)
    phi /= phi.sum()
    np.testing.assert_allclose(phi, oracle.gaussian_kernel1d(sigma, 0, radius))
    np.testing.assert_allclose(-phi * x / s2, oracle.gaussian_kernel1d(sigma, 1, radius))
    np.testing.assert_allclose(phi * (x * x / s2 - 1) / s2, oracle.gaussian_kernel1d(sigma, 2, radius))
    np.testing.assert_allclose(phi * (3 - x * x / s2) * x / (s2 * s2), oracle.gaussian_kernel1d(sigma, 3, radius))


def test_window_semantics():
    """out[p] <-> in[p + in_offset]: the slab-with-halo form the sharded path uses."""
    rng = np.random.default_rng(3)
    x = rng.random((20, 6, 5)).astype(np.float32)
    w = oracle.gaussian_kernel1d(1.0, 0, 4)
    full = oracle.correlate1d(x, w, axis=0)
    ext = x[3:17]                                   # slab 7..13 with 4 halo planes either side
    out = np.empty((6, 6, 5), np.float32)
    oracle._line_pass(ext, out, 0, w, w.size, 0, "reflect", 0.0, in_offset=4)
    np.testing.assert_array_equal(out, full[7:13])


def test_reference_executed_vectors():
    """The oracle against outputs of the reference ITSELF, executed on the CPU by compiling the kernel
    source its own generator emits (tests/golden/make_reference_vectors.py)."""
    n = {"checked": 0, "skipped": 0}
    for func, kwargs, x, ref in load_reference_vectors():
        got = call_filter(oracle, func, x, kwargs)
        n[check_against_reference(func, x, got, ref)] += 1
    assert n["checked"] > 1000 and n["skipped"] < 0.02 * n["checked"]


def test_complex_oracle_matches_scipy():
    """Complex input and / or weights (the reference's test_correlate1d_complex,
    tests/test_ndimage_vs_scipy.py:114-126, sweeps exactly this dtype matrix): the oracle's
    restatement by real components against scipy, bit for bit, with non-trivial imaginary parts."""
    import scipy.ndimage as sn
    rng = np.random.default_rng(0)
    types = [np.float32, np.float64, np.complex64, np.complex128]
    for dx, dh in itertools.product(types, types):
        cx, ch = np.dtype(dx).kind == "c", np.dtype(dh).kind == "c"
        if not (cx or ch):
            continue
        x = (rng.standard_normal((5, 9)) + (1j * rng.standard_normal((5, 9)) if cx else 0)).astype(dx)
        for len_h in (1, 2, 5, 12):
            h = (rng.standard_normal(len_h) + (1j * rng.standard_normal(len_h) if ch else 0)).astype(dh)
            for mode in ("reflect", "constant", "nearest", "mirror", "wrap"):
                cval = (0.5 - 2j) if cx else 0.5
                for fn in ("correlate1d", "convolve1d"):
                    got = getattr(oracle, fn)(x, h, axis=1, mode=mode, cval=cval)
                    want = getattr(sn, fn)(x, h, axis=1, mode=mode, cval=cval)
                    assert got.dtype == want.dtype
                    np.testing.assert_array_equal(got, want)
    with pytest.raises(RuntimeError):
        oracle.correlate1d(np.ones(4, np.complex64), [1.0, 2.0], output=np.float64)
    with pytest.raises(ValueError):
        oracle.correlate1d(np.ones(4), np.array([1.0, 2.0j]), cval=1j, mode="constant")


# the reference's literal known answers for the separable min / max filters
# (tests/test_ndimage.py:754-795 minimum_filter01-06, :830-873 maximum_filter01-06)
MINMAX_KATS = [
    ("minimum_filter", [1, 2, 3, 4, 5], [2], [1, 1, 2, 3, 4]),
    ("minimum_filter", [1, 2, 3, 4, 5], [3], [1, 1, 2, 3, 4]),
    ("minimum_filter", [3, 2, 5, 1, 4], [2], [3, 2, 2, 1, 1]),
    ("minimum_filter", [3, 2, 5, 1, 4], [3], [2, 2, 1, 1, 1]),
    ("minimum_filter", [[3, 2, 5, 1, 4], [7, 6, 9, 3, 5], [5, 8, 3, 7, 1]], [2, 3],
     [[2, 2, 1, 1, 1], [2, 2, 1, 1, 1], [5, 3, 3, 1, 1]]),
    ("maximum_filter", [1, 2, 3, 4, 5], [2], [1, 2, 3, 4, 5]),
    ("maximum_filter", [1, 2, 3, 4, 5], [3], [2, 3, 4, 5, 5]),
    ("maximum_filter", [3, 2, 5, 1, 4], [2], [3, 3, 5, 5, 4]),
    ("maximum_filter", [3, 2, 5, 1, 4], [3], [3, 5, 5, 5, 4]),
    ("maximum_filter", [[3, 2, 5, 1, 4], [7, 6, 9, 3, 5], [5, 8, 3, 7, 1]], [2, 3],
     [[3, 5, 5, 5, 4], [7, 9, 9, 9, 5], [8, 9, 9, 9, 7]]),
]


def test_minmax_oracle_known_answers_and_scipy():
    """minimum / maximum filters.  Reference known answers, then scipy bit for bit
    over dtype x mode x size x origin x axis (the oracle follows the reference's kernel: compare as
    double, C-cast store)."""
    for fn, x, size, want in MINMAX_KATS:
        np.testing.assert_array_equal(getattr(oracle, fn)(np.asarray(x), size), np.asarray(want))
    rng = np.random.default_rng(0)
    for dt in ["uint8", "int16", "uint16", "int64", "float32", "float64"]:
        x = (rng.random((4, 9, 6)) * 200 - (0 if dt[0] == "u" else 50)).astype(dt)
        for mode, size, fn in itertools.product(["reflect", "constant", "nearest", "mirror", "wrap"],
                                                [1, 2, 3, 7, 12], ["minimum", "maximum"]):
            for origin in sorted({-(size // 2), 0, (size - 1) // 2}):
                for axis in (0, 1, 2):
                    got = getattr(oracle, fn + "_filter1d")(x, size, axis=axis, mode=mode, cval=3.0, origin=origin)
                    want = getattr(sndi, fn + "_filter1d")(x, size, axis=axis, mode=mode, cval=3.0, origin=origin)
                    assert got.dtype == want.dtype
                    np.testing.assert_array_equal(got, want)
            got = getattr(oracle, fn + "_filter")(x, size=[size, 1, 3], mode=mode, cval=-2.0, output=np.float32)
            want = getattr(sndi, fn + "_filter")(x, size=[size, 1, 3], mode=mode, cval=-2.0, output=np.float32)
            np.testing.assert_array_equal(got, want)


# the reference's literal known answers for dense N-d correlate / convolve (tests/test_ndimage.py:214-341):
# (input, kernel, kwargs, expected correlate, expected convolve); "all" sweeps the 10 x 10 dtype matrix
A23 = [[1, 2, 3], [4, 5, 6]]
CORRELATE_ND_KATS = [
    (A23, [[1, 1], [1, 1]], {}, [[4, 6, 10], [10, 12, 16]], [[12, 16, 18], [18, 22, 24]], None),          # correlate11
    (A23, [[1, 0], [0, 1]], {}, [[2, 3, 5], [5, 6, 8]], [[6, 8, 9], [9, 11, 12]], "all"),                  # correlate12-14
    (A23, [[0.5, 0], [0, 0.5]], {"output": "float32"}, [[1, 1.5, 2.5], [2.5, 3, 4]], [[3, 4, 4.5], [4.5, 5.5, 6]], "in"),   # correlate16
    ([1, 2, 3], [1, 1], {"origin": -1}, [3, 5, 6], [2, 3, 5], None),                                        # correlate17
    (A23, [[1, 0], [0, 1]], {"output": "float32", "mode": "nearest", "origin": -1},
     [[6, 8, 9], [9, 11, 12]], [[2, 3, 5], [5, 6, 8]], "in"),                                               # correlate18
    (A23, [[1, 0], [0, 1]], {"output": "float32", "mode": "nearest", "origin": [-1, 0]},
     [[5, 6, 8], [8, 9, 11]], [[3, 5, 6], [6, 8, 9]], "in"),                                                # correlate19
]


def run_correlate_nd_kats(ns, to_dev=lambda a: a, to_np=lambda a: a):
    for x, k, kw, want_cor, want_cov, sweep in CORRELATE_ND_KATS:
        in_types = TYPES if sweep in ("all", "in") else ["int64"]
        out_types = TYPES if sweep == "all" else [None]
        for t1 in in_types:
            for t2 in out_types:
                kw2 = dict(kw)
                if t2 is not None:
                    kw2["output"] = np.dtype(t2)
                elif "output" in kw2:
                    kw2["output"] = np.dtype(kw2["output"])
                for fn, want in (("correlate", want_cor), ("convolve", want_cov)):
                    got = to_np(getattr(ns, fn)(to_dev(np.asarray(x, dtype=t1)), np.asarray(k, float), **kw2))
                    np.testing.assert_array_equal(got.astype(np.float64), np.asarray(want, float), err_msg=fn)
                    if "output" in kw2:
                        assert got.dtype == kw2["output"]


def test_correlate_nd_oracle_known_answers_and_scipy():
    """the oracle's N-d correlate against the reference's known answers and scipy,
    bit for bit (integer dtypes included: scipy's tap order and cast)."""
    run_correlate_nd_kats(oracle)
    e = oracle.correlate(np.zeros((1, 0)), np.ones((1, 2)))          # correlate10: empty in, empty out
    assert e.shape == (1, 0)
    rng = np.random.default_rng(0)
    for dt in ["uint8", "int16", "int64", "float32", "float64"]:
        for shape, wshapes in [((9, 11), [(3, 3), (2, 4), (5, 1)]), ((5, 6, 7), [(3, 3, 3), (2, 1, 4)]), ((12,), [(5,), (4,)])]:
            x = (rng.random(shape) * 100 - (0 if dt[0] == "u" else 30)).astype(dt)
            for ws in wshapes:
                w = rng.standard_normal(ws)
                w[tuple(0 for _ in ws)] = 0.0
                for mode in ["reflect", "constant", "nearest", "mirror", "wrap"]:
                    for origin in (0, [-(s // 2) for s in ws], [(s - 1) // 2 for s in ws]):
                        for fn in ("correlate", "convolve"):
                            got = getattr(oracle, fn)(x, w, mode=mode, cval=2.5, origin=origin)
                            want = getattr(sndi, fn)(x, w, mode=mode, cval=2.5, origin=origin)
                            assert got.dtype == want.dtype
                            np.testing.assert_array_equal(got, want)
    with pytest.raises(RuntimeError):
        oracle.correlate(np.ones((3, 3)), np.ones(3))
    with pytest.raises(ValueError):
        oracle.correlate(np.ones((3, 3)), np.ones((3, 3)), origin=2)
