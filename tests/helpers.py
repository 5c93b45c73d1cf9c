"""Shared test plumbing: the KAT runner and the numpy <-> device adapters."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TYPES = ["int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64"]


def load_kats():
    with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
        return json.load(f)["cases"]


def run_kat(case, call):
    """``call(func_name, input_ndarray, **kwargs) -> ndarray`` is the implementation under test."""
    in_types = TYPES if case["in_dtype"] == "all" else [case["in_dtype"] or "int64"]
    out_types = TYPES if case["out_dtype"] == "all" else [case["out_dtype"]]
    expected = np.asarray(case["expected"], dtype=np.float64)
    for t1 in in_types:
        x = np.asarray(case["input"], dtype=t1)
        for t2 in out_types:
            kw = dict(case["args"])
            if t2 is not None:
                kw["output"] = np.dtype(t2)
            got = call(case["func"], x, **kw)
            want_dtype = np.dtype(t2) if t2 is not None else x.dtype
            assert got.dtype == want_dtype, (case["id"], t1, t2, got.dtype)
            assert got.shape == expected.shape or (got.size == 0 and expected.size == 0), case["id"]
            if got.size == 0:
                continue
            if case["decimal"] is None:
                np.testing.assert_array_equal(got.astype(np.float64), expected, err_msg=case["id"])
            else:
                np.testing.assert_array_almost_equal(got.astype(np.float64), expected,
                                                     decimal=case["decimal"], err_msg=case["id"])


def to_device(x, device="cuda"):
    """numpy -> torch CUDA tensor, including the unsigned types torch cannot convert directly."""
    import torch
    x = np.ascontiguousarray(x)
    if x.dtype in (np.dtype("uint16"), np.dtype("uint32"), np.dtype("uint64")):
        signed = x.view(x.dtype.str.replace("u", "i"))
        t = torch.from_numpy(signed.copy()).to(device)
        return t.view({2: torch.uint16, 4: torch.uint32, 8: torch.uint64}[x.dtype.itemsize])
    return torch.from_numpy(x.copy()).to(device)


def to_host(t):
    """torch tensor -> numpy, including unsigned types."""
    import torch
    if t.dtype in (torch.uint16, torch.uint32, torch.uint64):
        signed = {torch.uint16: torch.int16, torch.uint32: torch.int32, torch.uint64: torch.int64}[t.dtype]
        a = t.contiguous().view(signed).cpu().numpy()
        return a.view(a.dtype.str.replace("i", "u"))
    return t.cpu().numpy()


def gpu_call(func, x, **kw):
    """Run one product-API function on the GPU with a numpy input, numpy result."""
    import cupyimg_b200
    from cupyimg_b200.scipy import ndimage as ndi
    fn = getattr(ndi, func, None) or getattr(cupyimg_b200, func)
    out = fn(to_device(x), **kw)
    return to_host(out)


def load_reference_vectors():
    """Cases produced by EXECUTING the reference on the CPU (tests/golden/make_reference_vectors.py):
    yields (func, kwargs, x, y_reference)."""
    path = os.path.join(HERE, "golden", "reference_vectors.npz")
    d = np.load(path)
    index = json.loads(str(d["index"]))
    xb, yb = d["x"].tobytes(), d["y"].tobytes()

    def get(blob, ref):
        off, shape, dt = ref
        dt = np.dtype(dt)
        n = int(np.prod(shape, dtype=np.int64))
        return np.frombuffer(blob, dt, n, off).reshape(shape).copy()
    for e in index:
        yield e["func"], e["kwargs"], get(xb, e["x"]), get(yb, e["y"])


def call_filter(ns, func, x, kwargs):
    """Call ``ns.<func>`` with the positional conventions of the scipy.ndimage API."""
    kw = dict(kwargs)
    fn = getattr(ns, func)
    if "weights" in kw:
        return fn(x, np.asarray(kw.pop("weights"), np.float64), **kw)
    for first in ("size", "sigma"):
        if first in kw:
            return fn(x, kw.pop(first), **kw)
    return fn(x, **kw)


def check_against_reference(func, x, got, ref):
    """How close an implementation that follows scipy must be to the executed reference.
    Exact where the two share arithmetic; otherwise the reference's documented deviations
    (SURVEY.md App. D): ascending summation order (float64 ulps), Gaussian taps rounded to float32
    for <= 32-bit inputs, 1/size uniform taps (integer results knowingly off: its own xfail test)."""
    assert got.dtype == ref.dtype and got.shape == ref.shape
    kind = ref.dtype.kind
    gaussian = func.startswith("gaussian")
    if func == "uniform_filter" and kind in "iu":
        return "skipped"
    if func == "gaussian_gradient_magnitude" and kind == "u":
        return "skipped"              # negative derivatives wrap in unsigned types: chaotic under tap rounding
    if kind in "iu":
        if gaussian:
            # float32-rounded taps may flip a truncation at an exact integer boundary
            diff = np.abs(got.astype(np.int64) - ref.astype(np.int64))
            assert diff.max() <= 1 and (diff != 0).mean() < 0.02, (func, diff.max())
        else:
            np.testing.assert_array_equal(got, ref)
    elif ref.dtype == np.float32:
        if gaussian or func == "uniform_filter":
            np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-5 * float(np.abs(ref).max()))
        else:
            np.testing.assert_array_equal(got, ref)
    else:
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * float(np.abs(ref).max()))
    return "checked"
