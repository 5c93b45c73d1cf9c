"""A/B check of the single-launch gradient magnitude (fused_ws.cu, GRAD) against the three-launch
path (SEPFILT_NO_WS=1) and the oracle, plus timings.  python tools/check_grad.py [quick]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.scipy import ndimage as ndi
from oracle import oracle


def run(fn, ws):
    if ws:
        os.environ.pop("SEPFILT_NO_WS", None)
    else:
        os.environ["SEPFILT_NO_WS"] = "1"
    try:
        return fn()
    finally:
        os.environ.pop("SEPFILT_NO_WS", None)


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps


bad = 0
shapes = [(40, 48, 64), (33, 9, 16), (64, 30, 132), (100, 77, 260), (20, 200, 128), (150, 14, 384), (70, 129, 516)]
for sigma in (1.5, 1.0, 0.5, 1.25, 0.75):
    for shape in shapes:
        for mode in ["reflect", "mirror", "nearest", "constant", "wrap"]:
            if sigma != 1.5 and mode in ("mirror", "nearest"):
                continue
            g = torch.Generator(device="cuda").manual_seed(sum(shape))
            x = torch.rand(shape, device="cuda", generator=g)
            _ffi.LAUNCHES = 0
            a = run(lambda: ndi.gaussian_gradient_magnitude(x, sigma, mode=mode), True)
            nl = _ffi.LAUNCHES
            b = run(lambda: ndi.gaussian_gradient_magnitude(x, sigma, mode=mode), False)
            torch.cuda.synchronize()
            d = float((a - b).abs().max())
            want = oracle.gaussian_gradient_magnitude(x.cpu().numpy(), sigma, mode=mode)
            e = float(np.abs(a.cpu().numpy().astype(np.float64) - want).max())
            flag = "" if (d < 2e-6 and e < 2e-6) else "   <-- MISMATCH"
            bad += bool(flag)
            print("s=%.2f %-18s %-9s launches %d  1-launch vs 3-launch %.2e  vs oracle %.2e%s" % (
                sigma, shape, mode, nl, d, e, flag), flush=True)
print("mismatches:", bad)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(1 if bad else 0)
for n in (512, 1024):
    x = torch.rand((n, n, n), device="cuda"); o = torch.empty_like(x)
    for sigma, mode in [(1.5, "reflect"), (1.5, "constant"), (1.0, "reflect")]:
        tw = run(lambda: timeit(lambda: ndi.gaussian_gradient_magnitude(x, sigma, output=o, mode=mode)), True)
        a = o.clone() if n == 512 else None
        t1 = run(lambda: timeit(lambda: ndi.gaussian_gradient_magnitude(x, sigma, output=o, mode=mode)), False)
        d = float((a - o).abs().max()) if a is not None else float("nan")
        print("%d^3 gradmag sigma %.1f %-9s  1-launch %.4f ms (%.0f Gvox/s, %.1f%% of 6545 GB/s)   3-launch %.4f ms   maxdiff %.2e" % (
            n, sigma, mode, tw, n ** 3 / tw / 1e6, 100 * n ** 3 * 8 / tw / 1e6 / 6545, t1, d), flush=True)
    del x, o
sys.exit(1 if bad else 0)
