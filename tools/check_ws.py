"""A/B check of the warp-specialised fused kernel (fused_ws.cu) against the round-1 fused kernel
(SEPFILT_NO_WS=1) and the oracle, plus timings.  python tools/check_ws.py [quick]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
from oracle import oracle


def run(fn, ws):
    os.environ["SEPFILT_WS"] = "1"
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


sigma = float(os.environ.get("SIGMA", "2.0"))
bad = 0
shapes = [(40, 48, 64), (33, 9, 16), (64, 30, 132), (100, 77, 260), (20, 200, 128), (150, 14, 384), (70, 129, 516)]
for shape in shapes:
    for mode in ["reflect", "mirror", "nearest", "constant"]:
        g = torch.Generator(device="cuda").manual_seed(sum(shape))
        x = torch.rand(shape, device="cuda", generator=g)
        a = run(lambda: ndi.gaussian_filter(x, sigma, mode=mode), True)
        b = run(lambda: ndi.gaussian_filter(x, sigma, mode=mode), False)
        torch.cuda.synchronize()
        d = float((a - b).abs().max())
        want = oracle.gaussian_filter(x.cpu().numpy(), sigma, mode=mode)
        e = float(np.abs(a.cpu().numpy().astype(np.float64) - want).max())
        flag = "" if (d < 2e-6 and e < 2e-6) else "   <-- MISMATCH"
        bad += bool(flag)
        print("%-18s %-9s ws-vs-r1 %.2e  ws-vs-oracle %.2e%s" % (shape, mode, d, e, flag), flush=True)
print("mismatches:", bad)
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    sys.exit(1 if bad else 0)
for n in (512,):
    x = torch.rand((n, n, n), device="cuda"); o = torch.empty_like(x)
    for mode in ["reflect", "constant", "nearest", "mirror"]:
        tw = run(lambda: timeit(lambda: ndi.gaussian_filter(x, sigma, output=o, mode=mode)), True)
        a = o.clone()
        t1 = run(lambda: timeit(lambda: ndi.gaussian_filter(x, sigma, output=o, mode=mode)), False)
        d = float((a - o).abs().max())
        print("%d^3 sigma %.1f %-9s  ws %.4f ms (%.0f Gvox/s, %.1f%% of 6545 GB/s)   r1 %.4f ms   maxdiff %.2e" % (
            n, sigma, mode, tw, n ** 3 / tw / 1e6, 100 * n ** 3 * 8 / tw / 1e6 / 6545, t1, d), flush=True)
    del x, o
sys.exit(1 if bad else 0)
