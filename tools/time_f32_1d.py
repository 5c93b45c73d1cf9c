"""Single-axis float32 passes on 512^3 (correlate1d through gaussian_filter1d), every axis, several radii."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
for sigma in (0.5, 1.0, 2.0, 3.0, 4.0):
    for axis in (0, 1, 2):
        for _ in range(3):
            ndi.gaussian_filter1d(x, sigma, axis=axis, output=o)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ndi.gaussian_filter1d(x, sigma, axis=axis, output=o)
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / 10
        print("gaussian_filter1d sigma=%.1f (%2d taps) axis %d  %.3f ms  %5.0f GB/s (%4.1f%% of 6545)" % (
            sigma, 2 * int(4 * sigma + 0.5) + 1, axis, ms, 512**3 * 8 / ms / 1e6, 100 * 512**3 * 8 / ms / 1e6 / 6545))
