"""One-GPU timings of the north_star scale configs at their stated size (CUDA events):
C4 gaussian_gradient_magnitude sigma=1.5 on 1024^3, C5 gaussian_filter sigma=4 on 2048^3,
C3 convolve1d x2 on 64 x 2048^2 uint16.   python tools/time_at_size.py [c4] [c5] [c3]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.scipy import ndimage as ndi

which = set(sys.argv[1:]) or {"c3", "c4", "c5"}


def timeit(name, fn, nvox, bytes_per_vox, reps):
    for _ in range(2):
        fn()
    _ffi.LAUNCHES = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-62s %9.3f ms %8.1f Gvoxel/s %7.0f GB/s (%4.1f%% of 6545)  launches/call %d" % (
        name, ms, nvox / ms / 1e6, nvox * bytes_per_vox / ms / 1e6, 100 * nvox * bytes_per_vox / ms / 1e6 / 6545,
        _ffi.LAUNCHES // reps), flush=True)


if "c4" in which:
    n = 1024
    x = torch.rand((n, n, n), device="cuda"); o = torch.empty_like(x)
    timeit("C4 gaussian_gradient_magnitude sigma=1.5 1024^3 f32", lambda: ndi.gaussian_gradient_magnitude(x, 1.5, output=o), n ** 3, 8, 5)
    timeit("   gaussian_filter sigma=2 1024^3 f32", lambda: ndi.gaussian_filter(x, 2.0, output=o), n ** 3, 8, 5)
    del x, o
if "c5" in which:
    n = 2048
    x = torch.empty((n, n, n), device="cuda")
    for z in range(0, n, 256):
        x[z:z + 256] = torch.rand((256, n, n), device="cuda")
    o = torch.empty_like(x)
    timeit("C5 gaussian_filter sigma=4 (33 taps) 2048^3 f32", lambda: ndi.gaussian_filter(x, 4.0, output=o), n ** 3, 8, 3)
    timeit("   gaussian_filter sigma=2 (17 taps) 2048^3 f32", lambda: ndi.gaussian_filter(x, 2.0, output=o), n ** 3, 8, 3)
    del x, o
    torch.cuda.empty_cache()
if "c3" in which:
    w = np.exp(-0.5 * (np.arange(-4, 5) / 1.5) ** 2); w /= w.sum()
    for nimg in (8, 64):
        img = torch.randint(0, 65536, (nimg, 2048, 2048), device="cuda", dtype=torch.int32).to(torch.uint16)
        t = torch.empty_like(img); oi = torch.empty_like(img)
        timeit("C3 convolve1d 9 taps axis=1 %dx2048^2 u16 mirror" % nimg, lambda: ndi.convolve1d(img, w, axis=1, output=t, mode="mirror"), nimg * 2048 * 2048, 4, 5)
        timeit("C3 convolve1d 9 taps axis=2 %dx2048^2 u16 mirror" % nimg, lambda: ndi.convolve1d(t, w, axis=2, output=oi, mode="mirror"), nimg * 2048 * 2048, 4, 5)
        del img, t, oi
