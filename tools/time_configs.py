"""Throughput of the BASELINE configs that fit one GPU (CUDA events, mean of 10)."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.scipy import ndimage as ndi

def timeit(name, fn, nvox, bytes_per_vox, reps=10):
    for _ in range(3):
        fn()
    _ffi.LAUNCHES = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-58s %8.3f ms %8.1f Gvoxel/s %7.0f GB/s (%4.1f%% of 6545)  launches/call %d" % (
        name, ms, nvox / ms / 1e6, nvox * bytes_per_vox / ms / 1e6, 100 * nvox * bytes_per_vox / ms / 1e6 / 6545, _ffi.LAUNCHES // reps))

x256 = torch.rand((256, 256, 256), device="cuda"); o256 = torch.empty_like(x256)
timeit("C1 uniform_filter size=5 256^3 f32 reflect", lambda: ndi.uniform_filter(x256, 5, output=o256), 256**3, 8)
x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
timeit("   uniform_filter size=5 512^3 f32 reflect", lambda: ndi.uniform_filter(x, 5, output=o), 512**3, 8)
timeit("   gaussian_filter sigma=1 (9 taps) 512^3 f32 reflect", lambda: ndi.gaussian_filter(x, 1.0, output=o), 512**3, 8)
timeit("C2 gaussian_filter sigma=2 (17 taps) 512^3 f32 reflect", lambda: ndi.gaussian_filter(x, 2.0, output=o), 512**3, 8)
timeit("   gaussian_filter sigma=4 (33 taps) 512^3 f32 reflect", lambda: ndi.gaussian_filter(x, 4.0, output=o), 512**3, 8)
timeit("   sobel axis=0 512^3 f32 reflect", lambda: ndi.sobel(x, 0, output=o), 512**3, 8)
timeit("   gaussian_gradient_magnitude sigma=1.5 512^3 f32", lambda: ndi.gaussian_gradient_magnitude(x, 1.5, output=o), 512**3, 8)
timeit("   correlate1d 17 taps axis=2 512^3 f32", lambda: ndi.gaussian_filter1d(x, 2.0, axis=2, output=o), 512**3, 8)
timeit("   correlate1d 17 taps axis=0 512^3 f32", lambda: ndi.gaussian_filter1d(x, 2.0, axis=0, output=o), 512**3, 8)
del x, o
import numpy as np
img = torch.randint(0, 65536, (8, 2048, 2048), device="cuda", dtype=torch.int32).to(torch.uint16)
oi = torch.empty_like(img)
w = np.exp(-0.5 * (np.arange(-4, 5) / 1.5) ** 2); w /= w.sum()
timeit("C3 convolve1d 9 taps axis=1 8x2048^2 u16 mirror (exact)", lambda: ndi.convolve1d(img, w, axis=1, output=oi, mode="mirror"), 8 * 2048 * 2048, 4, reps=5)
timeit("C3 convolve1d 9 taps axis=2 8x2048^2 u16 mirror (exact)", lambda: ndi.convolve1d(img, w, axis=2, output=oi, mode="mirror"), 8 * 2048 * 2048, 4, reps=5)

# SURVEY 8(f) ranks 2 and 3 (general per-element exact kernels)
x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
timeit("   maximum_filter size=5 512^3 f32 (3 exact window passes)", lambda: ndi.maximum_filter(x, 5, output=o), 512**3, 8, reps=5)
timeit("   maximum_filter1d size=5 axis=2 512^3 f32", lambda: ndi.maximum_filter1d(x, 5, axis=2, output=o), 512**3, 8, reps=5)
del x, o
im = torch.rand((8, 2048, 2048), device="cuda"); oi2 = torch.empty_like(im)
k33 = np.arange(9.0).reshape(1, 3, 3) / 36
timeit("   correlate 1x3x3 on 8x2048^2 f32 (exact N-d kernel)", lambda: ndi.correlate(im, k33, output=oi2), 8 * 2048 * 2048, 8, reps=5)
k55 = np.arange(25.0).reshape(1, 5, 5) / 300
timeit("   correlate 1x5x5 on 8x2048^2 f32 (exact N-d kernel)", lambda: ndi.correlate(im, k55, output=oi2), 8 * 2048 * 2048, 8, reps=5)
