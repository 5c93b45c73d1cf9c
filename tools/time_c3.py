"""C3 timing: convolve1d 9 taps on an 8 x 2048 x 2048 stack (one GPU's share), axes 1 and 2, a few dtypes.
SEPFILT_LIB selects the library build under test."""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi

def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps

w = np.exp(-0.5 * (np.arange(-4, 5) / 1.5) ** 2); w /= w.sum()
tag = os.path.basename(os.environ.get("SEPFILT_LIB", "default"))
npx = 8 * 2048 * 2048
for dt in (torch.uint16, torch.uint8, torch.float64):
    if dt == torch.float64:
        img = torch.rand((8, 2048, 2048), device="cuda", dtype=dt)
    else:
        img = torch.randint(0, 256 if dt == torch.uint8 else 65536, (8, 2048, 2048), device="cuda", dtype=torch.int32).to(dt)
    oi = torch.empty_like(img)
    es = img.element_size()
    for axis in (1, 2):
        ms = timeit(lambda: ndi.convolve1d(img, w, axis=axis, output=oi, mode="mirror", dtype_mode="ndimage"))
        print("%-22s %-8s axis %d  %.4f ms  %7.1f Gpx/s  %6.0f GB/s (%4.1f%% of 6545)" % (
            tag, str(dt).replace("torch.", ""), axis, ms, npx / ms / 1e6, npx * 2 * es / ms / 1e6, 100 * npx * 2 * es / ms / 1e6 / 6545))
