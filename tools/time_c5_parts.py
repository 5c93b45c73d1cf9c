"""C5's filter (gaussian sigma=4, 33 taps) on 512^3, pass by pass."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.scipy import ndimage as ndi

def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    _ffi.LAUNCHES = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-52s %.3f ms  launches/call %d  %.1f%% of 6545 GB/s at 8 B/voxel" % (
        name, ms, _ffi.LAUNCHES // reps, 100 * x.numel() * 8 / ms / 1e6 / 6545), flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.rand((n, n, n), device="cuda"); o = torch.empty_like(x)
for s in [(4, 4, 4), (4, 0, 0), (0, 4, 0), (0, 0, 4), (0, 4, 4), (4, 4, 0), (3, 3, 3), (2.5, 2.5, 2.5), (2, 2, 2)]:
    timeit("gaussian_filter sigma=%s" % (s,), lambda: ndi.gaussian_filter(x, s, output=o))
