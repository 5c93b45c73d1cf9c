"""Time gaussian_filter on n^3 f32 for several boundary modes (CUDA events, mean of 10)."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
x = torch.rand((n, n, n), device="cuda")
out = torch.empty_like(x)
for mode in sys.argv[3:] or ["reflect", "constant", "nearest", "wrap", "mirror"]:
    for _ in range(3):
        ndi.gaussian_filter(x, sigma, output=out, mode=mode)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ndi.gaussian_filter(x, sigma, output=out, mode=mode)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 10
    print("%-9s n=%d sigma=%g  %.3f ms  %.1f Gvoxel/s  %.0f GB/s" % (mode, n, sigma, ms, n**3 / ms / 1e6, n**3 * 8 / ms / 1e6))
