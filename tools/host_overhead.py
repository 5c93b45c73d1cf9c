"""Host-side cost of one API call (tiny volume, so the GPU work is negligible)."""
import sys, time
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
x = torch.rand((32, 32, 64), device="cuda"); out = torch.empty_like(x)
for name, fn in [("gaussian_filter (fused)", lambda: ndi.gaussian_filter(x, 2.0, output=out)),
                 ("gaussian_filter alloc out", lambda: ndi.gaussian_filter(x, 2.0)),
                 ("correlate1d f32", lambda: ndi.correlate1d(x, [1.0, 2.0, 1.0], output=out)),
                 ("uniform_filter", lambda: ndi.uniform_filter(x, 5, output=out))]:
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(500): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 500
    print("%-28s %.1f us per call" % (name, dt * 1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(300): ndi.gaussian_filter(x, 2.0, output=out)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
