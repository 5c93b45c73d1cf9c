"""Per-launch view of gaussian_gradient_magnitude sigma=1.5 on 512^3 f32 next to the plain filters."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi

def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    print("%-60s %.3f ms" % (name, a.elapsed_time(b) / reps))

x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
timeit("gaussian_filter sigma=1.5 (13 taps)", lambda: ndi.gaussian_filter(x, 1.5, output=o))
timeit("gaussian_filter sigma=1.5 order=(1,0,0)", lambda: ndi.gaussian_filter(x, 1.5, order=(1, 0, 0), output=o))
timeit("gaussian_filter sigma=0.75 (7 taps)", lambda: ndi.gaussian_filter(x, 0.75, output=o))
timeit("gaussian_gradient_magnitude sigma=1.5", lambda: ndi.gaussian_gradient_magnitude(x, 1.5, output=o))
timeit("gaussian_gradient_magnitude sigma=1.0", lambda: ndi.gaussian_gradient_magnitude(x, 1.0, output=o))
timeit("gaussian_gradient_magnitude sigma=2.0", lambda: ndi.gaussian_gradient_magnitude(x, 2.0, output=o))
