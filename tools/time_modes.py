"""Per-axis mode experiments on 512^3 (which edge costs what)."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
n = 512
x = torch.rand((n, n, n), device="cuda")
out = torch.empty_like(x)
for modes in (["constant"] * 3, ["reflect", "constant", "constant"], ["constant", "reflect", "constant"],
              ["constant", "constant", "reflect"], ["reflect"] * 3):
    for _ in range(3):
        ndi.gaussian_filter(x, 2.0, output=out, mode=modes)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ndi.gaussian_filter(x, 2.0, output=out, mode=modes)
    b.record(); b.synchronize()
    print("z,y,x = %-36s %.3f ms" % (modes, a.elapsed_time(b) / 10))
