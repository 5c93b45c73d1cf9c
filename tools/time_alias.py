"""Does the distance between the input and the output allocation matter?  (512 MiB apart = the bench layout.)"""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
N = 512
buf = torch.empty(3 * N**3 + (64 << 20), device="cuda")     # one arena: x at 0, out at N^3 + pad
x = buf[:N**3].view(N, N, N); x.uniform_()
for pad_kb in (0, 4, 32, 256, 1024, 2048 + 4, 8192 + 36, 32768 + 260):
    off = N**3 + pad_kb * 256          # floats
    out = buf[off:off + N**3].view(N, N, N)
    for _ in range(3):
        ndi.gaussian_filter(x, 2.0, output=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        ndi.gaussian_filter(x, 2.0, output=out)
    b.record(); b.synchronize()
    print("out = x + 512 MiB + %6d KiB   %.4f ms" % (pad_kb, a.elapsed_time(b) / 20))
