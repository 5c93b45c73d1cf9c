"""Dense 2-D correlate launches for ncu (python tools/prof_corr.py): 1x3x3 and 1x5x5 on 8 x 2048^2 f32, 3x3 on u16."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
im = torch.rand((8, 2048, 2048), device="cuda"); out = torch.empty_like(im)
k33 = np.arange(9.0).reshape(1, 3, 3) / 36
k55 = np.arange(25.0).reshape(1, 5, 5) / 300
for _ in range(3):
    ndi.correlate(im, k33, output=out)
    ndi.correlate(im, k55, output=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, k in (("3x3", k33), ("5x5", k55)):
    a.record()
    for _ in range(20):
        ndi.correlate(im, k, output=out)
    b.record(); b.synchronize()
    print("%s  %.4f ms per call (20 calls back to back)" % (name, a.elapsed_time(b) / 20))
