"""The three single-axis passes of C5's filter (gaussian sigma=4, 33 taps) on 512^3 for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
for _ in range(3):
    ndi.gaussian_filter(x, 4.0, output=o)
torch.cuda.synchronize()
