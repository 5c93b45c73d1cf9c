"""One gaussian_gradient_magnitude(sigma=1.5) launch on 512^3 for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
for _ in range(3):
    ndi.gaussian_gradient_magnitude(x, 1.5, output=o)
torch.cuda.synchronize()
