"""One fused gaussian_filter launch on n^3 for ncu (python tools/prof_fused.py [n] [sigma] [mode])."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
mode = sys.argv[3] if len(sys.argv) > 3 else "reflect"
x = torch.rand((n, n, n), device="cuda")
out = torch.empty_like(x)
for _ in range(3):
    ndi.gaussian_filter(x, sigma, output=out, mode=mode)
torch.cuda.synchronize()
