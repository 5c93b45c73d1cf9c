import sys
import numpy as np, torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
img = torch.randint(0, 65536, (8, 2048, 2048), device="cuda", dtype=torch.int32).to(torch.uint16)
oi = torch.empty_like(img)
w = np.exp(-0.5 * (np.arange(-4, 5) / 1.5) ** 2); w /= w.sum()
for _ in range(3):
    ndi.convolve1d(img, w, axis=1, output=oi, mode="mirror")
    ndi.convolve1d(img, w, axis=2, output=oi, mode="mirror")
torch.cuda.synchronize()
