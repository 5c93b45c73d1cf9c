"""Per-CTA cycle counts of the fused kernel (debug build: SEPFILT_LIB=build/dbg/libsepfilt_dbg.so)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.scipy import ndimage as ndi
mode = sys.argv[1] if len(sys.argv) > 1 else "reflect"
x = torch.rand((512, 512, 512), device="cuda")
out = torch.empty_like(x)
for _ in range(2):
    ndi.gaussian_filter(x, 2.0, output=out, mode=mode)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 148)()
_ffi.lib().sepfilt_debug_cycles(buf, 148)
c = np.array(buf[:]).reshape(37, 4)     # [tile_y][tile_x]
np.set_printoptions(linewidth=200)
print(mode, "cycles/1000 by tile_y (rows) x tile_x (cols)")
print((c / 1000).astype(int)[[0, 1, 2, 17, 34, 35, 36]])
print("min %d max %d mean %d" % (c.min() / 1000, c.max() / 1000, c.mean() / 1000))
