"""Per-CTA cycle counts of the fused kernel (debug build: SEPFILT_LIB=build/dbg/libsepfilt_dbg.so)."""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.scipy import ndimage as ndi
modes = sys.argv[1:4] if len(sys.argv) > 3 else [sys.argv[1] if len(sys.argv) > 1 else "reflect"] * 3
x = torch.rand((512, 512, 512), device="cuda")
out = torch.empty_like(x)
for _ in range(2):
    ndi.gaussian_filter(x, 2.0, output=out, mode=modes)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 4096)()
_ffi.lib().sepfilt_debug_cycles(buf, 4096)
allc = np.array(buf[:])
smid = (allc[:148] >> 48).astype(int)
allc[:148] &= (1 << 48) - 1
order = np.argsort(-allc[:148])
print("slowest CTAs (cycles/1000, blockIdx, smid):", [(int(allc[i] // 1000), int(i), int(smid[i])) for i in order[:24]])
print("fastest CTAs (cycles/1000, blockIdx, smid):", [(int(allc[i] // 1000), int(i), int(smid[i])) for i in order[-8:]])
print("distinct SMs used:", len(set(smid.tolist())))
c = allc[:148].reshape(37, 4)     # [tile_y][tile_x]
for name, off in (("y-warp wait", 1024), ("patch-warp TMA wait", 2048), ("patch work", 3072)):
    w = allc[off:off + 148].reshape(37, 4)
    print(name, "cycles/1000 rows 0,1,17,36:", (w / 1000).astype(int)[[0, 1, 17, 36]].tolist())
np.set_printoptions(linewidth=200)
print(modes, "cycles/1000 by tile_y (rows) x tile_x (cols)")
print((c / 1000).astype(int)[[0, 1, 2, 17, 34, 35, 36]])
print("min %d max %d mean %d" % (c.min() / 1000, c.max() / 1000, c.mean() / 1000))
