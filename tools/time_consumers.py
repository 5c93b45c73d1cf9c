"""Timings of the skimage-level consumers (CUDA events, mean of 10)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _ffi
from cupyimg_b200.skimage import feature as skfe, filters as skf, metrics as skm


def timeit(name, fn, nvox, reps=10):
    for _ in range(3):
        fn()
    _ffi.LAUNCHES = 0
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%-72s %8.3f ms %8.1f Gvoxel/s  launches/call %d" % (name, ms, nvox / ms / 1e6, _ffi.LAUNCHES // reps), flush=True)


for shape in [(4096, 4096), (256, 256, 256)]:
    n = int(np.prod(shape))
    x = torch.rand(shape, device="cuda"); y = (x + 0.05 * torch.randn(shape, device="cuda")).clamp(0, 1)
    tag = "x".join(map(str, shape)) + " f32"
    timeit("skimage.filters.gaussian sigma=2 %s" % tag, lambda: skf.gaussian(x, 2.0), n)
    timeit("structural_similarity (uniform 7) data_dtype=f32 %s" % tag,
           lambda: skm.structural_similarity(x, y, data_range=1.0, data_dtype=np.float32), n)
    timeit("structural_similarity (gaussian weights) data_dtype=f32 %s" % tag,
           lambda: skm.structural_similarity(x, y, data_range=1.0, data_dtype=np.float32, gaussian_weights=True), n)
    timeit("structure_tensor sigma=1.5 order=rc %s" % tag, lambda: skfe.structure_tensor(x, sigma=1.5, order="rc"), n)
