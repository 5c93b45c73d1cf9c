"""One launch each of the streaming single-axis kernels on 512^3 f32 for ncu."""
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
x = torch.rand((512, 512, 512), device="cuda"); o = torch.empty_like(x)
for _ in range(2):
    ndi.gaussian_filter1d(x, 2.0, axis=2, output=o)      # f32_stream_row_kernel<8>
    ndi.gaussian_filter1d(x, 1.0, axis=0, output=o)      # f32_stream_col_kernel<4>
    ndi.gaussian_filter1d(x, 2.0, axis=0, output=o)      # f32_stream_col_kernel<8>
    ndi.maximum_filter1d(x, 5, axis=1, output=o)         # minmax_stream_col_kernel<float, 5, max>
    ndi.maximum_filter1d(x, 5, axis=2, output=o)         # minmax_stream_row_kernel<float, 5, max>
torch.cuda.synchronize()
