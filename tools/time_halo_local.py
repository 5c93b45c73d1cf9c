"""The halo entry point with LOCAL arrays standing in for the neighbours' slabs (one GPU, no flags): the pure
kernel cost of the HALO instantiation next to the plain call, and a correctness check against the filter
of the concatenated volume."""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import _array, _ffi
from cupyimg_b200.scipy import ndimage as ndi
from cupyimg_b200.scipy.ndimage import filters as F


def halo_call(x, lo, hi, out, specs, dspecs=None, flags=None, epoch=0):
    inp, o = _array.ingest(x), _array.ingest(out)
    structs = [s.struct() for s in specs]
    arr = (_ffi.Pass * len(structs))(*[s[0] for s in structs])
    darr = None
    if dspecs is not None:
        ds = [s.struct() for s in dspecs]
        darr = (_ffi.Pass * len(ds))(*[s[0] for s in ds])
    h = _ffi.Halo()
    if lo is not None:
        h.lo, h.planes_lo = lo.data_ptr(), lo.shape[0]
    if hi is not None:
        h.hi, h.planes_hi = hi.data_ptr(), hi.shape[0]
    if flags is not None:
        h.ready_lo, h.ready_hi, h.epoch = flags.data_ptr(), flags.data_ptr() + 4, epoch
    rc = _ffi.lib().sepfilt_separable_f32_halo(inp.tensor(), o.tensor(), arr, len(structs), darr, 1 if dspecs is not None else 0,
                                               ctypes.byref(h), 0, 0.0, _array.current_stream(x.device))
    _ffi.check(rc)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps


n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
for sigma, grad in [(2.0, False), (1.0, False), (1.5, True)]:
    r = int(4 * sigma + 0.5)
    big = torch.rand((n + 2 * r, n, n), device="cuda")
    x, lo, hi = big[r:r + n], big[:r].clone(), big[r + n:].clone()
    out = torch.empty((n, n, n), device="cuda")
    probe = _array.ingest(x)
    smooth = F._gaussian_specs(probe, sigma, 0, "reflect", 4.0)
    deriv = F._gaussian_specs(probe, sigma, 1, "reflect", 4.0) if grad else None
    want = (ndi.gaussian_gradient_magnitude(big, sigma) if grad else ndi.gaussian_filter(big, sigma))[r:r + n]
    halo_call(x, lo, hi, out, smooth, deriv)
    torch.cuda.synchronize()
    print("sigma %.1f grad %d: halo call == filter of the concatenated volume: %s" % (sigma, grad, torch.equal(out, want)))
    flags = torch.full((4,), 7, dtype=torch.int32, device="cuda")
    t_h = timeit(lambda: halo_call(x, lo, hi, out, smooth, deriv))
    t_f = timeit(lambda: halo_call(x, lo, hi, out, smooth, deriv, flags, 7))
    xs = x.contiguous()
    t_p = timeit(lambda: (ndi.gaussian_gradient_magnitude(xs, sigma, output=out) if grad else ndi.gaussian_filter(xs, sigma, output=out)))
    print("   plain call %.4f ms   halo call %.4f ms   halo call with (already set) flags %.4f ms" % (t_p, t_h, t_f), flush=True)
