"""sepfilt_separable_f32_halo on ONE GPU: local arrays stand in for the neighbours' slabs, so the entry point the
multi-GPU path uses (cupyimg_b200/sharded.py, peer-memory backend) is covered by the single-GPU test tier too:
the slab filtered with halos == the same planes of the filtered concatenated volume, bit for bit; ready flags
(already set) are honoured; the kernel's last CTA writes the done flags and re-arms its counter."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _halo_call(x, lo, hi, out, specs, dspecs=None, flags=None, epoch=0, done=None):
    from cupyimg_b200 import _array, _ffi
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
    h.epoch = epoch
    if flags is not None:
        h.ready_lo, h.ready_hi = flags.data_ptr(), flags.data_ptr() + 4
    if done is not None:
        h.done_lo, h.done_hi, h.cta_counter = done.data_ptr(), done.data_ptr() + 4, done.data_ptr() + 8
    rc = _ffi.lib().sepfilt_separable_f32_halo(inp.tensor(), o.tensor(), arr, len(structs), darr,
                                               1 if dspecs is not None else 0, ctypes.byref(h), 0, 0.0,
                                               _array.current_stream(x.device))
    return rc


@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror"])
@pytest.mark.parametrize("sigma,grad", [(2.0, False), (1.0, False), (1.25, False), (2.5, False), (4.0, False), (1.5, True), (0.75, True)])
def test_halo_call_equals_concatenated_volume(mode, sigma, grad):
    from cupyimg_b200 import _array, _ffi
    from cupyimg_b200.scipy import ndimage as ndi
    from cupyimg_b200.scipy.ndimage import filters as F
    r = int(4 * sigma + 0.5)
    nz, ny, nx = 40, 72, 136
    g = torch.Generator(device="cuda").manual_seed(int(sigma * 100) + grad)
    big = torch.rand((nz + 2 * r, ny, nx), device="cuda", generator=g)
    x, lo, hi = big[r:r + nz], big[:r].clone(), big[r + nz:].clone()
    probe = _array.ingest(x)
    smooth = F._gaussian_specs(probe, sigma, 0, mode, 4.0)
    deriv = F._gaussian_specs(probe, sigma, 1, mode, 4.0) if grad else None
    fn = (lambda a: ndi.gaussian_gradient_magnitude(a, sigma, mode=mode)) if grad else (lambda a: ndi.gaussian_filter(a, sigma, mode=mode))
    # both neighbours / lower only / upper only (the ends of the global volume use the boundary mode)
    for use_lo, use_hi in [(True, True), (True, False), (False, True)]:
        vol = torch.cat(([lo] if use_lo else []) + [x] + ([hi] if use_hi else []))
        off = r if use_lo else 0
        want = fn(vol)[off:off + nz]
        out = torch.empty((nz, ny, nx), device="cuda")
        flags = torch.tensor([5, 5, 0, 0], dtype=torch.int32, device="cuda")
        done = torch.zeros(4, dtype=torch.int32, device="cuda")
        rc = _halo_call(x, lo if use_lo else None, hi if use_hi else None, out, smooth, deriv, flags, 5, done)
        assert rc == _ffi.OK, _ffi.last_error()
        torch.cuda.synchronize()
        assert torch.equal(out, want)
        assert done.tolist() == [5, 5, 0, 0]            # both done flags carry the epoch, the CTA counter is re-armed
        # a second launch through the same counter
        assert _halo_call(x, lo if use_lo else None, hi if use_hi else None, out, smooth, deriv, flags, 5, done) == _ffi.OK
        torch.cuda.synchronize()
        assert torch.equal(out, want) and done.tolist() == [5, 5, 0, 0]


def test_halo_call_declines_what_it_cannot_fuse():
    from cupyimg_b200 import _array, _ffi
    from cupyimg_b200.scipy.ndimage import filters as F
    x = torch.rand((40, 64, 64), device="cuda")
    out = torch.empty_like(x)
    lo = torch.rand((4, 64, 64), device="cuda")
    probe = _array.ingest(x)
    # halo thinner than the z radius
    specs = F._gaussian_specs(probe, 2.0, 0, "reflect", 4.0)
    assert _halo_call(x, lo, None, out, specs) == _ffi.ERR_UNSUPPORTED
    # radius 20 (sigma 5) has no fused kernel; `wrap` along y / x is declined at every radius
    specs = F._gaussian_specs(probe, 5.0, 0, "reflect", 4.0)
    assert _halo_call(x, torch.rand((20, 64, 64), device="cuda"), None, out, specs) == _ffi.ERR_UNSUPPORTED
    specs = F._gaussian_specs(probe, 2.0, 0, "wrap", 4.0)
    assert _halo_call(x, torch.rand((8, 64, 64), device="cuda"), None, out, specs) == _ffi.ERR_UNSUPPORTED


def test_stream_flag_operations():
    from cupyimg_b200 import _array, _ffi
    L = _ffi.lib()
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    s = _array.current_stream(flags.device)
    assert L.sepfilt_stream_write32(s, flags.data_ptr(), 7) == _ffi.OK
    assert L.sepfilt_stream_write32x2(s, flags.data_ptr() + 4, flags.data_ptr() + 8, 9) == _ffi.OK
    assert L.sepfilt_stream_write32x2(s, None, flags.data_ptr() + 12, 11) == _ffi.OK
    assert L.sepfilt_stream_wait32_geq(s, flags.data_ptr(), 7) == _ffi.OK      # already satisfied: the stream moves on
    torch.cuda.synchronize()
    assert flags.tolist() == [7, 9, 9, 11]
    assert L.sepfilt_stream_write32(s, flags.data_ptr() + 1, 1) == _ffi.ERR_INVALID
