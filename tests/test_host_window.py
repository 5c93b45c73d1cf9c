"""Host-volume pipeline with output windows (cupyimg_b200/host.py): the z-slab bookkeeping on the CPU, and on a
GPU the windowed pipeline against the filter of the whole volume (what bench.py's multi-GPU e2e leg runs: every
rank streams its own slab plus the overlap planes, no GPU-to-GPU traffic)."""
import numpy as np
import pytest


def test_slab_window_covers_the_volume_once():
    from cupyimg_b200.host import slab_window
    for nz, world, r in [(512, 1, 8), (4096, 8, 8), (2048, 8, 16), (100, 3, 7), (64, 4, 16)]:
        covered = []
        for rank in range(world):
            a, b, (wb, we) = slab_window(nz, world, rank, r)
            assert 0 <= a <= a + wb < a + we <= b <= nz
            assert wb == min(r, a + wb) and (b - a - we) == min(r, nz - (a + we))      # full halo except at the ends
            covered.extend(range(a + wb, a + we))
        assert covered == list(range(nz))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["reflect", "constant", "nearest", "mirror"])
def test_windowed_pipeline_equals_whole_volume(mode):
    torch = pytest.importorskip("torch")
    from cupyimg_b200 import host
    from cupyimg_b200.scipy import ndimage as ndi
    nz, ny, nx = 200, 40, 64
    g = torch.Generator().manual_seed(3)
    vol = torch.rand((nz, ny, nx), generator=g)
    want = ndi.gaussian_filter(vol.cuda(), 2.0, mode=mode).cpu()
    for world in (1, 2, 3, 5):
        for rank in range(world):
            a, b, win = host.slab_window(nz, world, rank, 8)
            hx = vol[a:b].contiguous().pin_memory()
            for chunk in (16, 32, 1000):                 # pipelined with several chunks / one shot
                got = host.gaussian_filter_host(hx, 2.0, mode=mode, chunk_planes=chunk,
                                                out_window=None if world == 1 else win)
                torch.cuda.synchronize()
                assert torch.equal(got, want[a + win[0]:a + win[1]]), (world, rank, chunk)


@pytest.mark.gpu
def test_window_argument_errors():
    torch = pytest.importorskip("torch")
    from cupyimg_b200 import host
    x = torch.rand((40, 16, 16)).pin_memory()
    with pytest.raises(ValueError):
        host.gaussian_filter_host(x, 1.0, out_window=(10, 50))
    with pytest.raises(Exception):
        host.gaussian_filter_host(x, 1.0, output=torch.empty((40, 16, 16)).pin_memory(), out_window=(10, 20))
