"""z-slab sharding on real GPUs: sharded result == single-GPU result, bit for bit, for every boundary mode and
for both halo backends (peer memory: neighbour planes pulled by copy engines / read in place by TMA from
symmetric memory; NCCL send/recv).  Needs >= 2 GPUs (skipped otherwise); on one GPU the world-size-1 plan is
still exercised and tests/test_halo_gpu.py covers the halo entry point with local stand-ins."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_dir, backend):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from cupyimg_b200 import sharded
        from cupyimg_b200.scipy import ndimage as ndi
        nz, ny, nx = 40, 48, 64
        g = torch.Generator(device="cpu").manual_seed(7)
        vol = torch.rand((nz * world, ny, nx), generator=g)
        ok = True
        for mode in ["reflect", "wrap", "constant", "mirror", "nearest"]:
            for sigma, radius in [(2.0, 8), (1.0, 4), (2.5, 10), (4.0, 16)]:
                want = ndi.gaussian_filter(vol.to(dev), sigma, mode=mode)
                x = vol[rank * nz:(rank + 1) * nz].to(dev)
                plan = sharded.ZSlabFilter(x.shape, radius=radius, mode=mode, device=dev, backend=backend)
                got = plan.gaussian_filter(x, sigma)
                torch.cuda.synchronize()
                ok = ok and torch.equal(got, want[rank * nz:(rank + 1) * nz])
                if backend == "p2p" and mode != "wrap":
                    ok = ok and plan.last_backend.startswith("peer memory")
                    # a second and third step through the same plan, the slab refilled in place in between
                    plan.begin_fill()
                    plan.slab.copy_(x * 0.5)
                    got_b = plan.gaussian_filter(plan.slab, sigma)
                    torch.cuda.synchronize()
                    ok = ok and torch.equal(got_b, ndi.gaussian_filter(vol.to(dev) * 0.5, sigma, mode=mode)[rank * nz:(rank + 1) * nz])
                got2 = plan.uniform_filter(x, 5)
                want2 = ndi.uniform_filter(vol.to(dev), 5, mode=mode)
                ok = ok and torch.equal(got2, want2[rank * nz:(rank + 1) * nz])
            # C4's filter: three derivative filters behind one halo exchange
            plan = sharded.ZSlabFilter(x.shape, radius=6, mode=mode, device=dev, backend=backend)
            got3 = plan.gaussian_gradient_magnitude(x, 1.5)
            want3 = ndi.gaussian_gradient_magnitude(vol.to(dev), 1.5, mode=mode)
            ok = ok and torch.equal(got3, want3[rank * nz:(rank + 1) * nz])
            got4 = plan.sobel(x, 0)
            want4 = ndi.sobel(vol.to(dev), 0, mode=mode)
            ok = ok and torch.equal(got4, want4[rank * nz:(rank + 1) * nz])
            # exact (float64-accumulate) staging of the same filter on an integer volume
            vi = (vol * 1000).to(torch.int32)
            plan_i = sharded.ZSlabFilter(x.shape, radius=6, mode=mode, device=dev, dtype=torch.int32, backend="auto")
            got5 = plan_i.gaussian_gradient_magnitude(vi[rank * nz:(rank + 1) * nz].to(dev), 1.5)
            want5 = ndi.gaussian_gradient_magnitude(vi.to(dev), 1.5, mode=mode)
            ok = ok and torch.equal(got5, want5[rank * nz:(rank + 1) * nz])
        with open(os.path.join(result_dir, "rank%d" % rank), "w") as f:
            f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("backend", ["p2p", "nccl"])
def test_zslab_matches_single_gpu(tmp_path, backend):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), backend), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("rank%d" % r)).read_text() == "ok"


def test_world_size_one_plan():
    import torch
    from cupyimg_b200 import sharded
    from cupyimg_b200.scipy import ndimage as ndi
    x = torch.rand((24, 32, 64), device="cuda")
    plan = sharded.ZSlabFilter(x.shape, radius=8, device="cuda")
    assert torch.equal(plan.gaussian_filter(x, 2.0), ndi.gaussian_filter(x, 2.0))
    assert torch.equal(plan.gaussian_gradient_magnitude(x, 1.5), ndi.gaussian_gradient_magnitude(x, 1.5))
    assert torch.equal(plan.sobel(x, 1), ndi.sobel(x, 1))
