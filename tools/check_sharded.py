"""Sharded == single-GPU (bit for bit) for both halo backends, plus step timings.
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/check_sharded.py [quick]"""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from cupyimg_b200 import sharded, _ffi
from cupyimg_b200.scipy import ndimage as ndi

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
bad = 0
nz, ny, nx = 40, 48, 64
g = torch.Generator(device="cpu").manual_seed(7)
vol = torch.rand((nz * world, ny, nx), generator=g).to(dev)
x = vol[rank * nz:(rank + 1) * nz].contiguous()
for backend in ["p2p", "nccl"]:
    for mode in ["reflect", "wrap", "constant", "mirror", "nearest"]:
        for name, radius, fs, fp in [
                ("gauss2", 8, lambda: ndi.gaussian_filter(vol, 2.0, mode=mode), lambda p: p.gaussian_filter(x, 2.0)),
                ("gauss1", 4, lambda: ndi.gaussian_filter(vol, 1.0, mode=mode), lambda p: p.gaussian_filter(x, 1.0)),
                ("gauss2.5", 10, lambda: ndi.gaussian_filter(vol, 2.5, mode=mode), lambda p: p.gaussian_filter(x, 2.5)),
                ("gauss4", 16, lambda: ndi.gaussian_filter(vol, 4.0, mode=mode), lambda p: p.gaussian_filter(x, 4.0)),
                ("uniform5", 4, lambda: ndi.uniform_filter(vol, 5, mode=mode), lambda p: p.uniform_filter(x, 5)),
                ("gradmag1.5", 6, lambda: ndi.gaussian_gradient_magnitude(vol, 1.5, mode=mode), lambda p: p.gaussian_gradient_magnitude(x, 1.5)),
                ("sobel0", 1, lambda: ndi.sobel(vol, 0, mode=mode), lambda p: p.sobel(x, 0))]:
            want = fs()[rank * nz:(rank + 1) * nz]
            plan = sharded.ZSlabFilter(x.shape, radius=radius, mode=mode, device=dev, backend=backend)
            _ffi.LAUNCHES = 0
            got = fp(plan)
            nl = _ffi.LAUNCHES
            got_b = fp(plan)                      # a second step through the same plan (epoch 2)
            torch.cuda.synchronize()
            ok = torch.equal(got, want) and torch.equal(got_b, want)
            t = torch.tensor([0 if ok else 1], device=dev); dist.all_reduce(t)
            bad += int(t.item() > 0)
            if rank == 0:
                print("%-5s %-9s %-11s launches/step %d  %s" % (backend, mode, name, nl, "ok" if t.item() == 0 else "MISMATCH"), flush=True)
            del plan
if rank == 0:
    print("mismatching cases:", bad, flush=True)
if not quick and bad == 0:
    n = 512
    xs = torch.rand((n, n, n), device=dev); out = torch.empty_like(xs)

    def timeit(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); b.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t1 = timeit(lambda: ndi.gaussian_filter(xs, 2.0, output=out))
    for backend in ["p2p", "nccl"]:
        plan = sharded.ZSlabFilter(xs.shape, radius=8, mode="reflect", device=dev, backend=backend)
        src = plan.slab if plan.p2p else xs
        if plan.p2p:
            src.copy_(xs)
        ts = timeit(lambda: plan.gaussian_filter(src, 2.0, output=out))
        if rank == 0:
            print("512^3 slab per rank, sigma 2, %d ranks: single-GPU call %.4f ms, sharded step (%s) %.4f ms -> weak efficiency %.3f" % (
                world, t1, backend, ts, t1 / ts), flush=True)
        del plan
    t1 = timeit(lambda: ndi.gaussian_gradient_magnitude(xs, 1.5, output=out))
    for backend in ["p2p", "nccl"]:
        plan = sharded.ZSlabFilter(xs.shape, radius=6, mode="reflect", device=dev, backend=backend)
        src = plan.slab if plan.p2p else xs
        if plan.p2p:
            src.copy_(xs)
        ts = timeit(lambda: plan.gaussian_gradient_magnitude(src, 1.5, output=out))
        if rank == 0:
            print("512^3 slab per rank, gradmag 1.5, %d ranks: single-GPU call %.4f ms, sharded step (%s) %.4f ms -> weak efficiency %.3f" % (
                world, t1, backend, ts, t1 / ts), flush=True)
        del plan
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
