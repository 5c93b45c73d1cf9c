#!/usr/bin/env python
"""bench.py — headline benchmark: gaussian_filter(sigma=2, truncate=4; 17 taps/axis) on a
512^3 float32 volume per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cpu] [--no-legs]

A "step" is one gaussian_filter call over one resident synthetic volume.  N > 1 (launched by
torchrun, one rank per GPU) is weak scaling: every rank owns one 512^3 z-slab of a (512*N) x 512 x 512
volume; the fused kernel reads the 8 halo planes per side straight from the neighbours' slabs over
NVLink (cupyimg_b200.sharded, peer-memory backend; NCCL send/recv where that does not apply).

The same JSON line carries, under "legs", the STRONG-scaling configurations BASELINE.json names, each on
the same global data for every N (synthetic planes are seeded by their global index):
  C5  gaussian_filter sigma=4 on 2048^3 f32, z-sharded          (32 GiB in + 32 GiB out in total)
  C4  gaussian_gradient_magnitude sigma=1.5 on 1024^3 f32, z-sharded
  C3  convolve1d (9 taps) along each axis of 64 x 2048 x 2048 uint16, mode mirror, batch-sharded, bit-exact
and "parity": the timed output checked against the CPU oracle on a sub-brick and, for N > 1, every rank's
slab against the single-GPU kernel run on the same planes, bit for bit.
Prints ONE JSON line (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NZ = NY = NX = 512
SIGMA, TRUNCATE, MODE = 2.0, 4.0, "reflect"
RADIUS = int(TRUNCATE * SIGMA + 0.5)
BYTES_PER_VOXEL = 8            # read f32 once + write f32 once per API call
METRIC = "gaussian_filter 512^3 f32 throughput"
UNIT = "Gvoxel/s"
SEED = 1234
# FP32 work of the fused kernel per output voxel on the 512^3 plan: 17 taps x (y pass incl. the
# x halo columns 144/128 + x pass + z pass), y and x passes also run on the 2 x 8 z-halo planes (528/512)
FMA_PER_VOXEL = 17 * ((144.0 / 128.0 + 1.0) * 528.0 / 512.0 + 1.0)
FMA_PER_CLK_SM = 128


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused kernel on this exact
    workload, from the committed `ncu --set full` capture (profiles/fused_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "fused_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and clock-event reasons read through NVML inside this process every 2 ms; only the
    samples taken while a timed region is open (`with sampler.region():`) are reported, so a 7 ms region
    still yields several samples under load."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.open, self.stop_flag, self.thread, self.h = [], False, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            if self.open:
                try:
                    self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                         nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0,
                                         nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
                except Exception:
                    pass
                time.sleep(0.002)
            else:
                time.sleep(0.0005)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    class _Region:
        def __init__(self, s):
            self.s = s

        def __enter__(self):
            self.s.open = True

        def __exit__(self, *a):
            self.s.open = False

    def region(self):
        return ClockSampler._Region(self)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.h is None:
            return out
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=2)
        if self.samples:
            bits = 0
            for s in self.samples:
                bits |= s[2]
            out = {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.max_mhz,
                   "sm_min_mhz": min(s[0] for s in self.samples),
                   "power_w_max": max(s[1] for s in self.samples),
                   "reasons": sorted(n for b, n in self.REASONS.items() if bits & b),
                   "samples": len(self.samples), "source": "NVML in-process, 2 ms period, timed regions only"}
        return out


# ----------------------------------------------------------------------------------------------------
# synthetic data: every plane (or image) is seeded by its GLOBAL index, so the global volume is the
# same for every N and any rank can regenerate a neighbour's planes for the parity checks
# ----------------------------------------------------------------------------------------------------
def synth_planes(torch, dev, tag, z0, z1, ny, nx, out=None, dtype=None):
    dtype = dtype or torch.float32
    if out is None:
        out = torch.empty((z1 - z0, ny, nx), dtype=dtype, device=dev)
    g = torch.Generator(device=dev)
    for z in range(z0, z1):
        g.manual_seed(SEED * 1000003 + tag * 100003 + z)
        if dtype == torch.float32:
            out[z - z0] = torch.rand((ny, nx), device=dev, generator=g)
        else:
            out[z - z0] = torch.randint(0, 4096, (ny, nx), device=dev, generator=g, dtype=torch.int32).to(dtype)
    return out


def _oracle_brick_check(np, oracle_fn, x_dev, out_dev, r, brick, edges):
    """Compare out[brick] with the oracle run on the input around the brick.  brick = (z0, z1, y0, y1, x0, x1) in the
    coordinates of x_dev; edges[a] = (lo_is_array_end, hi_is_array_end) tells where the boundary rule really
    applies — elsewhere r planes / rows / columns of margin are cut from the oracle's answer."""
    sl_in, sl_valid = [], []
    for a in range(3):
        b0, b1 = brick[2 * a], brick[2 * a + 1]
        n = x_dev.shape[a]
        lo = 0 if (edges[a][0] and b0 - r <= 0) else b0 - r
        hi = n if (edges[a][1] and b1 + r >= n) else b1 + r
        if lo < 0 or hi > n:
            raise ValueError("brick margin leaves the array")
        sl_in.append(slice(lo, hi))
        sl_valid.append(slice(b0 - lo, b1 - lo))
    sub = x_dev[tuple(sl_in)].cpu().numpy()
    want = oracle_fn(sub)[tuple(sl_valid)]
    got = out_dev[brick[0]:brick[1], brick[2]:brick[3], brick[4]:brick[5]].cpu().numpy()
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    tol = 1e-6 * float(np.abs(want).max()) + 1e-5 * np.abs(want)
    return float(diff.max()), float((diff / np.maximum(np.abs(want), 1e-30)).max()), bool((diff <= tol).all())


def _oracle_gaussian_mt(x, threads, sigma=SIGMA):
    """The CPU port (oracle/) with every 1-D pass split over `threads` host threads (disjoint
    line ranges; the ctypes call releases the GIL).  Bit-identical to the serial oracle."""
    from oracle import oracle
    oracle.THREADS = threads
    try:
        return oracle.gaussian_filter(x, sigma, truncate=TRUNCATE, mode=MODE)
    finally:
        oracle.THREADS = 1


def cpu_baseline(sample_planes, threads, repeats=1):
    """Time the CPU port on a bounded sample (sample_planes x 512 x 512) of the workload."""
    import numpy as np
    from oracle import oracle
    oracle.build()
    rng = np.random.default_rng(1234)
    x = rng.random((sample_planes, NY, NX), dtype=np.float32)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        _oracle_gaussian_mt(x, threads)
        best = min(best, time.perf_counter() - t0)
    return x.size / best / 1e9, best


def scipy_baseline(sample_planes):
    """scipy.ndimage.gaussian_filter itself (one thread: scipy.ndimage is single-threaded) on a bounded sample."""
    try:
        import numpy as np
        from scipy import ndimage as sndi
    except Exception as e:                                    # scipy missing on the box: say so
        return None, "scipy unavailable: %r" % (e,)
    rng = np.random.default_rng(1234)
    x = rng.random((sample_planes, NY, NX), dtype=np.float32)
    t0 = time.perf_counter()
    sndi.gaussian_filter(x, SIGMA, truncate=TRUNCATE, mode=MODE)
    t = time.perf_counter() - t0
    return x.size / t / 1e9, "%d x 512 x 512 z-slab, scipy.ndimage.gaussian_filter float32, 1 thread, %.1f s wall" % (sample_planes, t)


def run_reference(args):
    """--impl reference: the reference has no CPU implementation of its own and cannot run
    without CuPy, so this arm times the CPU port of the path (oracle/, scipy.ndimage's
    arithmetic) on all host cores, on a bounded z-slab sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_planes = 256 if cores >= 8 else 64
    cpu_baseline(16, cores)                       # page in, build
    for _ in range(max(args.warmup - 1, 0)):
        cpu_baseline(sample_planes, cores)
    t_all = 0.0
    vox = 0
    for _ in range(args.steps):
        v, t = cpu_baseline(sample_planes, cores)
        t_all += t
        vox += sample_planes * NY * NX
    value = vox / t_all / 1e9
    sample = "%d x %d x %d z-slab of the 512^3 volume per step (1/%d of the workload), oracle port on %d threads" % (
        sample_planes, NY, NX, NZ // sample_planes, cores)
    sv, ss = scipy_baseline(64)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "gaussian_filter sigma=2 truncate=4 (17 taps/axis) mode=reflect, 512^3 float32",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "scipy_ndimage_1thread": {"value": sv, "unit": UNIT, "cores": 1, "sample": ss}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# strong-scaling legs
# ----------------------------------------------------------------------------------------------------
def _time_steps(torch, dist, world, dev, fn, steps, warmup, sampler):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler.region():
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        b.synchronize()
    if world > 1:
        dist.barrier()
    ms = a.elapsed_time(b) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def _checksum(torch, dist, world, dev, out):
    """float64 checksum of the global result: per-plane sums (one float64 reduction per plane), gathered from all
    ranks and added on the host in global plane order — the same additions for every N, so the value is
    identical across rank counts exactly when the sharded outputs are bit-identical to the single-GPU ones."""
    per_plane = out.reshape(out.shape[0], -1).to(torch.float64).sum(dim=1)
    if world > 1:
        parts = [torch.zeros_like(per_plane) for _ in range(world)]
        dist.all_gather(parts, per_plane)
        per_plane = torch.cat(parts)
    tot = 0.0
    for v in per_plane.cpu().tolist():
        tot += v
    return tot


def leg_zsharded(torch, dist, np, rank, world, dev, sampler, name, n, kind, sigma, steps, warmup, tag):
    """Strong scaling of one z-sharded float32 volume n^3: gaussian_filter (kind 'gauss') or
    gaussian_gradient_magnitude (kind 'gradmag')."""
    from cupyimg_b200 import sharded, _ffi
    from cupyimg_b200.scipy import ndimage as ndi
    from cupyimg_b200.scipy.ndimage import filters as F
    from cupyimg_b200 import _array
    from oracle import oracle
    if n % world:
        return {"skipped": "n not divisible by the rank count"}
    nz = n // world
    r = int(TRUNCATE * sigma + 0.5)
    if world > 1 and nz < 2 * r:
        return {"skipped": "slab thinner than two halos"}
    z0 = rank * nz
    plan = sharded.ZSlabFilter((nz, n, n), radius=r, mode=MODE, device=dev) if world > 1 else None
    x = plan.slab if (plan is not None and plan.p2p) else torch.empty((nz, n, n), dtype=torch.float32, device=dev)
    synth_planes(torch, dev, tag, z0, z0 + nz, n, n, out=x)
    out = torch.empty((nz, n, n), dtype=torch.float32, device=dev)
    if kind == "gauss":
        single = lambda src, dst: ndi.gaussian_filter(src, sigma, output=dst, mode=MODE, truncate=TRUNCATE)
        step = (lambda: plan.gaussian_filter(x, sigma, truncate=TRUNCATE, output=out)) if world > 1 else (lambda: single(x, out))
        ofn = lambda a: oracle.gaussian_filter(a, sigma, truncate=TRUNCATE, mode=MODE)
    else:
        single = lambda src, dst: ndi.gaussian_gradient_magnitude(src, sigma, output=dst, mode=MODE, truncate=TRUNCATE)
        step = (lambda: plan.gaussian_gradient_magnitude(x, sigma, truncate=TRUNCATE, output=out)) if world > 1 else (lambda: single(x, out))
        ofn = lambda a: oracle.gaussian_gradient_magnitude(a, sigma, truncate=TRUNCATE, mode=MODE)
    _ffi.LAUNCHES = 0
    step()
    launches = _ffi.LAUNCHES
    ms = _time_steps(torch, dist, world, dev, step, steps, warmup, sampler)
    torch.cuda.synchronize()
    res = {"workload": name, "global_volume": [n, n, n], "slab_per_gpu": [nz, n, n], "ms_per_step": ms,
           "value": n ** 3 / ms / 1e6, "unit": UNIT, "launches_per_step_per_gpu": launches, "steps": steps,
           "scaling": "strong", "halo_backend": (plan.last_backend if plan is not None else "none"),
           "algorithmic_GBps_per_gpu": nz * n * n * 8 / ms / 1e6}
    # ---- parity: one brick per rank against the oracle (a corner brick on the first rank) ----
    oracle.THREADS = max(1, min(16, os.cpu_count() or 1))
    try:
        bz = min(24, nz - r) if world > 1 else 24
        if rank == 0:
            brick = (0, bz, 0, 32, n - 48, n)                 # z-low, y-low and x-high ends of the volume
        else:
            zb = r + (nz - 2 * r - bz) // 2 if nz - 2 * r >= bz else r
            brick = (zb, min(zb + bz, nz - r), n // 2 - 16, n // 2 + 16, 0, 48)
        edges = [(rank == 0, rank == world - 1), (True, True), (True, True)]
        mabs, mrel, ok = _oracle_brick_check(np, ofn, x, out, r, brick, edges)
    finally:
        oracle.THREADS = 1
    par = {"brick": list(brick), "max_abs": mabs, "max_rel": mrel, "within_tol": ok,
           "tol": "|a-b| <= 1e-6 max|b| + 1e-5 |b| (float32, north_star rtol 1e-5)"}
    # ---- sharded == single GPU, bit for bit: the single-GPU kernel on [halo | slab | halo] ----
    if world > 1:
        lo = r if rank > 0 else 0
        hi = r if rank < world - 1 else 0
        free, _ = torch.cuda.mem_get_info(dev)
        need = (nz + lo + hi + nz) * n * n * 4 * 1.6
        if need < free:
            ext = torch.empty((lo + nz + hi, n, n), dtype=torch.float32, device=dev)
            if lo:
                synth_planes(torch, dev, tag, z0 - lo, z0, n, n, out=ext[:lo])
            ext[lo:lo + nz].copy_(x)
            if hi:
                synth_planes(torch, dev, tag, z0 + nz, z0 + nz + hi, n, n, out=ext[lo + nz:])
            ref = torch.empty_like(out)
            inp = _array.ingest(ext)
            if kind == "gauss":
                specs = F._gaussian_specs(inp, sigma, 0, MODE, TRUNCATE)
                F._run_passes_window(inp, _array.ingest(ref), specs, 0.0, None, lo)
            else:
                smooth = F._gaussian_specs(inp, sigma, 0, MODE, TRUNCATE)
                deriv = F._gaussian_specs(inp, sigma, 1, MODE, TRUNCATE)
                F._gradient_magnitude_window(inp, _array.ingest(ref), smooth, deriv, 0.0, None, lo)
            torch.cuda.synchronize()
            eq = torch.tensor([1 if torch.equal(ref, out) else 0], device=dev)
            dist.all_reduce(eq, op=dist.ReduceOp.MIN)
            par["sharded_equals_single"] = bool(eq.item())
            del ext, ref
        else:
            par["sharded_equals_single"] = None
        flags = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        par["within_tol_all_ranks"] = bool(flags.item())
    par["checksum_f64"] = _checksum(torch, dist, world, dev, out)
    res["parity"] = par
    del x, out, plan
    torch.cuda.empty_cache()
    return res


def leg_c3(torch, dist, np, rank, world, dev, sampler, steps, warmup):
    """C3: convolve1d (9 taps) along each axis of a 64 x 2048 x 2048 uint16 stack, mode mirror; images are
    sharded over the ranks with no communication; bit-exact against the oracle on sample images."""
    from cupyimg_b200 import sharded, _ffi
    from cupyimg_b200.scipy import ndimage as ndi
    from oracle import oracle
    B, H, W = 64, 2048, 2048
    b0, b1 = sharded.batch_range(B, world, rank)
    nb = b1 - b0
    w = np.array([1, 4, 9, 15, 18, 15, 9, 4, 1], dtype=np.float64) / 76.0
    x = synth_planes(torch, dev, 3, b0, b1, H, W, dtype=torch.uint16)
    out = torch.empty_like(x)
    res = {"workload": "convolve1d 9 taps, 64 x 2048 x 2048 uint16, mode mirror, batch-sharded", "images_per_gpu": nb,
           "scaling": "strong", "unit": "Gpixel/s", "steps": steps}
    ok = True
    for axis in (1, 2):
        fn = lambda: ndi.convolve1d(x, w, axis=axis, output=out, mode="mirror")
        ms = _time_steps(torch, dist, world, dev, fn, steps, warmup, sampler)
        res["axis%d" % axis] = {"ms_per_step": ms, "value": B * H * W / ms / 1e6,
                                "algorithmic_GBps_per_gpu": nb * H * W * 4 / ms / 1e6}
        torch.cuda.synchronize()
        # bit-exact on the first and last image of this rank (full images: every edge and corner)
        for i in sorted(set([0, nb - 1])):
            want = oracle.convolve1d(x[i].cpu().numpy(), w, axis=axis - 1, mode="mirror")
            ok = ok and bool((out[i].cpu().numpy() == want).all())
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # axis 0 runs across images: only meaningful inside a rank's sub-stack, reported for completeness at N = 1
    res["parity"] = {"bit_exact_vs_oracle_all_ranks": bool(flag.item()),
                     "checked": "first and last image of every rank, both axes, all pixels",
                     "checksum_i64": None}
    s = out.to(torch.int64).sum()
    if world > 1:
        dist.all_reduce(s)
    res["parity"]["checksum_i64"] = int(s.item())
    del x, out
    torch.cuda.empty_cache()
    return res


def _bind_to_gpu_numa(index):
    """Run this rank (and first-touch its pinned buffers) on the CPUs NVML reports as local to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = [64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = _bind_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from cupyimg_b200 import _ffi, _array, sharded
    from cupyimg_b200 import host as host_api
    from cupyimg_b200.scipy import ndimage as ndi
    from cupyimg_b200.scipy.ndimage import filters as F
    from oracle import oracle
    oracle.build()

    plan = None
    if world > 1:
        plan = sharded.ZSlabFilter((NZ, NY, NX), radius=RADIUS, mode=MODE, device=dev)
    halo_p2p = plan is not None and plan.p2p
    x = plan.slab if (plan is not None and plan.p2p) else torch.empty((NZ, NY, NX), dtype=torch.float32, device=dev)
    synth_planes(torch, dev, 2, rank * NZ, (rank + 1) * NZ, NY, NX, out=x)
    out = torch.empty((NZ, NY, NX), dtype=torch.float32, device=dev)

    if world > 1:
        def step():
            plan.gaussian_filter(x, SIGMA, truncate=TRUNCATE, output=out)
    else:
        def step():
            ndi.gaussian_filter(x, SIGMA, output=out, mode=MODE, truncate=TRUNCATE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    _ffi.LAUNCHES = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with sampler.region():
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        ev1.synchronize()
    barrier()
    launches = _ffi.LAUNCHES
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    voxels = NZ * NY * NX * world
    value = voxels * args.steps / (ms * 1e-3) / 1e9

    # ---- sustained blocks: the same call for ~0.5 s in blocks of 50, so that the clock sampler sees the
    #      kernel under load (the power cap, not the timed 20 steps, decides the sustained rate)
    blocks = []
    with sampler.region():
        for _ in range(24):                # a FIXED count: every rank must run the same number of steps
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50):
                step()
            b.record()
            b.synchronize()
            blocks.append(a.elapsed_time(b) / 50)
    barrier()

    # ---- parity of the timed configuration ----
    oracle.THREADS = max(1, min(16, os.cpu_count() or 1))
    try:
        ofn = lambda a: oracle.gaussian_filter(a, SIGMA, truncate=TRUNCATE, mode=MODE)
        brick = (0, 24, 0, 32, NX - 48, NX) if rank == 0 else (RADIUS + 200, RADIUS + 224, 240, 272, 0, 48)
        edges = [(rank == 0, rank == world - 1), (True, True), (True, True)]
        mabs, mrel, ok = _oracle_brick_check(np, ofn, x, out, RADIUS, brick, edges)
    finally:
        oracle.THREADS = 1
    parity = {"brick": list(brick), "max_abs": mabs, "max_rel": mrel, "within_tol": ok,
              "tol": "|a-b| <= 1e-6 max|b| + 1e-5 |b| vs the CPU oracle (scipy.ndimage arithmetic)"}
    if world > 1:
        lo = RADIUS if rank > 0 else 0
        hi = RADIUS if rank < world - 1 else 0
        ext = torch.empty((lo + NZ + hi, NY, NX), dtype=torch.float32, device=dev)
        if lo:
            synth_planes(torch, dev, 2, rank * NZ - lo, rank * NZ, NY, NX, out=ext[:lo])
        ext[lo:lo + NZ].copy_(x)
        if hi:
            synth_planes(torch, dev, 2, (rank + 1) * NZ, (rank + 1) * NZ + hi, NY, NX, out=ext[lo + NZ:])
        ref = torch.empty_like(out)
        inp = _array.ingest(ext)
        F._run_passes_window(inp, _array.ingest(ref), F._gaussian_specs(inp, SIGMA, 0, MODE, TRUNCATE), 0.0, None, lo)
        torch.cuda.synchronize()
        eq = torch.tensor([1 if torch.equal(ref, out) else 0, 1 if ok else 0], device=dev)
        dist.all_reduce(eq, op=dist.ReduceOp.MIN)
        parity["sharded_equals_single"] = bool(eq[0].item())
        parity["within_tol_all_ranks"] = bool(eq[1].item())
        del ext, ref
    parity["checksum_f64"] = _checksum(torch, dist, world, dev, out)

    # ---- dominant kernel, timed alone with events on the launching stream ----
    kt = []
    for _ in range(max(args.steps, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ndi.gaussian_filter(x, SIGMA, output=out, mode=MODE, truncate=TRUNCATE)
        b.record()
        b.synchronize()
        kt.append(a.elapsed_time(b))
    kernel_ms_isolated = statistics.mean(kt)       # includes the host-side launch gap of an idle GPU
    _ffi.LAUNCHES = 0
    ndi.gaussian_filter(x, SIGMA, output=out, mode=MODE, truncate=TRUNCATE)
    per_call_launches = _ffi.LAUNCHES            # 1 = fused kernel, 3 = per-axis tiled passes
    # one launch per step and launches queued back to back: the device-timed step IS the kernel's
    # average duration (CUDA events on the launching stream over the timed region)
    kernel_ms = ms / args.steps if (world == 1 and per_call_launches == 1) else kernel_ms_isolated

    # ---- end to end: pinned host -> device -> filter -> host, every step, through the public host API ----
    e2e_steps = min(args.steps, 5)
    in_b, in_e, win = host_api.slab_window(NZ * world, world, rank, RADIUS)
    hx = torch.empty((in_e - in_b, NY, NX), dtype=torch.float32, pin_memory=True)
    hx[win[0]:win[1]].copy_(x)
    if win[0]:
        hx[:win[0]].copy_(synth_planes(torch, dev, 2, in_b, in_b + win[0], NY, NX))
    if in_e - in_b > win[1]:
        hx[win[1]:].copy_(synth_planes(torch, dev, 2, in_b + win[1], in_e, NY, NX))
    hy = torch.empty((NZ, NY, NX), dtype=torch.float32, pin_memory=True)

    def e2e_step():
        # the public host-volume API: z-chunks streamed H2D -> filter -> D2H on three streams; for N > 1 every
        # rank streams its own slab of the host volume (plus the 8 overlap planes per side) — no GPU-GPU traffic
        host_api.gaussian_filter_host(hx, SIGMA, output=hy, mode=MODE, truncate=TRUNCATE, chunk_planes=32,
                                      out_window=None if world == 1 else win)

    e2e_step()
    step()                                     # `out` again holds this rank's slab of the (sharded) result
    barrier()
    e2e_equal = bool(torch.equal(hy, out.cpu()))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, 0.0 if e2e_equal else 1.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0].item())
        e2e_equal = t[1].item() == 0.0
    e2e_value = voxels * e2e_steps / e2e_s / 1e9
    parity["e2e_equals_device_result"] = e2e_equal
    # the bound of this path: the same bytes as plain concurrent H2D + D2H copies on two streams, no kernels
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    dbuf_in = torch.empty(tuple(hx.shape), dtype=torch.float32, device=dev)
    dbuf_out = torch.empty((NZ, NY, NX), dtype=torch.float32, device=dev)

    def copy_step():
        with torch.cuda.stream(s_h2d):
            dbuf_in.copy_(hx, non_blocking=True)
        with torch.cuda.stream(s_d2h):
            hy.copy_(dbuf_out, non_blocking=True)

    copy_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copy_step()
    barrier()
    copy_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([copy_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_s = float(t.item())
    copy_only_value = voxels * e2e_steps / copy_s / 1e9
    del hx, hy, dbuf_in, dbuf_out

    # ---- strong-scaling legs (BASELINE.json configs[2..4]) ----
    legs = {}
    if not args.no_legs:
        del x, out, plan
        torch.cuda.empty_cache()
        ls = max(3, min(args.steps, 5))
        legs["C5_gaussian_sigma4_2048"] = leg_zsharded(torch, dist, np, rank, world, dev, sampler,
                                                       "gaussian_filter sigma=4 (33 taps/axis) 2048^3 f32", 2048, "gauss", 4.0, ls, 2, 5)
        legs["C4_gradmag_sigma1.5_1024"] = leg_zsharded(torch, dist, np, rank, world, dev, sampler,
                                                        "gaussian_gradient_magnitude sigma=1.5 1024^3 f32", 1024, "gradmag", 1.5, ls, 2, 4)
        legs["C3_convolve1d_u16_64x2048x2048"] = leg_c3(torch, dist, np, rank, world, dev, sampler, ls, 2)
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = NZ * NY * NX * BYTES_PER_VOXEL
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        tr = _ncu_traffic() if per_call_launches == 1 else None
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_tfma = sms * FMA_PER_CLK_SM * sm_mhz * 1e6 / 1e12
        ach_tfma = NZ * NY * NX * FMA_PER_VOXEL / (kernel_ms * 1e-3) / 1e12
        fp32 = {"fma_per_voxel": FMA_PER_VOXEL, "achieved_tfma": ach_tfma, "peak_tfma": peak_tfma,
                "frac": ach_tfma / peak_tfma,
                "peak_source": "%d SMs x 128 FMA/clk x %.0f MHz (median SM clock sampled under load)" % (sms, sm_mhz),
                "floor_ms": NZ * NY * NX * FMA_PER_VOXEL / (peak_tfma * 1e12) * 1e3}
        hbm_frac = achieved / peak
        roofline = {"bound": "fp32" if fp32["frac"] > hbm_frac else "hbm",
                    "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": hbm_frac, "traffic": tr["dram_bytes_per_launch"] if tr else None,
                    "traffic_source": tr["source"] if tr else None,
                    "kernel": "fws_kernel (csrc/fused_ws.cu, warp-specialised fused y/x/z pass; 1 launch per call)" if per_call_launches == 1 else
                              "gaussian_filter call = %s launches (per-axis tiled passes)" % per_call_launches,
                    "kernel_ms": kernel_ms, "kernel_ms_single_launch_idle_gpu": kernel_ms_isolated,
                    "kernel_ms_sustained_blocks": {"median": statistics.median(blocks) if blocks else None,
                                                   "first": blocks[0] if blocks else None, "n_blocks": len(blocks),
                                                   "note": "24 blocks of 50 calls: the 1000 W power cap lowers the SM "
                                                           "clock in long runs (clocks.reasons)"},
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0,
                    "fp32": fp32}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "gaussian_filter sigma=2 truncate=4 (17 taps/axis) mode=reflect, "
                                   "512^3 float32 per GPU (BASELINE.json configs[1])",
                       "volume_per_gpu": [NZ, NY, NX], "global_volume": [NZ * world, NY, NX],
                       "sharding": "none" if world == 1 else
                                   ("z-slabs; the fused kernel reads the 8 halo planes per side from the neighbours' slabs "
                                    "(symmetric memory, TMA over NVLink); ready / done flags by stream memory operations"
                                    if halo_p2p else "z-slabs, 8-plane halo exchange over NCCL send/recv"),
                       "l2": "input 512 MiB + output 512 MiB per step, both larger than the 126 MB L2; no flush",
                       "numa_bound_cpus": numa_cpus},
            "roofline": roofline, "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int((NZ * world + 2 * RADIUS * (world - 1)) * NY * NX * 4),
                    "d2h_bytes_per_step": NZ * NY * NX * 4 * world,
                    "steps": e2e_steps,
                    "copy_only": {"value": copy_only_value, "unit": UNIT, "frac": e2e_value / copy_only_value,
                                  "what": "the same H2D + D2H bytes as plain concurrent copies on two streams, no kernel: "
                                          "the PCIe / host-memory bound of this path on this box"},
                    "api": "cupyimg_b200.host.gaussian_filter_host (pinned host in / out, 32-plane chunks, halos filled "
                           "device-to-device, 3 streams)" + ("" if world == 1 else
                           "; every rank streams its slab of the host volume with 8 overlap planes per side (host.slab_window)")},
            "gpu_launches": launches, "clocks": clocks, "legs": legs,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            planes = 512 if cores >= 8 else 128
            cpu_baseline(16, cores)
            v, t = cpu_baseline(planes, cores)
            sv, ss = scipy_baseline(64)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d x 512 x 512 z-slab (%s of the workload), oracle port (scipy.ndimage "
                                              "arithmetic) on %d threads, %.1f s wall" % (
                                                  planes, "all" if planes == 512 else "1/4", cores, t),
                                    "scipy_ndimage_1thread": {"value": sv, "unit": UNIT, "cores": 1, "sample": ss}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the strong-scaling legs (C3 / C4 / C5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
