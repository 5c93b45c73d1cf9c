#!/usr/bin/env python
"""bench.py — headline benchmark: gaussian_filter(sigma=2, truncate=4; 17 taps/axis) on a
512^3 float32 volume per GPU (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one gaussian_filter call over one resident synthetic volume.  N > 1 (launched
by torchrun, one rank per GPU) is weak scaling: every rank owns one 512^3 z-slab of a
(512*N) x 512 x 512 volume and the slabs exchange 8-plane halos over NCCL inside the step.
Prints ONE JSON line (rank 0).  Nothing here reads /root/reference.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NZ = NY = NX = 512
SIGMA, TRUNCATE, MODE = 2.0, 4.0, "reflect"
BYTES_PER_VOXEL = 8            # read f32 once + write f32 once per API call
METRIC = "gaussian_filter 512^3 f32 throughput"
UNIT = "Gvoxel/s"


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
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.file = index, None, None

    def start(self):
        try:
            self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.file.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.file.name)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def _oracle_gaussian_mt(x, threads):
    """The CPU port (oracle/) with every 1-D pass split over `threads` host threads (disjoint
    line ranges; the ctypes call releases the GIL).  Bit-identical to the serial oracle."""
    from oracle import oracle
    oracle.THREADS = threads
    try:
        return oracle.gaussian_filter(x, SIGMA, truncate=TRUNCATE, mode=MODE)
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "gaussian_filter sigma=2 truncate=4 (17 taps/axis) mode=reflect, 512^3 float32",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from cupyimg_b200 import _ffi
    from cupyimg_b200.scipy import ndimage as ndi

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.rand((NZ, NY, NX), device=dev, generator=g)
    out = torch.empty_like(x)

    if world > 1:
        from cupyimg_b200 import sharded
        plan = sharded.ZSlabFilter(x.shape, radius=int(TRUNCATE * SIGMA + 0.5), mode=MODE, device=dev)

        def step():
            plan.gaussian_filter(x, SIGMA, truncate=TRUNCATE, output=out)
    else:
        def step():
            ndi.gaussian_filter(x, SIGMA, output=out, mode=MODE, truncate=TRUNCATE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler (an nvidia-smi child polling every 200 ms) starts BEFORE the warm-up: its process
    # start-up and NVML initialisation otherwise land inside the few-millisecond timed region and perturb it
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.5)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    _ffi.LAUNCHES = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
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
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end: pinned host -> device -> filter -> host, every step ----
    e2e_steps = min(args.steps, 5)
    hx = torch.empty((NZ, NY, NX), dtype=torch.float32, pin_memory=True)
    hx.copy_(x)
    hy = torch.empty((NZ, NY, NX), dtype=torch.float32, pin_memory=True)
    dx = torch.empty_like(x)

    from cupyimg_b200 import host as host_api

    def e2e_step():
        if world > 1:
            dx.copy_(hx, non_blocking=True)
            plan.gaussian_filter(dx, SIGMA, truncate=TRUNCATE, output=out)
            hy.copy_(out, non_blocking=True)
        else:
            # the public host-volume API: z-chunks streamed H2D -> filter -> D2H on three streams
            host_api.gaussian_filter_host(hx, SIGMA, output=hy, mode=MODE, truncate=TRUNCATE, chunk_planes=32)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = voxels * e2e_steps / e2e_s / 1e9

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = NZ * NY * NX * BYTES_PER_VOXEL
        # one API call = the algorithmic bytes; with the fused kernel it is one launch, otherwise
        # the per-axis launches share the call and the slowest of them is reported below
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        tr = _ncu_traffic() if per_call_launches == 1 else None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": tr["dram_bytes_per_launch"] if tr else None,
                    "traffic_source": tr["source"] if tr else None,
                    "kernel": "fused3d_f32 (1 launch per call)" if per_call_launches == 1 else
                              "gaussian_filter call = %s launches (per-axis tiled passes)" % per_call_launches,
                    "kernel_ms": kernel_ms, "kernel_ms_single_launch_idle_gpu": kernel_ms_isolated,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "gaussian_filter sigma=2 truncate=4 (17 taps/axis) mode=reflect, "
                                   "512^3 float32 per GPU (BASELINE.json configs[1])",
                       "volume_per_gpu": [NZ, NY, NX], "global_volume": [NZ * world, NY, NX],
                       "sharding": "none" if world == 1 else "z-slabs, 8-plane halo exchange over NCCL send/recv",
                       "l2": "input 512 MiB + output 512 MiB per step, both larger than the 126 MB L2; no flush"},
            "roofline": roofline, "e2e": {"value": e2e_value, "unit": UNIT,
                                           "h2d_bytes_per_step": NZ * NY * NX * 4 * world,
                                           "d2h_bytes_per_step": NZ * NY * NX * 4 * world,
                                           "steps": e2e_steps,
                                           "api": "cupyimg_b200.host.gaussian_filter_host (pinned host in / out, "
                                                  "32-plane chunks, halos filled device-to-device, 3 streams)" if world == 1 else
                                                  "pinned H2D copy + sharded.ZSlabFilter.gaussian_filter + D2H copy"},
            "gpu_launches": launches, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            planes = 512 if cores >= 8 else 128
            cpu_baseline(16, cores)
            v, t = cpu_baseline(planes, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d x 512 x 512 z-slab (%s of the workload), oracle port (scipy.ndimage "
                                              "arithmetic) on %d threads, %.1f s wall" % (
                                                  planes, "all" if planes == 512 else "1/4", cores, t)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
