"""Host-side cost of one sharded step (python -m torch.distributed.run --nproc-per-node N tools/host_step_time.py):
wall time to ENQUEUE 200 steps without synchronising, next to the GPU time of the same steps."""
import os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
from cupyimg_b200 import sharded
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
plan = sharded.ZSlabFilter((512, 512, 512), radius=8, mode="reflect", device=dev)
x = plan.slab if plan.p2p else torch.empty((512, 512, 512), device=dev)
x.uniform_(); out = torch.empty_like(x)
for _ in range(10):
    plan.gaussian_filter(x, 2.0, output=out)
torch.cuda.synchronize(); dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); a.record()
for _ in range(200):
    plan.gaussian_filter(x, 2.0, output=out)
t1 = time.perf_counter(); b.record(); b.synchronize()
print("rank %d: host enqueue %.1f us/step, GPU %.1f us/step (%s)" % (rank, (t1 - t0) / 200 * 1e6, a.elapsed_time(b) / 200 * 1e3, plan.last_backend[:40]), flush=True)
dist.barrier(); dist.destroy_process_group()
