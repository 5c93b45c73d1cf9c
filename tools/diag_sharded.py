"""Where does a sharded step spend its time?  torchrun --nproc-per-node 2 tools/diag_sharded.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from cupyimg_b200 import sharded, _array
from cupyimg_b200.scipy import ndimage as ndi
from cupyimg_b200.scipy.ndimage import filters as F
x = torch.rand((512, 512, 512), device=dev); out = torch.empty_like(x)
plan = sharded.ZSlabFilter(x.shape, radius=8, mode="reflect", device=dev)
specs = F._gaussian_specs(_array.ingest(x), 2.0, 0, "reflect", 4.0)
compute = sharded._cuda_compute(specs, 0.0, None)

def timeit(name, fn, n=20):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    if rank == 0: print("%-44s %.3f ms" % (name, a.elapsed_time(b) / n), flush=True)

r, nz = 8, 512
timeit("full sharded step", lambda: plan.gaussian_filter(x, 2.0, output=out))
timeit("single-GPU filter of the slab (reference)", lambda: ndi.gaussian_filter(x, 2.0, output=out))
z0 = r if plan.has_lo else 0; z1 = nz - r if plan.has_hi else nz
timeit("interior window launch only", lambda: compute(x, out[z0:z1], z0))
def exch():
    for q in plan.post_exchange(x): q.wait()
timeit("halo exchange only (NCCL + 2 copies)", exch)
def strips():
    if plan.has_lo: compute(plan.lo_ext, out[:r], r)
    if plan.has_hi: compute(plan.hi_ext, out[nz - r:], r)
timeit("boundary strip launch(es) only", strips)
def python_only():
    F._gaussian_specs(_array.ingest(x), 2.0, 0, "reflect", 4.0)
timeit("python: building the pass specs", python_only)
dist.destroy_process_group()
