"""Does torch symmetric memory work on this box?  torchrun --nproc-per-node 2 tools/symm_probe.py"""
import os
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm
from torch._C._distributed_c10d import _SymmetricMemory as SM

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
buf = symm.empty((64, 512, 512), dtype=torch.float32, device=dev)
flags = symm.empty(64, dtype=torch.int32, device=dev)
hb = symm.rendezvous(buf, dist.group.WORLD)
hf = symm.rendezvous(flags, dist.group.WORLD)
print(rank, "rendezvous ok in %.2f s" % (time.time() - t0), "ptrs", [hex(p) for p in hb.buffer_ptrs], flush=True)
buf.fill_(float(rank + 1)); flags.zero_()
torch.cuda.synchronize(); dist.barrier()
peer = (rank + 1) % world
pbuf = hb.get_buffer(peer, (64, 512, 512), torch.float32)
pflags = hf.get_buffer(peer, (64,), torch.int32)
# remote flag write by stream memop
try:
    SM.stream_write_value32(pflags, 3, 41 + rank)
    torch.cuda.synchronize(); dist.barrier()
    print(rank, "flag written by peer via stream_write_value32:", int(flags[3].item()), flush=True)
except Exception as e:
    print(rank, "stream_write_value32 on peer failed:", repr(e)[:200], flush=True)
# copy-engine pull of peer planes
loc = torch.empty((16, 512, 512), device=dev)
for _ in range(3):
    loc.copy_(pbuf[:16])
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    loc.copy_(pbuf[:16])
b.record(); b.synchronize()
ms = a.elapsed_time(b) / 20
print(rank, "peer pull 16 MiB: %.3f ms = %.0f GB/s, value %.1f" % (ms, 16.78 / ms, float(loc[0, 0, 0])), flush=True)
dist.barrier()
dist.destroy_process_group()
