"""End-to-end (pinned host -> GPU -> pinned host) gaussian_filter sigma=2 on 512^3 f32 for several chunk sizes."""
import sys, time
import torch
sys.path.insert(0, ".")
from cupyimg_b200 import host
hx = torch.rand((512, 512, 512)).pin_memory()
hy = torch.empty((512, 512, 512), dtype=torch.float32, pin_memory=True)
for c in [int(a) for a in sys.argv[1:]] or [16, 32, 64, 128]:
    host.gaussian_filter_host(hx, 2.0, output=hy, chunk_planes=c)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        host.gaussian_filter_host(hx, 2.0, output=hy, chunk_planes=c)
    torch.cuda.synchronize()
    t = (time.perf_counter() - t0) / 5
    print("chunk_planes %4d  %.2f ms  %.2f Gvoxel/s  (%.1f GB/s each way)" % (c, t * 1e3, 512**3 / t / 1e9, 512**3 * 4 / t / 1e9))
# plain copies for reference: H2D alone, D2H alone, both concurrently
d = torch.empty((512, 512, 512), device="cuda"); d2 = torch.empty_like(d)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n
t = timed(lambda: d.copy_(hx, non_blocking=True)); print("H2D alone   %.2f ms  %.1f GB/s" % (t * 1e3, 0.536870912 / t))
t = timed(lambda: hy.copy_(d2, non_blocking=True)); print("D2H alone   %.2f ms  %.1f GB/s" % (t * 1e3, 0.536870912 / t))
def both():
    with torch.cuda.stream(s1): d.copy_(hx, non_blocking=True)
    with torch.cuda.stream(s2): hy.copy_(d2, non_blocking=True)
t = timed(both); print("H2D + D2H concurrently  %.2f ms  %.1f GB/s each way" % (t * 1e3, 0.536870912 / t))
