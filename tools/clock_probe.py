"""SM clock / power / throttle reasons (NVML, sampled in-process every few ms) while one filter call is
looped back to back, with the per-block mean time: is the kernel running at the clock the roofline assumes?
python tools/clock_probe.py [n] [sigma] [mode] [seconds]"""
import sys
import threading
import time
import torch
import pynvml
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
mode = sys.argv[3] if len(sys.argv) > 3 else "reflect"
secs = float(sys.argv[4]) if len(sys.argv) > 4 else 2.0
what = sys.argv[5] if len(sys.argv) > 5 else "gauss"
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
x = torch.rand((n, n, n), device="cuda"); o = torch.empty_like(x)
fn = (lambda: ndi.gaussian_filter(x, sigma, output=o, mode=mode)) if what == "gauss" else \
     (lambda: ndi.gaussian_gradient_magnitude(x, sigma, output=o, mode=mode))
for _ in range(3):
    fn()
torch.cuda.synchronize()
samples, stop = [], False


def sampler():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.003)


th = threading.Thread(target=sampler); th.start()
time.sleep(0.05)
t0 = time.perf_counter()
blocks = []
while time.perf_counter() - t0 < secs:
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        fn()
    b.record(); b.synchronize()
    blocks.append((time.perf_counter() - t0, a.elapsed_time(b) / 50))
t1 = time.perf_counter()
time.sleep(0.05)
stop = True; th.join()
load = [s for s in samples if t0 + 0.01 <= s[0] <= t1]
idle = [s for s in samples if s[0] < t0 or s[0] > t1 + 0.02]
print("%s n=%d sigma=%g mode=%s: %d blocks of 50 calls" % (what, n, sigma, mode, len(blocks)))
for i in sorted(set([0, 1, 2, len(blocks) // 4, len(blocks) // 2, len(blocks) - 1])):
    print("  block at %.3f s: %.4f ms/call" % blocks[i])
clk = sorted(s[1] for s in load)
pw = sorted(s[2] for s in load)
reasons = 0
for s in load:
    reasons |= s[3]
print("  under load: %d samples, SM MHz min/median/max %d/%d/%d, power W median/max %.0f/%.0f, reasons bitmask 0x%x" % (
    len(load), clk[0], clk[len(clk) // 2], clk[-1], pw[len(pw) // 2], pw[-1], reasons))
if idle:
    print("  idle: SM MHz %s" % sorted(set(s[1] for s in idle)))
print("  max SM clock %d MHz, power limit %.0f W" % (pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM),
                                                   pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0))
