"""Run the differential fuzz over many seeds (python tools/fuzz_many.py FIRST LAST)."""
import sys
sys.path.insert(0, "tools")
import fuzz_fused
a, b = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(a, b):
    bad += fuzz_fused.run(seed, 60, verbose=False)
    bad += fuzz_fused.run_exact(1000 + seed, 60, verbose=False)
print("seeds %d..%d failures: %d" % (a, b, bad))
