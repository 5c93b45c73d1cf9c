"""Randomised differential test of the fused / streaming float32 paths against the oracle."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from cupyimg_b200.scipy import ndimage as ndi
from oracle import oracle
modes = ["reflect", "constant", "nearest", "mirror", "wrap"]


def run(seed, iterations, verbose=True):
  rng = np.random.default_rng(seed)
  bad = 0
  for it in range(iterations):
      nd = int(rng.choice([2, 3, 3, 3]))
      shape = tuple(int(rng.integers(1, 90)) for _ in range(nd - 2)) + (int(rng.integers(2, 400)), 4 * int(rng.integers(1, 90)))
      if rng.random() < 0.2:
          shape = shape[:-1] + (int(rng.integers(2, 300)),)          # unaligned rows: per-axis fallback
      x = rng.random(shape).astype(np.float32)
      mode = [str(rng.choice(modes)) for _ in range(nd)] if rng.random() < 0.5 else str(rng.choice(modes))
      kind = rng.choice(["gauss", "gauss", "uniform", "sobel", "gradmag", "gauss1d", "max"])
      xd = torch.from_numpy(x).cuda()
      try:
          if kind == "gauss":
              sig = [float(rng.choice([0.0, 0.6, 1.0, 1.5, 2.0, 2.2, 3.0])) for _ in range(nd)]
              want = oracle.gaussian_filter(x, sig, mode=mode); got = ndi.gaussian_filter(xd, sig, mode=mode)
          elif kind == "uniform":
              sz = [int(rng.integers(1, 8)) for _ in range(nd)]
              want = oracle.uniform_filter(x, sz, mode=mode); got = ndi.uniform_filter(xd, sz, mode=mode)
          elif kind == "sobel":
              ax = int(rng.integers(0, nd))
              want = oracle.sobel(x, ax, mode=mode); got = ndi.sobel(xd, ax, mode=mode)
          elif kind == "gradmag":
              s = float(rng.choice([0.7, 1.0, 1.5, 2.0]))
              want = oracle.gaussian_gradient_magnitude(x, s, mode=mode); got = ndi.gaussian_gradient_magnitude(xd, s, mode=mode)
          elif kind == "gauss1d":
              ax = int(rng.integers(0, nd)); s = float(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0, 4.0])); m = mode if isinstance(mode, str) else mode[0]
              want = oracle.gaussian_filter1d(x, s, axis=ax, mode=m); got = ndi.gaussian_filter1d(xd, s, axis=ax, mode=m)
          else:
              sz = [int(rng.integers(1, 10)) for _ in range(nd)]
              want = oracle.maximum_filter(x, sz, mode=mode); got = ndi.maximum_filter(xd, sz, mode=mode)
      except Exception as e:                                        # shapes smaller than a radius etc. must fail alike
          print("EXC", kind, shape, mode, type(e).__name__, e); bad += 1; continue
      got = got.cpu().numpy()
      err = np.abs(got.astype(np.float64) - want)
      tol = 2e-6 * max(float(np.abs(want).max()), 1e-30) + 1e-5 * np.abs(want) + (1e-6 if kind in ("sobel", "gradmag") else 0)
      if kind == "max":
          ok = np.array_equal(got, want)
      else:
          ok = bool((err <= tol).all())
      if not ok:
          bad += 1
          print("MISMATCH", kind, shape, mode, float(err.max()))
  if verbose:
      print("fuzz done, failures:", bad)
  return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 80) else 0)
