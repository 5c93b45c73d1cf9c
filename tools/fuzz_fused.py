"""Randomised differential test of the fused / streaming float32 paths against the oracle."""
import os
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


def run_exact(seed, iterations, verbose=True):
    """Integer / float64 paths: bit-exact against the oracle."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
    from helpers import to_device, to_host
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(iterations):
        dt = str(rng.choice(["uint8", "int16", "uint16", "int32", "float64", "int8", "uint32"]))
        nd = int(rng.choice([1, 2, 3, 3]))
        shape = tuple(int(rng.integers(1, 40)) for _ in range(nd - 1)) + (int(rng.choice([int(rng.integers(1, 300)), 8 * int(rng.integers(1, 40))])),)
        if nd >= 2 and rng.random() < 0.6:
            shape = shape[:-2] + (int(rng.integers(2, 200)), shape[-1])
        if np.dtype(dt).kind == "f":
            x = (rng.standard_normal(shape) * 100).astype(dt)
        else:
            info = np.iinfo(dt)
            x = rng.integers(max(info.min, -30000), min(info.max, 30000), shape, endpoint=True).astype(dt)
        mode = str(rng.choice(modes))
        kind = str(rng.choice(["conv1d", "conv1d", "gauss", "uniform", "sobel", "min", "max", "corr_nd", "laplace"]))
        xd = to_device(x)
        kw = dict(mode=mode, cval=float(rng.choice([0.0, 3.0, -2.5, 7.7])))
        try:
            if kind == "conv1d":
                r = int(rng.integers(1, 9)); ax = int(rng.integers(0, nd))
                w = oracle.gaussian_kernel1d(max(r / 3.0, 0.5), int(rng.choice([0, 0, 1])), r)
                org = int(rng.choice([0, 0, -1, 1])) if r > 1 else 0
                want = oracle.convolve1d(x, w, axis=ax, origin=org, **kw); got = ndi.convolve1d(xd, w, axis=ax, origin=org, **kw)
            elif kind == "gauss":
                s = float(rng.choice([0.7, 1.0, 1.5, 2.0])); want = oracle.gaussian_filter(x, s, **kw); got = ndi.gaussian_filter(xd, s, **kw)
            elif kind == "uniform":
                sz = int(rng.integers(1, 7)); want = oracle.uniform_filter(x, sz, **kw); got = ndi.uniform_filter(xd, sz, **kw)
            elif kind == "sobel":
                ax = int(rng.integers(0, nd)); want = oracle.sobel(x, ax, **kw); got = ndi.sobel(xd, ax, **kw)
            elif kind == "laplace":
                want = oracle.laplace(x, **kw); got = ndi.laplace(xd, **kw)
            elif kind in ("min", "max"):
                sz = [int(rng.integers(1, 10)) for _ in range(nd)]
                f = "minimum_filter" if kind == "min" else "maximum_filter"
                want = getattr(oracle, f)(x, sz, **kw); got = getattr(ndi, f)(xd, sz, **kw)
            else:
                ws = tuple(int(rng.integers(1, 5)) for _ in range(nd)); w = rng.standard_normal(ws)
                want = oracle.correlate(x, w, **kw); got = ndi.correlate(xd, w, **kw)
        except Exception as e:
            print("EXC", kind, dt, shape, mode, type(e).__name__, e); bad += 1; continue
        got = to_host(got)
        if got.dtype != want.dtype or not np.array_equal(got, want):
            if dt == "float64" and kind == "uniform" and np.allclose(
                    got, want, rtol=1e-12, atol=1e-15 * float(np.abs(x).max()) * max(shape)):
                continue        # scipy's running sum drifts along the line; the kernel sums each window exactly
            bad += 1
            print("MISMATCH", kind, dt, shape, mode, kw)
    if verbose:
        print("exact fuzz done, failures:", bad)
    return bad
