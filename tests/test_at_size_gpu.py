"""Parity at the sizes BASELINE.json states (VERDICT r1 "weak" 1): the exact call bench.py times
(512^3, C2), gaussian_gradient_magnitude sigma=1.5 at 1024^3 (C4) and gaussian_filter sigma=4 at
2048^3 (C5: 8.6 G voxels, 32 GiB in + 32 GiB out, element indices beyond 2^31 / 2^32 — the volumes
for which the reference switches its index type, /root/reference/cupyimg/scipy/ndimage/_util.py:122-134),
plus one GPU's share and the whole stack of C3 (64 x 2048^2 uint16, bit-exact).

scipy needs minutes for these volumes, so the oracle runs on sub-bricks: every corner, faces, the
interior and bricks straddling element index 2^31 / 2^32.  A brick is cut with a halo of one filter
radius where the volume continues and ends exactly at the array face where it does not, so the
oracle applies the boundary rule in the same places as the full-volume filter does.
"""
import itertools

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def _free_gib():
    import torch
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def brick_origins(shape, b, extra=()):
    """All 8 corners, the 6 face centres, the centre, plus ``extra`` origins."""
    nz, ny, nx = shape
    ends = [(0, (n - b) // 2, n - b) for n in (nz, ny, nx)]
    out = [(ends[0][i], ends[1][j], ends[2][k]) for i, j, k in itertools.product((0, 2), repeat=3)]
    mid = (ends[0][1], ends[1][1], ends[2][1])
    for ax in range(3):
        for side in (0, 2):
            o = list(mid)
            o[ax] = ends[ax][side]
            out.append(tuple(o))
    out.append(mid)
    out.extend(extra)
    return out


def check_bricks(x, y, origins, b, halo, cpu_filter, rtol, atol_rel):
    """Compare y[brick] with cpu_filter(x[brick + halo]) cropped to the brick.  Returns (max_abs, max_rel)."""
    shape = tuple(x.shape)
    max_abs = max_rel = 0.0
    for o in origins:
        lo = [max(a - halo, 0) for a in o]
        hi = [min(a + b + halo, n) for a, n in zip(o, shape)]
        sub = x[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]].cpu().numpy()
        want = cpu_filter(sub)
        off = [a - l for a, l in zip(o, lo)]
        want = want[off[0]:off[0] + b, off[1]:off[1] + b, off[2]:off[2] + b].astype(np.float64)
        got = y[o[0]:o[0] + b, o[1]:o[1] + b, o[2]:o[2] + b].cpu().numpy().astype(np.float64)
        err = np.abs(got - want)
        tol = atol_rel * np.abs(want).max() + rtol * np.abs(want)
        assert (err <= tol).all(), "brick at %s: max err %.3g (tol %.3g)" % (o, err.max(), tol.min())
        max_abs = max(max_abs, float(err.max()))
        nz = np.abs(want) > 1e-3 * np.abs(want).max()
        max_rel = max(max_rel, float((err[nz] / np.abs(want[nz])).max()))
    return max_abs, max_rel


def test_c2_bench_call_512():
    """bench.py's call: gaussian_filter(sigma=2) on 512^3 float32, mode reflect (one fused launch, the
    148-CTA plan); rtol 1e-5 (north_star) with atol 1e-6 max|b|."""
    import torch
    from cupyimg_b200.scipy import ndimage as ndi
    n = 512
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand((n, n, n), device="cuda", generator=g)
    y = torch.empty_like(x)
    ndi.gaussian_filter(x, 2.0, output=y, mode="reflect", truncate=4.0)
    check_bricks(x, y, brick_origins(x.shape, 40, extra=[(100, 3, 470), (250, 500 - 40, 9)]), 40, 8,
                 lambda s: oracle.gaussian_filter(s, 2.0, mode="reflect"), 1e-5, 1e-6)
    # the other boundary modes at the same size (same tile plan, different edge handling)
    # (wrap: a wrapped halo comes from the far side of the array, see test_c2_wrap_512_periodic)
    for mode in ("constant", "nearest", "mirror"):
        ndi.gaussian_filter(x, 2.0, output=y, mode=mode)
        corners = brick_origins(x.shape, 40)[:8]
        check_bricks(x, y, corners, 40, 8, lambda s, m=mode: oracle.gaussian_filter(s, 2.0, mode=m), 1e-5, 1e-6)


def test_c2_wrap_512_periodic():
    """mode='wrap' at 512^3: filtering a volume rolled by (a, b, c) == rolling the filtered volume."""
    import torch
    from cupyimg_b200.scipy import ndimage as ndi
    n = 512
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.rand((n, n, n), device="cuda", generator=g)
    y = ndi.gaussian_filter(x, 2.0, mode="wrap")
    xr = torch.roll(x, shifts=(37, 250, 101), dims=(0, 1, 2))
    yr = ndi.gaussian_filter(xr, 2.0, mode="wrap")
    del xr
    d = (torch.roll(y, shifts=(37, 250, 101), dims=(0, 1, 2)) - yr).abs().max()
    assert float(d) <= 2e-6
    # and a brick in the interior against the oracle
    check_bricks(x, y, [(200, 210, 220)], 40, 8, lambda s: oracle.gaussian_filter(s, 2.0, mode="reflect"), 1e-5, 1e-6)


def test_c4_gradient_magnitude_1024():
    """C4: gaussian_gradient_magnitude(sigma=1.5) on 1024^3 float32 (4 GiB in, 4 GiB out)."""
    import torch
    from cupyimg_b200.scipy import ndimage as ndi
    if _free_gib() < 14:
        pytest.skip("needs 14 GiB of free device memory")
    n = 1024
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand((n, n, n), device="cuda", generator=g)
    y = torch.empty_like(x)
    ndi.gaussian_gradient_magnitude(x, 1.5, output=y)
    # element index 2^30 (byte offset 2^32) is plane z = 1024: not reached; byte offset 2^31 is z = 512
    extra = [(512 - 20, 0, 0), (512 - 20, 1024 - 40, 1024 - 40), (300, 17, 900)]
    mx = check_bricks(x, y, brick_origins(x.shape, 40, extra), 40, 6,
                      lambda s: oracle.gaussian_gradient_magnitude(s, 1.5), 1e-5, 2e-6)
    print("C4 1024^3 max abs / max rel:", mx)


def test_c5_gaussian_sigma4_2048():
    """C5: gaussian_filter(sigma=4; 33 taps) on 2048^3 float32 on ONE GPU: 32 GiB in + 32 GiB out,
    element indices up to 2^33 — the first run of the 64-bit paths at size."""
    import torch
    from cupyimg_b200.scipy import ndimage as ndi
    if _free_gib() < 110:
        pytest.skip("needs 110 GiB of free device memory")
    n = 2048
    x = torch.empty((n, n, n), device="cuda")
    g = torch.Generator(device="cuda").manual_seed(1234)
    for z in range(0, n, 256):                       # generated per slab: torch.rand's temporary stays small
        x[z:z + 256] = torch.rand((256, n, n), device="cuda", generator=g)
    y = torch.empty_like(x)
    ndi.gaussian_filter(x, 4.0, output=y)
    torch.cuda.synchronize()
    b = 32
    # element index 2^31 = plane 512, 2^32 = plane 1024, 2^32 + 2^31 = plane 1536 (x = y = 0)
    extra = [(512 - 16, 0, 0), (1024 - 16, 0, 0), (1536 - 16, 0, 0), (1024 - 16, n - b, n - b), (700, 1000, 5)]
    mx = check_bricks(x, y, brick_origins(x.shape, b, extra), b, 16,
                      lambda s: oracle.gaussian_filter(s, 4.0), 1e-5, 1e-6)
    print("C5 2048^3 max abs / max rel:", mx)
    # size-independent property on the whole volume: reflect preserves the mean
    mean_in = float(sum(x[z:z + 256].double().sum() for z in range(0, n, 256))) / n ** 3
    mean_out = float(sum(y[z:z + 256].double().sum() for z in range(0, n, 256))) / n ** 3
    assert abs(mean_in - mean_out) < 1e-6
    del x, y
    torch.cuda.empty_cache()


def test_c3_uint16_stack_bit_exact():
    """C3: convolve1d along axis 1 then axis 2 of a 64 x 2048 x 2048 uint16 stack, mode mirror:
    the whole stack on one GPU (512 MiB), bit-exact against the oracle on 6 of the 64 images
    (first, last and four in between) — the kernels see the full (64, 2048, 2048) geometry."""
    import torch
    from cupyimg_b200.scipy import ndimage as ndi
    w = oracle.gaussian_kernel1d(1.5, 0, 4)
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.randint(0, 65536, (64, 2048, 2048), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
    y = ndi.convolve1d(ndi.convolve1d(x, w, axis=1, mode="mirror"), w, axis=2, mode="mirror")
    assert y.dtype == torch.uint16
    for i in (0, 1, 17, 31, 32, 63):
        img = x[i].view(torch.int16).cpu().numpy().view(np.uint16)
        want = oracle.convolve1d(oracle.convolve1d(img, w, axis=0, mode="mirror"), w, axis=1, mode="mirror")
        got = y[i].view(torch.int16).cpu().numpy().view(np.uint16)
        assert np.array_equal(got, want), "image %d differs" % i
