"""world_size 2 / 3 gloo tests (CPU) of the z-slab sharding plan: halo exchange, windows,
end-rank boundary handling and the wrap ring.  The per-slab compute is injected — here the
CPU oracle (the checker); on GPUs it is the CUDA path (tests/test_sharded_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle

SIGMA = 1.0          # radius 4
NZ, NY, NX = 12, 9, 10


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_window_compute(mode):
    w = oracle.gaussian_kernel1d(SIGMA, 0, int(4 * SIGMA + 0.5))[::-1]

    def compute(src, dst, in_offset0):
        s = src.numpy()
        out = np.empty(tuple(dst.shape), np.float32)
        oracle._line_pass(s, out, 0, w, w.size, 0, mode, 0.0, in_offset=in_offset0)
        for axis in (1, 2):
            out = oracle.correlate1d(out, w, axis=axis, mode=mode)
        dst.copy_(torch.from_numpy(out))
    return compute


def _oracle_gradmag_window_compute(mode, sigma):
    """Per-slab gradient magnitude by the oracle: the three derivative filters on the windowed slab."""
    lw = int(4 * sigma + 0.5)
    g = oracle.gaussian_kernel1d(sigma, 0, lw)[::-1]
    d = oracle.gaussian_kernel1d(sigma, 1, lw)[::-1]

    def compute(src, dst, in_offset0):
        s = src.numpy()
        acc = np.zeros(tuple(dst.shape), np.float32)
        for a in range(3):
            w = [d if b == a else g for b in range(3)]
            out = np.empty(tuple(dst.shape), np.float32)
            oracle._line_pass(s, out, 0, w[0], w[0].size, 0, mode, 0.0, in_offset=in_offset0)
            for axis in (1, 2):
                out = oracle.correlate1d(out, w[axis], axis=axis, mode=mode)
            acc = acc + out * out
        dst.copy_(torch.from_numpy(np.sqrt(acc)))
    return compute


def _worker(rank, world, port, mode, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cupyimg_b200 import sharded
        rng = np.random.default_rng(99)
        vol = rng.random((NZ * world, NY, NX), dtype=np.float32)
        want = oracle.gaussian_filter(vol, SIGMA, mode=mode)
        x = torch.from_numpy(vol[rank * NZ:(rank + 1) * NZ].copy())
        plan = sharded.ZSlabFilter(x.shape, radius=4, mode=mode, device="cpu")
        out = plan.gaussian_filter(x, SIGMA, compute=_oracle_window_compute(mode))
        ok = np.array_equal(out.numpy(), want[rank * NZ:(rank + 1) * NZ])
        # a second call reuses the plan's halo buffers
        out2 = plan.gaussian_filter(x, SIGMA, compute=_oracle_window_compute(mode))
        ok = ok and np.array_equal(out2.numpy(), out.numpy())
        # gradient magnitude (C4's filter): one halo exchange serves the three derivative filters
        want3 = oracle.gaussian_gradient_magnitude(vol, SIGMA, mode=mode)
        out3 = plan.gaussian_gradient_magnitude(x, SIGMA, compute=_oracle_gradmag_window_compute(mode, SIGMA))
        ok = ok and np.array_equal(out3.numpy(), want3[rank * NZ:(rank + 1) * NZ])
        with open(os.path.join(result_dir, "rank%d" % rank), "w") as f:
            f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["reflect", "wrap", "constant", "mirror", "nearest"])
def test_zslab_plan_matches_global_filter(world, mode, tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, mode, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("rank%d" % r)).read_text() == "ok"


def test_batch_range_partitions():
    from cupyimg_b200 import sharded
    for n, w in [(64, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [sharded.batch_range(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_thin_slab_is_rejected():
    from cupyimg_b200 import sharded
    plan = sharded.ZSlabFilter((4, 8, 8), radius=8, device="cpu")     # world 1: fine, no exchange
    assert plan.world == 1
