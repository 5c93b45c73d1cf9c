"""Host-resident volumes: stream z-chunks through the GPU with copies and compute overlapped.

A caller of the reference moves a volume to the device (``cupy.asarray``), filters it and moves
the result back (``cupy.asnumpy``): three serial steps, with the PCIe transfers dominating.  For
volumes that live in host memory this module pipelines the same work: the volume is cut into
z-chunks, each chunk is copied to the device on one stream — every plane crosses PCIe exactly once:
the r halo planes the z pass needs are filled device-to-device from the neighbouring chunk buffers —
filtered on a second stream through the windowed C-ABI call (``in_offset0``: halo planes are read,
never written) and copied back on a third, with four chunks in flight.  H2D and D2H run concurrently
(full-duplex PCIe) and hide the kernels.

The result is bit-identical to filtering the whole volume at once: a chunk's window never touches
the ends of its device buffer except where those are the ends of the volume, so the boundary rule
is applied in exactly the same places.
"""
import torch

from . import _array
from .scipy.ndimage import filters as _filters

__all__ = ["gaussian_filter_host", "uniform_filter_host", "slab_window"]


def slab_window(nz_global, world, rank, radius):
    """z-slab of rank ``rank`` of a host volume sharded over ``world`` GPUs with overlapping input slabs:
    returns ``(in_begin, in_end, out_window)`` — the rank loads global planes [in_begin, in_end) (its own
    planes plus up to ``radius`` planes of each neighbour) and filters ``out_window`` (in slab coordinates).
    No GPU-to-GPU traffic is needed on this path: the overlap planes cross PCIe twice instead."""
    z0, z1 = (rank * nz_global) // world, ((rank + 1) * nz_global) // world
    a, b = max(0, z0 - radius), min(nz_global, z1 + radius)
    return a, b, (z0 - a, z1 - a)


def _run_chunked(x, out, specs, cval, dtype_mode, chunk_planes, device, out_window=None):
    if x.device.type != "cpu" or out.device.type != "cpu":
        raise TypeError("host pipeline expects CPU tensors (pinned memory recommended)")
    if x.dim() != 3:
        raise _array.OutputShapeError("host pipeline expects 3-D volumes")
    nz_in = x.shape[0]
    wb, we = (0, nz_in) if out_window is None else (int(out_window[0]), int(out_window[1]))
    if not (0 <= wb < we <= nz_in):
        raise ValueError("out_window must satisfy 0 <= begin < end <= input planes")
    if tuple(out.shape) != (we - wb,) + tuple(x.shape[1:]):
        raise _array.OutputShapeError("host pipeline: output shape does not match the (windowed) input")
    if out.dtype != x.dtype:
        raise RuntimeError("host pipeline keeps the input dtype")
    nz, ny, nx = we - wb, x.shape[1], x.shape[2]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    r = max([s.radius() for s in specs if s.axis == 0], default=0)
    zmodes = [s.mode for s in specs if s.axis == 0]
    wrap = bool(zmodes) and zmodes[0] == 4
    windowed = (wb, we) != (0, nz_in)
    # planes below / above the window that the z pass reads (present in x): they replace the boundary rule
    pre, post = min(r, wb), min(r, nz_in - we)
    if windowed and (wrap or nz <= 2 * r or nz <= chunk_planes):
        a, b = wb - pre, we + post
        d = x[a:b].to(dev, non_blocking=True)
        o = torch.empty((nz, ny, nx), dtype=x.dtype, device=dev)
        if wrap and (pre < r or post < r):
            raise ValueError("a wrapped z axis needs the full halo on both sides of the window")
        _filters._run_passes_window(_array.ingest(d), _array.ingest(o), specs, cval, dtype_mode, wb - a)
        out.copy_(o)
        return out
    if wrap or nz <= 2 * r or nz <= chunk_planes:
        # thin volumes and wrap-around halos: one shot (still correct, just not pipelined)
        d = x.to(dev, non_blocking=True)
        o = torch.empty_like(d)
        _filters._run_passes(_array.ingest(d), _array.ingest(o), specs, cval, dtype_mode)
        out.copy_(o)
        return out
    C = int(chunk_planes)
    if C < 2 * r:
        C = 2 * r                                            # a chunk must hold its neighbours' halos
    NB = 4
    chunks = [(z0, min(z0 + C, we)) for z0 in range(wb, we, C)]      # in INPUT plane coordinates
    # a chunk thinner than the halo cannot serve its neighbours: merge it into its predecessor
    merged = [chunks[0]]
    for a, b in chunks[1:]:
        if b - a < max(r, 1):
            merged[-1] = (merged[-1][0], b)
        else:
            merged.append((a, b))
    chunks = merged
    C = max(b - a for a, b in chunks)
    # device chunk buffers laid out [r lower halo | C own planes | r upper halo].  Every plane crosses
    # PCIe exactly once: the halos are filled device-to-device from the neighbouring chunk buffers, so a
    # chunk is filtered one chunk late, when its successor has landed.
    ibuf = [torch.empty((C + 2 * r, ny, nx), dtype=x.dtype, device=dev) for _ in range(NB)]
    obuf = [torch.empty((C, ny, nx), dtype=x.dtype, device=dev) for _ in range(NB)]
    s_in, s_run, s_out = (torch.cuda.Stream(dev) for _ in range(3))
    ev_in, ev_run, ev_out = {}, {}, {}
    start = torch.cuda.Event()
    start.record(torch.cuda.current_stream(dev))
    for s in (s_in, s_run, s_out):
        s.wait_event(start)
    n = len(chunks)

    def load(i):
        z0, z1 = chunks[i]
        with torch.cuda.stream(s_in):
            if i >= NB:
                s_in.wait_event(ev_run[i - NB + 1])             # this buffer's chunk AND its successor are filtered
            ibuf[i % NB][r:r + (z1 - z0)].copy_(x[z0:z1], non_blocking=True)
            if i == 0 and pre:                                  # planes below the window come from the host too
                ibuf[0][r - pre:r].copy_(x[wb - pre:wb], non_blocking=True)
            if i == n - 1 and post:
                ibuf[i % NB][r + (z1 - z0):r + (z1 - z0) + post].copy_(x[we:we + post], non_blocking=True)
            ev_in[i] = torch.cuda.Event()
            ev_in[i].record(s_in)

    def run(i):
        z0, z1 = chunks[i]
        b, m = i % NB, z1 - z0
        with torch.cuda.stream(s_run):
            s_run.wait_event(ev_in[min(i + 1, n - 1)])          # own planes and the successor's have landed
            buf = ibuf[b]
            lo = 0
            if i > 0:
                pz0, pz1 = chunks[i - 1]
                buf[:r].copy_(ibuf[(i - 1) % NB][r + (pz1 - pz0) - r:r + (pz1 - pz0)], non_blocking=True)
            else:
                lo = r - pre                                    # pre == r inside a sharded volume, 0 at its start
            hi = r + m
            if i + 1 < n:
                buf[r + m:r + m + r].copy_(ibuf[(i + 1) % NB][r:2 * r], non_blocking=True)
                hi = r + m + r
            else:
                hi = r + m + post
            if i >= NB:
                s_run.wait_event(ev_out[i - NB])                # the copy that drained this output buffer is done
            dst = obuf[b][:m]
            _filters._run_passes_window(_array.ingest(buf[lo:hi]), _array.ingest(dst), specs, cval, dtype_mode, r - lo)
            ev_run[i] = torch.cuda.Event()
            ev_run[i].record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_run[i])
            out[z0 - wb:z1 - wb].copy_(dst, non_blocking=True)
            ev_out[i] = torch.cuda.Event()
            ev_out[i].record(s_out)

    for i in range(n):
        load(i)
        if i >= 1:
            run(i - 1)
    run(n - 1)
    done = torch.cuda.Event()
    done.record(s_out)
    torch.cuda.current_stream(dev).wait_event(done)
    for s in (s_in, s_run):
        torch.cuda.current_stream(dev).wait_stream(s)
    return out


def gaussian_filter_host(input, sigma, order=0, output=None, mode="reflect", cval=0.0, truncate=4.0, *,
                         chunk_planes=32, device=None, dtype_mode=None, out_window=None):
    """``gaussian_filter`` for a 3-D CPU tensor, streamed through the GPU in z-chunks.
    Returns a CPU tensor (``output`` or a new pinned one).  The call returns once the work is
    enqueued on the current stream; synchronise that stream before reading the result.
    ``out_window=(begin, end)``: filter only those input planes (the planes outside the window serve as
    z halo in place of the boundary rule) — one rank's share of a volume sharded with :func:`slab_window`."""
    x = input
    if output is None:
        oshape = x.shape if out_window is None else (out_window[1] - out_window[0],) + tuple(x.shape[1:])
        output = torch.empty(oshape, dtype=x.dtype, pin_memory=True)
    probe = _array.DevArray(0, x.shape, [s * x.element_size() for s in x.stride()],
                            _array.to_numpy_dtype(x.dtype), 0, None)
    specs = _filters._gaussian_specs(probe, sigma, order, mode, truncate)
    return _run_chunked(x, output, specs, cval, dtype_mode, chunk_planes, device, out_window)


def uniform_filter_host(input, size=3, output=None, mode="reflect", cval=0.0, origin=0, *,
                        chunk_planes=32, device=None, dtype_mode=None):
    """``uniform_filter`` for a 3-D CPU tensor, streamed through the GPU in z-chunks."""
    x = input
    if output is None:
        output = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    sizes = _filters._normalize_sequence(size, 3)
    origins = _filters._normalize_sequence(origin, 3)
    modes = _filters._normalize_sequence(mode, 3)
    specs = [_filters._PassSpec(a, None, _filters._check_origin(o, int(s)), _filters._check_mode(m),
                                uniform=True, size=int(s))
             for a, (s, o, m) in enumerate(zip(sizes, origins, modes)) if s > 1]
    return _run_chunked(x, output, specs, cval, dtype_mode, chunk_planes, device)
