"""Host-resident volumes: stream z-chunks through the GPU with copies and compute overlapped.

A caller of the reference moves a volume to the device (``cupy.asarray``), filters it and moves
the result back (``cupy.asnumpy``): three serial steps, with the PCIe transfers dominating.  For
volumes that live in host memory this module pipelines the same work: the volume is cut into
z-chunks, each chunk (plus the r halo planes the z pass needs, taken straight from the host
array) is copied to the device on one stream, filtered on a second stream through the windowed
C-ABI call (``in_offset0``: halo planes are read, never written) and copied back on a third, with
three chunks in flight.  H2D and D2H run concurrently (full-duplex PCIe) and hide the kernels.

The result is bit-identical to filtering the whole volume at once: a chunk's window never touches
the ends of its device buffer except where those are the ends of the volume, so the boundary rule
is applied in exactly the same places.
"""
import torch

from . import _array
from .scipy.ndimage import filters as _filters

__all__ = ["gaussian_filter_host", "uniform_filter_host"]


def _run_chunked(x, out, specs, cval, dtype_mode, chunk_planes, device):
    if x.device.type != "cpu" or out.device.type != "cpu":
        raise TypeError("host pipeline expects CPU tensors (pinned memory recommended)")
    if x.dim() != 3 or tuple(out.shape) != tuple(x.shape):
        raise _array.OutputShapeError("host pipeline expects 3-D volumes of equal shape")
    if out.dtype != x.dtype:
        raise RuntimeError("host pipeline keeps the input dtype")
    nz, ny, nx = x.shape
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    r = max([s.radius() for s in specs if s.axis == 0], default=0)
    zmodes = [s.mode for s in specs if s.axis == 0]
    wrap = bool(zmodes) and zmodes[0] == 4
    if wrap or nz <= 2 * r or nz <= chunk_planes:
        # thin volumes and wrap-around halos: one shot (still correct, just not pipelined)
        d = x.to(dev, non_blocking=True)
        o = torch.empty_like(d)
        _filters._run_passes(_array.ingest(d), _array.ingest(o), specs, cval, dtype_mode)
        out.copy_(o)
        return out
    C = int(chunk_planes)
    NB = 3
    ibuf = [torch.empty((C + 2 * r, ny, nx), dtype=x.dtype, device=dev) for _ in range(NB)]
    obuf = [torch.empty((C, ny, nx), dtype=x.dtype, device=dev) for _ in range(NB)]
    s_in, s_run, s_out = (torch.cuda.Stream(dev) for _ in range(3))
    ev_in, ev_run, ev_out = {}, {}, {}
    start = torch.cuda.Event()
    start.record(torch.cuda.current_stream(dev))
    for s in (s_in, s_run, s_out):
        s.wait_event(start)
    chunks = [(z0, min(z0 + C, nz)) for z0 in range(0, nz, C)]
    for i, (z0, z1) in enumerate(chunks):
        lo, hi = max(z0 - r, 0), min(z1 + r, nz)
        b = i % NB
        with torch.cuda.stream(s_in):
            if i >= NB:
                s_in.wait_event(ev_run[i - NB])                 # the kernel that read this buffer is done
            src = ibuf[b][:hi - lo]
            src.copy_(x[lo:hi], non_blocking=True)
            ev_in[i] = torch.cuda.Event()
            ev_in[i].record(s_in)
        with torch.cuda.stream(s_run):
            s_run.wait_event(ev_in[i])
            if i >= NB:
                s_run.wait_event(ev_out[i - NB])                # the copy that drained this buffer is done
            dst = obuf[b][:z1 - z0]
            _filters._run_passes_window(_array.ingest(src), _array.ingest(dst), specs, cval, dtype_mode, z0 - lo)
            ev_run[i] = torch.cuda.Event()
            ev_run[i].record(s_run)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_run[i])
            out[z0:z1].copy_(dst, non_blocking=True)
            ev_out[i] = torch.cuda.Event()
            ev_out[i].record(s_out)
    done = torch.cuda.Event()
    done.record(s_out)
    torch.cuda.current_stream(dev).wait_event(done)
    for s in (s_in, s_run):
        torch.cuda.current_stream(dev).wait_stream(s)
    return out


def gaussian_filter_host(input, sigma, order=0, output=None, mode="reflect", cval=0.0, truncate=4.0, *,
                         chunk_planes=64, device=None, dtype_mode=None):
    """``gaussian_filter`` for a 3-D CPU tensor, streamed through the GPU in z-chunks.
    Returns a CPU tensor (``output`` or a new pinned one).  The call returns once the work is
    enqueued on the current stream; synchronise that stream before reading the result."""
    x = input
    if output is None:
        output = torch.empty(x.shape, dtype=x.dtype, pin_memory=True)
    probe = _array.DevArray(0, x.shape, [s * x.element_size() for s in x.stride()],
                            _array.to_numpy_dtype(x.dtype), 0, None)
    specs = _filters._gaussian_specs(probe, sigma, order, mode, truncate)
    return _run_chunked(x, output, specs, cval, dtype_mode, chunk_planes, device)


def uniform_filter_host(input, size=3, output=None, mode="reflect", cval=0.0, origin=0, *,
                        chunk_planes=64, device=None, dtype_mode=None):
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
