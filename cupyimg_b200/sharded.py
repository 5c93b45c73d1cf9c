"""Multi-GPU sharding of the separable filters: one process per GPU (torch.distributed).

* 3-D volumes shard into contiguous z-slabs (slowest axis).  A separable filter of radius
  r along z needs r planes of RAW input from each z-neighbour, once, before compute — the
  x / y passes are local.  :class:`ZSlabFilter` posts the halo planes as NCCL send/recv
  pairs on a side stream, launches the interior planes (which need no halo) on the compute
  stream at once, and launches the two r-plane boundary strips when the halos have landed,
  so the exchange hides behind the interior compute.  Output stays sharded.
* Batched 2-D stacks (filtered axes never split) shard over the batch axis with no
  communication: :func:`batch_range`.

The reference is single-GPU (no collectives anywhere); this is the
B200 scale-out of its per-axis loops (filters.py:651-662, :777-789).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _array
from .scipy.ndimage import filters as _filters


def batch_range(n_items, world_size, rank):
    """[begin, end) of the images rank ``rank`` owns out of ``n_items`` (no data-path collective)."""
    return (rank * n_items) // world_size, ((rank + 1) * n_items) // world_size


def _cuda_compute(specs, cval, dtype_mode):
    def compute(src, dst, in_offset0):
        _filters._run_passes_window(_array.ingest(src), _array.ingest(dst), specs, cval, dtype_mode, in_offset0)
    return compute


class ZSlabFilter:
    """Halo-exchange plan for one z-slab of a volume sharded over ``group``.

    slab_shape: (nz_local, ny, nx) — the same on every rank.
    radius:     largest tap radius along z the plan must serve.
    mode:       boundary mode along z (the global volume's ends; ``wrap`` closes the ring).
    """

    def __init__(self, slab_shape, radius, mode="reflect", device=None, dtype=torch.float32, group=None,
                 cval=0.0):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nz, self.ny, self.nx = (int(s) for s in slab_shape)
        self.r = int(radius)
        zmode = mode if isinstance(mode, str) else mode[0]
        self.zmode = _filters._check_mode(zmode)
        self.mode = mode
        self.cval = cval
        if self.world > 1 and self.nz < 2 * self.r:
            raise ValueError("slab of %d planes is too thin for radius %d: use fewer ranks (replicas only)"
                             % (self.nz, self.r))
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.wrap = self.zmode == 4
        self.has_lo = self.world > 1 and (self.rank > 0 or self.wrap)
        self.has_hi = self.world > 1 and (self.rank < self.world - 1 or self.wrap)
        r = self.r
        shape = (3 * r, self.ny, self.nx)
        # [halo | 2r own planes] and [2r own planes | halo]: all a boundary strip of r planes reads
        self.lo_ext = torch.empty(shape, dtype=dtype, device=self.device) if self.has_lo and r else None
        self.hi_ext = torch.empty(shape, dtype=dtype, device=self.device) if self.has_hi and r else None
        self.on_cuda = self.device.type == "cuda"
        self.comm_stream = torch.cuda.Stream(self.device) if self.on_cuda else None
        if self.world > 1 and self.on_cuda:
            # the interior launch overlaps the NCCL send/recv kernels: leave them a few SMs, otherwise a
            # grid sized to fill every SM in one wave waits for the SMs NCCL holds and runs two waves
            os.environ.setdefault("SEPFILT_RESERVE_SMS", "8")

    # -- halo exchange ---------------------------------------------------------------
    def _peer(self, step):
        return (self.rank + step) % self.world

    def _global_rank(self, r):
        return dist.get_global_rank(self.group, r) if self.group is not None else r

    def post_exchange(self, x):
        """Send my first / last r planes, receive the neighbours' into lo_ext / hi_ext."""
        r, nz = self.r, self.nz
        if r == 0 or self.world == 1:
            return []
        ops = []
        if self.has_hi:
            ops.append(dist.P2POp(dist.isend, x[nz - r:], self._global_rank(self._peer(+1)), self.group))
        if self.has_lo:
            ops.append(dist.P2POp(dist.irecv, self.lo_ext[:r], self._global_rank(self._peer(-1)), self.group))
        if self.has_lo:
            ops.append(dist.P2POp(dist.isend, x[:r], self._global_rank(self._peer(-1)), self.group))
        if self.has_hi:
            ops.append(dist.P2POp(dist.irecv, self.hi_ext[2 * r:], self._global_rank(self._peer(+1)), self.group))
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if self.has_lo:
            self.lo_ext[r:].copy_(x[:2 * r], non_blocking=True)
        if self.has_hi:
            self.hi_ext[:2 * r].copy_(x[nz - 2 * r:], non_blocking=True)
        return reqs

    # -- compute ---------------------------------------------------------------------
    def run(self, x, output, compute):
        """``compute(src, dst, in_offset0)`` filters the planes of ``src`` that ``dst`` windows."""
        r, nz = self.r, self.nz
        if tuple(x.shape) != (nz, self.ny, self.nx) or tuple(output.shape) != tuple(x.shape):
            raise _array.OutputShapeError("slab shape does not match the plan")
        if self.world == 1 or r == 0:
            compute(x, output, 0)
            return output
        z0 = r if self.has_lo else 0
        z1 = nz - r if self.has_hi else nz
        if self.on_cuda:
            main = torch.cuda.current_stream(self.device)
            self.comm_stream.wait_stream(main)            # x must be complete before it is sent
            with torch.cuda.stream(self.comm_stream):
                reqs = self.post_exchange(x)
                for q in reqs:
                    q.wait()
                done = torch.cuda.Event()
                done.record(self.comm_stream)
            compute(x, output[z0:z1], z0)                  # interior: overlaps the exchange
            main.wait_event(done)
        else:
            for q in self.post_exchange(x):
                q.wait()
            compute(x, output[z0:z1], z0)
        if self.has_lo:
            compute(self.lo_ext, output[:r], r)
        if self.has_hi:
            compute(self.hi_ext, output[nz - r:], r)
        return output

    # -- filters ---------------------------------------------------------------------
    def _check_radius(self, specs):
        for s in specs:
            if s.axis == 0 and s.radius() > self.r:
                raise ValueError("filter radius %d along z exceeds the plan's halo %d" % (s.radius(), self.r))

    def gaussian_filter(self, x, sigma, order=0, output=None, truncate=4.0, dtype_mode=None, compute=None):
        """Sharded ``gaussian_filter``: ``x`` is this rank's z-slab; returns this rank's slab of the result."""
        if output is None:
            output = torch.empty_like(x)
        if compute is None:
            specs = _filters._gaussian_specs(_array.ingest(x), sigma, order, self.mode, truncate)
            self._check_radius(specs)
            compute = _cuda_compute(specs, self.cval, dtype_mode)
        return self.run(x, output, compute)

    def separable(self, x, specs, output=None, dtype_mode=None, compute=None):
        """Sharded run of any list of 1-D passes (``filters._PassSpec``, increasing axis order)."""
        if output is None:
            output = torch.empty_like(x)
        if compute is None:
            self._check_radius(specs)
            compute = _cuda_compute(list(specs), self.cval, dtype_mode)
        return self.run(x, output, compute)

    def sobel(self, x, axis=-1, output=None, dtype_mode=None, compute=None, smooth=(1.0, 2.0, 1.0)):
        """Sharded ``sobel`` (``smooth=(1, 1, 1)``: ``prewitt``), reference filters.py:828-941."""
        axis = axis % 3
        modes = [_filters._check_mode(m) for m in _filters._normalize_sequence(self.mode, 3)]
        specs = [_filters._PassSpec(a, np.array([-1.0, 0.0, 1.0]) if a == axis else np.asarray(smooth, np.float64),
                                    0, modes[a]) for a in range(3)]
        return self.separable(x, specs, output, dtype_mode, compute)

    def gaussian_gradient_magnitude(self, x, sigma, output=None, truncate=4.0, dtype_mode=None, compute=None):
        """Sharded ``gaussian_gradient_magnitude`` (reference filters.py:1207-1252): the three
        derivative filters share ONE halo exchange of raw planes; square / sum / sqrt stay local."""
        if output is None:
            output = torch.empty_like(x)
        if compute is None:
            if not isinstance(self.mode, str):
                raise ValueError("sharded gradient magnitude takes a single boundary mode (a sequence means one "
                                 "mode per DERIVATIVE axis, filters.py:1175-1201)")
            inp = _array.ingest(x)
            smooth = _filters._gaussian_specs(inp, sigma, 0, self.mode, truncate)
            deriv = _filters._gaussian_specs(inp, sigma, 1, self.mode, truncate)
            if len(smooth) != 3 or len(deriv) != 3:
                raise ValueError("sharded gradient magnitude needs sigma > 0 on every axis")
            self._check_radius(deriv)
            cval = self.cval

            def compute(src, dst, in_offset0):
                _filters._gradient_magnitude_window(_array.ingest(src), _array.ingest(dst), smooth, deriv,
                                                    cval, dtype_mode, in_offset0)
        return self.run(x, output, compute)

    def uniform_filter(self, x, size=3, origin=0, output=None, dtype_mode=None, compute=None):
        """Sharded ``uniform_filter``."""
        if output is None:
            output = torch.empty_like(x)
        if compute is None:
            sizes = _filters._normalize_sequence(size, 3)
            origins = _filters._normalize_sequence(origin, 3)
            modes = _filters._normalize_sequence(self.mode, 3)
            specs = [_filters._PassSpec(a, None, _filters._check_origin(o, int(s)), _filters._check_mode(m),
                                        uniform=True, size=int(s))
                     for a, (s, o, m) in enumerate(zip(sizes, origins, modes)) if s > 1]
            self._check_radius(specs)
            compute = _cuda_compute(specs, self.cval, dtype_mode)
        return self.run(x, output, compute)
