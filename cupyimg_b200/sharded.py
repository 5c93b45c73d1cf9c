"""Multi-GPU sharding of the separable filters: one process per GPU (torch.distributed).

* 3-D volumes shard into contiguous z-slabs (slowest axis).  A separable filter of radius
  r along z needs r planes of RAW input from each z-neighbour, once, before compute — the
  x / y passes are local.  :class:`ZSlabFilter` posts the halo planes as NCCL send/recv
  pairs on a side stream, launches the interior planes (which need no halo) on the compute
  stream at once, and launches the two r-plane boundary strips when the halos have landed,
  so the exchange hides behind the interior compute.  Output stays sharded.
* Float32 volumes whose whole filter runs in one fused launch take the peer-memory path instead
  (``backend="p2p"``, the default where it applies): the slab lives in symmetric memory
  (``torch.distributed._symmetric_memory``: CUDA VMM allocations mapped into every rank of the node), and the
  fused kernel loads the r neighbour planes it needs with TMA straight from the neighbour's slab over
  NVLink.  No halo buffer, no copy and no communication kernel: the one-wave grid keeps every SM.
  Readiness travels as 32-bit flags written by stream memory operations (``sepfilt_stream_write32``):
  "my slab is complete" before the launch, "I have read your planes" after it.
* Batched 2-D stacks (filtered axes never split) shard over the batch axis with no
  communication: :func:`batch_range`.

The reference is single-GPU (no collectives anywhere); this is the
B200 scale-out of its per-axis loops (filters.py:651-662, :777-789).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _array, _ffi
from .scipy.ndimage import filters as _filters

# flag slots of the peer-memory protocol (32-bit words in each rank's symmetric flag array)
_LO_READY, _HI_READY, _LO_DONE, _HI_DONE, _LO_PAD, _HI_PAD, _CTA_COUNTER = 0, 1, 2, 3, 4, 5, 8


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
                 cval=0.0, backend="auto"):
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
        self.comm_stream2 = torch.cuda.Stream(self.device) if self.on_cuda else None
        # ---- peer-memory backend: slab + flags in symmetric memory ----
        if backend not in ("auto", "p2p", "nccl"):
            raise ValueError("backend must be 'auto', 'p2p' or 'nccl'")
        self.slab = None
        self.epoch = 0
        self.p2p = False
        self.last_backend = "none"        # how the last run got its halos (reported by bench.py)
        if backend != "nccl" and self.world > 1 and self.on_cuda and dtype == torch.float32 and r > 0:
            try:
                self._init_p2p()
                self.p2p = True
            except Exception:
                if backend == "p2p":
                    raise
        elif backend == "p2p" and self.world > 1:
            raise ValueError("the peer-memory backend needs CUDA float32 slabs and a z radius > 0")

    # -- peer-memory backend ---------------------------------------------------------
    def _init_p2p(self):
        import torch.distributed._symmetric_memory as symm
        group = self.group if self.group is not None else dist.group.WORLD
        self.slab = symm.empty((self.nz, self.ny, self.nx), dtype=torch.float32, device=self.device)
        self._flags = symm.empty(16, dtype=torch.int32, device=self.device)
        self._slab_hdl = symm.rendezvous(self.slab, group)
        self._flags_hdl = symm.rendezvous(self._flags, group)
        self._flags.zero_()
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                       # every rank's flags are zero before anyone writes one
        self._slab_ptrs = [int(p) for p in self._slab_hdl.buffer_ptrs]
        self._flag_ptrs = [int(p) for p in self._flags_hdl.buffer_ptrs]
        self._plane_bytes = self.ny * self.nx * 4
        # "pull": the copy engines fetch the neighbours' r planes into local pads first, each plane
        # crossing NVLink once; "direct": the kernel's TMA boxes read them in place (every tile re-reads its
        # x / y halo over NVLink, which the local L2 does not cache: measured slower on 512^2 planes)
        self.p2p_mode = "auto"
        self._view_cache = {}
        r = self.r
        shape = (self.nz, self.ny, self.nx)
        lower, upper = self._peer(-1), self._peer(+1)
        self._peer_lo = self._slab_hdl.get_buffer(lower, shape, torch.float32)[self.nz - r:] if self.has_lo else None
        self._peer_hi = self._slab_hdl.get_buffer(upper, shape, torch.float32)[:r] if self.has_hi else None
        self._pad_lo = torch.empty((r, self.ny, self.nx), dtype=torch.float32, device=self.device) if self.has_lo else None
        self._pad_hi = torch.empty((r, self.ny, self.nx), dtype=torch.float32, device=self.device) if self.has_hi else None

    def _write_flag(self, rank, slot, value, stream):
        _ffi.check(_ffi.lib().sepfilt_stream_write32(stream, self._flag_ptrs[rank] + 4 * slot, value))

    def begin_fill(self):
        """Hold the current stream until both neighbours have finished reading this rank's slab in the
        previous step.  Call before overwriting ``self.slab`` (``run`` does when it copies the input in)."""
        if not self.p2p or self.epoch == 0:
            return
        stream = _array.current_stream(self.device)
        mine = self._flag_ptrs[self.rank]
        if self.has_lo:
            _ffi.check(_ffi.lib().sepfilt_stream_wait32_geq(stream, mine + 4 * _LO_DONE, self.epoch))
        if self.has_hi:
            _ffi.check(_ffi.lib().sepfilt_stream_wait32_geq(stream, mine + 4 * _HI_DONE, self.epoch))

    def _run_p2p(self, x, output, specs, dspecs, cval):
        """One fused launch over the whole slab; the r planes beyond each end are read from the neighbours'
        slabs.  Returns False (nothing enqueued) when the library has no fused kernel for the request."""
        inp, out = _array.ingest(self.slab), _array.ingest(output)
        if not _filters._fused_candidate(inp, out, specs, False, gradmag=dspecs is not None):
            return False
        structs = [s.struct() for s in specs]
        arr = (_ffi.Pass * len(structs))(*[s[0] for s in structs])
        darr = None
        if dspecs is not None:
            dstructs = [s.struct() for s in dspecs]
            darr = (_ffi.Pass * len(dstructs))(*[s[0] for s in dstructs])
        L = _ffi.lib()
        if not L.sepfilt_separable_f32_supported(inp.tensor(), out.tensor(), arr, len(structs),
                                                 1 if dspecs is not None else 0, float(cval)):
            return False
        if x.data_ptr() != self.slab.data_ptr():
            self.begin_fill()
            self.slab.copy_(x, non_blocking=True)
        stream = _array.current_stream(self.device)
        lower, upper = self._peer(-1), self._peer(+1)
        epoch = self.epoch + 1
        self.epoch = epoch
        r, nz = self.r, self.nz
        mine = self._flag_ptrs[self.rank]
        # the warp-specialised kernel (gradient magnitude) waits for a neighbour's flag only when its march reaches
        # that neighbour's planes and spreads the reads over its z segments: in place is faster there
        # plain filters of radius > 8 (8-row tiles that re-read a tall box, a long pipeline fill): two boundary strips cost
        # 40 % of a 256-plane slab there, so the pads are pulled as above but ONE launch covers the whole slab and waits
        # for each pad's flag only when its march reaches that pad ("pull1")
        wide = dspecs is None and max(sp.radius() for sp in specs) > 8
        mode = self.p2p_mode if self.p2p_mode != "auto" else ("direct" if dspecs is not None else ("pull1" if wide else "pull"))
        if dspecs is None and os.environ.get("SEPFILT_P2P_PLAIN"):      # A/B aid: pull / pull1 / direct for the plain filters
            mode = os.environ["SEPFILT_P2P_PLAIN"]
        grad = 1 if dspecs is not None else 0
        # my slab is complete at this point of the stream: tell the ranks that read it (one batched memop)
        _ffi.check(L.sepfilt_stream_write32x2(
            stream, self._flag_ptrs[lower] + 4 * _HI_READY if self.has_lo else None,
            self._flag_ptrs[upper] + 4 * _LO_READY if self.has_hi else None, epoch))

        def halo_launch(src, dst, halo, in_offset0):
            rc = L.sepfilt_separable_f32_halo(_array.ingest(src).tensor(), _array.ingest(dst).tensor(), arr, len(structs),
                                              darr, grad, ctypes.byref(halo), int(in_offset0), float(cval), stream)
            if rc == _ffi.ERR_UNSUPPORTED:
                return False
            _ffi.check(rc)
            _ffi.count_launch(1)
            return True

        if mode == "direct":
            # ONE launch over the whole slab: TMA reads the neighbours' planes in place; the kernel's last CTA
            # sets the neighbours' done flags (no stream operation after the launch)
            halo = _ffi.Halo()
            halo.epoch = epoch
            halo.cta_counter = mine + 4 * _CTA_COUNTER
            if self.has_lo:
                halo.lo, halo.planes_lo = self._slab_ptrs[lower] + (nz - r) * self._plane_bytes, r
                halo.ready_lo = mine + 4 * _LO_READY
                halo.done_lo = self._flag_ptrs[lower] + 4 * _HI_DONE       # I am the lower rank's UPPER neighbour
            if self.has_hi:
                halo.hi, halo.planes_hi = self._slab_ptrs[upper], r
                halo.ready_hi = mine + 4 * _HI_READY
                halo.done_hi = self._flag_ptrs[upper] + 4 * _LO_DONE
            ok = halo_launch(self.slab, output, halo, 0)
            if ok:
                self.last_backend = "peer memory: the fused kernel's TMA reads the neighbour planes in place over NVLink"
            else:                                   # nothing was launched: release the neighbours by hand
                if self.has_lo:
                    self._write_flag(lower, _HI_DONE, epoch, stream)
                if self.has_hi:
                    self._write_flag(upper, _LO_DONE, epoch, stream)
            return ok

        # "pull": copy engines fetch the neighbours' r planes into local pads on side streams (a stream memory
        # operation holds each copy until the neighbour's slab is complete — no SM involved), while the INTERIOR
        # planes [r, nz - r), which need no halo, are filtered at once; the two r-plane boundary strips follow,
        # each from the slab's first / last 2r planes plus one pad.  The whole exchange hides behind the interior.
        main = torch.cuda.current_stream(self.device)
        sides = ((self.has_lo, self.comm_stream, _LO_READY, self._pad_lo, self._peer_lo, lower, _HI_DONE),
                 (self.has_hi, self.comm_stream2, _HI_READY, self._pad_hi, self._peer_hi, upper, _LO_DONE))
        for present, side, slot, pad, peer_planes, peer, done_slot in sides:
            if not present:
                continue
            side.wait_stream(main)                                   # the previous strip launch has finished reading the pad
            with torch.cuda.stream(side):
                cs = _array.current_stream(self.device)
                _ffi.check(L.sepfilt_stream_wait32_geq(cs, mine + 4 * slot, epoch))
                pad.copy_(peer_planes, non_blocking=True)
                if mode == "pull1":                                  # the pad is filled: the kernel may read it
                    _ffi.check(L.sepfilt_stream_write32(cs, mine + 4 * (_LO_PAD if slot == _LO_READY else _HI_PAD), epoch))
                # the neighbour's planes have been read: it may overwrite its slab
                _ffi.check(L.sepfilt_stream_write32(cs, self._flag_ptrs[peer] + 4 * done_slot, epoch))
        if mode == "pull1":
            halo = _ffi.Halo()
            halo.epoch = epoch
            if self.has_lo:
                halo.lo, halo.planes_lo, halo.ready_lo = self._pad_lo.data_ptr(), r, mine + 4 * _LO_PAD
            if self.has_hi:
                halo.hi, halo.planes_hi, halo.ready_hi = self._pad_hi.data_ptr(), r, mine + 4 * _HI_PAD
            ok = halo_launch(self.slab, output, halo, 0)
            if ok:                                   # (the next step's pulls wait for this launch: side.wait_stream(main) above)
                self.last_backend = ("peer memory: copy engines pull the neighbour planes over NVLink into local pads; ONE launch "
                                     "reads them when its march gets there (flag written behind each copy)")
            return ok
        z0 = r if self.has_lo else 0
        z1 = nz - r if self.has_hi else nz
        # the views of a step are the same every step: slice and ingest them once per output buffer (host time per step
        # matters here — eight ranks enqueue ~10 stream operations per 0.3 ms step)
        key = (output.data_ptr(), z0, z1)
        views = self._view_cache.get(key)
        if views is None:
            if len(self._view_cache) >= 2:                 # (an entry keeps its output buffer alive)
                self._view_cache.clear()
            views = self._view_cache[key] = (
                _array.ingest(self.slab), _array.ingest(output[z0:z1]) if z1 > z0 else None,
                self.slab[:2 * r], output[:r], self.slab[nz - 2 * r:], output[nz - r:], output)
        ok = True
        if z1 > z0:
            ok = _filters._try_fused(views[0], views[1], specs, cval, dspecs=dspecs, in_offset0=z0)
        if ok and self.has_lo:
            main.wait_stream(self.comm_stream)
            halo = _ffi.Halo()
            halo.lo, halo.planes_lo = self._pad_lo.data_ptr(), r
            ok = halo_launch(views[2], views[3], halo, 0)
        if ok and self.has_hi:
            main.wait_stream(self.comm_stream2)
            halo = _ffi.Halo()
            halo.hi, halo.planes_hi = self._pad_hi.data_ptr(), r
            ok = halo_launch(views[4], views[5], halo, r)
        if ok:
            self.last_backend = ("peer memory: copy engines pull the neighbour planes over NVLink behind the interior launch, "
                                 "boundary strips from slab + pad")
        return ok

    def _try_p2p(self, x, output, specs, dspecs, dtype_mode):
        if not self.p2p or dtype_mode == "ndimage" or self.world == 1:
            return False
        if tuple(x.shape) != (self.nz, self.ny, self.nx) or tuple(output.shape) != tuple(x.shape):
            raise _array.OutputShapeError("slab shape does not match the plan")
        if x.dtype != torch.float32 or output.dtype != torch.float32:
            return False
        if not any(s.axis == 0 for s in specs):
            return False                          # no z pass: nothing to exchange, the plain call serves it
        return self._run_p2p(x, output, specs, dspecs, self.cval)

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
        self.last_backend = "nccl send/recv of raw halo planes overlapped with the interior launches"
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
            if self._try_p2p(x, output, specs, None, dtype_mode):
                return output
            compute = _cuda_compute(specs, self.cval, dtype_mode)
        return self.run(x, output, compute)

    def separable(self, x, specs, output=None, dtype_mode=None, compute=None):
        """Sharded run of any list of 1-D passes (``filters._PassSpec``, increasing axis order)."""
        if output is None:
            output = torch.empty_like(x)
        if compute is None:
            self._check_radius(specs)
            if self._try_p2p(x, output, list(specs), None, dtype_mode):
                return output
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
            if self._try_p2p(x, output, smooth, deriv, dtype_mode):
                return output

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
            if self._try_p2p(x, output, specs, None, dtype_mode):
                return output
            compute = _cuda_compute(specs, self.cval, dtype_mode)
        return self.run(x, output, compute)
