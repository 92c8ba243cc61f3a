"""
Multi-GPU on one box: one process per GPU (torch.distributed, NCCL over NVLink).

The reference's only multi-GPU mechanism is a job farm of OS processes fed over
ssh pipes (distribute.py:131-248).  On an 8 x B200 box the same two ways of
sharding are available in-process:

  * stills: every GPU runs a disjoint share of the frame's samples with its own
    RNG streams into a private float4 histogram; ``HistReducer`` sums them onto
    the root with one NCCL reduce, and the root runs the filter chain.  With
    ``HistReducer(root=None)`` (all-reduce) + ``BandFilter`` the filter chain is
    sharded too: every GPU filters a band of rows plus the halo the chain's
    stencils reach -- bit-identical to filtering the whole frame on one GPU.
    The bands are either gathered on the root (which converts and copies the
    frame out alone) or, with ``BandFilter(shared=SharedFrame(...))``, each GPU
    converts its own band and copies it over its own PCIe link into one
    page-locked shared-memory host frame: no device exchange after the reduce.
    ``HistReducer(integer_sums=True)`` reduces unscaled integer level sums, which
    add up exactly in any order, so every rank holds the same bits.
  * animations: whole frames are independent (points are re-seeded every frame),
    so ``partition_frames`` deals frames round-robin and no collective is needed.

torch is used only for process-group plumbing and the collective.
"""
import os

import numpy as np


def env_rank_world():
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)),
            int(os.environ.get('LOCAL_RANK', os.environ.get('RANK', 0))))


def init_process_group(backend=None):
    """Join the process group described by the torchrun environment."""
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            import torch
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            import torch
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def sample_share(total_samples, rank, world, unit=65536):
    """(first, count): the unit-aligned share of a frame's samples for `rank`."""
    units = (int(total_samples) + unit - 1) // unit
    lo, hi = units * rank // world, units * (rank + 1) // world
    first = lo * unit
    return first, max(min(hi * unit, int(total_samples)) - first, 0)


def partition_frames(frames, rank, world):
    """Round-robin frame assignment: frame k goes to rank k mod world."""
    return [f for k, f in enumerate(frames) if k % world == rank]


def seed_slice(rank, world, nstreams=262144):
    """
    Disjoint RNG streams per rank: the multiplier table is split into `world`
    contiguous slices and each rank repeats its slice to fill its stream table
    with independently drawn state/carry (different seeds per repeat).
    """
    per = nstreams // world
    return rank * per, per


def make_rank_seeds(rank, world, host_seed, nstreams=262144):
    from . import mwc
    mults = mwc.load_mults()
    lo, per = seed_slice(rank, world, nstreams)
    rs = np.random.RandomState((int(host_seed) * 1000003 + rank * 7919 + 1) % (2 ** 32))
    seeds = np.empty((nstreams, 3), np.uint32)
    seeds[:, 0] = np.resize(mults[lo:lo + per], nstreams)
    seeds[:, 1] = rs.randint(1, 0x7fffffff, size=nstreams)
    seeds[:, 2] = rs.randint(1, 0x7fffffff, size=nstreams)
    return seeds


class NativeComm(object):
    """
    The library's own NCCL communicator (cb_comm_*): the exchange steps then run as
    C-ABI calls on the renderer's stream, and torch.distributed is only the courier
    for the 128-byte unique id (any backend; a file or a socket would do as well).
    """
    def __init__(self, rank=None, world=None, exchange_id=None):
        import ctypes
        from . import _native as N
        env_rank, env_world, _ = env_rank_world()
        self.rank = env_rank if rank is None else rank
        self.world = env_world if world is None else world
        N.ensure_init()
        N.preload_nccl()
        ident = (ctypes.c_uint8 * 128)()
        if self.rank == 0:
            N.check(N.lib().cb_comm_unique_id(ident))
        raw = bytes(ident)
        if self.world > 1:
            raw = (exchange_id or self._broadcast)(raw)
        ident = (ctypes.c_uint8 * 128)(*raw)
        handle = ctypes.c_void_p()
        N.check(N.lib().cb_comm_create(ident, self.rank, self.world, ctypes.byref(handle)))
        self.handle, self._N = handle, N

    @staticmethod
    def _broadcast(raw):
        import torch.distributed as dist
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def hist_reduce(self, buf, dim, root, stream):
        """root: a rank, or None for every rank (all-reduce)."""
        N = self._N
        N.check(N.lib().cb_hist_reduce(self.handle, int(buf), N.byref(dim),
                                       -1 if root is None else root, stream.handle))

    def band_gather(self, buf, dim, root, stream):
        N = self._N
        N.check(N.lib().cb_band_gather(self.handle, int(buf), N.byref(dim), root, stream.handle))

    def close(self):
        if self.handle:
            self._N.check(self._N.lib().cb_comm_destroy(self.handle))
            self.handle = None


class HistReducer(object):
    """
    ``RenderManager.hist_hook``: sums the per-GPU float4 histograms onto
    ``root`` before the filter chain runs there; ``root=None`` leaves the sum on
    every GPU (all-reduce), which ``BandFilter`` needs.
    """
    def __init__(self, root=0, comm=None, integer_sums=False):
        """comm: a NativeComm to run the collective through the C ABI; None = torch.

        integer_sums: the iterate stage leaves the integer level sums unscaled
        (``RenderManager._iter`` asks this attribute), they are reduced as they are, and
        the render manager divides by 255 after the collective (``_combine``).  Sums below 2^24 then add
        up exactly in any order, so every GPU holds bit-identical copies of the histogram
        whatever reduction order NCCL chose (with pre-scaled float sums the copies differ
        in the last bit of a few bins: 1 byte of a 33 M byte 4K frame, 8 GPUs)."""
        self.root, self.comm, self.integer_sums = root, comm, bool(integer_sums)
        if comm is None:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
        self._events = []

    def tensor_view(self, buf, nfloats):
        view = buf.view((int(nfloats),), '<f4')
        return self.torch.as_tensor(view, device='cuda')

    def __call__(self, fb, dim, stream):
        if self.comm is not None:
            from . import _native as N
            e0, e1 = N.Event().record(stream), N.Event()
            self.comm.hist_reduce(fb.d_front, dim, self.root, stream)
            e1.record(stream)
            self._events.append((e0, e1))
            return
        if not self.dist.is_initialized() or self.dist.get_world_size() == 1:
            return
        torch = self.torch
        t = self.tensor_view(fb.d_front, 4 * dim.ah * dim.astride)
        # Run the collective in stream order on the renderer's own stream (wrapped
        # as a torch ExternalStream): no host synchronisation on either side.
        ext = torch.cuda.ExternalStream(stream.handle.value)
        with torch.cuda.stream(ext):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            if self.root is None:
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            else:
                self.dist.reduce(t, dst=self.root, op=self.dist.ReduceOp.SUM)
            e1.record()
        self._events.append((e0, e1))

    def mean_reduce_ms(self, last=None):
        """Device time of the recorded reduces (call after a synchronize)."""
        evs = self._events[-last:] if last else self._events
        return float(np.mean([_elapsed(a, b) for a, b in evs])) if evs else None


def _elapsed(e0, e1):
    """ms between two recorded events (torch.cuda.Event or the library's Event)."""
    return e0.elapsed_time(e1) if hasattr(e0, 'elapsed_time') else e1.time_since(e0)


def band_rows(aheight, rank, world):
    """
    (row0, row1): the band of accumulation rows `rank` filters.  Bands are equally
    tall (a multiple of 16 rows, what the filter kernels' grids need) so that one
    gather moves them; when `aheight` is not a multiple of the band height the last
    bands slide up and overlap their neighbour -- overlapping rows are computed
    identically by both owners.
    """
    assert aheight % 16 == 0 and aheight > 0
    band = min(16 * -(-(aheight // 16) // world), aheight)
    row0 = min(rank * band, aheight - band)
    return row0, row0 + band


def chain_reach(filts, gprof, tc):
    """Rows of context the filter chain needs around a band, rounded up to 16."""
    rows = sum(f.reach(gprof, getattr(gprof.filters, f.name, None), tc) for f in filts)
    return 16 * -(-rows // 16)


class BandFilter(object):
    """
    ``RenderManager.band_filter``: the filter chain sharded by rows over the GPUs of
    a still.  Each GPU holds the complete histogram (``HistReducer(root=None)``),
    runs the unchanged chain on ``FramebufferBand`` = its band plus ``chain_reach``
    rows either side, and the bands proper are gathered into the root's ``d_front``.
    Rows inside the halo are wrong near the band's artificial edges and are never
    used; the band itself is bit-identical to the single-GPU result.
    """
    def __init__(self, rank, world, root=0, comm=True, shared=None):
        """comm: True = torch.distributed gather, a NativeComm = cb_band_gather through
        the C ABI, False = no exchange (one process playing a rank, for tests).
        shared: a SharedFrame -- then nothing is gathered on the device: every GPU
        converts its own band (Output.convert(rows=...)) and copies it to its place in
        the shared host frame over its own PCIe link (RenderManager.queue_frame)."""
        self.rank, self.world, self.root, self.comm = rank, world, root, comm
        self.shared = shared
        if shared is not None:
            self.comm = comm = False
        if comm is True:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
        self._events = []

    def filter_band(self, fb, filts, gprof, dim, tc, stream):
        """Run the chain on this rank's band + halo; returns (row0, row1)."""
        from .render import FramebufferBand
        row0, row1 = band_rows(dim.ah, self.rank, self.world)
        halo = chain_reach(filts, gprof, tc)
        band = FramebufferBand(fb, dim, max(row0 - halo, 0), min(row1 + halo, dim.ah))
        for filt in filts:
            filt.apply(band, gprof, getattr(gprof.filters, filt.name), band.dim, tc, stream)
        return row0, row1

    def __call__(self, fb, filts, gprof, dim, tc, stream):
        row0, row1 = self.filter_band(fb, filts, gprof, dim, tc, stream)
        if not self.comm or self.world == 1:
            return
        if self.comm is not True:
            from . import _native as N
            e0, e1 = N.Event().record(stream), N.Event()
            self.comm.band_gather(fb.d_front, dim, self.root, stream)
            e1.record(stream)
            self._events.append((e0, e1))
            return
        torch, dist = self.torch, self.dist
        full = torch.as_tensor(fb.d_front.view((dim.ah, 4 * dim.astride), '<f4'), device='cuda')
        ext = torch.cuda.ExternalStream(stream.handle.value)
        with torch.cuda.stream(ext):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            dest = None
            if self.rank == self.root:
                dest = [full[slice(*band_rows(dim.ah, r, self.world))]
                        for r in range(self.world)]
            dist.gather(full[row0:row1], dest, dst=self.root)
            e1.record()
        self._events.append((e0, e1))

    def mean_gather_ms(self, last=None):
        evs = self._events[-last:] if last else self._events
        return float(np.mean([_elapsed(a, b) for a, b in evs])) if evs else None

    def output_rows(self, dim, gutter=12):
        """Rows of the cropped output frame that this rank's band covers."""
        row0, row1 = band_rows(dim.ah, self.rank, self.world)
        if self.rank == 0:
            row0 = 0
        if self.rank == self.world - 1:
            row1 = dim.ah
        return min(max(row0 - gutter, 0), dim.h), min(max(row1 - gutter, 0), dim.h)


class SharedFrame(object):
    """
    One host frame in POSIX shared memory, mapped and page-locked by every process of the
    job: each GPU copies the rows it produced straight into it.  Rank 0 creates the
    segment; the name is derived from the rendezvous port so no exchange is needed
    beyond a barrier.
    """
    _seq = 0

    def __init__(self, shape, dtype, rank, world, barrier=None):
        import mmap
        import os
        from . import _native as N
        SharedFrame._seq += 1
        self.rank = rank
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.path = '/dev/shm/cuburn_b200_%s_%d' % (os.environ.get('MASTER_PORT', str(os.getppid())),
                                                    SharedFrame._seq)
        if rank == 0:
            fd = os.open(self.path, os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, self.nbytes)
        if barrier is not None:
            barrier()
        if rank != 0:
            fd = os.open(self.path, os.O_RDWR)
        self._mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        self.array = np.frombuffer(self._mm, dtype).reshape(shape)
        N.check(N.lib().cb_host_register(self.array.ctypes.data, self.nbytes))
        if barrier is not None:
            barrier()
        if rank == 0:
            os.unlink(self.path)            # the mappings keep the segment alive

    def close(self):
        """Unpin the frame; the mapping itself goes away with its last numpy view."""
        from . import _native as N
        if self.array is not None:
            N.check(N.lib().cb_host_unregister(self.array.ctypes.data))
            self.array = None
            self._mm = None

    def __del__(self):
        # a mapping that goes away while still registered poisons its address range: the
        # next segment mapped there fails to register ("already mapped")
        try:
            self.close()
        except Exception:
            pass


def gather_host_bands(frame, rank, world, root=0):
    """gloo/CPU twin of BandFilter's gather: `frame` is [aheight, ...]; the root's copy
    ends up with every rank's band rows."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(frame)
    row0, row1 = band_rows(frame.shape[0], rank, world)
    dest = None
    if rank == root:
        dest = [torch.empty_like(t[slice(*band_rows(frame.shape[0], r, world))])
                for r in range(world)]
    dist.gather(t[row0:row1].contiguous(), dest, dst=root)
    if rank == root:
        for r, part in enumerate(dest):
            a, b = band_rows(frame.shape[0], r, world)
            t[a:b] = part
    return t.numpy()


def reduce_host_hist(hist, root=0):
    """gloo/CPU twin of the reduce, used by the world_size-2 CPU tests."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(hist))
    dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
    return t.numpy()
