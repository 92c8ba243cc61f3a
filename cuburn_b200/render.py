"""
Render orchestration: device buffers, the per-genome compiled module, and the
per-frame launch sequence  interpolate -> iterate -> filter chain -> convert ->
copy to host.

Keeps the reference surface (cuburn/render.py): ``Framebuffers`` (with
``calc_dim / set_dim / alloc / free / flip / flip_side`` and ``gutter == 12``),
``DevSrc``, ``DevInfo``, ``Renderer(gnm, gprof, keep=False, arch=None)`` with
``.packer .lib .cubin .mod .filts .out``, ``RenderManager.queue_frame(rdr, gnm,
gprof, tc, copy=True) -> (evt, h_out)`` where ``evt.query() / synchronize() /
time()``.  Everything below the Python surface goes through the C ABI
(_native.py); nothing here computes pixels on the CPU.
"""
from collections import namedtuple

import numpy as np

from . import _native as N
from . import filters, output, mwc
from .code import itergen, packer as packer_mod
from .genome.util import palette_decode

RenderedImage = namedtuple('RenderedImage', 'buf idx gpu_time')
Dimensions = N.Dims

# Samples one CTA processes per work unit (256 threads x 128 rounds); sample
# ranges handed to cb_iterate are aligned to this.
UNIT_SAMPLES = 32768
ITER_THREADS = 256
# layout of RenderManager.d_hot (cb_hot_scan): scratch, tags, counters
HOT_TAGS_OFF = 8 * 1024 * 8
HOT_COUNT_OFF = HOT_TAGS_OFF + 1028 * 4
HOT_BYTES = HOT_COUNT_OFF + 16


class DurationEvent(N.Event):
    """An event that remembers a prior event to measure from (render.py:26-38)."""
    def __init__(self, prior):
        super().__init__()
        self._prior = prior

    def time(self):
        return self.time_since(self._prior)


class Framebuffers(object):
    """
    The large device allocations (render.py:40-170).  ``d_front`` / ``d_back``
    / ``d_left`` / ``d_right`` hold one float4 per accumulation bin.  A filter
    may use any of them as long as its result ends up in ``d_front``;
    ``Output.convert`` writes ``d_back``.

    The chaos game accumulates straight into ``d_front`` (float4 atomics), so
    the reference's packed integer side buffers and hotspot flag planes
    (``d_uleft`` / ``d_uright``) have no job here; the attributes exist for
    API compatibility and stay ``None``.
    """
    gutter = 12

    # RNG streams / trajectories available to any kernel.  The multiplier
    # table has 262144 entries (= 1024 ring slots x 256 in the reference,
    # render.py:100-104).
    nstreams = 262144

    @classmethod
    def calc_dim(cls, width, height):
        return N.calc_dim(width, height)

    def __init__(self, seed=None):
        N.ensure_init()
        self.stream = N.Stream()
        self.pool = N.PinnedPool()
        self._clear()
        seeds = mwc.make_seeds(self.nstreams, host_seed=seed)
        self.d_seeds = N.to_device(seeds)
        # two trajectories per RNG stream (the iterate kernel's POINTS)
        self._len_d_points = 2 * self.nstreams * 16
        self.d_points = N.DeviceBuffer(self._len_d_points)
        N.fill32(self.d_points, self._len_d_points // 4, np.float32(np.nan))
        N.check(N.lib().cb_device_sync())

    def reseed(self, seed):
        """Reset every RNG stream as ``make_seeds(nstreams, host_seed=seed)``."""
        N.memcpy_htod(self.d_seeds, mwc.make_seeds(self.nstreams, host_seed=seed))

    def _clear(self):
        self.nbins = None
        self.d_front = self.d_back = self.d_left = self.d_right = None
        self.d_uleft = self.d_uright = None

    def free(self, stream=None):
        if stream is not None:
            stream.synchronize()
        else:
            N.check(N.lib().cb_device_sync())
        for p in (self.d_front, self.d_back, self.d_left, self.d_right):
            if p is not None:
                p.free()
        self._clear()

    def alloc(self, dim, stream=None):
        nbins = dim.ah * dim.astride
        if self.nbins is not None and self.nbins >= nbins:
            return
        if self.nbins is not None:
            self.free(stream)
        try:
            self.d_front = N.DeviceBuffer(16 * nbins)
            self.d_back = N.DeviceBuffer(16 * nbins)
            self.d_left = N.DeviceBuffer(16 * nbins)
            self.d_right = N.DeviceBuffer(16 * nbins)
            self.nbins = nbins
        except MemoryError:
            self.free(stream)
            raise

    def set_dim(self, width, height, stream=None):
        dim = self.calc_dim(width, height)
        self.alloc(dim, stream)
        return dim

    def flip(self):
        self.d_front, self.d_back = self.d_back, self.d_front

    def flip_side(self):
        self.d_left, self.d_right = self.d_right, self.d_left
        self.d_uleft, self.d_uright = self.d_uright, self.d_uleft


class FramebufferBand(object):
    """
    Rows [row0, row1) of a ``Framebuffers`` seen as a framebuffer of their own: the
    same four float4 planes, offset, with ``dim`` narrowed to the band.  Filters run
    on it unchanged (their stencils clamp at the band's edges exactly as they do at
    the frame's), and ``flip`` / ``flip_side`` act on the parent, so the parent ends
    with its planes in the roles the chain left them in.
    """
    def __init__(self, fb, dim, row0, row1):
        assert 0 <= row0 < row1 <= dim.ah and row0 % 16 == 0 and (row1 - row0) % 16 == 0
        self.fb, self.row0, self.row1 = fb, row0, row1
        self.dim = N.Dims(dim.w, dim.h, dim.aw, row1 - row0, dim.astride)
        self._off, self._len = 16 * row0 * dim.astride, 16 * (row1 - row0) * dim.astride
        self.gutter, self.pool = fb.gutter, fb.pool

    def _plane(self, name):
        return N.DeviceSlice(getattr(self.fb, name), self._off, self._len)

    d_front = property(lambda self: self._plane('d_front'))
    d_back = property(lambda self: self._plane('d_back'))
    d_left = property(lambda self: self._plane('d_left'))
    d_right = property(lambda self: self._plane('d_right'))

    def flip(self):
        self.fb.flip()

    def flip_side(self):
        self.fb.flip_side()


class DevSrc(object):
    """Device copies of the genome's interpolation sources (render.py:172-190)."""
    max_knots = packer_mod.MAX_KNOTS
    max_params = 1024

    def __init__(self):
        self.d_times = N.DeviceBuffer(4 * self.max_knots * self.max_params)
        self.d_knots = N.DeviceBuffer(4 * self.max_knots * self.max_params)
        self.d_ptimes = N.DeviceBuffer(4 * self.max_knots)
        self.d_pals = N.DeviceBuffer(4 * 4 * 256 * self.max_knots)
        self.d_row_mag = N.DeviceBuffer(4 * self.max_params)
        self.d_program = N.DeviceBuffer(4 * packer_mod.PROG_WIDTH * 2 * self.max_params)


class DevInfo(object):
    """Per-frame temporal samples on the device (render.py:192-223)."""
    palette_width = 256
    palette_height = 64
    ntemporal_samples = 1024
    # unrecorded settling rounds for freshly seeded trajectories
    fuse = 32

    def __init__(self):
        nts = self.ntemporal_samples
        self.d_params = N.DeviceBuffer(nts * DevSrc.max_params * 4)
        self.d_vals = N.DeviceBuffer(nts * DevSrc.max_params * 4)
        self.d_palette = N.DeviceBuffer(self.palette_height * self.palette_width * 16)
        self.d_palette_packed = N.DeviceBuffer(self.palette_height * self.palette_width * 8)


class Renderer(object):
    """
    A genome structure compiled for the device, plus its filter chain and
    output module (render.py:225-251).  Two variants of the iterate module
    exist per structure: parameters in __constant__ memory (stills: one block
    per launch, every parameter a constant-bank operand) and parameters staged
    in shared memory per temporal sample (motion blur).  Each is compiled on
    first use and cached by generated source, so genomes that share a structure
    share modules.
    """
    MAX_MODREFS = 40
    _modrefs = {}
    # trajectories per thread of the production modules: None = itergen.best_points
    # (two for the motion-blur variant of mid-sized genomes, else one)
    points = None

    @classmethod
    def _module(cls, src):
        mod = cls._modrefs.get(src)
        if mod is None:
            names, hdrs = itergen.load_headers()
            mod = N.Module(src, 'iter.cu', hdrs, names, itergen.NVRTC_OPTIONS)
            if len(cls._modrefs) > cls.MAX_MODREFS:
                cls._modrefs.clear()
            cls._modrefs[src] = mod
        return mod

    @classmethod
    def _points(cls, pk, params_const):
        return cls.points if cls.points is not None else itergen.best_points(pk, params_const)

    @classmethod
    def compile(cls, gnm, arch=None, keep=False, params_const=False, acc_packed=False,
                hot_bins=False, big_grid=False):
        pk = itergen.GenomePacker(gnm)
        src = itergen.generate_source(pk, params_const, acc_packed=acc_packed, hot_bins=hot_bins,
                                      points=cls._points(pk, params_const),
                                      extra_defines={'RED_POLICY': '1'} if big_grid else None)
        mod = cls._module(src)
        if keep:
            import os, tempfile
            base = os.path.join(tempfile.gettempdir(), 'iter_kern')
            with open(base + '.cu', 'w') as fp:
                fp.write(src)
            with open(base + '.cubin', 'wb') as fp:
                fp.write(mod.cubin)
        return pk, src, mod

    def __init__(self, gnm, gprof, keep=False, arch=None):
        self._gnm_structure, self._keep = gnm, keep
        self.packer, self.lib, self.mod = self.compile(gnm, arch=arch, keep=keep)
        self._variants = {(False, False, False): self.mod}
        # hot-bin state of this genome: None = not yet probed, else whether the last
        # probe found bins hot enough to privatise (RenderManager._iter)
        self.hot = None
        self.filts = filters.create(gprof)
        self.out = output.get_output_for_profile(gprof)
        self._grid = {}

    @property
    def cubin(self):
        return self.mod.cubin

    def variant(self, params_const, acc_packed=False, hot_bins=False, big_grid=False):
        """The iterate module for (parameters in __constant__ memory?, packed u64
        accumulation?, private shared-memory cells for hot bins?, reductions with an L2
        evict-last priority -- for grids that are a large part of L2 or exceed it?),
        compiled on first use."""
        key = (bool(params_const), bool(acc_packed), bool(hot_bins))
        if big_grid:
            key += (True,)
        if key not in self._variants:
            self._variants[key] = self.compile(self._gnm_structure, params_const=key[0],
                                               acc_packed=key[1], hot_bins=key[2],
                                               big_grid=bool(big_grid))[2]
        return self._variants[key]

    @property
    def mod_const(self):
        """The still variant (parameters in __constant__ memory)."""
        return self.variant(True)

    def grid_ctas(self, nstreams, mod=None):
        """Persistent grid: every SM filled to the kernel's occupancy."""
        mod = mod or self.mod
        if id(mod) not in self._grid:
            info = mod.kernel_info('cb_iter', ITER_THREADS)
            sms = N.device_info(N._initialised or 0)['sm_count']
            self.kernel_info = info
            self._grid[id(mod)] = max(1, min(info['ctas_per_sm'] * sms,
                                             nstreams // ITER_THREADS))
        return self._grid[id(mod)]


class RenderManager(object):
    """Queues frames on the device (render.py:253-434)."""
    def __init__(self, seed=None, rank=0, world=1):
        N.ensure_init()
        self.fb = Framebuffers(seed=seed)
        if world > 1:
            # disjoint RNG streams per GPU (multigpu.make_rank_seeds)
            from .multigpu import make_rank_seeds
            N.memcpy_htod(self.fb.d_seeds, make_rank_seeds(rank, world,
                                                           1 if seed is None else seed,
                                                           self.fb.nstreams))
        self.src_a, self.src_b = DevSrc(), DevSrc()
        self.info_a, self.info_b = DevInfo(), DevInfo()
        self.stream_a, self.stream_b = N.Stream(), N.Stream()
        self.filt_evt = self.copy_evt = None
        import collections
        self._pinned = collections.deque(maxlen=4)
        # hot-bin table: u64 scratch[8][1024] | int32 tags[1024 + 1, padded] | int32 count[4]
        self.d_hot = N.DeviceBuffer(HOT_BYTES)
        N.fill32(self.d_hot, HOT_BYTES // 4, 0)
        self._hot_probe = None          # (event, pinned count, renderer) of the last scan
        self._hot_counts = self.fb.pool.allocate((8,), 'i4')       # pinned ring of results
        self._hot_seq = 0
        self.d_tickets = N.DeviceBuffer(16)     # unit / sweep counters of a cb_iterate launch
        # share of the frame's samples this manager renders (multi-GPU stills)
        self.sample_share = (rank, world)
        self.hist_hook = None

    # -- upload ----------------------------------------------------------------
    def _copy(self, rdr, gnm):
        """H2D: knot rows, palettes and the precalc program (render.py:264-285)."""
        pk = rdr.packer
        if pk.nrows > DevSrc.max_params or pk.nslots > DevSrc.max_params:
            raise ValueError('genome needs %d rows / %d slots; limit is %d'
                             % (pk.nrows, pk.nslots, DevSrc.max_params))
        pool, s, src = self.fb.pool, self.stream_a, self.src_a
        palsrc = dict((v[0], palette_decode(v[1:])) for v in gnm['palette'])
        if len(palsrc) > DevSrc.max_knots:
            raise ValueError('too many palettes')
        ptimes, pvals = zip(*sorted(palsrc.items()))
        program = pk.program_array()
        # one pinned staging block per frame, carved into the six arrays (64-byte aligned)
        times, knots, palettes, palette_times, mag, prog = _carve(pool, (
            ((pk.nrows, pk_width()), 'f4'), ((pk.nrows, pk_width()), 'f4'),
            ((len(palsrc), 256, 4), 'f4'), ((DevSrc.max_knots,), 'f4'),
            ((pk.nrows,), 'i4'), (program.shape, 'i4')))
        pk.pack(gnm, times, knots)
        N.memcpy_htod(src.d_times, times, s)
        N.memcpy_htod(src.d_knots, knots, s)

        palettes[:] = pvals
        palette_times.fill(1e9)
        palette_times[:len(ptimes)] = ptimes
        N.memcpy_htod(src.d_pals, palettes, s)
        N.memcpy_htod(src.d_ptimes, palette_times, s)

        mag[:] = pk.row_mag
        prog[:] = program
        N.memcpy_htod(src.d_row_mag, mag, s)
        N.memcpy_htod(src.d_program, prog, s)
        # keep the staging arrays of the last few frames alive: their async H2D copies
        # may still be queued behind earlier frames on the alternating streams
        self._pinned.append((times, knots, palettes, palette_times, mag, prog))

    # -- interpolate -------------------------------------------------------------
    def _interp(self, rdr, gnm, dim, ts, td):
        L, s, info, src, pk = N.lib(), self.stream_a, self.info_a, self.src_a, rdr.packer
        N.check(L.cb_interp_palette(
            info.d_palette.ptr, self.fb.d_seeds.ptr, src.d_ptimes.ptr, src.d_pals.ptr,
            np.float32(ts), np.float32(td / info.palette_height), info.palette_height,
            s.handle))
        N.check(L.cb_palette_pack(info.d_palette_packed.ptr, info.d_palette.ptr,
                                  info.palette_height, s.handle))
        nts = info.ntemporal_samples
        N.check(L.cb_interp_rows(
            info.d_vals.ptr, src.d_times.ptr, src.d_knots.ptr, src.d_row_mag.ptr,
            pk.nrows, np.float32(ts), np.float32(td / nts), nts, s.handle))
        N.check(L.cb_interp_params(
            info.d_params.ptr, pk.param_stride, info.d_vals.ptr, pk.nrows,
            src.d_program.ptr, len(pk.program), N.byref(dim), nts, s.handle))

    # -- iterate -------------------------------------------------------------------
    def frame_samples(self, gprof, dim, tc):
        """Samples for the whole frame (render.py:331), and this manager's share."""
        total = int(gprof.spp(tc) * dim.w * dim.h)
        rank, world = self.sample_share
        units = (total + UNIT_SAMPLES - 1) // UNIT_SAMPLES
        lo = units * rank // world
        hi = units * (rank + 1) // world
        first = lo * UNIT_SAMPLES
        n = min(hi * UNIT_SAMPLES, total) - first
        return total, first, max(n, 0)

    # Accumulate in the slice-balancing layout (see device/iter_kernel.cuh) into the
    # side buffer and unswizzle into d_front afterwards; False = accumulate straight
    # into d_front in linear layout.  'auto': swizzle while the histogram is (about)
    # L2-sized -- measured on B200: 1080p +30 % (sparse flames), 4K +15 %, but 8K
    # (512 MiB, HBM-bound sector read-modify-write) -30 % because scattering
    # neighbouring bins destroys sector locality.
    swizzle = 'auto'

    def _l2(self):
        """Bytes of L2 on this device."""
        if not hasattr(self, '_l2_cached'):
            self._l2_cached = N.device_info(N._initialised or 0)['l2_bytes']
        return self._l2_cached

    def _use_swizzle(self, nbins):
        if self.swizzle == 'auto':
            return 16 * nbins <= 1.5 * self._l2()
        return bool(self.swizzle)

    # How samples are accumulated.  'auto': float4 reductions straight into the
    # histogram while it is (about) L2-sized; beyond 1.5 x L2 the reference's packed
    # u64 cells (half the bytes per bin, so twice the bins stay L2-resident) with an
    # in-kernel drain and a flush pass: measured +44 % (G6F) / +33 % (G24H) at 8K,
    # -15 % at 4K, -25 % at 1080p.  'float4' / 'packed' force one.
    accumulate = 'auto'

    def _use_packed(self, nbins):
        if self.accumulate == 'auto':
            return 16 * nbins > 1.5 * self._l2()
        return self.accumulate == 'packed'

    # A bin that collects more than ~0.4 % of the samples is bound by the rate of one
    # histogram address (~6.5e8 reductions/s).  The frame's first ``1/hot_pilot`` of
    # samples is rendered as a pilot; if cb_hot_scan finds a bin above ``hot_trigger``,
    # the rest of the frame runs the HOT_BINS variant, which accumulates every bin above
    # ``hot_share`` in shared memory (it costs ~6 % on flames without such bins).  'auto':
    # probe every genome once (one host sync on a renderer's first frame), afterwards
    # follow the previous frame's scan without synchronising.  False: never; True:
    # always run the pilot and the HOT_BINS variant.
    hot_bins = 'auto'
    hot_share = 1.0 / 2048
    hot_trigger = 1.0 / 512
    hot_pilot = 64
    hot_min_waves = 4               # frames shorter than this many waves of units: never
    hot_recheck = 8                 # 'auto', genome without hot bins: pilot every n-th frame

    # Exact sums on the float4 path (device/iter_kernel.cuh, spill_sweep): the kernel adds
    # integer palette levels, and sweeps the grid once per ``spill_interval`` samples,
    # moving every bin that holds >= ``spill_count`` samples into a second grid.  A sum
    # stays exact while it is below 2^24 = 65793 samples of level 255, so a bin would have
    # to collect 65793 - spill_count samples between two sweeps (1/1088 of all samples at
    # the defaults) before a float add rounds; much hotter bins are what ``hot_bins`` is for.
    # (The CTAs of a wave begin their units -- and load their windows -- together, so sweeps
    # more frequent than one per wave of the grid, 1024 x 32768 = 2^25 samples, add nothing.)
    # Grids that do not stay L2-resident are swept less often -- a sweep pulls every sector of
    # the grid through L2, the untouched background included, as scattered DRAM traffic:
    # 0.26 ms per sweep of the 4K grid (measured, tools/stage_parts.py), 10 x its streaming
    # time -- so there the bound on a bin's rounding error only improves by the number of
    # sweeps (8 per 4K / 4000 spp frame, ~1 % of its time).  False: no sweeps (sums round like
    # any float32 running sum, relative error up to n * 2^-25 for a bin of n samples).
    spill = True
    spill_interval = 1 << 26
    spill_count = 4096.0
    spill_max_window = 1024         # SWEEP_MAX_WINDOW of device/iter_kernel.cuh

    def _spill_window(self, nbins, n):
        """Bins every unit examines, for a launch of n samples."""
        if not self.spill or n <= 0:
            return 0
        nunits = (n + UNIT_SAMPLES - 1) // UNIT_SAMPLES
        sweeps = max(1, -(-n // self.spill_interval))
        if 16 * nbins > 0.6 * self._l2():
            sweeps = min(sweeps, max(1, int(0.002 * n / nbins)))
        return int(min(-(-nbins * sweeps // nunits), self.spill_max_window, nbins))

    def _launch_iter(self, mod, rdr, info, d_acc, swz, dim, first, n, total, fuse, packed,
                     hot, s, first_round=0):
        window = 0 if packed else self._spill_window(dim.ah * dim.astride, n)
        args = N.IterArgs(
            spill=self.fb.d_right.ptr if window else 0, spill_bins=window,
            spill_count=self.spill_count, tickets=self.d_tickets.ptr,
            dynamic=int(self.schedule == 'dynamic'),
            hist=int(d_acc), swizzle_bins=swz, seeds=self.fb.d_seeds.ptr,
            points=self.fb.d_points.ptr, params=info.d_params.ptr,
            palette=info.d_palette.ptr, dim=dim, param_stride=rdr.packer.param_stride,
            nts=info.ntemporal_samples, pal_rows=info.palette_height,
            fuse_rounds=fuse, first_sample=first, nsamples=n, total_samples=total,
            cells=self.fb.d_left.ptr if packed else 0,
            palette_packed=info.d_palette_packed.ptr,
            hot_tags=self.d_hot.ptr + HOT_TAGS_OFF if hot else 0, first_round=first_round)
        N.check(N.lib().cb_iterate(mod.handle, N.byref(args),
                                   self.iter_grid or rdr.grid_ctas(self.fb.nstreams, mod),
                                   s.handle))

    # persistent CTAs per cb_iterate launch; None: fill the GPU at the module's occupancy
    iter_grid = None

    # How units of 32768 samples are dealt out to the persistent CTAs.  'dynamic': a CTA
    # claims its next unit when it is ready for it -- the CTAs of an SM run at very
    # different speeds (the warp schedulers favour the oldest warps), and with static
    # ownership the slow ones finish the launch alone.  'static': CTA b runs units
    # b, b + grid, ...: every RNG stream draws the same samples in every run, so a frame
    # is a pure function of its seeds (reproducible frames; what the parity tests use).
    schedule = 'dynamic'

    def _hot_decision(self, rdr, nunits, packed, grid):
        """(run the pilot + scan?, use the HOT_BINS variant?) for this frame."""
        if packed or self.hot_bins is False or nunits < self.hot_min_waves * grid:
            return False, False
        if self.hot_bins is True:
            return True, True
        # 'auto': take in the result of the last scan if it has completed
        if self._hot_probe is not None and self._hot_probe[2] is rdr:
            evt, count, _ = self._hot_probe
            if evt.query():
                rdr.hot = bool(count[0] > 0)
        if rdr.hot is None:
            return True, None               # first frame of this genome: probe and wait
        if not rdr.hot:
            # nothing hot last time: look again every hot_recheck frames only (the pilot
            # splits the frame into two launches, ~0.1 ms per 1080p frame)
            rdr._hot_quiet = getattr(rdr, '_hot_quiet', 0) + 1
            if rdr._hot_quiet % self.hot_recheck:
                return False, False
        return True, rdr.hot

    def _iter(self, rdr, gnm, gprof, dim, tc):
        s, info = self.stream_a, self.info_a
        nbins = dim.ah * dim.astride
        packed = self._use_packed(nbins)
        swz = (nbins // 65536) * 65536 if (not packed and self._use_swizzle(nbins)) else 0
        # float4 path: integer level sums in d_left (+ swept bins in d_right), scaled and
        # brought into linear layout in d_front by cb_hist_finish; packed path: u64 cells
        # in d_left, drained into d_front
        d_acc = self.fb.d_front if packed else self.fb.d_left
        if not packed and self.spill:
            N.fill32(self.fb.d_right, 4 * nbins, 0, s)
        # the grid the samples go to is cleared last: what of it fits stays in L2
        N.fill32(d_acc, 4 * nbins, 0, s)
        if packed:
            N.fill32(self.fb.d_left, 2 * nbins, 0, s)           # u64 cells
        total, first, n = self.frame_samples(gprof, dim, tc)
        # without motion blur all temporal samples are identical: use the variant
        # that reads one parameter block from __constant__ memory
        still = gprof.frame_width(tc) == 0
        nunits = (n + UNIT_SAMPLES - 1) // UNIT_SAMPLES
        # grids from about half of L2 up: reductions carry an evict-last priority (measured:
        # 4K float4 -4 %, 8K packed cells -3 %, 1080p +2 % -- so not there)
        big = (8 if packed else 16) * nbins > 0.5 * self._l2()
        mod = rdr.variant(still, packed, big_grid=big)
        grid = self.iter_grid or rdr.grid_ctas(self.fb.nstreams, mod)
        pilot, hot = self._hot_decision(rdr, nunits, packed, grid)
        if still:
            mod.set_global('c_params', info.d_params.ptr, 4 * rdr.packer.nslots, s)
        fuse, first_round = info.fuse, 0
        n_frame = n
        probe = None
        if pilot:
            # whole waves of the persistent grid, about 1/hot_pilot of the frame
            npilot = grid * max(1, round(nunits / float(self.hot_pilot * grid))) * UNIT_SAMPLES
            self._launch_iter(mod, rdr, info, d_acc, swz, dim, first, npilot, total, fuse,
                              packed, False, s)
            N.check(N.lib().cb_hot_scan(
                self.d_hot.ptr + HOT_TAGS_OFF, self.d_hot.ptr + HOT_COUNT_OFF, self.d_hot.ptr,
                int(d_acc), self.fb.d_right.ptr if self.spill else 0, swz,
                np.float32(max(32.0, self.hot_share * npilot)),
                np.float32(max(32.0, self.hot_trigger * npilot)), N.byref(dim), s.handle))
            self._hot_seq = (self._hot_seq + 1) % 8
            probe = self._hot_counts[self._hot_seq:self._hot_seq + 1]
            if hot is None:
                # first frame of this genome: the variant depends on the answer
                N.memcpy_dtoh(probe, N.DeviceSlice(self.d_hot, HOT_COUNT_OFF, 4), s)
                s.synchronize()
                hot = rdr.hot = bool(probe[0] > 0)
                probe = None
            # the pilot is whole waves: every CTA has run the same number of rounds
            ppt = itergen.points_per_thread(rdr.packer, rdr._points(rdr.packer, still))
            first_round = fuse + (npilot // UNIT_SAMPLES // grid) * \
                (UNIT_SAMPLES // (ITER_THREADS * ppt))
            first, n, fuse = first + npilot, n - npilot, 0
        if hot:
            mod = rdr.variant(still, packed, True, big_grid=big)
            if still:
                mod.set_global('c_params', info.d_params.ptr, 4 * rdr.packer.nslots, s)
        if n > 0:
            self._launch_iter(mod, rdr, info, d_acc, swz, dim, first, n, total, fuse, packed,
                              bool(hot), s, first_round)
        if probe is not None:
            # read the scan's verdict back behind the main launch (nothing between the pilot
            # and the main launch may wait on the host); the next frame picks it up
            N.memcpy_dtoh(probe, N.DeviceSlice(self.d_hot, HOT_COUNT_OFF, 4), s)
            self._hot_probe = (N.Event().record(s), probe, rdr)
        self._raw_levels = False
        if packed:
            N.check(N.lib().cb_flush_packed(self.fb.d_front.ptr, self.fb.d_left.ptr,
                                            N.byref(dim), s.handle))
        else:
            # a multi-GPU hook that wants to sum the integer levels exactly gets them
            # unscaled; _combine divides by 255 after it
            raw = self._raw_levels = bool(getattr(self.hist_hook, 'integer_sums', False))
            N.check(N.lib().cb_hist_finish(
                self.fb.d_front.ptr, int(d_acc), self.fb.d_right.ptr if self.spill else 0, swz,
                np.float32(1.0 if raw else 1.0 / 255.0), N.byref(dim), s.handle))
        self.last_iter_samples = n_frame
        self.last_iter_hot = bool(hot)

    # -- combine (multi-GPU stills) -----------------------------------------------------
    _raw_levels = False

    def _combine(self, dim):
        """Run ``hist_hook`` (the per-GPU histograms become one) and bring the result to
        the (sum / 255, count) form if the iterate stage left integer level sums for it."""
        if self.hist_hook is not None:
            self.hist_hook(self.fb, dim, self.stream_a)
        if self._raw_levels:
            N.check(N.lib().cb_hist_finish(self.fb.d_front.ptr, self.fb.d_front.ptr, 0, 0,
                                           np.float32(1.0 / 255.0), N.byref(dim),
                                           self.stream_a.handle))
            self._raw_levels = False

    # -- filter ----------------------------------------------------------------------
    # multi-GPU stills: a multigpu.BandFilter makes every GPU filter a band of rows
    # (with halos) of the combined histogram and gathers the bands on the root
    band_filter = None

    def _filter(self, rdr, gprof, dim, tc):
        if self.band_filter is not None:
            self.band_filter(self.fb, rdr.filts, gprof, dim, tc, self.stream_a)
            return
        for filt in rdr.filts:
            params = getattr(gprof.filters, filt.name)
            filt.apply(self.fb, gprof, params, dim, tc, self.stream_a)

    # -- frame -----------------------------------------------------------------------
    def queue_frame(self, rdr, gnm, gprof, tc, copy=True, frame_seed=None):
        """
        Queue one frame; returns ``(evt, h_out)`` (render.py:374-434).  ``evt``
        completes when ``h_out`` (pinned host memory in the output module's
        format) is valid.  Not thread-safe.

        ``frame_seed`` (an addition): re-seed every RNG stream for this frame, so a
        frame's sample set does not depend on which frames were rendered before it
        on this GPU -- what a frame-partitioned multi-GPU animation needs to be
        reproducible.
        """
        timing_event = N.Event().record(self.stream_b)
        if frame_seed is not None:
            seeds = self.fb.pool.allocate((self.fb.nstreams, 3), 'u4')
            rank, world = self.sample_share
            if world > 1:
                # a sample-split still: every GPU needs its own streams for the frame
                from .multigpu import make_rank_seeds
                seeds[:] = make_rank_seeds(rank, world, int(frame_seed), self.fb.nstreams)
            else:
                seeds[:] = mwc.make_seeds(self.fb.nstreams, host_seed=int(frame_seed))
            # the seed table is shared with the previous frame (which ran on stream_b and
            # dithers its output from it): order this upload after that frame's conversion
            # -- not after its D2H copy, which only reads the converted frame
            if self.filt_evt:
                self.stream_a.wait_for_event(self.filt_evt)
            N.memcpy_htod(self.fb.d_seeds, seeds, self.stream_a)
            self._pinned.append((seeds,))
        dim = self.fb.set_dim(gprof.width, gprof.height, self.stream_b)

        td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
        ts = tc - 0.5 * td

        if copy:
            self.src_a, self.src_b = self.src_b, self.src_a
            self._copy(rdr, gnm)
        # the previous frame (other stream) owns the framebuffers and the RNG streams
        # until its conversion has run: cb_interp_palette advances the same seed table
        # that frame's cb_iter / cb_convert read and write back
        if self.filt_evt:
            self.stream_a.wait_for_event(self.filt_evt)
        self._interp(rdr, gnm, dim, ts, td)
        self._iter(rdr, gnm, gprof, dim, tc)
        # multi-GPU stills: combine per-GPU histograms before filtering
        self._combine(dim)
        if self.copy_evt:
            self.stream_a.wait_for_event(self.copy_evt)
        self._filter(rdr, gprof, dim, tc)
        rows = out = None
        if self.band_filter is not None and getattr(self.band_filter, 'shared', None) is not None:
            # a still sharded by rows: this GPU converts and copies out its own band
            rows, out = self.band_filter.output_rows(dim, self.fb.gutter), \
                self.band_filter.shared.array
        rdr.out.convert(self.fb, gprof, dim, self.stream_a, rows=rows)
        self.filt_evt = N.Event().record(self.stream_a)
        h_out = rdr.out.copy(self.fb, dim, self.fb.pool, self.stream_a, rows=rows, out=out)
        self.copy_evt = DurationEvent(timing_event).record(self.stream_a)

        self.info_a, self.info_b = self.info_b, self.info_a
        self.stream_a, self.stream_b = self.stream_b, self.stream_a
        return self.copy_evt, h_out


def _carve(pool, specs):
    """One pinned allocation from ``pool`` cut into arrays of the given (shape, dtype),
    each starting on a 64-byte boundary; the block lives as long as any of them."""
    import math
    sizes = [math.prod(shape) * np.dtype(dt).itemsize for shape, dt in specs]
    offsets, total = [], 0
    for n in sizes:
        offsets.append(total)
        total += (n + 63) // 64 * 64
    block = pool.allocate((max(total, 1),), 'u1')
    return [block[off:off + n].view(dt).reshape(shape)
            for off, n, (shape, dt) in zip(offsets, sizes, specs)]


def frame_pipeline(rmgr, rdr, gnm, gprof, times, wait=None):
    """
    Render the frames at ``times`` one frame ahead of the host: frame k + 1 is queued
    before frame k is waited for, so the device never idles while the host encodes or
    writes (the loop of the reference's callers, main.py:62-76 and distribute.py:104-118,
    as one generator).  Yields ``(n, evt, h_out)`` per finished frame in order, ``n``
    counting from 1.  ``wait(evt)`` replaces the blocking ``evt.synchronize()`` (e.g. to
    poll instead).
    """
    ahead = None
    for n, tc in enumerate(list(times) + [None]):
        queued = rmgr.queue_frame(rdr, gnm, gprof, tc) if tc is not None else None
        if ahead is not None:
            evt, h_out = ahead
            if wait is not None:
                wait(evt)
            else:
                evt.synchronize()
            yield n, evt, h_out
        ahead = queued


def pk_width():
    return packer_mod.MAX_KNOTS
