"""
Host-side orchestration of a frame (cuburn_b200/render.py, the counterpart of
cuburn/render.py:253-434) on the CPU: which C-ABI calls ``RenderManager.queue_frame``
issues, in which order and with which arguments, for the BASELINE configurations --
against a recording stand-in for the library (tests/fake_native.py).  No pixel is computed
here; the same paths run on the device in tests/test_render_gpu.py / test_iter_gpu.py.
"""
import numpy as np
import pytest

import fake_native
from helpers import still_profile

pytestmark = pytest.mark.production_schedule        # the shipped defaults are what is tested


@pytest.fixture
def lib(built, monkeypatch):
    from cuburn_b200 import render
    fake = fake_native.install(monkeypatch)
    monkeypatch.setattr(render.Renderer, '_modrefs', {})
    yield fake
    fake_native.uninstall()


def _frame(lib, rmgr, rdr, gnm, gprof, tc, **kw):
    start = lib.mark()
    evt, h_out = rmgr.queue_frame(rdr, gnm, gprof, tc, **kw)
    return start, evt, h_out


BILATERAL = ['cb_bilateral_direction'] * 8
DEFAULT_CHAIN = (['cb_yuv_to_rgb'] + BILATERAL + ['cb_logscale', 'cb_apply_gamma_full_hi'] +
                 ['cb_full_blur'] * 4 + ['cb_smearclip'])
INTERP = ['cb_interp_palette', 'cb_palette_pack', 'cb_interp_rows', 'cb_interp_params']


def _kernels(lib, start):
    return [n for n in lib.names(start) if n in fake_native.KERNEL_CALLS]


def test_1080p_still_is_41_launches_in_the_documented_order(lib):
    """BASELINE config 2.  First frame of a genome: pilot + hot-bin scan + one host read;
    afterwards interpolation (4), two fills, ONE cb_iterate, cb_hist_finish, the default
    chain (1 + 8 x 3 + 1 + 6) and the conversion: 41 kernel launches."""
    from cuburn_b200 import samples, render
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 2000
    gprof, tc = still_profile(gnm, w, h, spp)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.calc_dim(w, h)
    nbins = dim.ah * dim.astride
    assert nbins == 2155008

    # frame 1: nothing is known about the genome -> pilot, scan, synchronous verdict
    s0, evt, out = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    assert out.shape == (h, w, 4) and out.dtype == np.uint8
    k = _kernels(lib, s0)
    assert k == INTERP + ['cb_fill32', 'cb_fill32', 'cb_iterate', 'cb_hot_scan', 'cb_iterate',
                          'cb_hist_finish'] + DEFAULT_CHAIN + ['cb_convert_rows']
    pilot, main = lib.iterations[-2:]
    n = w * h * spp
    assert pilot['nsamples'] + main['nsamples'] == n == rmgr.last_iter_samples
    assert pilot['first_sample'] == 0 and main['first_sample'] == pilot['nsamples']
    assert pilot['nsamples'] % (main['grid'] * render.UNIT_SAMPLES) == 0      # whole waves
    assert abs(pilot['nsamples'] / n - 1 / 64.) < 0.01
    assert pilot['fuse_rounds'] == 32 and main['fuse_rounds'] == 0 and main['first_round'] > 32
    assert pilot['hot_tags'] == main['hot_tags'] == 0 and rdr.hot is False
    assert 'cb_stream_sync' in lib.names(s0)

    # frame 2: the documented steady state
    s1, evt, out = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    k = _kernels(lib, s1)
    assert k == INTERP + ['cb_fill32', 'cb_fill32', 'cb_iterate', 'cb_hist_finish'] + \
        DEFAULT_CHAIN + ['cb_convert_rows']
    assert lib.kernel_launches(s1) == 41
    assert 'cb_stream_sync' not in lib.names(s1) and 'cb_device_sync' not in lib.names(s1)
    it = lib.iterations[-1]
    assert it['nsamples'] == it['total_samples'] == n and it['first_sample'] == 0
    assert it['grid'] == 1024                       # min(148 SMs x 8, 262144 streams / 256)
    assert it['dynamic'] == 1 and it['fuse_rounds'] == 32 and it['nts'] == 1024
    assert it['swizzle_bins'] == nbins // 65536 * 65536
    # integer level sums go to d_left, swept bins to d_right, cb_hist_finish -> d_front
    fb = rmgr.fb
    assert it['cells'] == 0 and it['spill_bins'] == 1024 and it['spill_count'] == 4096.0
    fin = lib.args_of('cb_hist_finish', s1)[0]
    assert fin[3] == it['swizzle_bins'] and float(fin[4]) == np.float32(1 / 255.)
    # (the chain flips planes afterwards: compare with the roles at launch time)
    planes = {fb.d_front.ptr, fb.d_back.ptr, fb.d_left.ptr, fb.d_right.ptr}
    assert {it['hist'], it['spill'], fin[0]} <= planes and len({it['hist'], it['spill'], fin[0]}) == 3
    assert (fin[1], fin[2]) == (it['hist'], it['spill'])
    fills = lib.args_of('cb_fill32', s1)
    assert [f[1] for f in fills] == [4 * nbins, 4 * nbins] and fills[1][0] == it['hist']
    # a still reads its one parameter block from __constant__ memory
    assert len(lib.args_of('cb_module_set_global', s1)) == 1
    assert lib.args_of('cb_module_set_global', s1)[0][1] == b'c_params'
    # the frame leaves through one D2H copy of w x h x 4 bytes
    d2h = lib.args_of('cb_memcpy_d2h', s1)
    assert [a[2] for a in d2h] == [w * h * 4]

    # frames 3..9 skip the pilot, the 8th frame after the probe looks again
    pilots = []
    for _ in range(7):
        s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
        pilots.append('cb_hot_scan' in lib.names(s))
    assert pilots == [False] * 6 + [True]


def test_frames_alternate_streams_and_wait_for_the_previous_conversion(lib):
    """render.py:401-434: two streams alternate; frame k + 1 may upload its genome at once
    but interpolates (the seed table is shared) only after frame k's conversion."""
    from cuburn_b200 import samples, render
    gnm = samples.g3()
    gprof, tc = still_profile(gnm, 640, 360, 256)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    streams = []
    for i in range(3):
        s, evt, out = _frame(lib, rmgr, rdr, gnm, gprof, tc)
        names = lib.names(s)
        streams.append(lib.iterations[-1]['stream'])
        if i:
            assert 'cb_stream_wait_event' in names
            assert names.index('cb_stream_wait_event') < names.index('cb_interp_palette')
            assert names.index('cb_memcpy_h2d') < names.index('cb_stream_wait_event')
        else:
            assert 'cb_stream_wait_event' not in names
        assert evt.query() and evt.time() == 41.0      # the stand-in's clock: 1 ms per launch;
        # (a frame this short -- under four waves of units -- never runs the hot-bin pilot)
    assert streams[0] != streams[1] and streams[0] == streams[2]
    # six pinned staging arrays per frame (knots, times, palettes, ...) and the output
    h2d = lib.args_of('cb_memcpy_h2d', s)
    assert len(h2d) == 6
    # ... all of them in page-locked memory (an asynchronous copy from pageable memory would
    # be staged synchronously by the driver), 64-byte aligned
    pinned = [(a, a + n) for a, (n, _) in lib._host.items()]
    for dst, src, nbytes, stream in h2d:
        assert any(lo <= src.value and src.value + nbytes <= hi for lo, hi in pinned)
        assert src.value % 64 == 0 and stream is not None
    # ... and so is the frame handed back
    assert any(lo <= out.ctypes.data < hi for lo, hi in pinned)


def test_copy_false_reuses_the_uploaded_genome(lib):
    """queue_frame(copy=False) (render.py:374,415-417): no packing, no H2D, and the
    interpolation reads the device copy the previous call uploaded."""
    from cuburn_b200 import samples, render
    gnm = samples.g3()
    gprof, tc = still_profile(gnm, 640, 360, 256)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    s0, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    src0 = lib.args_of('cb_interp_rows', s0)[0][1:4]
    s1, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc, copy=False)
    assert lib.args_of('cb_memcpy_h2d', s1) == []
    assert lib.args_of('cb_interp_rows', s1)[0][1:4] == src0
    assert lib.kernel_launches(s1) == 41
    s2, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)           # copy=True: the other buffer
    assert len(lib.args_of('cb_memcpy_h2d', s2)) == 6
    assert lib.args_of('cb_interp_rows', s2)[0][1:4] != src0


def test_motion_blur_stages_parameters_per_temporal_sample(lib):
    from cuburn_b200 import samples, render, profile
    gnm = samples.g6f(animated=True)
    gprof = profile.wrap(dict(width=640, height=360, spp=300, fps=24, duration=1.0), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    assert gprof.frame_width(tc) > 0
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    assert lib.args_of('cb_module_set_global', s) == []
    still_mod = rdr.variant(True)
    assert lib.iterations[-1]['module'] is not still_mod.handle
    assert lib.iterations[-1]['module'] is rdr.variant(False).handle
    # the temporal window handed to the interpolation: td = frame_width / (fps * duration)
    rows = lib.args_of('cb_interp_rows', s)[0]
    td = gprof.frame_width(tc) / round(gprof.fps * gprof.duration)
    assert float(rows[5]) == np.float32(tc - 0.5 * td) and float(rows[6]) == np.float32(td / 1024)


def test_8k_accumulates_in_packed_cells_with_evict_last_reductions(lib):
    """BASELINE config 5: the float4 grid (512 MiB) is beyond 1.5 x L2 -> the reference's
    packed u64 cells in d_left, drained into d_front, cb_flush_packed at the end; no
    swizzle, no sweep, no hot-bin pilot; the module variant carries the L2 policy."""
    from cuburn_b200 import samples, render
    gnm = samples.g24h()
    w, h, spp = 7680, 4320, 25
    gprof, tc = still_profile(gnm, w, h, spp)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    s, _, out = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    dim = rmgr.fb.calc_dim(w, h)
    nbins = dim.ah * dim.astride
    k = _kernels(lib, s)
    assert k == INTERP + ['cb_fill32', 'cb_fill32', 'cb_iterate', 'cb_flush_packed'] + \
        DEFAULT_CHAIN + ['cb_convert_rows']
    it = lib.iterations[-1]
    fills = lib.args_of('cb_fill32', s)
    assert [f[1] for f in fills] == [4 * nbins, 2 * nbins]
    assert it['cells'] == fills[1][0] and it['hist'] == fills[0][0] and it['cells'] != it['hist']
    assert it['swizzle_bins'] == 0 and it['spill'] == 0 and it['spill_bins'] == 0
    assert it['palette_packed'] != 0
    assert it['module'] is rdr.variant(True, True, big_grid=True).handle
    assert (True, True, False, True) in rdr._variants
    flush = lib.args_of('cb_flush_packed', s)[0]
    assert (flush[0], flush[1]) == (it['hist'], it['cells'])
    assert out.shape == (h, w, 4)
    # 4K: float4 reductions, swizzled, with the policy; sweeps limited beyond 0.6 x L2
    gprof4, tc4 = still_profile(samples.g6f(), 3840, 2160, 4000)
    rdr4 = render.Renderer(samples.g6f(), gprof4)
    rdr4.hot = False
    s, _, _ = _frame(lib, rmgr, rdr4, samples.g6f(), gprof4, tc4)
    it = lib.iterations[-1]
    assert it['cells'] == 0 and it['swizzle_bins'] == 8487424 // 65536 * 65536
    assert 0 < it['spill_bins'] < 64
    assert it['module'] is rdr4.variant(True, False, big_grid=True).handle


def test_a_hot_genome_switches_to_shared_memory_cells(lib):
    """cb_hot_scan's verdict (count of bins above 1/512 of the pilot's samples) steers the
    variant: read synchronously on a genome's first frame, asynchronously afterwards."""
    from cuburn_b200 import samples, render
    import ctypes
    gnm = samples.g2m()
    gprof, tc = still_profile(gnm, 1920, 1080, 100)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    verdict = [3]

    def scan_result(dst, src, nbytes):
        if nbytes == 4 and src == rmgr.d_hot.ptr + render.HOT_COUNT_OFF:
            ctypes.c_int32.from_address(dst).value = verdict[0]
    lib.d2h_hook = scan_result
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    pilot, main = lib.iterations[-2:]
    assert rdr.hot is True and rmgr.last_iter_hot
    assert pilot['hot_tags'] == 0 and main['hot_tags'] == rmgr.d_hot.ptr + render.HOT_TAGS_OFF
    assert main['module'] is rdr.variant(True, False, True).handle
    assert pilot['module'] is rdr.variant(True, False, False).handle
    scan = lib.args_of('cb_hot_scan', s)[0]
    assert float(scan[6]) == np.float32(pilot['nsamples'] / 2048.)      # listed above this
    assert float(scan[7]) == np.float32(pilot['nsamples'] / 512.)       # trigger
    # a hot genome is scanned on every frame, without another host synchronisation
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    assert 'cb_hot_scan' in lib.names(s) and 'cb_stream_sync' not in lib.names(s)
    assert lib.iterations[-1]['hot_tags'] != 0
    # ... and cools down when a later scan finds nothing
    verdict[0] = 0
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)     # this frame's scan says "cold"
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)     # the next one follows it
    assert rdr.hot is False and lib.iterations[-1]['hot_tags'] == 0


class _Hook(object):
    integer_sums = True

    def __init__(self, lib):
        self.lib, self.at = lib, None

    def __call__(self, fb, dim, stream):
        self.at = self.lib.mark()
        self.front = fb.d_front.ptr


def test_multi_gpu_still_renders_its_share_and_reduces_integer_sums(lib):
    """Rank r of N takes units [r U / N, (r + 1) U / N) with its own RNG streams; with an
    integer-sum reducer the histogram is handed over unscaled and divided by 255 after."""
    from cuburn_b200 import samples, render, multigpu
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 2000
    gprof, tc = still_profile(gnm, w, h, spp)
    total = w * h * spp
    seen = []
    for rank in range(4):
        start = lib.mark()
        rmgr = render.RenderManager(seed=9, rank=rank, world=4)
        # rank-specific seed table: a second 3 MiB upload after the default one
        seeds = [a for a in lib.args_of('cb_memcpy_h2d', start) if a[2] == 262144 * 12]
        assert len(seeds) == 2
        rdr = render.Renderer(gnm, gprof)
        rdr.hot = False
        rmgr.hist_hook = hook = _Hook(lib)
        s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
        it = lib.iterations[-1]
        first, count = multigpu.sample_share(total, rank, 4, render.UNIT_SAMPLES)
        assert (it['first_sample'], it['nsamples'], it['total_samples']) == (first, count, total)
        seen.append((first, count))
        fin = lib.args_of('cb_hist_finish', s)
        assert len(fin) == 2
        assert float(fin[0][4]) == 1.0 and float(fin[1][4]) == np.float32(1 / 255.)
        assert fin[1][0] == fin[1][1] == hook.front and fin[1][2] == 0 and fin[1][3] == 0
        names = lib.names(s)
        first_fin = names.index('cb_hist_finish')
        assert s + first_fin < hook.at <= s + names.index('cb_hist_finish', first_fin + 1)
    assert seen[0][0] == 0 and sum(c for _, c in seen) == total
    assert all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(3))


def test_banded_still_filters_converts_and_copies_only_its_rows(lib, monkeypatch):
    """BandFilter(shared=...): the chain runs on this rank's band + halo (a narrowed dim and
    offset planes), conversion and the D2H copy cover the band's output rows only, into the
    shared host frame."""
    from cuburn_b200 import samples, render, multigpu
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 200
    gprof, tc = still_profile(gnm, w, h, spp)
    dim = render.Framebuffers.calc_dim(w, h)
    frame = np.zeros((h, w, 4), np.uint8)
    shared = type('Shared', (), {'array': frame, 'close': lambda self: None})()
    rank, world = 1, 4
    rmgr = render.RenderManager(seed=9, rank=rank, world=world)
    rdr = render.Renderer(gnm, gprof)
    rdr.hot = False
    rmgr.band_filter = bf = multigpu.BandFilter(rank, world, shared=shared)
    s, _, out = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    assert out is frame
    row0, row1 = multigpu.band_rows(dim.ah, rank, world)
    halo = multigpu.chain_reach(rdr.filts, gprof, tc)
    assert halo == 160
    lo, hi = max(row0 - halo, 0), min(row1 + halo, dim.ah)
    yuv = lib.args_of('cb_yuv_to_rgb', s)[0]
    band_dim = tuple(yuv[2]._obj)
    assert band_dim == (dim.w, dim.h, dim.aw, hi - lo, dim.astride)
    fb = rmgr.fb
    offsets = {yuv[0] - p for p in (fb.d_front.ptr, fb.d_back.ptr, fb.d_left.ptr, fb.d_right.ptr)}
    assert 16 * lo * dim.astride in offsets
    conv = lib.args_of('cb_convert_rows', s)[0]
    rows = bf.output_rows(dim, 12)
    assert (conv[-3], conv[-2]) == rows == (row0 - 12, row1 - 12)
    d2h = lib.args_of('cb_memcpy_d2h', s)
    assert len(d2h) == 1 and d2h[0][2] == (rows[1] - rows[0]) * w * 4
    assert d2h[0][0].value == frame[rows[0]:].ctypes.data
    # ... out of the converted plane at the band's byte offset
    assert d2h[0][1] - rows[0] * w * 4 in (fb.d_front.ptr, fb.d_back.ptr, fb.d_left.ptr,
                                           fb.d_right.ptr)


def test_native_collectives_run_in_stream_order_around_the_filter_chain(lib):
    """The library's own NCCL entry points (cb_comm.cu) as reducer and band gather: integer
    sums out, cb_hist_reduce (all ranks), the 1/255 scale, the band's chain, cb_band_gather
    to the root -- all on the frame's stream, no host synchronisation in between."""
    from cuburn_b200 import samples, render, multigpu
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 400
    gprof, tc = still_profile(gnm, w, h, spp)
    rank, world = 1, 2
    comm = multigpu.NativeComm(rank, world, exchange_id=lambda raw: raw)
    assert lib.names().count('cb_comm_create') == 1 and 'cb_comm_unique_id' not in lib.names()
    rmgr = render.RenderManager(seed=9, rank=rank, world=world)
    rdr = render.Renderer(gnm, gprof)
    rdr.hot = False
    rmgr.hist_hook = multigpu.HistReducer(root=None, comm=comm, integer_sums=True)
    rmgr.band_filter = multigpu.BandFilter(rank, world, root=0, comm=comm)
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    names = [n for n in lib.names(s) if n in fake_native.KERNEL_CALLS or n.startswith(
        ('cb_hist_reduce', 'cb_band_gather', 'cb_stream_sync', 'cb_device_sync'))]
    assert names == INTERP + ['cb_fill32', 'cb_fill32', 'cb_iterate', 'cb_hist_finish',
                              'cb_hist_reduce', 'cb_hist_finish'] + DEFAULT_CHAIN + \
        ['cb_band_gather', 'cb_convert_rows']
    red = lib.args_of('cb_hist_reduce', s)[0]
    it = lib.iterations[-1]
    assert red[0] is comm.handle and red[3] == -1 and red[4].value == it['stream']
    gat = lib.args_of('cb_band_gather', s)[0]
    assert gat[0] is comm.handle and gat[3] == 0 and gat[4].value == it['stream']
    assert tuple(gat[2]._obj) == tuple(rmgr.fb.calc_dim(w, h))          # the whole frame's dims
    assert rmgr.hist_hook.mean_reduce_ms() == 0.0                       # events bracket it
    comm.close()
    assert lib.names().count('cb_comm_destroy') == 1


def test_allocation_failure_frees_what_was_allocated_and_raises_memory_error(lib):
    """render.py:133-147: if one of the four planes cannot be allocated the others are
    released and MemoryError reaches the caller; a smaller frame then still renders."""
    from cuburn_b200 import samples, render, _native as N
    gnm = samples.g3()
    rmgr = render.RenderManager(seed=5)
    big = 16 * render.Framebuffers.calc_dim(7680, 4320).nbins
    real_malloc, seen = lib._do_malloc, []

    def malloc(nbytes, out):
        if int(nbytes) == big:
            seen.append(int(nbytes))
            if len(seen) == 3:
                return N.CB_ERR_NOMEM           # the third plane does not fit
        return real_malloc(nbytes, out)
    lib._do_malloc = malloc
    gprof, tc = still_profile(gnm, 7680, 4320, 10)
    rdr = render.Renderer(gnm, gprof)
    s = lib.mark()
    with pytest.raises(MemoryError):
        rmgr.queue_frame(rdr, gnm, gprof, tc)
    assert len(seen) == 3 and lib.names(s).count('cb_free') == 2
    fb = rmgr.fb
    assert fb.nbins is None and fb.d_front is fb.d_back is fb.d_left is fb.d_right is None
    assert 'cb_iterate' not in lib.names(s)
    gprof, tc = still_profile(gnm, 640, 360, 256)
    rdr = render.Renderer(gnm, gprof)
    s, _, out = _frame(lib, rmgr, rdr, gnm, gprof, tc)
    assert out.shape == (360, 640, 4) and lib.kernel_launches(s) == 41


def test_frame_seed_reseeds_after_the_previous_conversion(lib):
    """queue_frame(frame_seed=k): a fresh seed table per frame (rank-specific when the
    frame's samples are split), uploaded once the previous frame no longer dithers from it."""
    from cuburn_b200 import samples, render, mwc, multigpu
    gnm = samples.g3()
    gprof, tc = still_profile(gnm, 640, 360, 256)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    _frame(lib, rmgr, rdr, gnm, gprof, tc)
    captured = []
    real_h2d = lib._do_memcpy_h2d if hasattr(lib, '_do_memcpy_h2d') else None
    assert real_h2d is None

    def grab(dst, src, nbytes, stream):
        if nbytes == 262144 * 12:
            captured.append(np.ctypeslib.as_array(
                (np.ctypeslib.ctypes.c_uint32 * (262144 * 3)).from_address(src.value)).copy())
        return 0
    lib._do_memcpy_h2d = grab
    s, _, _ = _frame(lib, rmgr, rdr, gnm, gprof, tc, frame_seed=0)
    names = lib.names(s)
    assert names.index('cb_stream_wait_event') < names.index('cb_memcpy_h2d')
    assert len(captured) == 1
    assert np.array_equal(captured[0].reshape(-1, 3), mwc.make_seeds(262144, host_seed=0))
    rmgr.sample_share = (2, 4)
    _frame(lib, rmgr, rdr, gnm, gprof, tc, frame_seed=7)
    assert np.array_equal(captured[1].reshape(-1, 3), multigpu.make_rank_seeds(2, 4, 7, 262144))
