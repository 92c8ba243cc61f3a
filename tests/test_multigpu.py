"""
Multi-GPU host logic: sample shares, frame partition, disjoint RNG
streams, and the histogram reduce -- exercised on CPU with two gloo ranks; the
NCCL path runs on the GPU box (bench.py --gpus N, test marked gpu below).
"""
import os
import socket

import numpy as np
import pytest


def test_sample_share_covers_frame_exactly():
    from cuburn_b200 import multigpu
    for total in (1, 65535, 65536, 65537, 1920 * 1080 * 2000, 3840 * 2160 * 4000 + 17):
        for world in (1, 2, 4, 8):
            parts = [multigpu.sample_share(total, r, world) for r in range(world)]
            assert sum(n for _, n in parts) == total
            pos = 0
            for first, n in parts:
                assert first % 65536 == 0
                if n:
                    assert first == pos
                    pos += n
            if total >= world * 65536:
                sizes = [n for _, n in parts]
                assert max(sizes) - min(sizes) <= 65536


def test_partition_frames_round_robin():
    from cuburn_b200 import multigpu
    frames = list(range(720))
    parts = [multigpu.partition_frames(frames, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == frames
    assert all(len(p) == 90 for p in parts) and parts[3][:3] == [3, 11, 19]


def test_rank_seeds_are_disjoint(built):
    from cuburn_b200 import multigpu, mwc
    world = 8
    seeds = [multigpu.make_rank_seeds(r, world, host_seed=5, nstreams=4096) for r in range(world)]
    mults = [set(s[:, 0].tolist()) for s in seeds]
    for a in range(world):
        for b in range(a + 1, world):
            assert not (mults[a] & mults[b])
    table = set(mwc.load_mults().tolist())
    assert all(m <= table for m in mults)
    assert not np.array_equal(seeds[0][:, 1], seeds[1][:, 1])
    again = multigpu.make_rank_seeds(3, world, host_seed=5, nstreams=4096)
    assert np.array_equal(again, seeds[3])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cuburn_b200 import multigpu, samples
    from oracle import flame_ref as R
    r, w, _ = multigpu.init_process_group('gloo')
    assert (r, w) == (rank, world)
    gnm = samples.g3()
    ev = R.GenomeEval(gnm, 160, 90, 0.5, 0.0)
    total = 160 * 90 * 40
    first, n = multigpu.sample_share(total, rank, world)
    seeds = multigpu.make_rank_seeds(rank, world, host_seed=9, nstreams=32768)
    pal, _ = R.palette_table(gnm, ev.ts, 0.0, multigpu.make_rank_seeds(0, world, 9, 32768))
    hist, _ = R.iterate(ev, pal, seeds[16384:], n, ntraj=64, nthreads=1)
    np.save(os.path.join(out_dir, 'hist_%d.npy' % rank), hist)
    summed = multigpu.reduce_host_hist(hist.copy(), root=0)
    if rank == 0:
        np.save(os.path.join(out_dir, 'reduced.npy'), summed)
        np.save(os.path.join(out_dir, 'share.npy'), np.array([first, n]))
    frames = multigpu.partition_frames(list(range(10)), rank, world)
    assert frames == list(range(rank, 10, world))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_hist_reduce_world2(built, tmp_path):
    """The reduce equals the sum of the per-rank histograms; shares add up."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    h0 = np.load(tmp_path / 'hist_0.npy')
    h1 = np.load(tmp_path / 'hist_1.npy')
    red = np.load(tmp_path / 'reduced.npy')
    assert np.array_equal(red, h0 + h1)
    assert not np.array_equal(h0, h1)            # independent streams
    total = 160 * 90 * 40
    inside = red[..., 3].sum()
    assert 0.9 * total < inside <= total


def test_band_rows_cover_the_grid_with_equal_bands():
    from cuburn_b200 import multigpu
    for ah in (16, 208, 384, 752, 1104, 2192, 4352):
        for world in (1, 2, 3, 4, 8):
            bands = [multigpu.band_rows(ah, r, world) for r in range(world)]
            assert len({b - a for a, b in bands}) == 1            # one gather moves them
            assert all(a % 16 == 0 and (b - a) % 16 == 0 and 0 <= a < b <= ah for a, b in bands)
            covered = np.zeros(ah, bool)
            for a, b in bands:
                covered[a:b] = True
            assert covered.all()
            assert bands[0][0] == 0 and bands[-1][1] == ah


def test_output_rows_assemble_the_whole_frame(built):
    """BandFilter.output_rows: the rows each GPU converts and copies into the shared host
    frame cover every output row, and a rank only writes rows its own band produced
    (rows written by two ranks -- overlapping bands -- hold identical bytes)."""
    from cuburn_b200 import _native as N, multigpu
    for w, h in ((640, 360), (1920, 1080), (3840, 2160), (7680, 4320), (320, 180), (64, 40),
                 (100, 8), (33, 17)):
        dim = N.calc_dim(w, h)
        for world in (1, 2, 3, 4, 5, 7, 8):
            writers = np.zeros(dim.h, int)
            for r in range(world):
                bf = multigpu.BandFilter(r, world, comm=False)
                a, b = bf.output_rows(dim, 12)
                assert 0 <= a <= b <= dim.h
                row0, row1 = multigpu.band_rows(dim.ah, r, world)
                lo = 0 if r == 0 else row0              # edge ranks also own the gutter rows
                hi = dim.ah if r == world - 1 else row1
                assert a == b or (lo <= a + 12 and b - 1 + 12 < hi), (w, h, world, r)
                writers[a:b] += 1
            assert writers.min() >= 1, (w, h, world)


def test_shared_frame_is_one_segment_mapped_by_every_rank(built, monkeypatch):
    """SharedFrame: rank 0 creates the POSIX shared-memory segment, the others open it by
    name between two barriers, every rank page-locks its mapping (cb_host_register -- a
    recording stand-in here) and the name is gone once all have it.  Two threads play the
    ranks."""
    import threading
    import fake_native
    from cuburn_b200 import multigpu
    lib = fake_native.install(monkeypatch)
    monkeypatch.setenv('MASTER_PORT', '29%03d' % (os.getpid() % 1000))
    monkeypatch.setattr(multigpu.SharedFrame, '_seq', 0)
    gate = threading.Barrier(2)
    frames, errors = [None, None], []
    expect = '/dev/shm/cuburn_b200_%s_1' % os.environ['MASTER_PORT']

    def wait():
        gate.wait(30)                   # a stuck rank fails the test instead of hanging it

    def rank0():
        try:
            frames[0] = multigpu.SharedFrame((90, 160, 4), 'u1', 0, 2, barrier=wait)
        except Exception as e:          # pragma: no cover
            errors.append(e)
            gate.abort()
    t = threading.Thread(target=rank0, daemon=True)
    t.start()
    import time
    for _ in range(1000):               # rank 0 creates and sizes the segment, then waits
        if errors or (os.path.exists(expect) and os.path.getsize(expect) == 90 * 160 * 4):
            break
        time.sleep(0.01)
    assert not errors and os.path.getsize(expect) == 90 * 160 * 4
    multigpu.SharedFrame._seq = 0       # "another process": its own frame counter
    frames[1] = multigpu.SharedFrame((90, 160, 4), 'u1', 1, 2, barrier=wait)
    t.join(30)
    assert not errors, errors
    a, b = frames
    assert a.path == b.path and not os.path.exists(a.path)      # unlinked, mappings live on
    assert a.array.shape == b.array.shape == (90, 160, 4) and a.nbytes == 90 * 160 * 4
    a.array[10:20] = 7                                           # rank 0's band ...
    b.array[50:60] = 9
    assert (b.array[10:20] == 7).all() and (a.array[50:60] == 9).all()     # ... seen by rank 1
    reg = lib.args_of('cb_host_register')
    assert sorted(x[1] for x in reg) == [a.nbytes, a.nbytes] and reg[0][0] != reg[1][0]
    pa, pb = a.array.ctypes.data, b.array.ctypes.data
    a.close()
    b.close()
    b.close()                                                    # idempotent
    assert sorted(x[0] for x in lib.args_of('cb_host_unregister')) == sorted([pa, pb])
    assert a.array is None
    fake_native.uninstall()


def test_c_abi_band_rows_match_the_host_mirror(built):
    import ctypes
    from cuburn_b200 import _native as N, multigpu
    L = N.lib()
    r0, r1 = ctypes.c_int(), ctypes.c_int()
    for ah in (16, 208, 384, 752, 1104, 2192, 4352):
        for world in (1, 2, 3, 4, 8):
            for rank in range(world):
                N.check(L.cb_band_rows(ah, rank, world, ctypes.byref(r0), ctypes.byref(r1)))
                assert (r0.value, r1.value) == multigpu.band_rows(ah, rank, world)
    assert L.cb_band_rows(100, 0, 1, ctypes.byref(r0), ctypes.byref(r1)) == N.CB_ERR_INVALID
    assert L.cb_band_rows(208, 2, 2, ctypes.byref(r0), ctypes.byref(r1)) == N.CB_ERR_INVALID
    # the exchange entry points refuse a null communicator instead of crashing
    dim = N.calc_dim(320, 180)
    assert L.cb_hist_reduce(None, 16, N.byref(dim), 0, None) == N.CB_ERR_INVALID
    assert L.cb_band_gather(None, 16, N.byref(dim), 0, None) == N.CB_ERR_INVALID
    ver = ctypes.c_int()
    bundled = N.preload_nccl()             # PyTorch's NCCL first, so torch can still load
    if L.cb_comm_version(ctypes.byref(ver)) == 0:       # NCCL present on this box
        assert ver.value >= 21800
    if bundled:
        import torch                                    # noqa: F401  (must still import)
        assert torch.cuda.nccl.version() >= (2, 18)


@pytest.mark.gpu
def test_native_comm_world1_is_the_identity(native, built):
    """cb_comm_* on one GPU: a one-rank NCCL communicator, reduce / all-reduce leave
    the histogram as it is and the band gather has nothing to move."""
    N = native
    from cuburn_b200 import multigpu, render
    comm = multigpu.NativeComm(rank=0, world=1)
    fb = render.Framebuffers(seed=1)
    dim = fb.set_dim(320, 180)
    hist = np.random.RandomState(3).rand(dim.ah, dim.astride, 4).astype(np.float32)
    s = N.Stream()
    N.memcpy_htod(fb.d_front, hist, s)
    for root in (0, None):
        multigpu.HistReducer(root=root, comm=comm)(fb, dim, s)
    comm.band_gather(fb.d_front, dim, 0, s)
    s.synchronize()
    assert np.array_equal(N.from_device(fb.d_front, hist.shape, np.float32), hist)
    comm.close()


def _chain_and_profile():
    from cuburn_b200 import samples, profile, filters
    gnm = samples.g3()
    gprof = profile.wrap(dict(width=1920, height=1080, spp=100, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    return filters.create(gprof), gprof, tc


def _random_hist(ah, astride, seed=5):
    rs = np.random.RandomState(seed)
    dens = rs.gamma(0.3, 40.0, size=(ah, astride)).astype(np.float32)
    dens[rs.rand(ah, astride) < 0.3] = 0
    col = rs.rand(ah, astride, 3).astype(np.float32)
    hist = np.empty((ah, astride, 4), np.float32)
    hist[..., 0] = dens * col[..., 0]
    hist[..., 1] = dens * (col[..., 1] * 0.6 + 0.2)
    hist[..., 2] = dens * (col[..., 2] * 0.6 + 0.2)
    hist[..., 3] = dens
    return hist


def _oracle_band(hist, rank, world, halo):
    from cuburn_b200 import multigpu
    from oracle import filters_ref as F
    ah = hist.shape[0]
    a, b = multigpu.band_rows(ah, rank, world)
    e0, e1 = max(a - halo, 0), min(b + halo, ah)
    out = F.default_chain(hist[e0:e1], 1920, 1080, 0.5, 100)
    return a, b, out[a - e0:b - e0]


def test_chain_reach_is_enough_halo_for_the_oracle_chain():
    """Filtering a band plus `chain_reach` rows of halo with the CPU restatement of
    the default chain gives exactly the rows full-frame filtering gives; a 16-row halo
    does not (so the test can see a missing halo)."""
    from cuburn_b200 import multigpu
    from oracle import filters_ref as F
    filts, gprof, tc = _chain_and_profile()
    halo = multigpu.chain_reach(filts, gprof, tc)
    assert halo == 160
    hist = _random_hist(544, 64)
    full = F.default_chain(hist, 1920, 1080, 0.5, 100)
    for rank in range(3):
        a, b, band = _oracle_band(hist, rank, 3, halo)
        assert np.array_equal(band.view(np.uint32), full[a:b].view(np.uint32)), rank
    a, b, band = _oracle_band(hist, 1, 3, 16)
    assert not np.array_equal(band.view(np.uint32), full[a:b].view(np.uint32))


def _band_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cuburn_b200 import multigpu
    multigpu.init_process_group('gloo')
    filts, gprof, tc = _chain_and_profile()
    hist = _random_hist(400, 64)
    a, b, band = _oracle_band(hist, rank, world, multigpu.chain_reach(filts, gprof, tc))
    frame = np.full_like(hist, np.nan)
    frame[a:b] = band
    frame = multigpu.gather_host_bands(frame, rank, world, root=0)
    if rank == 0:
        np.save(os.path.join(out_dir, 'banded.npy'), frame)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_banded_filter_world2(built, tmp_path):
    """Two ranks filter one band each (oracle chain) and gather on the root: the
    assembled frame equals the frame filtered whole."""
    import torch.multiprocessing as mp
    from oracle import filters_ref as F
    mp.spawn(_band_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    full = F.default_chain(_random_hist(400, 64), 1920, 1080, 0.5, 100)
    banded = np.load(tmp_path / 'banded.npy')
    assert np.array_equal(banded.view(np.uint32), full.view(np.uint32))


@pytest.mark.gpu
def test_banded_filter_chain_is_bit_identical(native, built):
    """Every band of the sharded filter chain (BandFilter.filter_band on one GPU,
    playing each rank in turn) equals the same rows of the frame filtered whole."""
    N = native
    from cuburn_b200 import samples, render, profile, multigpu
    gnm = samples.g6f()
    gprof = profile.wrap(dict(width=256, height=1000, spp=60, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=4)
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(256, 1000)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    s = rmgr.stream_a
    s.synchronize()
    shape = (dim.ah, dim.astride, 4)
    hist = N.from_device(rmgr.fb.d_front, shape, np.float32)
    assert hist[..., 3].sum() > 0.5 * 256 * 1000 * 60
    rmgr._filter(rdr, gprof, dim, tc)
    s.synchronize()
    full = N.from_device(rmgr.fb.d_front, shape, np.float32)
    assert np.isfinite(full).all() and full[..., 3].max() > 0
    world = 4
    halo = multigpu.chain_reach(rdr.filts, gprof, tc)
    assert dim.ah > multigpu.band_rows(dim.ah, 0, world)[1] + 2 * halo     # a real interior band
    for rank in range(world):
        N.memcpy_htod(rmgr.fb.d_front, hist, s)
        for plane in (rmgr.fb.d_back, rmgr.fb.d_left, rmgr.fb.d_right):
            N.fill32(plane, 4 * dim.ah * dim.astride, np.float32(np.nan), s)
        rmgr.band_filter = multigpu.BandFilter(rank, world, comm=False)
        rmgr._filter(rdr, gprof, dim, tc)
        s.synchronize()
        a, b = multigpu.band_rows(dim.ah, rank, world)
        band = N.from_device(rmgr.fb.d_front, shape, np.float32)[a:b]
        assert np.array_equal(band.view(np.uint32), full[a:b].view(np.uint32)), rank
    rmgr.band_filter = None


@pytest.mark.gpu
def test_hist_view_is_zero_copy(native, built):
    """torch sees the library's device buffer through __cuda_array_interface__."""
    import torch
    N = native
    buf = N.DeviceBuffer(4 * 1024)
    N.fill32(buf, 1024, np.float32(1.5))
    N.check(N.lib().cb_device_sync())
    t = torch.as_tensor(buf.view((1024,), '<f4'), device='cuda')
    assert t.data_ptr() == buf.ptr and float(t.sum()) == 1536.0
    t += 1
    torch.cuda.synchronize()
    assert np.all(N.from_device(buf, (1024,), np.float32) == 2.5)


def test_rank_seeds_distinguish_seed_zero_and_one(built):
    from cuburn_b200 import multigpu
    a = multigpu.make_rank_seeds(0, 2, 0, 4096)
    b = multigpu.make_rank_seeds(0, 2, 1, 4096)
    assert not np.array_equal(a[:, 1:], b[:, 1:])


@pytest.mark.gpu
def test_frame_seed_keeps_rank_streams_disjoint(native, built):
    """A sample-split still rendered with ``frame_seed``: the two ranks must draw different
    sample sets (identical seeds would make the reduced histogram two copies of one share),
    and the same (rank, frame_seed) must repeat exactly."""
    N = native
    from cuburn_b200 import samples, render
    from helpers import still_profile
    gnm = samples.g3()
    gprof, tc = still_profile(gnm, 320, 180, 64)
    seeds = {}
    for rank in (0, 1, 1):
        rmgr = render.RenderManager(seed=3, rank=rank, world=2)
        rdr = render.Renderer(gnm, gprof)
        evt, _ = rmgr.queue_frame(rdr, gnm, gprof, tc, frame_seed=0)
        evt.synchronize()
        dim = rmgr.fb.calc_dim(320, 180)
        # the filters ran in place; what identifies the sample set is the RNG state left behind
        s = N.from_device(rmgr.fb.d_seeds, (rmgr.fb.nstreams, 3), np.uint32)
        if rank in seeds:
            assert np.array_equal(seeds[rank], s)
        seeds[rank] = s
    assert not np.array_equal(seeds[0][:, 0], seeds[1][:, 0])       # disjoint multipliers
    assert not np.array_equal(seeds[0][:, 1:], seeds[1][:, 1:])


@pytest.mark.gpu
@pytest.mark.parametrize('w,h', [(256, 1000), (3200, 500)])
def test_banded_output_into_a_shared_frame_equals_the_whole_frame(native, built, w, h):
    """Every rank filters its band, converts its own output rows (cb_convert_rows) and
    copies them into one page-locked shared-memory frame: the assembled frame is
    byte-identical to the frame rendered on one GPU.  One process plays the ranks in
    turn on the same histogram.  The wide frame runs the TMA-staged tile kernel for the
    x-major bilateral directions, in the bands as in the whole frame."""
    N = native
    from cuburn_b200 import samples, render, profile, multigpu
    gnm = samples.g6f()
    gprof = profile.wrap(dict(width=w, height=h, spp=40, frame_width=0, start=1, end=2), gnm)
    tc = profile.enumerate_times(gprof)[0][1][0]
    rmgr = render.RenderManager(seed=4)
    rdr = render.Renderer(gnm, gprof)
    # float reductions commute only up to rounding: pin the histogram so that every frame
    # below is filtered from the same bits (the hook stands where the NCCL reduce would)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    rmgr.stream_a.synchronize()
    hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
    rmgr.hist_hook = lambda fb, dim_, stream: N.memcpy_htod(fb.d_front, hist, stream)
    rmgr.fb.reseed(77)
    evt, whole = rmgr.queue_frame(rdr, gnm, gprof, tc)
    evt.synchronize()
    whole = np.array(whole)
    assert whole[..., :3].max() > 0
    world = 4
    shared = multigpu.SharedFrame(whole.shape, whole.dtype, 0, 1)
    shared.array[:] = 0
    covered = np.zeros(h, bool)
    for rank in range(world):
        rmgr.fb.reseed(77)                                  # same samples, same dither
        rmgr.band_filter = multigpu.BandFilter(rank, world, comm=False, shared=shared)
        assert rmgr.band_filter.comm is False
        evt, out = rmgr.queue_frame(rdr, gnm, gprof, tc)
        evt.synchronize()
        assert out is shared.array
        r0, r1 = rmgr.band_filter.output_rows(rmgr.fb.calc_dim(w, h), 12)
        covered[r0:r1] = True
        assert np.array_equal(shared.array[r0:r1], whole[r0:r1]), rank
    rmgr.band_filter = None
    assert covered.all()
    assert np.array_equal(shared.array, whole)
    shared.close()
