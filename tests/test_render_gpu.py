"""
End to end through the reference-facing API (RenderManager.queue_frame): frame
PSNR against the oracle (north star: >= 40 dB), double-buffered queueing,
encoders, and re-use of one compiled module across genomes of equal structure.
"""
import io

import numpy as np
import pytest

from helpers import still_profile, frame_window, psnr

pytestmark = pytest.mark.gpu


def _oracle_frame(gnm, w, h, spp, tc, td, seed, threads=1):
    from cuburn_b200 import mwc
    from oracle import flame_ref as R, filters_ref as F, output_ref as O
    ev = R.GenomeEval(gnm, w, h, tc, td)
    seeds = mwc.make_seeds(262144, host_seed=seed)
    pal, seeds = R.palette_table(gnm, tc - 0.5 * td, td, seeds)
    hist, _ = R.iterate(ev, pal, seeds, w * h * spp)
    pix = F.default_chain(hist, w, h, gnm['camera']['scale'], spp, threads=threads)
    o8, _ = O.convert('rgba_u8', pix, w, h, seeds)
    return o8.reshape(h, w, 4)


@pytest.mark.production_schedule
@pytest.mark.parametrize('gname,spp', [('G3', 1500), ('G6F', 4000)])
def test_frame_psnr_vs_oracle(native, built, gname, spp):
    """T10.  Both sides are Monte-Carlo renders with different sample sets, so the
    comparison needs enough samples for the noise floor to sit below 40 dB."""
    from cuburn_b200 import samples, render
    gnm = samples.GENOMES[gname]()
    w, h = 320, 180
    gprof, tc = still_profile(gnm, w, h, spp)
    rmgr = render.RenderManager(seed=31)
    rdr = render.Renderer(gnm, gprof)
    evt, frame = rmgr.queue_frame(rdr, gnm, gprof, tc)
    evt.synchronize()
    assert evt.query() and evt.time() > 0
    want = _oracle_frame(gnm, w, h, spp, tc, 0.0, seed=77)
    assert frame.shape == want.shape == (h, w, 4) and frame.dtype == np.uint8
    val = psnr(frame[..., :3], want[..., :3])
    assert val >= 40.0, val


@pytest.mark.production_schedule
def test_frame_psnr_vs_oracle_at_1080p(native, built):
    """BASELINE config 2 end to end against the oracle at its real size: the 1080p /
    2000 spp G6F frame through queue_frame (device interpolation, chaos game, filter
    chain, RGBA8) vs the oracle's chaos game (4.1e9 iterations on the host cores), numpy
    filter chain (row strips in threads) and output conversion: >= 40 dB."""
    import os
    from cuburn_b200 import samples, render
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 2000
    gprof, tc = still_profile(gnm, w, h, spp)
    rmgr = render.RenderManager(seed=31)
    rdr = render.Renderer(gnm, gprof)
    evt, frame = rmgr.queue_frame(rdr, gnm, gprof, tc)
    evt.synchronize()
    frame = np.array(frame)
    rmgr.fb.free()
    want = _oracle_frame(gnm, w, h, spp, tc, 0.0, seed=77, threads=os.cpu_count() or 8)
    assert frame.shape == want.shape == (h, w, 4)
    val = psnr(frame[..., :3], want[..., :3])
    assert val >= 40.0, val


@pytest.mark.production_schedule
def test_full_size_1080p_still_properties(native, built):
    """BASELINE config 2 at its real size (1920x1080, 2000 spp, G6F), where the oracle
    is too slow to run: properties that do not depend on the size.
      * conservation: the density plane sums to an integer, every launched sample is
        either in the grid or was rejected by the bounds test (a few per cent);
      * every bin count is a whole number and colour sums stay inside the palette's
        range per sample;
      * the swizzled accumulation layout is a permutation: linear and swizzled
        renders of the same seeds agree in their totals and statistically per block;
      * two renders with independent RNG streams agree to >= 40 dB as 8-bit frames."""
    N = native
    from cuburn_b200 import samples, render
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 2000
    gprof, tc = still_profile(gnm, w, h, spp)
    rmgr = render.RenderManager(seed=5)
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    shape = (dim.ah, dim.astride, 4)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    hists = {}
    for layout in (True, False):
        rmgr.fb.reseed(5)
        rmgr.swizzle = layout
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        hists[layout] = N.from_device(rmgr.fb.d_front, shape, np.float32).astype(np.float64)
    rmgr.swizzle = 'auto'
    n = w * h * spp
    for hist in hists.values():
        count = hist[..., 3]
        total = count.sum()
        assert total == np.floor(total) and 0.90 * n <= total <= n
        assert np.array_equal(count, np.floor(count))
        live = count > 0
        for ch in range(3):           # (Y, U + 0.5, V + 0.5) of 8-bit colours, / 255
            mean = hist[..., ch][live] / count[live]
            assert mean.min() >= -1e-6 and mean.max() <= 1.0 + 1e-6
    a, b = hists[True], hists[False]
    assert abs(a[..., 3].sum() - b[..., 3].sum()) <= 2e-4 * n
    blk = lambda x: x[:1104, :1920, 3].reshape(69, 16, 60, 32).sum((1, 3))
    assert np.abs(blk(a) - blk(b)).sum() / blk(a).sum() < 0.01
    frames = []
    for seed in (101, 202):
        evt, frame = rmgr.queue_frame(rdr, gnm, gprof, tc, frame_seed=seed)
        evt.synchronize()
        frames.append(np.array(frame))
    assert frames[0].shape == (h, w, 4) and frames[0][..., :3].max() > 200
    assert not np.array_equal(frames[0], frames[1])
    assert psnr(frames[0][..., :3], frames[1][..., :3]) >= 40.0


@pytest.mark.production_schedule
@pytest.mark.parametrize('gname,w,h,spp,modes', [
    ('G6F', 3840, 2160, 4000, ('float4',)),                  # BASELINE config 3
    ('G24H', 7680, 4320, 2000, ('packed', 'float4'))])       # BASELINE config 5
def test_full_size_4k_8k_conservation(native, built, gname, w, h, spp, modes):
    """BASELINE configs 3 and 5 at their real sizes: every sample is accounted for
    (whole-number density, sum within the bounds-test losses of the launched count),
    colour sums stay in range, and at 8K the packed-u64 accumulation (what 'auto'
    picks there) and the float4 one agree in their totals and per 64x64 block."""
    N = native
    from cuburn_b200 import samples, render
    gnm = samples.GENOMES[gname]()
    gprof, tc = still_profile(gnm, w, h, spp)
    rmgr = render.RenderManager(seed=6)
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    nbins = dim.ah * dim.astride
    assert rmgr._use_packed(nbins) == (modes[0] == 'packed')         # what 'auto' would do
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc, 0.0)
    n = w * h * spp
    blocks = {}
    for mode in modes:
        rmgr.accumulate = mode
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        assert rmgr.last_iter_samples == n
        hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
        count = hist[..., 3].astype(np.float64)
        total = count.sum()
        assert total == np.floor(total) and 0.85 * n <= total <= n, (total, n)
        assert np.array_equal(count, np.floor(count))
        live = count > 0
        for ch in range(3):
            mean = hist[..., ch][live] / count[live]
            assert mean.min() >= -1e-4 and mean.max() <= 1.0 + 1e-4
        hh, ww = dim.ah // 64 * 64, dim.astride // 64 * 64
        blocks[mode] = count[:hh, :ww].reshape(hh // 64, 64, ww // 64, 64).sum((1, 3))
        del hist, count
    rmgr.accumulate = 'auto'
    if len(modes) == 2:
        a, b = blocks[modes[0]], blocks[modes[1]]
        assert abs(a.sum() - b.sum()) <= 2e-4 * n
        assert np.abs(a - b).sum() / a.sum() < 0.01
    rmgr.fb.free()


def test_queue_frame_pipelining_and_encode(native, built):
    from cuburn_b200 import samples, render
    from PIL import Image
    gnm = samples.g6f(animated=True)
    from cuburn_b200 import profile
    gprof = profile.wrap(dict(width=320, height=180, spp=50, fps=24, duration=1.0), gnm)
    times = [t[0] for _, t in profile.enumerate_times(gprof)][:4]
    rmgr = render.RenderManager(seed=2)
    rdr = render.Renderer(gnm, gprof)
    # software pipelining as in main.py:63-76: queue k+1 before waiting on k
    pending, frames = None, []
    for t in times + [None]:
        nxt = rmgr.queue_frame(rdr, gnm, gprof, t) if t is not None else None
        if pending is not None:
            pending[0].synchronize()
            frames.append(np.array(pending[1]))
        pending = nxt
    assert len(frames) == 4
    assert all(f.shape == (180, 320, 4) and f[..., :3].max() > 0 for f in frames)
    # the flame rotates: consecutive frames differ, but not wildly
    d = [np.abs(frames[i].astype(int) - frames[i + 1].astype(int)).mean() for i in range(3)]
    assert all(0.05 < x < 60 for x in d), d
    media, logs = rdr.out.encode(frames[0])
    assert list(media) == ['.jpg']
    img = Image.open(io.BytesIO(media['.jpg'].read()))
    assert img.size == (320, 180)
    assert rdr.out.encode(None) == ({}, [])


def test_motion_blur_uses_all_temporal_samples(native, built):
    """With a wide shutter the frame is the average over the shutter interval: it
    differs from the still at the centre time and is smoother."""
    N = native
    from cuburn_b200 import samples, render, profile
    gnm = samples.g6f(animated=True)
    still = profile.wrap(dict(width=320, height=180, spp=400, frame_width=0, fps=24, duration=1.0), gnm)
    blur = profile.wrap(dict(width=320, height=180, spp=400, frame_width=3.0, fps=24, duration=1.0), gnm)
    rmgr = render.RenderManager(seed=8)
    out = []
    for gp in (still, blur):
        rdr = render.Renderer(gnm, gp)
        dim = rmgr.fb.set_dim(320, 180)
        ts, td = frame_window(gp, 0.3)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, ts, td)
        rmgr._iter(rdr, gnm, gp, dim, 0.3)
        rmgr.stream_a.synchronize()
        out.append(N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)[..., 3])
    a, b = out
    assert abs(a.sum() - b.sum()) / a.sum() < 0.05
    # blurred density is spread over more bins
    assert (b > 0).sum() > 1.05 * (a > 0).sum()


def test_module_shared_by_structure(native, built):
    from cuburn_b200 import samples, render
    g1, g2 = samples.g6f(), samples.g6f()
    g2['xforms']['3']['color'] = 0.123
    g2['camera']['scale'] = 0.4
    gprof, tc = still_profile(g1, 320, 180, 10)
    r1, r2 = render.Renderer(g1, gprof), render.Renderer(g2, gprof)
    assert r1.mod is r2.mod and len(r1.cubin) > 1000
    g3 = samples.g6f()
    del g3['final_xform']
    assert render.Renderer(g3, gprof).mod is not r1.mod


def test_errors_are_loud(native, built):
    N = native
    with pytest.raises(N.CompileError) as e:
        N.Module('this is not CUDA', 'bad.cu', [], [], ['--gpu-architecture=sm_100a'])
    assert 'bad.cu' in str(e.value)
    with pytest.raises(ValueError):
        N.check(N.lib().cb_den_blur(0, 0, 99, 0, None, N.byref(N.calc_dim(64, 64)), None))
    with pytest.raises(MemoryError):
        N.DeviceBuffer(1 << 46)


def test_resize_and_reuse_manager(native, built):
    """One manager renders different sizes back to back (Framebuffers.alloc grows,
    never shrinks; dims are always recomputed) and different genomes."""
    N = native
    from cuburn_b200 import samples, render
    rmgr = render.RenderManager(seed=6)
    sizes = [(320, 180), (640, 360), (200, 88), (640, 360)]
    frames = []
    for (w, h), gname in zip(sizes, ('G3', 'G6F', 'G3', 'G6F')):
        gnm = samples.GENOMES[gname]()
        gprof, tc = still_profile(gnm, w, h, 40)
        rdr = render.Renderer(gnm, gprof)
        evt, frame = rmgr.queue_frame(rdr, gnm, gprof, tc)
        evt.synchronize()
        frames.append(np.array(frame))
        assert frame.shape == (h, w, 4) and frame[..., :3].max() > 0
        dim = rmgr.fb.calc_dim(w, h)
        assert rmgr.fb.nbins >= dim.ah * dim.astride
    assert rmgr.fb.nbins == 384 * 672            # grew to the largest request, kept it
    # nothing of a larger earlier frame bleeds into a smaller later one
    assert frames[2][:, -1, 3].max() <= 255 and np.isfinite(frames[2]).all()
    rmgr.fb.free()
    assert rmgr.fb.d_front is None and rmgr.fb.nbins is None


def test_other_filter_chains_and_outputs(native, built):
    """colorclip / haloclip / plainclip / logencode chains and 16-bit + planar outputs
    run through queue_frame and give sane frames."""
    from cuburn_b200 import samples, render, profile
    gnm = samples.g6f()
    rmgr = render.RenderManager(seed=12)
    base = dict(width=320, height=180, spp=100, frame_width=0, start=1, end=2)
    cases = [
        (['bilateral', 'logscale', 'colorclip'], dict(type='png'), (180, 320, 4), np.uint8),
        (['logscale', 'haloclip'], dict(type='tiff'), (180, 320, 4), np.uint16),
        (['logscale', 'plainclip'], dict(type='raw', pix_fmt='yuv444p10'), (3, 180, 320), np.uint16),
        (['logscale', 'smearclip'], dict(type='raw', pix_fmt='yuv420p10'), (180 * 320 * 3 // 2,), np.uint16),
        (['logscale', 'colorclip', 'logencode'], dict(type='raw', pix_fmt='yuv444p12'), (3, 180, 320), np.uint16),
    ]
    for order, out, shape, dt in cases:
        prof = dict(base, filter_order=order, output=out)
        gprof = profile.wrap(prof, gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        rdr = render.Renderer(gnm, gprof)
        assert [f.name for f in rdr.filts] == ['yuv'] + order
        evt, frame = rmgr.queue_frame(rdr, gnm, gprof, tc)
        evt.synchronize()
        assert frame.shape == shape and frame.dtype == dt
        assert frame.max() > frame.min()
        media, logs = rdr.out.encode(frame)
        assert len(media) == 1 and len(next(iter(media.values())).read()) > 1000


def test_main_cli_renders_flam3_file(native, built, tmp_path):
    """The reference's command line end to end: flam3 XML in, JPEG + raw preview out."""
    import os
    import subprocess
    import sys
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xml = tmp_path / 'spark.flam3'
    xml.write_text('''<flame size="640 360" center="0 0" scale="120" brightness="4" gamma="4">
      <color index="0" rgb="255 120 20"/><color index="128" rgb="40 200 255"/><color index="255" rgb="250 250 120"/>
      <xform weight="1" color="0" linear="0.6" spherical="0.25" coefs="0.5 0.1 -0.1 0.5 -0.6 0.2"/>
      <xform weight="1" color="0.5" linear="0.5" julian="0.4" julian_power="3" julian_dist="1.1" coefs="0.45 -0.25 0.25 0.45 0.5 -0.1"/>
      <xform weight="0.7" color="1" linear="0.8" sinusoidal="0.3" coefs="0.55 0 0 0.55 0.1 0.55" post="0.95 0 0 0.95 0 0.05"/>
      <finalxform color="0" color_speed="0" linear="0.9" eyefish="0.15" coefs="1 0 0 1 0 0"/>
    </flame>''')
    raw = tmp_path / 'preview.raw'
    r = subprocess.run([sys.executable, os.path.join(root, 'main.py'), str(xml), '-P', 'preview',
                        '--still', '--spp', '150', '-o', str(tmp_path), '--raw', str(raw)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = tmp_path / 'spark_00002.jpg'             # --still renders frame 2 (Q4)
    assert out.exists(), os.listdir(tmp_path)
    img = np.array(Image.open(out))
    assert img.shape == (360, 640, 3) and img.max() > 100 and (img.sum(axis=2) > 30).mean() > 0.02
    assert raw.stat().st_size == 640 * 360 * 4
    assert 'spark_00002' in r.stderr and 'ms' in r.stderr
    # --resume skips the finished frame
    r2 = subprocess.run([sys.executable, os.path.join(root, 'main.py'), str(xml), '-P', 'preview',
                         '--still', '--spp', '150', '-o', str(tmp_path), '--resume'],
                        stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r2.returncode == 0 and 'spark_00002' not in r2.stderr


def test_video_outputs_stream_rendered_frames(native, built, tmp_path):
    """Frames rendered through queue_frame reach the x264 / vpxenc pipes in the pixel
    format each encoder expects (a stand-in script plays the encoder)."""
    import stat
    from test_encoders import FAKE
    from cuburn_b200 import samples, render, profile
    fake = tmp_path / 'fakeenc'
    fake.write_text(FAKE)
    fake.chmod(fake.stat().st_mode | stat.S_IXUSR)
    gnm = samples.g6f(animated=True)
    for otype, suffix, frame_bytes in (
            (dict(type='x264', command=str(fake)), '.h264', 320 * 180 * 3 * 2),
            (dict(type='vp9', command=str(fake), pix_fmt='yuv420p10'), '.webm', 320 * 180 * 3),
            (dict(type='vp8', command=str(fake)), '.webm', 320 * 180 * 3 // 2),
            (dict(type='prores', command=str(fake)), '.mov', 320 * 180 * 3 * 2)):
        gprof = profile.wrap(dict(width=320, height=180, spp=40, fps=24, duration=1.0,
                                  output=otype), gnm)
        times = [t[0] for _, t in profile.enumerate_times(gprof)][:3]
        rmgr = render.RenderManager(seed=2)
        rdr = render.Renderer(gnm, gprof)
        for t in times:
            evt, buf = rmgr.queue_frame(rdr, gnm, gprof, t)
            evt.synchronize()
            assert rdr.out.encode(buf) == ({}, [])
        media, logs = rdr.out.encode(None)
        data = media[suffix].read()
        assert len(data) == 3 * frame_bytes, otype
        assert np.frombuffer(data, 'u1').std() > 1       # an image, not a constant


def test_job_farm_renders_frames_through_local_workers(native, built, tmp_path):
    """distribute.py end to end: the dispatcher starts a worker process per job on this
    GPU, each renders its frame and streams the JPEG back; a second run resumes (nothing
    left to do)."""
    import os
    import subprocess
    import sys
    from PIL import Image
    from cuburn_b200 import samples
    from cuburn_b200.genome.util import json_encode
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    flame = tmp_path / 'whirl.json'
    flame.write_text(json_encode(samples.g6f(animated=True)))
    cmd = [sys.executable, os.path.join(root, 'distribute.py'), 'dispatch', str(flame),
           '--worker', 'localhost/0', 'localhost/0', '-P', 'preview', '--spp', '60',
           '--duration', '1', '--fps', '4', '--skip', '0', '-o', str(tmp_path)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    frames = sorted(tmp_path.glob('whirl_*.jpg'))
    assert [f.name for f in frames] == ['whirl_%05d.jpg' % i for i in (1, 2, 3, 4)]
    imgs = [np.array(Image.open(f)).astype(int) for f in frames]
    assert all(im.shape == (360, 640, 3) and im.max() > 100 for im in imgs)
    assert np.abs(imgs[0] - imgs[2]).mean() > 0.05              # the flame moves
    assert not list(tmp_path.glob('*.tmp')) and 'whirl_00003' in r.stderr
    r2 = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r2.returncode == 0 and 'whirl_0000' not in r2.stderr


def test_frame_seed_makes_frames_order_independent(native, built):
    """T11 (animation): with per-frame seeds a frame does not depend on what was rendered
    before it, so frames dealt round-robin to different GPUs equal a sequential render
    (up to the float-add order of the histogram)."""
    from cuburn_b200 import samples, render, profile, multigpu
    gnm = samples.g6f(animated=True)
    gprof = profile.wrap(dict(width=320, height=180, spp=60, fps=24, duration=1.0), gnm)
    times = [t[0] for _, t in profile.enumerate_times(gprof)][:4]

    def render_frames(order):
        rmgr = render.RenderManager(seed=1)
        rdr = render.Renderer(gnm, gprof)
        out = {}
        for k in order:
            evt, buf = rmgr.queue_frame(rdr, gnm, gprof, times[k], frame_seed=1000 + k)
            evt.synchronize()
            out[k] = np.array(buf)
        return out
    seq = render_frames([0, 1, 2, 3])
    rank0 = render_frames(multigpu.partition_frames([0, 1, 2, 3], 0, 2))
    rank1 = render_frames(multigpu.partition_frames([0, 1, 2, 3], 1, 2))
    assert sorted(rank0) == [0, 2] and sorted(rank1) == [1, 3]
    for k in range(4):
        got = rank0[k] if k in rank0 else rank1[k]
        diff = np.abs(got.astype(int) - seq[k].astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01, (k, diff.max(), (diff > 0).mean())
    # without frame seeds the RNG streams simply continue: a frame depends on history
    rm = render.RenderManager(seed=1)
    rd = render.Renderer(gnm, gprof)
    a = np.array(_sync(rm.queue_frame(rd, gnm, gprof, times[1])))
    rm2 = render.RenderManager(seed=1)
    _sync(rm2.queue_frame(rd, gnm, gprof, times[0]))
    b = np.array(_sync(rm2.queue_frame(rd, gnm, gprof, times[1])))
    assert (a != b).mean() > 0.05


def _sync(evt_buf):
    evt_buf[0].synchronize()
    return evt_buf[1]


def test_custom_filter_plugs_into_the_chain(native, built):
    """Filter.register / apply contract (cuburn/filters.py:23-43): a user filter that
    leaves its result in fb.d_front takes part in the chain."""
    N = native
    from cuburn_b200 import samples, render, profile
    from cuburn_b200 import filters as F
    from cuburn_b200.genome import specs

    @F.Filter.register('invert_test')
    class Invert(F.Filter):
        calls = 0

        def apply(self, fb, gprof, params, dim, tc, stream=None):
            # out = 1 - in on the tone-mapped image, via the C ABI: logencode is the only
            # stock kernel with a free dst, so compose: back = front; front = plainclip(...)
            Invert.calls += 1
            N.check(N.lib().cb_memcpy_d2d(fb.d_back.ptr, fb.d_front.ptr,
                                          16 * dim.ah * dim.astride, stream.handle))
            fb.flip()
    specs.filters['invert_test'] = {}
    specs.prof_filters['invert_test'] = {}
    try:
        gnm = samples.g3()
        gprof = profile.wrap(dict(width=160, height=90, spp=50, frame_width=0, start=1, end=2,
                                  filter_order=['logscale', 'invert_test', 'colorclip']), gnm)
        tc = profile.enumerate_times(gprof)[0][1][0]
        rmgr = render.RenderManager(seed=3)
        rdr = render.Renderer(gnm, gprof)
        assert [f.name for f in rdr.filts] == ['yuv', 'logscale', 'invert_test', 'colorclip']
        frame = np.array(_sync(rmgr.queue_frame(rdr, gnm, gprof, tc)))
        assert Invert.calls == 1 and frame[..., :3].max() > 0
    finally:
        F.Filter.filter_map.pop('invert_test', None)
        specs.filters.pop('invert_test', None)
        specs.prof_filters.pop('invert_test', None)
