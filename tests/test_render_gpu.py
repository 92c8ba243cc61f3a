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


def _oracle_frame(gnm, w, h, spp, tc, td, seed):
    from cuburn_b200 import mwc
    from oracle import flame_ref as R, filters_ref as F, output_ref as O
    ev = R.GenomeEval(gnm, w, h, tc, td)
    seeds = mwc.make_seeds(262144, host_seed=seed)
    pal, seeds = R.palette_table(gnm, tc - 0.5 * td, td, seeds)
    hist, _ = R.iterate(ev, pal, seeds, w * h * spp)
    pix = F.default_chain(hist, w, h, gnm['camera']['scale'], spp)
    o8, _ = O.convert('rgba_u8', pix, w, h, seeds)
    return o8.reshape(h, w, 4)


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
