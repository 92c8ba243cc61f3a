"""
T9: every filter kernel on synthetic float4 fields (point sources, ramps, empty
regions, structure next to the edges) -- device vs the numpy oracle.

Tolerance: the device kernels use SFU fast-math intrinsics (as the reference
does), the oracle uses libm; relative 1e-4 of the field's scale for the
pointwise and blur kernels, 2e-3 for the 31-tap bilateral whose weights chain
five approximate transcendentals per tap.  Zero / empty-bin handling is exact.
"""
import ctypes

import numpy as np
import pytest

from helpers import upload_field

pytestmark = pytest.mark.gpu

W, H = 200, 88        # -> astride 224, ah 112: small enough for the numpy oracle


def _field(seed=0, kind='mixed'):
    from cuburn_b200 import _native as N
    dim = N.calc_dim(W, H)
    rs = np.random.RandomState(seed)
    f = np.zeros((dim.ah, dim.astride, 4), np.float32)
    yy, xx = np.mgrid[0:dim.ah, 0:dim.astride]
    if kind in ('mixed', 'dense'):
        den = 40 * np.exp(-((xx - 90) ** 2 + (yy - 50) ** 2) / 600.0)
        den += 5 * (np.sin(xx / 7.0) > 0.3) * (yy > 20) * (yy < 70)
        den += np.where((xx < 6) | (yy < 5) | (xx > dim.astride - 7), 12.0, 0.0)   # edges
        if kind == 'dense':
            den += 1.0
        den[30:60, 130:170] = 0                                       # an empty hole
        den *= rs.uniform(0.7, 1.3, den.shape)
        f[..., 3] = den
        for ch, ph in enumerate((0.0, 2.0, 4.0)):
            f[..., ch] = den * (0.5 + 0.45 * np.sin(xx / 23.0 + yy / 31.0 + ph))
    if kind in ('mixed', 'points'):
        for (py, px, v) in ((10, 10, 300.0), (60, 150, 5000.0), (100, 200, 50.0), (55, 3, 800.0)):
            f[py, px] += np.array([0.8 * v, 0.3 * v, 0.1 * v, v], np.float32)
    return dim, f.astype(np.float32)


def _run(N, fn, dim, *args):
    N.check(fn(*args, N.byref(dim), None))
    N.check(N.lib().cb_device_sync())


def _close(got, want, rel, what):
    scale = max(float(np.abs(want).max()), 1e-6)
    bad = np.abs(got - want) > rel * scale + rel * np.abs(want)
    assert not np.isnan(got).any(), what
    assert bad.mean() == 0, (what, float(np.abs(got - want).max()), scale, float(bad.mean()))


def test_yuv_to_rgb(native, built):
    N = native
    from oracle import filters_ref as F
    dim, f = _field(1)
    f[..., 1] += 0.5 * f[..., 3]
    f[..., 2] += 0.5 * f[..., 3]
    src, dst = upload_field(N, f), N.DeviceBuffer(f.nbytes)
    _run(N, N.lib().cb_yuv_to_rgb, dim, dst.ptr, src.ptr)
    got = N.from_device(dst, f.shape, np.float32)
    _close(got, F.yuv_to_rgb(f), 1e-5, 'yuv_to_rgb')
    assert got[..., :3].min() >= 0 and np.array_equal(got[..., 3], f[..., 3])


@pytest.mark.parametrize('pattern', list(range(16)))
def test_directional_blurs(native, built, pattern):
    N = native
    from oracle import filters_ref as F
    from cuburn_b200.filters import gauss_coefs
    dim, f = _field(2)
    src = upload_field(N, f)
    L = N.lib()
    for up in (0, 1):
        coefs = gauss_coefs(1 if up == 0 else 0.7)
        cf = F.gauss_coefs(1 if up == 0 else 0.7)
        assert np.array_equal(np.array(list(coefs), np.float32), cf)
        d1 = N.DeviceBuffer(f.nbytes // 4)
        N.check(L.cb_den_blur(d1.ptr, src.ptr, pattern, up, coefs, N.byref(dim), None))
        got = N.from_device(d1, f.shape[:2], np.float32)
        _close(got, F.blur7(f[..., 3], pattern, up, cf), 1e-5, 'den_blur %d' % pattern)
        s1 = upload_field(N, f[..., 0].copy())
        N.check(L.cb_den_blur_1c(d1.ptr, s1.ptr, pattern, up, coefs, N.byref(dim), None))
        got = N.from_device(d1, f.shape[:2], np.float32)
        _close(got, F.blur7(f[..., 0], pattern, up, cf), 1e-5, 'den_blur_1c %d' % pattern)
        d4 = N.DeviceBuffer(f.nbytes)
        N.check(L.cb_full_blur(d4.ptr, src.ptr, pattern, up, coefs, N.byref(dim), None))
        got = N.from_device(d4, f.shape, np.float32)
        _close(got, F.blur7(f, pattern, up, cf), 1e-5, 'full_blur %d' % pattern)


def test_blur_preserves_mass_and_point_source(native, built):
    """A unit impulse far from the edges spreads into exactly the 7 sheared taps."""
    N = native
    from oracle import filters_ref as F
    from cuburn_b200.filters import gauss_coefs
    dim = N.calc_dim(W, H)
    f = np.zeros((dim.ah, dim.astride, 4), np.float32)
    f[50, 100] = 1.0
    src, d4 = upload_field(N, f), N.DeviceBuffer(f.nbytes)
    for pattern in (0, 2, 5, 9):
        N.check(N.lib().cb_full_blur(d4.ptr, src.ptr, pattern, 0, gauss_coefs(1),
                                     N.byref(dim), None))
        got = N.from_device(d4, f.shape, np.float32)[..., 3]
        assert abs(got.sum() - 1.0) < 1e-5
        ys, xs = np.nonzero(got)
        want = sorted((50 - F.shear_offset(pattern, r)[1], 100 - F.shear_offset(pattern, r)[0])
                      for r in range(-3, 4))
        assert sorted(set(zip(ys.tolist(), xs.tolist()))) == sorted(set(want))


@pytest.mark.parametrize('pattern', [0, 1, 3, 6, 7])
@pytest.mark.parametrize('kind', ['mixed', 'points'])
def test_bilateral_pass(native, built, pattern, kind):
    N = native
    from oracle import filters_ref as F
    from cuburn_b200.filters import gauss_coefs
    dim, f = _field(3, kind)
    L = N.lib()
    src = upload_field(N, f)
    d_a, d_b, d_out = N.DeviceBuffer(f.nbytes // 4), N.DeviceBuffer(f.nbytes // 4), N.DeviceBuffer(f.nbytes)
    c1 = gauss_coefs(1)
    args = dict(sstd=6 * W / 1920. * 4, cstd=0.05, dstd=1.5, dpow=0.8, gspeed=4.0)
    N.check(L.cb_den_blur(d_a.ptr, src.ptr, pattern, 0, c1, N.byref(dim), None))
    N.check(L.cb_den_blur_1c(d_b.ptr, d_a.ptr, pattern, 1, c1, N.byref(dim), None))
    N.check(L.cb_bilateral(d_out.ptr, src.ptr, d_b.ptr, pattern, 15,
                           np.float32(args['sstd']), np.float32(args['cstd']),
                           np.float32(args['dstd']), np.float32(args['dpow']),
                           np.float32(args['gspeed']), N.byref(dim), None))
    N.check(L.cb_device_sync())
    got = N.from_device(d_out, f.shape, np.float32)
    want = F.bilateral_pass(f, pattern, 15, args['sstd'], args['cstd'], args['dstd'],
                            args['dpow'], args['gspeed'])
    assert np.all(np.isfinite(got))
    scale = float(np.abs(want).max())
    err = np.abs(got - want)
    assert (err > 2e-3 * scale + 2e-3 * np.abs(want)).mean() < 1e-3, float(err.max())
    # weighted mean of non-negative data stays non-negative; mass roughly conserved
    assert got.min() >= 0
    assert abs(got[..., 3].sum() / f[..., 3].sum() - 1) < 0.35


@pytest.mark.parametrize('pattern', [0, 2, 4, 5, 6, 7])
@pytest.mark.parametrize('kind', ['mixed', 'points', 'dense'])
def test_bilateral_direction_fused(native, built, pattern, kind):
    """The restructured direction pass (hoisted per-pixel terms, one exp2 per tap)
    against the oracle's statement of the reference recipe."""
    N = native
    from oracle import filters_ref as F
    from cuburn_b200.filters import gauss_coefs
    dim, f = _field(6, kind)
    src = upload_field(N, f)
    scratch, d_out = N.DeviceBuffer(f.nbytes), N.DeviceBuffer(f.nbytes)
    args = dict(sstd=6 * W / 1920. * 4, cstd=0.05, dstd=1.5, dpow=0.8, gspeed=4.0)
    N.check(N.lib().cb_bilateral_direction(
        d_out.ptr, src.ptr, scratch.ptr, pattern, 15, gauss_coefs(1),
        np.float32(args['sstd']), np.float32(args['cstd']), np.float32(args['dstd']),
        np.float32(args['dpow']), np.float32(args['gspeed']), N.byref(dim), None))
    N.check(N.lib().cb_device_sync())
    got = N.from_device(d_out, f.shape, np.float32)
    want = F.bilateral_pass(f, pattern, 15, args['sstd'], args['cstd'], args['dstd'],
                            args['dpow'], args['gspeed'])
    assert np.all(np.isfinite(got)) and got.min() >= 0
    scale = float(np.abs(want).max())
    err = np.abs(got - want)
    assert (err > 2e-3 * scale + 2e-3 * np.abs(want)).mean() < 1e-3, float(err.max())
    # and against the unfused reference-shaped kernels on the device
    d_a, d_b, d_ref = N.DeviceBuffer(f.nbytes // 4), N.DeviceBuffer(f.nbytes // 4), N.DeviceBuffer(f.nbytes)
    L, c1 = N.lib(), gauss_coefs(1)
    N.check(L.cb_den_blur(d_a.ptr, src.ptr, pattern, 0, c1, N.byref(dim), None))
    N.check(L.cb_den_blur_1c(d_b.ptr, d_a.ptr, pattern, 1, c1, N.byref(dim), None))
    N.check(L.cb_bilateral(d_ref.ptr, src.ptr, d_b.ptr, pattern, 15, np.float32(args['sstd']),
                           np.float32(args['cstd']), np.float32(args['dstd']),
                           np.float32(args['dpow']), np.float32(args['gspeed']),
                           N.byref(dim), None))
    N.check(L.cb_device_sync())
    ref = N.from_device(d_ref, f.shape, np.float32)
    assert (np.abs(got - ref) > 1e-3 * scale + 1e-3 * np.abs(ref)).mean() < 1e-3


@pytest.mark.parametrize('pattern,twin', [(0, 1), (4, 7), (6, 5)])
def test_tile_kernel_equals_the_window_kernel_on_the_transposed_field(native, built, pattern, twin):
    """The x-major directions run over a shared-memory tile staged by bulk asynchronous
    copies (k_bilateral_tile); transposing the field turns them into the y-major
    directions of k_bilateral_window, which walk the same records in the same order:
    (1,0) <-> (0,1), (1,.5) <-> (.5,1), (1,-.5) <-> (-.5,1).  Square 224 x 224 grid with
    structure at all four edges, so interior tiles, clamped edge blocks and the
    vertically wrapping sheared blocks are all exercised."""
    N = native
    from cuburn_b200.filters import gauss_coefs
    n = 224
    dim = N.Dims(n - 24, n - 24, n, n, n)
    rs = np.random.RandomState(pattern)
    yy, xx = np.mgrid[0:n, 0:n]
    den = 30 * np.exp(-((xx - 80) ** 2 + (yy - 130) ** 2) / 900.0) + rs.gamma(0.5, 4.0, (n, n))
    den[rs.rand(n, n) < 0.25] = 0
    den[:, :3] += 20; den[:2, :] += 15; den[-3:, :] += 9; den[:, -2:] += 11
    f = np.zeros((n, n, 4), np.float32)
    f[..., 3] = den
    for ch in range(3):
        f[..., ch] = den * rs.uniform(0.2, 0.9, (n, n))
    args = (15, gauss_coefs(1), np.float32(6 * 4.0), np.float32(0.05), np.float32(1.5),
            np.float32(0.8), np.float32(4.0))

    def run(field, pat):
        src, scratch, out = upload_field(N, field), N.DeviceBuffer(field.nbytes), N.DeviceBuffer(field.nbytes)
        N.check(N.lib().cb_bilateral_direction(out.ptr, src.ptr, scratch.ptr, pat, *args,
                                               N.byref(dim), None))
        N.check(N.lib().cb_device_sync())
        return N.from_device(out, field.shape, np.float32)
    import os
    os.environ['CB_BILAT_TILE'] = '1'           # small planes default to the per-pixel kernel
    try:
        got = run(f, pattern)
    finally:
        del os.environ['CB_BILAT_TILE']
    untiled = run(f, pattern)
    assert np.abs(got - untiled).max() <= 2e-6 * float(np.abs(untiled).max())
    want = run(np.ascontiguousarray(f.transpose(1, 0, 2)), twin).transpose(1, 0, 2)
    assert np.isfinite(got).all() and got[..., 3].max() > 1
    scale = float(np.abs(want).max())
    assert np.abs(got - want).max() <= 2e-6 * scale, float(np.abs(got - want).max() / scale)


def test_pointwise_tonemap_kernels(native, built):
    N = native
    from oracle import filters_ref as F
    L = N.lib()
    dim, f = _field(4, 'mixed')
    f[..., :3] = np.maximum(f[..., :3], 0)

    def dev(fn, *args, inplace=True, extra=None):
        buf = upload_field(N, f)
        bufs = [buf.ptr] + ([e.ptr for e in extra] if extra else [])
        N.check(fn(*bufs, *args, N.byref(dim), None))
        N.check(L.cb_device_sync())
        return N.from_device(buf, f.shape, np.float32)

    # logscale (in place, zero-density bins must come out exactly zero)
    k1, k2 = F.logscale_consts(4, 0.28, W, H, 256)
    buf = upload_field(N, f)
    N.check(L.cb_logscale(buf.ptr, buf.ptr, k1, k2, N.byref(dim), None))
    got = N.from_device(buf, f.shape, np.float32)
    _close(got, F.logscale(f, k1, k2), 1e-4, 'logscale')
    assert np.all(got[f[..., 3] == 0] == 0)

    lf = F.logscale(f, k1, k2)
    f_save, f[...] = f.copy(), lf
    gam, lin, lingam = F.calc_lingam(4, 0.01)
    got = dev(L.cb_plainclip, np.float32(gam - 1), lin, lingam, np.float32(1.3))
    _close(got, F.plainclip(lf, 1.3), 1e-4, 'plainclip')
    for vib, hp in ((1.0, -1.0), (0.6, 2.0), (0.8, -0.4)):
        got = dev(L.cb_colorclip, np.float32(vib), np.float32(hp), gam, lin, lingam)
        _close(got, F.colorclip(lf, vib, hp), 2e-4, 'colorclip %g %g' % (vib, hp))
        assert got[..., :3].max() <= 1.0 and got[..., 3].max() <= 1.0
    # logencode writes to a second buffer
    src, dst = upload_field(N, lf + 1e-3), N.DeviceBuffer(f.nbytes)
    N.check(L.cb_logencode(dst.ptr, src.ptr, np.float32(2.2), N.byref(dim), None))
    got = N.from_device(dst, f.shape, np.float32)
    _close(got, F.logencode(lf + np.float32(1e-3)), 1e-4, 'logencode')
    f[...] = f_save


def test_smearclip_and_haloclip_recipes(native, built):
    """The multi-kernel recipes through the Filter classes vs the oracle recipes."""
    N = native
    from oracle import filters_ref as F
    from cuburn_b200 import filters, profile, samples

    class FB(object):
        pass
    dim, f = _field(5, 'mixed')
    k1, k2 = F.logscale_consts(4, 0.28, W, H, 256)
    lf = F.logscale(f, k1, k2)
    gnm = samples.g6f()
    gprof = profile.wrap(dict(width=W, height=H), gnm)
    for name, want in (('smearclip', F.smearclip(lf)), ('haloclip', F.haloclip(lf))):
        fb = FB()
        fb.d_front = upload_field(N, lf)
        fb.d_back, fb.d_left, fb.d_right = (N.DeviceBuffer(lf.nbytes) for _ in range(3))
        fb.flip = lambda: None
        filt = filters.Filter.filter_map[name]()
        filt.apply(fb, gprof, getattr(gprof.filters, name), dim, 0.5, None)
        N.check(N.lib().cb_device_sync())
        got = N.from_device(fb.d_front, lf.shape, np.float32)
        _close(got, want, 2e-4, name)
        assert np.all(got[lf[..., 3] <= 0] == 0) or name == 'smearclip'


def test_filter_registry_contract():
    from cuburn_b200 import filters, profile
    gprof = profile.wrap(dict(filter_order=['logscale', 'colorclip']), {'type': 'animation'})
    chain = filters.create(gprof)
    assert [c.name for c in chain] == ['yuv', 'logscale', 'colorclip']
    assert set(filters.Filter.filter_map) >= {'yuv', 'bilateral', 'logscale', 'haloclip',
                                              'smearclip', 'colorclip', 'plainclip', 'logencode'}
    with pytest.raises(KeyError):          # 'de' is schema-only in the reference too (Q8)
        filters.create(profile.wrap(dict(filter_order=['de']), {'type': 'animation'}))
