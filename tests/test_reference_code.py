"""
The oracle (and, on the GPU box, the device) against the reference's OWN device
code compiled for the CPU from /root/reference by oracle/build_ref.py:
all 95 variation bodies, catmull_rom / catmull_rom_mag, the YUV helpers.

Where the reference tree is mounted the library is (re)built; elsewhere the
prebuilt oracle/_ref/libref_kernels.so that travelled with the snapshot is used;
if neither exists the tests skip.
"""
import ctypes

import numpy as np
import pytest

from helpers import single_xform_genome


def _ref():
    from oracle import build_ref
    if build_ref.lib() is None:
        pytest.skip('reference kernels library not available')
    return build_ref


# oracle argument order per variation (see the comment above `variation` in chaos.c)
def _oracle_args(name, pv_values, pa_values):
    from cuburn_b200.genome.variations import var_param_order
    special = {'waves': [('pa', 'xy'), ('pa', 'yy'), ('pv', 'dx2'), ('pv', 'dy2')],
               'popcorn': [('pa', 'xo'), ('pa', 'yo')], 'rings': [('pa', 'xo')],
               'fan': [('pa', 'xo'), ('pa', 'yo')],
               'perspective': [('pv', 'mdist'), ('pv', 'sin'), ('pv', 'cos')],
               'julian': [('pv', 'power'), ('pv', 'cn')], 'juliascope': [('pv', 'power'), ('pv', 'cn')],
               'curve': [('pv', 'xamp'), ('pv', 'yamp'), ('pv', 'x2'), ('pv', 'y2')]}
    order = special.get(name, [('pv', p) for p in var_param_order[name]])
    return [(pv_values if k == 'pv' else pa_values)[n] for k, n in order]


def _names():
    from cuburn_b200.genome.variations import VAR_TABLE
    return [n for _, n, _ in VAR_TABLE]


_VALUES = dict(low=0.4, high=1.2, waves=5.0, a=1.1, b=-0.7, c=0.9, d=1.3, x=0.6, y=0.3, val=0.7,
               mdist=2.2, sin=0.58, cos=1.8, power=3.0, cn=0.21, angle=0.35, slices=5.0,
               rotation=0.3, thickness=0.6, sides=5.0, circle=0.8, corners=1.2, c1=0.5, c2=0.2,
               rot=0.7, twist=7.0, rnd=0.3, m=5.0, n1=1.4, n2=1.2, n3=0.8, holes=0.1, petals=5.0,
               eccentricity=0.7, height=0.8, width=1.3, shift=0.3, size=0.4, r=1.2, i=0.3,
               xamp=0.3, yamp=-0.2, x2=1.5, y2=0.5, beta=0.7, space=0.3, spin=0.8, separation=0.8,
               frequency=2.0, amplitude=1.1, damping=0.3, xinside=0.2, yinside=-0.1, xsize=0.7,
               ysize=1.3, warp=0.4, hole=0.1, count=3.0, swirl=0.2, inside=0.4, outside=-0.3,
               scalex=0.3, scaley=0.2, freqx=2.5, freqy=3.5, spread=0.4, re_a=0.9, im_a=0.1,
               re_b=0.2, im_b=-0.1, re_c=0.1, im_c=0.3, re_d=1.0, im_d=0.2, dx2=4.0, dy2=9.0,
               dist=1.3)
_PA = dict(xx=0.8, xy=-0.3, xo=0.31, yx=0.25, yy=0.9, yo=-0.17)


def _lattice():
    gx, gy = np.meshgrid(np.linspace(-1.7, 1.7, 41), np.linspace(-1.3, 1.9, 37))
    return (np.ascontiguousarray(gx.ravel() + 0.013, np.float32),
            np.ascontiguousarray(gy.ravel() - 0.007, np.float32))


@pytest.mark.parametrize('name', _names())
def test_oracle_variation_equals_reference_code(built, name):
    """oracle/chaos.c `variation` vs the reference's body of the same variation:
    both plain C with libm on the CPU, so they must agree to rounding."""
    B = _ref()
    from cuburn_b200 import mwc
    from oracle import flame_ref as R
    m = B.meta()['variations'][name]
    pv = [_VALUES[n] for n in m['pv']]
    pa = [_PA[n] for n in m['pa']]
    txs, tys = _lattice()
    seeds = mwc.make_seeds(txs.size, host_seed=5)
    w = 0.8
    rtx, rty, rox, roy, rseeds = B.ref_variation(name, txs, tys, w, seeds, pv, pa)
    args = _oracle_args(name, _VALUES, _PA)
    otx, oty, oox, ooy, oseeds = R.variation(R.var_number(name), args, w, txs, tys, seeds)
    assert np.array_equal(rseeds, oseeds), 'RNG consumption differs from the reference'
    for got, want in ((otx, rtx), (oty, rty), (oox, rox), (ooy, roy)):
        both_nan = np.isnan(got) & np.isnan(want)
        same_inf = np.isinf(got) & np.isinf(want) & (np.sign(got) == np.sign(want))
        close = np.abs(got - want) <= 2e-6 * (1 + np.abs(want))
        assert np.all(both_nan | same_inf | close), (name, float(np.nanmax(np.abs(got - want))))


def test_catmull_rom_equals_reference_code(built):
    B = _ref()
    from oracle import flame_ref as R
    rs = np.random.RandomState(11)
    tt = np.linspace(0, 1, 513).astype(np.float32)
    worst_mag = 0.0
    for trial in range(60):
        spec = [rs.randn(), 2 * rs.randn(), rs.randn(), 2 * rs.randn()]
        for t in np.sort(rs.uniform(0.02, 0.98, rs.randint(0, 24))):
            spec += [float(np.round(t, 4)), float(rs.randn())]
        if trial % 3 == 0:
            spec = [abs(v) + 0.07 if i % 2 == 1 and i > 3 or i in (0, 2) else v
                    for i, v in enumerate(spec)]
        t, k = R.normalize_spline(spec, 1.0)
        # plain domain: identical operation order => identical bits
        assert np.array_equal(R.catmull_rom(t, k, tt), B.ref_catmull_rom(t, k, tt))
        # magnitude domain: libm log2f/exp2f vs the deterministic float64 routines
        a, b = R.catmull_rom(t, k, tt, mag=True), B.ref_catmull_rom(t, k, tt, mag=True)
        ok = np.isfinite(a) & np.isfinite(b)
        rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-6)
        worst_mag = max(worst_mag, float(rel.max()))
        assert rel.max() < 5e-6
    assert worst_mag > 0 or True


def test_yuv_helpers_equal_reference_code(built):
    B = _ref()
    from oracle import filters_ref as F
    rs = np.random.RandomState(2)
    pix = rs.uniform(0, 3, (4096, 4)).astype(np.float32)
    pix[:100, 3] = 0
    want = B.ref_yuvo2rgb(pix)
    got = F.yuv_to_rgb(pix.reshape(64, 64, 4)).reshape(-1, 4)
    assert np.allclose(got, want, rtol=0, atol=1e-6)
    # palette-side RGB -> YUV as used by the oracle palette table
    rgb = rs.uniform(0, 1, (1000, 3)).astype(np.float32)
    yuv = B.ref_rgb2yuv(rgb)
    y = (np.float32(0.299) * rgb[:, 0] + np.float32(0.587) * rgb[:, 1]) + np.float32(0.114) * rgb[:, 2]
    assert np.array_equal(y, yuv[:, 0])


@pytest.mark.gpu
@pytest.mark.parametrize('name', _names())
def test_device_variation_vs_reference_code(native, built, name):
    """The device library's variation against the reference's body of the same
    variation (compiled for the CPU) on the same lattice and RNG state."""
    B = _ref()
    N = native
    from cuburn_b200 import mwc
    from cuburn_b200.code import itergen
    from test_iter_gpu import _PARAMS, _DISCONTINUOUS, _module_for, _interp_once
    g = single_xform_genome(name, _PARAMS.get(name), weight=0.8,
                            extra_vars={'linear': {'weight': 0.25}} if name == 'pre_blur' else None)
    g['xforms']['0']['color_speed'] = 0.0
    pk, mod = _module_for(N, g)
    d_par, dim = _interp_once(N, pk, g, 640, 360)
    par = N.from_device(d_par, (pk.param_stride,), np.float32)
    xs, ys = _lattice()
    n = xs.size
    seeds = mwc.make_seeds(n, host_seed=77)
    d_x, d_y, d_c, d_s = (N.to_device(a) for a in (xs, ys, np.zeros(n, np.float32), seeds))
    c = ctypes
    mod.launch('cb_probe_xform', ((n + 255) // 256,), (256,),
               [c.c_uint64(d_par.ptr), c.c_uint64(d_x.ptr), c.c_uint64(d_y.ptr),
                c.c_uint64(d_c.ptr), c.c_uint64(d_s.ptr), c.c_int(n), c.c_float(0.0), c.c_int(0),
                c.c_int(0), c.c_uint64(0), c.c_uint64(0)])
    N.check(N.lib().cb_device_sync())
    gx, gy = N.from_device(d_x, (n,), np.float32), N.from_device(d_y, (n,), np.float32)
    gseeds = N.from_device(d_s, (n, 3), np.uint32)

    # the same xform through the reference's code: pre-affine, then the variation(s)
    S = lambda *p: par[pk.slot('xforms', '0', *p)]
    pa = {k: S('pre_affine', k) for k in ('xx', 'xy', 'xo', 'yx', 'yy', 'yo')}
    tx = (pa['xx'] * xs + pa['xy'] * ys + pa['xo']).astype(np.float32)
    ty = (pa['yx'] * xs + pa['yy'] * ys + pa['yo']).astype(np.float32)
    ox, oy = np.zeros(n, np.float32), np.zeros(n, np.float32)
    rseeds = seeds
    for v in sorted(g['xforms']['0']['variations']):
        m = B.meta()['variations'][v]
        pv = [S('variations', v, p) for p in m['pv']]
        pav = [pa[p] for p in m['pa']]
        tx, ty, dox, doy, rseeds = B.ref_variation(v, tx, ty, float(S('variations', v, 'weight')),
                                                   rseeds, pv, pav)
        ox, oy = ox + dox, oy + doy
    assert np.array_equal(gseeds, rseeds), 'RNG consumption differs from the reference'
    ok = np.isfinite(ox) & np.isfinite(oy) & (np.abs(ox) < 1e4) & (np.abs(oy) < 1e4)
    err = np.maximum(np.abs(gx - ox), np.abs(gy - oy)) / (1.0 + np.maximum(np.abs(ox), np.abs(oy)))
    badfrac = np.mean(err[ok] > 2e-4)
    assert badfrac <= (0.02 if name in _DISCONTINUOUS else 0.0), (name, badfrac, float(err[ok].max()))


# ---- cumulative xform densities (code/iter.py:12-30) -----------------------------------
def test_precalc_densities_equal_reference_code(built):
    """The reference's precalc_densities template, rendered for 3 and 6 xforms and
    compiled for the CPU, against the oracle's restatement: the same float32
    operations in the same order, so the cumulative densities agree bit for bit."""
    B = _ref()
    from oracle import flame_ref as R
    from cuburn_b200 import samples
    rs = np.random.RandomState(11)
    for make, kind in ((samples.g3, 'densities3'), (samples.g6f, 'densities6')):
        for trial in range(40):
            g = make()
            names = sorted(g['xforms'])
            weights = rs.uniform(0.01, 5.0, len(names)) * rs.choice([1.0, 1e-3, 30.0], len(names))
            for n, wgt in zip(names, weights):
                g['xforms'][n]['weight'] = float(wgt)
            ev = R.GenomeEval(g, 640, 360, 0.5, 0.0)
            want = B.ref_precalc(kind, dict(('in_xf_%s_weight' % n, np.float32(wgt))
                                            for n, wgt in zip(names, weights)))
            assert sorted(want) == ['out_cp_den_%s' % n for n in names[:-1]]
            for n in names[:-1]:
                got = np.float32(ev.values['xforms.%s.density' % n][0])
                assert got.view(np.uint32) == np.float32(want['out_cp_den_%s' % n]).view(np.uint32), \
                    (kind, trial, n, got, want)


# ---- precalc hunks (code/iter.py:56-95, variation precalcs) ----------------------------
def test_precalc_equals_reference_code(built):
    B = _ref()
    from oracle import flame_ref as R
    from cuburn_b200 import samples
    rs = np.random.RandomState(4)
    for trial in range(25):
        g = samples.g3()
        cam = dict(rotation=float(rs.uniform(-400, 400)), scale=float(rs.uniform(0.05, 3)),
                   center=dict(x=float(rs.uniform(-2, 2)), y=float(rs.uniform(-2, 2))))
        aff = dict(angle=float(rs.uniform(-720, 720)), spread=float(rs.uniform(-180, 180)),
                   magnitude=dict(x=float(rs.uniform(0.05, 2)), y=float(rs.uniform(0.05, 2))),
                   offset=dict(x=float(rs.uniform(-1, 1)), y=float(rs.uniform(-1, 1))))
        g['camera'] = cam
        g['xforms']['0']['pre_affine'] = aff
        g['xforms']['0']['variations'] = {
            'waves': {'weight': 0.1}, 'linear': {'weight': 0.5},
            'perspective': {'weight': 0.1, 'angle': float(rs.uniform(-2, 2)), 'dist': float(rs.uniform(0.1, 4))},
            'julian': {'weight': 0.1, 'power': float(rs.uniform(0.5, 6)), 'dist': float(rs.uniform(0.1, 3))},
            'curve': {'weight': 0.1, 'xlength': float(rs.uniform(0.1, 3)), 'ylength': float(rs.uniform(0.1, 3))}}
        w, h = 1920, 1080
        ev = R.GenomeEval(g, w, h, 0.5, 0.0)
        v = {k: float(a[0]) for k, a in ev.values.items()}
        d = ev.dim

        def close(got, want, what):
            assert abs(got - want) <= 2e-6 * max(1.0, abs(want)), (what, got, want)
        out = B.ref_precalc('camera', {'in_cam_rotation': cam['rotation'], 'in_cam_scale': cam['scale'],
                                       'in_cam_center_x': cam['center']['x'],
                                       'in_cam_center_y': cam['center']['y']}, w, d['aw'], d['ah'])
        for c in ('xx', 'xy', 'xo', 'yx', 'yy', 'yo'):
            close(v['camera.' + c], out['out_cam_' + c], 'camera.' + c)
        out = B.ref_precalc('affine', {'in_px_angle': aff['angle'], 'in_px_spread': aff['spread'],
                                       'in_px_magnitude_x': aff['magnitude']['x'],
                                       'in_px_magnitude_y': aff['magnitude']['y'],
                                       'in_px_offset_x': aff['offset']['x'],
                                       'in_px_offset_y': aff['offset']['y']})
        for c in ('xx', 'xy', 'xo', 'yx', 'yy', 'yo'):
            close(v['xforms.0.pre_affine.' + c], out['out_px_' + c], 'affine.' + c)
        vs = g['xforms']['0']['variations']
        out = B.ref_precalc('waves', {'in_px_pre_affine_offset_x': aff['offset']['x'],
                                      'in_px_pre_affine_offset_y': aff['offset']['y']})
        close(v['xforms.0.variations.waves.dx2'], out['out_pv_dx2'], 'dx2')
        close(v['xforms.0.variations.waves.dy2'], out['out_pv_dy2'], 'dy2')
        out = B.ref_precalc('perspective', {'in_pv_angle': vs['perspective']['angle'],
                                            'in_pv_dist': vs['perspective']['dist']})
        for n in ('mdist', 'sin', 'cos'):
            close(v['xforms.0.variations.perspective.' + n], out['out_pv_' + n], n)
        out = B.ref_precalc('julian', {'in_pv_dist': vs['julian']['dist'], 'in_pv_power': vs['julian']['power']})
        close(v['xforms.0.variations.julian.cn'], out['out_pv_cn'], 'cn')
        out = B.ref_precalc('curve', {'in_pv_xlength': vs['curve']['xlength'],
                                      'in_pv_ylength': vs['curve']['ylength']})
        close(v['xforms.0.variations.curve.x2'], out['out_pv_x2'], 'x2')
        close(v['xforms.0.variations.curve.y2'], out['out_pv_y2'], 'y2')


# ---- palette (code/interp.py:372-433) -------------------------------------------------------
@pytest.mark.parametrize('npal', [1, 3])
def test_palette_equals_reference_code(built, npal):
    """Same stream assignment (block r -> slot r), same arithmetic => identical levels."""
    B = _ref()
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    g = samples.g6f()
    if npal == 3:
        g['palette'] = [[0.0] + samples.make_palette('spectrum'), [0.5] + samples.make_palette('fire'),
                        [1.0] + samples.make_palette('ocean')]
    ts, td = 0.35, 0.2
    seeds = mwc.make_seeds(262144, host_seed=3)
    mine, mseeds = R.palette_table(g, ts, td, seeds)
    pals = sorted((float(p[0]), R.decode_palette(p[1:])) for p in g['palette'])
    ptimes = np.full(32, 1e9, np.float32)
    ptimes[:len(pals)] = [p[0] for p in pals]
    src = np.zeros((32, 256, 4), np.float32)
    for i, p in enumerate(pals):
        src[i] = p[1]
    packed, rseeds = B.ref_palette(ptimes, src, seeds, np.float32(ts), np.float32(td / 64))
    lo, hi = packed[..., 0].astype(np.uint64), packed[..., 1].astype(np.uint64)
    # out.y = (1 << 22) | (y << 4);  out.x = (u << 18) | v   (interp.py:428-429)
    y = (hi >> np.uint64(4)) & np.uint64(0xff)
    u = (lo >> np.uint64(18)) & np.uint64(0xff)
    v = lo & np.uint64(0xff)
    assert np.all((hi >> np.uint64(22)) == 1)
    lv = np.round(mine[..., :3] * 255).astype(np.uint64)
    assert np.array_equal(lv[..., 0], y) and np.array_equal(lv[..., 1], u) and np.array_equal(lv[..., 2], v)
    assert np.array_equal(mseeds[:16384], rseeds[:16384])


# ---- filters (code/filters.py) ----------------------------------------------------------------
def _field(seed, w=200, h=88, kind='mixed'):
    from oracle import flame_ref as R
    d = R.calc_dim(w, h)
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:d['ah'], 0:d['astride']]
    den = 40 * np.exp(-((xx - 90) ** 2 + (yy - 50) ** 2) / 600.0)
    den += 5 * (np.sin(xx / 7.0) > 0.3) * (yy > 20) * (yy < 70)
    den += np.where((xx < 6) | (yy < 5) | (xx > d['astride'] - 7), 12.0, 0.0)
    den[30:60, 130:170] = 0
    den *= rs.uniform(0.7, 1.3, den.shape)
    f = np.zeros((d['ah'], d['astride'], 4), np.float32)
    f[..., 3] = den
    for ch, ph in enumerate((0.0, 2.0, 4.0)):
        f[..., ch] = den * (0.5 + 0.45 * np.sin(xx / 23.0 + yy / 31.0 + ph))
    if kind == 'points':
        f[...] = 0
    for (py, px, val) in ((10, 10, 300.0), (60, 150, 5000.0), (100, 200, 50.0), (55, 3, 800.0)):
        f[py, px] += np.array([0.8 * val, 0.3 * val, 0.1 * val, val], np.float32)
    return d, f


def _close(got, want, rel, what):
    scale = max(float(np.abs(want).max()), 1e-6)
    bad = np.abs(got - want) > rel * scale + rel * np.abs(want)
    assert not np.isnan(got).any(), what
    assert bad.mean() == 0, (what, float(np.abs(got - want).max()), scale)


def test_pointwise_filters_equal_reference_code(built):
    B = _ref()
    from oracle import filters_ref as F
    d, f = _field(1)
    shape = f.shape[:2]
    out = np.zeros_like(f)
    g = f.copy()
    g[..., 1] += 0.5 * g[..., 3]
    g[..., 2] += 0.5 * g[..., 3]
    B.ref_filter('yuv_to_rgb', out, g, shape=shape)
    _close(F.yuv_to_rgb(g), out, 1e-6, 'yuv_to_rgb')
    k1, k2 = F.logscale_consts(4, 0.28, 200, 88, 256)
    B.ref_filter('logscale', out, f, k1, k2, shape=shape)
    lf = F.logscale(f, k1, k2)
    _close(lf, np.nan_to_num(out), 1e-5, 'logscale')
    gam, lin, lingam = F.calc_lingam(4, 0.01)
    pix = lf.copy()
    B.ref_filter('plainclip', pix, gam - 1, lin, lingam, 1.3, shape=shape)
    _close(F.plainclip(lf, 1.3), pix, 1e-5, 'plainclip')
    for vib, hp in ((1.0, -1.0), (0.6, 2.0), (0.8, -0.4)):
        pix = lf.copy()
        B.ref_filter('colorclip', pix, vib, hp, gam, lin, lingam, shape=shape)
        _close(F.colorclip(lf, vib, hp), pix, 1e-5, 'colorclip')
    B.ref_filter('logencode', out, lf + np.float32(1e-3), 2.2, shape=shape)
    _close(F.logencode(lf + np.float32(1e-3)), out, 1e-5, 'logencode')


@pytest.mark.parametrize('pattern', list(range(16)))
def test_blurs_equal_reference_code(built, pattern):
    B = _ref()
    from oracle import filters_ref as F
    d, f = _field(2)
    shape = f.shape[:2]
    for up, stdev in ((0, 1), (1, 0.7)):
        coefs = F.gauss_coefs(stdev)
        B.ref_set_gauss(coefs)
        o1 = np.zeros(shape, np.float32)
        B.ref_filter('den_blur', o1, f, pattern, up, shape=shape)
        _close(F.blur7(f[..., 3], pattern, up, coefs), o1, 1e-6, 'den_blur')
        plane = np.ascontiguousarray(f[..., 0])
        B.ref_filter('den_blur_1c', o1, plane, pattern, up, shape=shape)
        _close(F.blur7(plane, pattern, up, coefs), o1, 1e-6, 'den_blur_1c')
        o4 = np.zeros_like(f)
        B.ref_filter('full_blur', o4, f, pattern, up, shape=shape)
        _close(F.blur7(f, pattern, up, coefs), o4, 1e-6, 'full_blur')


@pytest.mark.parametrize('pattern', [0, 1, 2, 5, 7])
@pytest.mark.parametrize('kind', ['mixed', 'points'])
def test_bilateral_equals_reference_code(built, pattern, kind):
    B = _ref()
    from oracle import filters_ref as F
    d, f = _field(3, kind=kind)
    shape = f.shape[:2]
    args = dict(sstd=6 * 200 / 1920. * 4, cstd=0.05, dstd=1.5, dpow=0.8, gspeed=4.0)
    coefs = F.gauss_coefs(1)
    B.ref_set_gauss(coefs)
    b0, b1 = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
    B.ref_filter('den_blur', b0, f, pattern, 0, shape=shape)
    B.ref_filter('den_blur_1c', b1, b0, pattern, 1, shape=shape)
    out = np.zeros_like(f)
    B.ref_filter('bilateral', out, f, b1, pattern, 15, args['sstd'], args['cstd'], args['dstd'],
                 args['dpow'], args['gspeed'], shape=shape)
    want = F.bilateral_pass(f, pattern, 15, args['sstd'], args['cstd'], args['dstd'], args['dpow'],
                            args['gspeed'])
    scale = float(np.abs(out).max())
    err = np.abs(want - out)
    assert (err > 1e-4 * scale + 1e-4 * np.abs(out)).mean() < 1e-4, float(err.max())


def test_clip_recipes_equal_reference_code(built):
    """smearclip and haloclip as the host recipes chain them (cuburn/filters.py:110-163)."""
    B = _ref()
    from oracle import filters_ref as F
    d, f = _field(5)
    shape = f.shape[:2]
    k1, k2 = F.logscale_consts(4, 0.28, 200, 88, 256)
    lf = F.logscale(f, k1, k2)
    gam, lin, lingam = F.calc_lingam(4, 0.01)
    # smearclip
    B.ref_set_gauss(F.gauss_coefs(0.7))
    a, b = np.zeros_like(lf), np.zeros_like(lf)
    B.ref_filter('apply_gamma_full_hi', a, lf.copy(), gam - 1, shape=shape)
    for pattern, (src, dst) in zip((2, 3, 0, 1), ((a, b), (b, a), (a, b), (b, a))):
        B.ref_filter('full_blur', dst, src, pattern, 0, shape=shape)
    pix = lf.copy()
    B.ref_filter('smearclip', pix, a, gam - 1, lin, lingam, shape=shape)
    _close(F.smearclip(lf), pix, 2e-5, 'smearclip')
    # haloclip
    B.ref_set_gauss(F.gauss_coefs(1))
    p0, p1 = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
    B.ref_filter('apply_gamma', p0, lf.copy(), 0.1, shape=shape)
    B.ref_filter('den_blur_1c', p1, p0, 2, 0, shape=shape)
    B.ref_filter('den_blur_1c', p0, p1, 3, 0, shape=shape)
    pix = lf.copy()
    B.ref_filter('haloclip', pix, p0, np.float32(1 / 4.0 - 1), shape=shape)
    _close(F.haloclip(lf), pix, 2e-5, 'haloclip')


# ---- pixel formats (code/output.py) ---------------------------------------------------------------
@pytest.mark.parametrize('fmt', ['rgba_u8', 'rgba_u16', 'yuv444p', 'yuv444p10', 'yuv420p10', 'yuv444p12'])
def test_output_equals_reference_code(built, fmt):
    """The reference hands RNG streams to blocks through a ring buffer; ours are
    owned by pixel lanes.  Same maths, different dither draws: levels agree to +-1,
    and exactly wherever no dither is involved."""
    B = _ref()
    from cuburn_b200 import mwc
    from oracle import flame_ref as R, output_ref as O
    w, h = 128, 72                       # multiples of the 32 x 8 launch tile
    d = R.calc_dim(w, h)
    rs = np.random.RandomState(6)
    src = rs.uniform(-0.2, 1.3, (d['ah'], d['astride'], 4)).astype(np.float32)
    src[rs.rand(d['ah'], d['astride']) < 0.2] = 0
    src[..., 3] = np.abs(src[..., 3])
    seeds = mwc.make_seeds(262144, host_seed=9)
    ref = B.ref_convert(fmt, src, w, h, seeds).astype(np.int64)
    mine = O.convert(fmt, src, w, h, seeds)[0].reshape(-1).astype(np.int64)
    assert ref.shape == mine.shape
    diff = np.abs(ref - mine)
    crop = src[12:12 + h, 12:12 + w]
    if fmt == 'yuv444p10':
        # the reference stores U without dither or clamp (a negative float -> u16
        # conversion); we clamp.  Compare U only where it is in range.
        cb = (-0.168736 * crop[..., 0] - 0.331264 * crop[..., 1] + 0.5 * crop[..., 2] + 0.5).reshape(-1)
        diff[w * h:2 * w * h][(cb < 0) | (cb > 1)] = 0
    if fmt == 'yuv420p10':
        # the reference's bounds tests are `>`: its x = w/2 threads overwrite chroma
        # column 0 of the next row and its y = h/2 threads write Cb values over row 0
        # of the Cr plane (code/output.py:164,186-189); ignore those cells
        for plane in (0, 1):
            base = w * h + plane * (w * h // 4)
            view = diff[base: base + w * h // 4].reshape(h // 2, w // 2)
            view[:, 0] = 0
            if plane == 1:
                view[0, :] = 0
    assert diff.max() <= 1, int(diff.max())
    if fmt in ('rgba_u8', 'rgba_u16'):
        zero = (crop <= 0).reshape(-1)
        assert np.all(ref[zero] == 0) and np.all(mine[zero] == 0)
        peak = 255 if fmt == 'rgba_u8' else 65535
        sat = (crop >= 1.0).reshape(-1)
        assert np.all(ref[sat] == peak) and np.all(mine[sat] == peak)


# ---- device vs the reference's code, directly (GPU box) ----------------------------------------
@pytest.mark.gpu
def test_device_palette_vs_reference_code(native, built):
    B = _ref()
    N = native
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    g = samples.g6f()
    g['palette'] = [[0.0] + samples.make_palette('spectrum'), [0.4] + samples.make_palette('fire'),
                    [1.0] + samples.make_palette('ocean')]
    ts, td = 0.2, 0.5
    seeds = mwc.make_seeds(262144, host_seed=4)
    pals = sorted((float(p[0]), R.decode_palette(p[1:])) for p in g['palette'])
    ptimes = np.full(32, 1e9, np.float32)
    ptimes[:3] = [p[0] for p in pals]
    src = np.zeros((32, 256, 4), np.float32)
    for i, p in enumerate(pals):
        src[i] = p[1]
    d_pal, d_seeds = N.DeviceBuffer(64 * 256 * 16), N.to_device(seeds)
    d_pt, d_src = N.to_device(ptimes), N.to_device(src)
    N.check(N.lib().cb_interp_palette(d_pal.ptr, d_seeds.ptr, d_pt.ptr, d_src.ptr,
                                      np.float32(ts), np.float32(td / 64), 64, None))
    N.check(N.lib().cb_device_sync())
    dev = N.from_device(d_pal, (64, 256, 4), np.float32)
    packed, rseeds = B.ref_palette(ptimes, src, seeds, np.float32(ts), np.float32(td / 64))
    lo, hi = packed[..., 0].astype(np.uint64), packed[..., 1].astype(np.uint64)
    lv = np.round(dev[..., :3] * 255).astype(np.uint64)
    assert np.array_equal(lv[..., 0], (hi >> np.uint64(4)) & np.uint64(0xff))
    assert np.array_equal(lv[..., 1], (lo >> np.uint64(18)) & np.uint64(0xff))
    assert np.array_equal(lv[..., 2], lo & np.uint64(0xff))
    assert np.array_equal(N.from_device(d_seeds, seeds.shape, np.uint32), rseeds)


@pytest.mark.gpu
def test_device_filters_vs_reference_code(native, built):
    """Device kernels (fast-math SFU intrinsics) vs the reference's kernels on the CPU."""
    B = _ref()
    N = native
    from oracle import filters_ref as F
    from cuburn_b200.filters import gauss_coefs
    L = N.lib()
    d, f = _field(8)
    dim = N.calc_dim(200, 88)
    shape = f.shape[:2]
    up = lambda a: N.to_device(np.ascontiguousarray(a, np.float32))
    # logscale + colorclip + plainclip
    k1, k2 = F.logscale_consts(4, 0.28, 200, 88, 256)
    ref = np.zeros_like(f)
    B.ref_filter('logscale', ref, f, k1, k2, shape=shape)
    ref = np.nan_to_num(ref)
    buf = up(f)
    N.check(L.cb_logscale(buf.ptr, buf.ptr, k1, k2, N.byref(dim), None))
    _close(N.from_device(buf, f.shape, np.float32), ref, 1e-4, 'logscale')
    gam, lin, lingam = F.calc_lingam(4, 0.01)
    for vib, hp in ((1.0, -1.0), (0.7, 1.5)):
        want = ref.copy()
        B.ref_filter('colorclip', want, vib, hp, gam, lin, lingam, shape=shape)
        buf = up(ref)
        N.check(L.cb_colorclip(buf.ptr, np.float32(vib), np.float32(hp), gam, lin, lingam,
                               N.byref(dim), None))
        _close(N.from_device(buf, f.shape, np.float32), want, 2e-4, 'colorclip')
    # blurs, all three kernels, a few directions
    for pattern in (0, 3, 6, 9, 14):
        c = F.gauss_coefs(1)
        B.ref_set_gauss(c)
        want = np.zeros_like(f)
        B.ref_filter('full_blur', want, f, pattern, 0, shape=shape)
        src, dst = up(f), N.DeviceBuffer(f.nbytes)
        N.check(L.cb_full_blur(dst.ptr, src.ptr, pattern, 0, gauss_coefs(1), N.byref(dim), None))
        _close(N.from_device(dst, f.shape, np.float32), want, 1e-5, 'full_blur')
    # one bilateral direction: the fused device pass vs the reference's three kernels
    for pattern in (1, 4):
        args = dict(sstd=6 * 200 / 1920. * 4, cstd=0.05, dstd=1.5, dpow=0.8, gspeed=4.0)
        B.ref_set_gauss(F.gauss_coefs(1))
        b0, b1 = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
        B.ref_filter('den_blur', b0, f, pattern, 0, shape=shape)
        B.ref_filter('den_blur_1c', b1, b0, pattern, 1, shape=shape)
        want = np.zeros_like(f)
        B.ref_filter('bilateral', want, f, b1, pattern, 15, args['sstd'], args['cstd'],
                     args['dstd'], args['dpow'], args['gspeed'], shape=shape)
        src, scratch, dst = up(f), N.DeviceBuffer(f.nbytes), N.DeviceBuffer(f.nbytes)
        N.check(L.cb_bilateral_direction(
            dst.ptr, src.ptr, scratch.ptr, pattern, 15, gauss_coefs(1), np.float32(args['sstd']),
            np.float32(args['cstd']), np.float32(args['dstd']), np.float32(args['dpow']),
            np.float32(args['gspeed']), N.byref(dim), None))
        N.check(L.cb_device_sync())
        got = N.from_device(dst, f.shape, np.float32)
        scale = float(np.abs(want).max())
        assert (np.abs(got - want) > 2e-3 * scale + 2e-3 * np.abs(want)).mean() < 1e-3
