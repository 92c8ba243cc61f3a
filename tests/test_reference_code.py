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
                c.c_uint64(d_c.ptr), c.c_uint64(d_s.ptr), c.c_int(n), c.c_float(0.0), c.c_int(0)])
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
