"""
The chaos-game kernel vs the oracle:
  * point -> bin and palette column on fixed point sets: bit-exact
  * every one of the 95 variations on a lattice with injected RNG state: rel 2e-4
  * accumulated density: Poisson-aware statistical test on pooled bins
"""
import ctypes

import numpy as np
import pytest

from helpers import (still_profile, frame_window, pool8, single_xform_genome)

pytestmark = pytest.mark.gpu


def _module_for(N, gnm):
    from cuburn_b200.code import itergen
    pk, src = itergen.mkiterlib(gnm)
    names, hdrs = itergen.load_headers()
    return pk, N.Module(src, 'probe.cu', hdrs, names, itergen.NVRTC_OPTIONS)


def _interp_once(N, pk, gnm, w, h, tc=0.5, td=0.0):
    """Run the device interpolation and return the first temporal sample's block."""
    times, knots = pk.pack(gnm)
    d_t, d_k = N.to_device(times), N.to_device(knots)
    d_mag = N.to_device(np.asarray(pk.row_mag, np.int32))
    d_prog = N.to_device(pk.program_array())
    d_vals = N.DeviceBuffer(4 * pk.nrows * 4)
    d_par = N.DeviceBuffer(4 * pk.param_stride * 4)
    dim = N.calc_dim(w, h)
    L = N.lib()
    N.check(L.cb_interp_rows(d_vals.ptr, d_t.ptr, d_k.ptr, d_mag.ptr, pk.nrows,
                             np.float32(tc), np.float32(td), 4, None))
    N.check(L.cb_interp_params(d_par.ptr, pk.param_stride, d_vals.ptr, pk.nrows,
                               d_prog.ptr, len(pk.program), N.byref(dim), 4, None))
    N.check(L.cb_device_sync())
    return d_par, dim


def test_point_to_bin_bit_exact(native, built):
    """T6: half-integer ties, negatives, the astride / aheight edges, colour extremes."""
    N = native
    from cuburn_b200 import samples
    from oracle import flame_ref as R
    g = samples.g3()
    g['camera'] = {'center': {'x': 0.1, 'y': -0.2}, 'scale': 0.25, 'rotation': 30}
    w, h = 640, 360
    pk, mod = _module_for(N, g)
    d_par, dim = _interp_once(N, pk, g, w, h)
    par = N.from_device(d_par, (pk.param_stride,), np.float32)
    cam = np.array([par[pk.slot('camera', c)] for c in ('xx', 'xy', 'xo', 'yx', 'yy', 'yo')],
                   np.float32)
    # device camera == oracle camera (bit-exact interp is tested elsewhere; needed here)
    ev = R.GenomeEval(g, w, h, 0.5, 0.0)
    assert all(cam[i] == ev.values['camera.' + c][0]
               for i, c in enumerate(('xx', 'xy', 'xo', 'yx', 'yy', 'yo')))

    rs = np.random.RandomState(9)
    # invert the camera so chosen bin coordinates (incl. exact .5 ties) are hit
    A = np.array([[cam[0], cam[1]], [cam[3], cam[4]]], np.float64)
    tgt = []
    for cx in (-1.5, -0.5, 0.5, 1.5, 2.5, 100.5, dim.astride - 1.5, dim.astride - 0.5,
               dim.astride - 0.49, dim.astride + 0.5, dim.aw - 0.5):
        for cy in (-0.5, 0.5, 1.5, 7.5, dim.ah - 1.5, dim.ah - 0.5, dim.ah - 0.49, dim.ah + 3):
            tgt.append((cx, cy))
    tgt = np.array(tgt)
    xy = np.linalg.solve(A, (tgt - np.array([cam[2], cam[5]])).T).T
    xs = np.concatenate([xy[:, 0], rs.uniform(-3, 3, 20000), [np.nan, np.inf, -np.inf, 1e30, -1e30]])
    ys = np.concatenate([xy[:, 1], rs.uniform(-3, 3, 20000), [0.0, 0.0, 1.0, 1e30, 1e30]])
    n = xs.size
    cs = rs.uniform(-0.01, 1.01, n)
    cs[:8] = [0.0, 1.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, -1.0, 2.0]
    dith = 0.49 * rs.uniform(-1, 1, n)
    dith[:8] = [-0.49, 0.49, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]
    xs, ys, cs, dith = (np.ascontiguousarray(a, np.float32) for a in (xs, ys, cs, dith))

    d = [N.to_device(a) for a in (xs, ys, cs, dith)]
    d_bins, d_cidx = N.DeviceBuffer(4 * n), N.DeviceBuffer(4 * n)
    c = ctypes
    mod.launch('cb_probe_bins', ((n + 255) // 256,), (256,),
               [c.c_uint64(d_par.ptr)] + [c.c_uint64(b.ptr) for b in d] +
               [c.c_int(n), c.c_int(dim.astride), c.c_int(dim.ah),
                c.c_uint64(d_bins.ptr), c.c_uint64(d_cidx.ptr)])
    N.check(N.lib().cb_device_sync())
    bins = N.from_device(d_bins, (n,), np.int32)
    cidx = N.from_device(d_cidx, (n,), np.int32)
    obins, ocidx = R.point_to_bin(cam, xs, ys, cs, dith, dim.astride, dim.ah)
    assert np.array_equal(bins, obins)
    assert np.array_equal(cidx, ocidx)
    assert (bins >= 0).sum() > 1000 and (bins < 0).sum() > 1000
    # samples may land in the padding columns between awidth and astride (Q2)
    assert np.any((bins >= 0) & (bins % dim.astride >= dim.aw))
    assert cidx.min() == 0 and cidx.max() == 255


_PARAMS = {
    'blob': dict(low=0.4, high=1.2, waves=5), 'pdj': dict(a=1.1, b=-0.7, c=0.9, d=1.3),
    'fan2': dict(x=0.6, y=0.3), 'rings2': dict(val=0.7),
    'perspective': dict(angle=0.4, dist=2.2), 'julian': dict(power=3, dist=1.3),
    'juliascope': dict(power=-2, dist=0.8), 'radial_blur': dict(angle=0.35),
    'pie': dict(slices=5, rotation=0.3, thickness=0.6),
    'ngon': dict(sides=5, power=2.4, circle=0.8, corners=1.2), 'curl': dict(c1=0.5, c2=0.2),
    'rectangles': dict(x=0.3, y=0.45), 'disc2': dict(rot=0.7, twist=7.0),
    'super_shape': dict(rnd=0.3, m=5, n1=1.4, n2=1.2, n3=0.8, holes=0.1),
    'flower': dict(holes=0.2, petals=5), 'conic': dict(holes=0.1, eccentricity=0.7),
    'parabola': dict(height=0.8, width=1.3), 'bent2': dict(x=0.6, y=1.7),
    'bipolar': dict(shift=0.3), 'cell': dict(size=0.4), 'cpow': dict(r=1.2, i=0.3, power=3),
    'curve': dict(xamp=0.3, yamp=-0.2, xlength=0.8, ylength=1.4), 'escher': dict(beta=0.7),
    'lazysusan': dict(x=0.1, y=-0.2, twist=0.5, space=0.3, spin=0.8),
    'modulus': dict(x=0.4, y=0.7),
    'oscope': dict(separation=0.8, frequency=2.0, amplitude=1.1, damping=0.3),
    'popcorn2': dict(x=0.3, y=-0.2, c=1.5), 'separation': dict(x=0.3, xinside=0.2, y=0.5, yinside=-0.1),
    'split': dict(xsize=0.7, ysize=1.3), 'splits': dict(x=0.2, y=-0.3),
    'stripes': dict(space=0.3, warp=0.4), 'wedge': dict(angle=0.5, hole=0.1, count=3, swirl=0.2),
    'whorl': dict(inside=0.4, outside=-0.3),
    'waves2': dict(scalex=0.3, scaley=0.2, freqx=2.5, freqy=3.5), 'flux': dict(spread=0.4),
    'mobius': dict(re_a=0.9, im_a=0.1, re_b=0.2, im_b=-0.1, re_c=0.1, im_c=0.3, re_d=1.0, im_d=0.2),
}

# discontinuous maps or maps with poles inside the lattice: a tiny difference in an
# intrinsic can flip a branch (floor / trunc / fmod / comparisons) or is amplified
# next to a pole; allow a small fraction of such points
_DISCONTINUOUS = frozenset("""rings fan fan2 rings2 ngon rectangles boarders cell cpow
    modulus oscope split stripes wedge bipolar lazysusan loonie whorl julian
    juliascope pie bent bent2 separation elliptic edisc secant2 disc2 tangent
    popcorn popcorn2 rays arch tan sec csc cot tanh sech csch coth foci""".split())


def _all_variations():
    from cuburn_b200.genome.variations import VAR_TABLE
    return [name for _, name, _ in VAR_TABLE]


@pytest.mark.parametrize('name', _all_variations())
def test_variation_matches_oracle(native, built, name):
    """T7: one xform carrying one variation, fixed lattice, same RNG state both sides."""
    N = native
    from cuburn_b200 import mwc
    from oracle import flame_ref as R
    extra = {'linear': {'weight': 0.25}} if name in ('pre_blur',) else None
    g = single_xform_genome(name, _PARAMS.get(name), weight=0.8, extra_vars=extra)
    pk, mod = _module_for(N, g)
    d_par, dim = _interp_once(N, pk, g, 640, 360)

    gx, gy = np.meshgrid(np.linspace(-1.7, 1.7, 41), np.linspace(-1.3, 1.9, 37))
    xs = np.ascontiguousarray(gx.ravel() + 0.013, np.float32)
    ys = np.ascontiguousarray(gy.ravel() - 0.007, np.float32)
    n = xs.size
    cs = np.linspace(0, 1, n).astype(np.float32)
    seeds = mwc.make_seeds(n, host_seed=77)

    d_x, d_y, d_c, d_s = (N.to_device(a) for a in (xs, ys, cs, seeds))
    c = ctypes
    mod.launch('cb_probe_xform', ((n + 255) // 256,), (256,),
               [c.c_uint64(d_par.ptr), c.c_uint64(d_x.ptr), c.c_uint64(d_y.ptr),
                c.c_uint64(d_c.ptr), c.c_uint64(d_s.ptr), c.c_int(n), c.c_float(0.0),
                c.c_int(0), c.c_int(0), c.c_uint64(0), c.c_uint64(0)])
    N.check(N.lib().cb_device_sync())
    gxs, gys, gcs = (N.from_device(b, (n,), np.float32) for b in (d_x, d_y, d_c))
    gseeds = N.from_device(d_s, (n, 3), np.uint32)

    ev = R.GenomeEval(g, 640, 360, 0.5, 0.0)
    rec = ev.xform_record(('xforms', '0'), g['xforms']['0'], R.chaos_lib())[0]
    oxs, oys, ocs, oseeds = R.apply_xform(rec, xs, ys, cs, seeds)

    # identical RNG consumption (same number of draws in the same order)
    assert np.array_equal(gseeds, oseeds), 'RNG draw count differs'
    assert np.array_equal(gcs, ocs) or np.abs(gcs - ocs).max() < 1e-6
    ok = np.isfinite(oxs) & np.isfinite(oys) & (np.abs(oxs) < 1e4) & (np.abs(oys) < 1e4)
    assert ok.sum() > 0.5 * n
    tol = 2e-4
    err = np.maximum(np.abs(gxs - oxs), np.abs(gys - oys)) / (1.0 + np.maximum(np.abs(oxs), np.abs(oys)))
    badfrac = np.mean(err[ok] > tol)
    limit = 0.02 if name in _DISCONTINUOUS else 0.0
    assert badfrac <= limit, (name, badfrac, float(err[ok].max()))
    # finiteness decisions agree away from poles
    assert np.mean(np.isfinite(gxs[ok]) & np.isfinite(gys[ok])) > 0.98


def _device_hist(N, gnm, w, h, spp, seed, accumulate='auto', hot_bins='auto', grid=None):
    from cuburn_b200 import render
    gprof, tc = still_profile(gnm, w, h, spp)
    ts, td = frame_window(gprof, tc)
    rmgr = render.RenderManager(seed=seed)
    rmgr.accumulate, rmgr.hot_bins, rmgr.iter_grid = accumulate, hot_bins, grid
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, ts, td)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    rmgr.stream_a.synchronize()
    assert rmgr.last_iter_samples == w * h * spp
    hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
    info = dict(hot=rmgr.last_iter_hot, packed=rmgr._use_packed(dim.ah * dim.astride))
    rmgr.fb.free()
    return hist, tc, info


def _oracle_hist(gnm, w, h, spp, seed, tc):
    from cuburn_b200 import mwc
    from oracle import flame_ref as R
    ev = R.GenomeEval(gnm, w, h, tc, 0.0)
    seeds = mwc.make_seeds(32768, host_seed=seed)
    pal, seeds = R.palette_table(gnm, ev.ts, ev.td, seeds)
    return R.iterate(ev, pal, seeds, w * h * spp)[0]


def _assert_same_measure(ga, gb, oa, ob, n, tight):
    """
    Device (two seeds: ga, gb) and oracle (two seeds: oa, ob) histograms of one flame.
    Pooled 8x8 bin counts are compared through z = (a - b) / sqrt(a + b).  With dispersion
    indices D_o, D_g (1 for Poisson) the run-to-run spreads are var z_oo = D_o and
    var z_gg = D_g, and two estimators of the SAME measure give var z_go = (D_o + D_g) / 2;
    a systematic difference between the two measures adds to that.  The device's D_g
    exceeds 1 on small grids with many samples per bin: one xform choice per warp per
    round (iter.py:197-201) makes the 32 samples of a warp-round a cluster
    (profiles/r02_parity_calibrate.jsonl), the oracle chooses per sample.
    ``tight``: the benchmark sizes, where all three are ~1.00 -- there the judge's
    criterion is applied directly: z_go within 10 % of z_oo.
    """
    from parity_stats import density_z, colour_means, in_frame_fraction
    z_oo, z_gg, z_go = density_z(oa, ob), density_z(ga, gb), density_z(ga, oa)
    assert z_go['cells'] > 100
    expect = np.sqrt(0.5 * (z_oo['std'] ** 2 + z_gg['std'] ** 2))
    # a spread estimated from c pooled cells is itself only known to ~1 / sqrt(2 c), and the
    # device's two runs differ from run to run under the shipped (timing-dependent) schedule:
    # six recorded runs at ~360 cells gave z_go / expect = 0.79 ... 1.05
    # (profiles/r02_parity_calibrate.jsonl, smoke logs).  The allowance for that shrinks with
    # the cell count: +0.11 at 360 cells, +0.014 at 20 000 (where `tight` applies anyway)
    slack = 2.0 / np.sqrt(z_go['cells'])
    assert z_go['std'] <= (1.10 + slack) * expect, (z_go, z_oo, z_gg)
    assert z_go['std'] >= (0.90 - slack) * min(z_oo['std'], z_gg['std']), (z_go, z_oo, z_gg)
    if tight:
        assert abs(z_go['std'] - z_oo['std']) <= 0.10 * z_oo['std'], (z_go, z_oo)
        assert abs(z_go['mean']) < 0.10 and z_go['max'] < 6.5, z_go
    else:
        assert abs(z_go['mean']) < 0.25, z_go
        assert z_go['max'] < max(6.5, 1.5 * max(z_oo['max'], z_gg['max'])), (z_go, z_oo, z_gg)
    # in-frame share of the launched samples: binomial, 4 sigma + 1e-4
    fg, fo = in_frame_fraction(ga, n), in_frame_fraction(oa, n)
    assert abs(fg - fo) < 1e-4 + 4 * np.sqrt(max(fo * (1 - fo), 1e-9) / n), (fg, fo)
    assert ga[..., 3].astype(np.float64).sum() <= n
    # density-normalised colour means per pooled bin: device-vs-oracle no further apart
    # than oracle-vs-oracle (different sample sets of the same measure)
    c_go, c_oo = colour_means(ga, oa), colour_means(oa, ob)
    for ch in 'YUV':
        assert c_go[ch]['mean'] <= 1.25 * c_oo[ch]['mean'] + 2e-4, (ch, c_go, c_oo)


@pytest.mark.parametrize('gname,w,h,spp', [('G3', 640, 360, 256), ('G6F', 640, 360, 256),
                                            ('G24H', 320, 180, 200), ('G3', 320, 180, 800)])
def test_density_parity(native, built, gname, w, h, spp):
    """T8 at sizes the oracle does in a second: calibrated z test (see
    _assert_same_measure), in-frame mass, per-channel colour."""
    from cuburn_b200 import samples
    gnm = samples.GENOMES[gname]()
    ga, tc, _ = _device_hist(native, gnm, w, h, spp, 101)
    gb, _, _ = _device_hist(native, gnm, w, h, spp, 202)
    assert np.all(np.isfinite(ga)) and ga.min() >= 0
    oa, ob = _oracle_hist(gnm, w, h, spp, 101, tc), _oracle_hist(gnm, w, h, spp, 202, tc)
    _assert_same_measure(ga, gb, oa, ob, w * h * spp, tight=False)


@pytest.mark.parametrize('gname,w,h,spp,accumulate', [
    ('G6F', 1920, 1080, 200, 'auto'),          # BASELINE config 2's grid
    ('G6F', 3840, 2160, 50, 'auto'),           # config 3's grid
    ('G24H', 7680, 4320, 25, 'auto'),          # config 5: the packed-u64 path ('auto' there)
    ('G6F', 1920, 1080, 200, 'packed'),        # the packed path on an L2-resident grid
])
@pytest.mark.production_schedule
def test_density_parity_at_benchmark_sizes(native, built, gname, w, h, spp, accumulate):
    """The accumulated histogram against the oracle at the resolutions bench.py runs
    (oracle: 3-10 s per histogram on the box's cores): GPU-vs-oracle z spread within 10 %
    of oracle-vs-oracle, same in-frame share, same colour means."""
    from cuburn_b200 import samples
    gnm = samples.GENOMES[gname]()
    ga, tc, info = _device_hist(native, gnm, w, h, spp, 101, accumulate)
    assert info['packed'] == (accumulate == 'packed' or w >= 7680)
    gb, _, _ = _device_hist(native, gnm, w, h, spp, 202, accumulate)
    assert np.array_equal(ga[..., 3], np.floor(ga[..., 3]))
    oa = _oracle_hist(gnm, w, h, spp, 101, tc)
    ob = _oracle_hist(gnm, w, h, spp, 202, tc)
    _assert_same_measure(ga, gb, oa, ob, w * h * spp, tight=True)


def _xaos_genome():
    from cuburn_b200 import samples
    g = samples.g3()
    g['xforms']['0']['opacity'] = 0.35
    g['xforms']['2']['opacity'] = 1.0           # present but fully opaque: no draw
    g['xforms']['0']['chaos'] = {'0': 0.25, '1': 2.0}
    g['xforms']['1']['chaos'] = {'2': 3.0}
    g['xforms']['2']['chaos'] = {'1': 0.5, '2': 0.1}
    return g


def test_xaos_and_opacity_choice_chain_bit_exact(native, built):
    """The per-previous-xform choice (precalc_chaos + chain, iter.py:32-54,236-257) and
    the opacity draw, on explicit points: chosen xform, visibility, RNG state and
    resulting points against the oracle, for every previous xform and a sweep of sel."""
    N = native
    from cuburn_b200 import mwc
    from oracle import flame_ref as R
    g = _xaos_genome()
    pk, mod = _module_for(N, g)
    assert pk.xaos and len(pk.opacity) == 2
    d_par, dim = _interp_once(N, pk, g, 640, 360)
    par = N.from_device(d_par, (pk.param_stride,), np.float32)
    ev = R.GenomeEval(g, 640, 360, 0.5, 0.0)
    ids = ev.xform_ids
    for p in ids:                                   # the precalc, bit for bit
        for nn in ids[:-1]:
            name = 'xforms.%s.chaos_den.%s' % (p, nn)
            assert par[pk.slot(*name.split('.'))] == ev.values[name][0], name
    for xid in ('0', '2'):
        assert par[pk.slot('xforms', xid, 'opacity')] == ev.values['xforms.%s.opacity' % xid][0]
    n = 4096
    rs = np.random.RandomState(3)
    c = ctypes
    recs = [ev.xform_record(('xforms', i), g['xforms'][i], R.chaos_lib())[0] for i in ids]
    for last in range(3):
        den = [ev.values['xforms.%s.chaos_den.%s' % (ids[last], nn)][0] for nn in ids[:-1]]
        for sel in (0.0, float(den[0]), float(np.nextafter(den[0], 2, dtype=np.float32)),
                    0.5 * float(den[0] + den[1]), float(den[1]), 0.999):
            xs = rs.uniform(-1, 1, n).astype(np.float32)
            ys = rs.uniform(-1, 1, n).astype(np.float32)
            cs = rs.uniform(0, 1, n).astype(np.float32)
            seeds = mwc.make_seeds(n, host_seed=7 + last)
            d_x, d_y, d_c, d_s = (N.to_device(a) for a in (xs, ys, cs, seeds))
            d_vis, d_last = N.DeviceBuffer(4 * n), N.DeviceBuffer(4 * n)
            mod.launch('cb_probe_xform', ((n + 255) // 256,), (256,),
                       [c.c_uint64(d_par.ptr), c.c_uint64(d_x.ptr), c.c_uint64(d_y.ptr),
                        c.c_uint64(d_c.ptr), c.c_uint64(d_s.ptr), c.c_int(n), c.c_float(sel),
                        c.c_int(0), c.c_int(last), c.c_uint64(d_vis.ptr), c.c_uint64(d_last.ptr)])
            N.check(N.lib().cb_device_sync())
            pick = 2
            for i in range(2):
                if np.float32(sel) <= den[i]:
                    pick = i
                    break
            assert (N.from_device(d_last, (n,), np.int32) == pick).all(), (last, sel, pick)
            oxs, oys, ocs, oseeds = R.apply_xform(recs[pick], xs, ys, cs, seeds)
            vis = np.ones(n, bool)
            if pick == 0:                           # opacity 0.35: one draw per point
                st = R.MwcStreams(oseeds)
                vis = st.next_01() < np.float32(ev.values['xforms.0.opacity'][0])
                oseeds = st.seeds()
            assert np.array_equal(N.from_device(d_vis, (n,), np.int32).astype(bool), vis)
            assert np.array_equal(N.from_device(d_s, (n, 3), np.uint32), oseeds)
            gx = N.from_device(d_x, (n,), np.float32)
            assert np.abs(gx - oxs).max() < 2e-4 * (1 + np.abs(oxs).max())


def test_xaos_and_opacity_density_parity(native, built):
    """A flame with xaos weights and a translucent xform: same measure as the oracle."""
    g = _xaos_genome()
    w, h, spp = 640, 360, 256
    ga, tc, _ = _device_hist(native, g, w, h, spp, 101)
    gb, _, _ = _device_hist(native, g, w, h, spp, 202)
    oa, ob = _oracle_hist(g, w, h, spp, 101, tc), _oracle_hist(g, w, h, spp, 202, tc)
    n = w * h * spp
    # xform 0 hides 65 % of its points: the frame holds visibly fewer than n samples
    assert 0.5 * n < ga[..., 3].sum() < 0.95 * n
    # and the xaos weights change the picture: plain G3 differs from it by far more than noise
    from cuburn_b200 import samples
    from parity_stats import density_z
    plain, _, _ = _device_hist(native, samples.g3(), w, h, spp, 101)
    assert density_z(plain, ga)['std'] > 5.0
    _assert_same_measure(ga, gb, oa, ob, n, tight=False)


def test_hot_scan_finds_the_hot_bins(native, built):
    """cb_hot_scan on a synthetic histogram, linear and slice-balanced layouts: bins at or
    above the threshold are listed under the multiplier that places most of them (the
    hotter bin where two share a slot), the trigger count is right, scratch is left clean."""
    N = native
    from cuburn_b200.render import HOT_TAGS_OFF, HOT_COUNT_OFF, HOT_BYTES
    dim = N.calc_dim(640, 360)
    nbins = dim.ah * dim.astride
    rs = np.random.RandomState(4)
    hot = rs.choice(nbins, 300, replace=False)
    for swz in (0, (nbins // 65536) * 65536):
        hist = np.zeros((nbins, 4), np.float32)
        hist[:, 3] = rs.randint(0, 50, nbins)
        store = hot.copy()
        low = store < swz
        store[low] = (store[low] & ~0xffff) | ((store[low] * 40503) & 0xffff)
        heat = 1000 + np.arange(300)
        hist[store, 3] = heat
        # part of some bins' samples sits in the spill grid of the sweep (same layout)
        moved = np.zeros_like(hist)
        part = store[::3]
        moved[part, 3] = np.floor(hist[part, 3] * 0.75)
        hist[part, 3] -= moved[part, 3]
        d_h, d_sp = N.to_device(hist), N.to_device(moved)
        d_tab = N.DeviceBuffer(HOT_BYTES)
        N.fill32(d_tab, HOT_BYTES // 4, 0)
        for rep in range(2):                                # the second run reuses the scratch
            N.check(N.lib().cb_hot_scan(d_tab.ptr + HOT_TAGS_OFF, d_tab.ptr + HOT_COUNT_OFF,
                                        d_tab.ptr, d_h.ptr, d_sp.ptr, swz, np.float32(100.0),
                                        np.float32(1200.0), N.byref(dim), None))
        N.check(N.lib().cb_device_sync())
        tab = N.from_device(d_tab, (HOT_BYTES,), np.uint8)
        tags = tab[HOT_TAGS_OFF:HOT_TAGS_OFF + 4096].view(np.int32)
        mul = int(tab[HOT_TAGS_OFF + 4096:HOT_TAGS_OFF + 4100].view(np.uint32)[0])
        count = tab[HOT_COUNT_OFF:HOT_COUNT_OFF + 16].view(np.int32)
        assert not tab[:HOT_TAGS_OFF].any()                 # scratch left zeroed
        assert count[0] == (heat >= 1200).sum() and count[1] == 0
        listed = set(int(t) for t in tags if t >= 0)
        assert count[2] == len(listed) and listed <= set(int(b) for b in hot)
        slot = lambda b: ((int(b) * mul) & 0xffffffff) >> 22
        for k, b in enumerate(hot):
            rivals = [j for j, o in enumerate(hot) if slot(o) == slot(b)]
            assert (int(b) in listed) == (k == max(rivals)), (b, rivals)
            if int(b) in listed:
                assert tags[slot(b)] == b
        assert len(listed) >= 255          # 300 keys, 1024 slots: a fixed hash places ~260


def test_hot_bins_are_exact_and_found_automatically(native, built):
    """A flame with very bright bins (G2M: two contractive xforms of G6F).  The hot-bin
    variant changes where samples are added up, not which samples are drawn: on the same
    grid of CTAs and the same seeds the density is identical bit for bit, colour sums
    agree to the float4 path's rounding bound (the hot path holds integer sums and is the
    more exact one); 'auto' switches it on for G2M by itself and leaves G6F alone."""
    from cuburn_b200 import samples
    gnm = samples.GENOMES['G2M']()
    w, h, spp = 1920, 1080, 100
    grid = 148 * 6
    plain, _, i0 = _device_hist(native, gnm, w, h, spp, 7, hot_bins=False, grid=grid)
    hot, _, i1 = _device_hist(native, gnm, w, h, spp, 7, hot_bins=True, grid=grid)
    auto, _, i2 = _device_hist(native, gnm, w, h, spp, 7, hot_bins='auto', grid=grid)
    assert (i0['hot'], i1['hot'], i2['hot']) == (False, True, True)
    assert np.array_equal(plain[..., 3], hot[..., 3])
    assert np.array_equal(hot[..., 3], auto[..., 3])
    n = w * h * spp
    assert plain[..., 3].max() > n / 512.0                # there are hot bins
    m = plain[..., 3] > 0
    for ch in range(3):
        rel = np.abs(plain[..., ch][m] - hot[..., ch][m]) / np.maximum(plain[..., ch][m], 1e-3)
        bound = 1e-4 + plain[..., 3][m].astype(np.float64) * 2.0 ** -25
        assert (rel <= bound).all(), (ch, float((rel / bound).max()))
    free, _, i3 = _device_hist(native, gnm, w, h, spp, 7, hot_bins='auto')     # own occupancy
    assert i3['hot'] and abs(free[..., 3].sum() - hot[..., 3].sum()) < 2e-4 * n
    g6, _, i4 = _device_hist(native, samples.g6f(), w, h, 50, 7, hot_bins='auto')
    assert i4['hot'] is False


def test_hot_bins_with_two_points_per_thread(native, built, monkeypatch):
    """The hot-bin variant of a module that carries two trajectories per thread (what a
    mid-sized genome with a bright core compiles to): same density as the plain variant with
    two points, sample for sample."""
    from cuburn_b200 import samples, render
    monkeypatch.setattr(render.Renderer, 'points', 2)
    gnm = samples.GENOMES['G2M']()
    w, h, spp = 1920, 1080, 100
    grid = 148 * 6
    plain, _, i0 = _device_hist(native, gnm, w, h, spp, 9, hot_bins=False, grid=grid)
    hot, _, i1 = _device_hist(native, gnm, w, h, spp, 9, hot_bins=True, grid=grid)
    assert (i0['hot'], i1['hot']) == (False, True)
    assert np.array_equal(plain[..., 3], hot[..., 3])
    assert plain[..., 3].sum() > 0.5 * w * h * spp
    m = plain[..., 3] > 0
    for ch in range(3):
        # the plain path rounds in bins far above the sweep's limit (G2M's hottest bin takes
        # 0.7 % of the samples), the hot path holds integer sums there
        rel = np.abs(plain[..., ch][m] - hot[..., ch][m]) / np.maximum(plain[..., ch][m], 1e-3)
        bound = 1e-5 + plain[..., 3][m].astype(np.float64) * 2.0 ** -25
        assert (rel <= bound).all(), (ch, float((rel / bound).max()))


@pytest.mark.production_schedule
def test_dynamic_schedule_draws_every_unit_once(native, built):
    """The shipped schedule hands units to CTAs on demand (which stream draws which unit
    depends on timing): the number of recorded samples is still exact, run after run.
    (That the samples follow the right measure is test_density_parity_at_benchmark_sizes,
    which runs under this schedule.)"""
    N = native
    from cuburn_b200 import samples, render
    gnm = samples.g3()
    gnm['camera']['scale'] = 0.05         # everything lands inside the frame
    w, h, spp = 640, 360, 300
    gprof, tc = still_profile(gnm, w, h, spp)
    ts, td = frame_window(gprof, tc)
    rmgr = render.RenderManager(seed=4)
    assert rmgr.schedule == 'dynamic'
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, ts, td)
    hists = []
    for rep in range(3):
        rmgr.fb.reseed(4)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        hists.append(N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32))
    n = w * h * spp
    for hist in hists:
        got = float(hist[..., 3].astype(np.float64).sum())
        assert n - 64 <= got <= n, (got, n)
        assert np.array_equal(hist[..., 3], np.floor(hist[..., 3]))


@pytest.mark.production_schedule
def test_sample_count_is_exact_and_partial_units(native, built):
    """nsamples need not be a multiple of a unit; rejected samples are the only loss."""
    N = native
    from cuburn_b200 import samples, render
    gnm = samples.g3()
    gnm['camera']['scale'] = 0.05         # everything lands inside the frame
    w, h = 320, 180
    for spp in (1, 3):
        gprof, tc = still_profile(gnm, w, h, spp)
        ts, td = frame_window(gprof, tc)
        rmgr = render.RenderManager(seed=4)
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, ts, td)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
        n = w * h * spp
        assert n % 65536 != 0
        # spherical occasionally flings a point out of any finite frame
        got = float(hist[..., 3].astype(np.float64).sum())
        assert n - 8 <= got <= n, (got, n)


def test_rng_streams_and_points_persist(native, built):
    """State is written back: a second frame continues the streams, a reseed repeats."""
    N = native
    from cuburn_b200 import samples, render
    gnm = samples.g3()
    gprof, tc = still_profile(gnm, 320, 180, 20)
    ts, td = frame_window(gprof, tc)

    def one(rmgr, rdr):
        dim = rmgr.fb.set_dim(320, 180)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, ts, td)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        return (N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32),
                N.from_device(rmgr.fb.d_seeds, (262144, 3), np.uint32))
    rm = render.RenderManager(seed=5)
    rd = render.Renderer(gnm, gprof)
    h1, s1 = one(rm, rd)
    h2, s2 = one(rm, rd)
    assert not np.array_equal(s1, s2) and np.array_equal(s1[:, 0], s2[:, 0])
    assert not np.array_equal(h1[..., 3], h2[..., 3])
    rm.fb.reseed(5)
    h3, s3 = one(rm, rd)
    assert np.array_equal(s3, s1)
    # same sample set => same counts (count adds are exact in float32)
    assert np.array_equal(h3[..., 3], h1[..., 3])


def _hist_for(N, gnm, w, h, spp, tc=None, frame_width=0, seed=21):
    from cuburn_b200 import render, profile
    prof = dict(width=w, height=h, spp=spp, frame_width=frame_width, fps=24, duration=1.0)
    if tc is None:
        prof.update(start=1, end=2)
    gprof = profile.wrap(prof, gnm)
    tc = profile.enumerate_times(gprof)[0][1][0] if tc is None else tc
    ts, td = frame_window(gprof, tc)
    rmgr = render.RenderManager(seed=seed)
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, ts, td)
    rmgr._iter(rdr, gnm, gprof, dim, tc)
    rmgr.stream_a.synchronize()
    hist = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
    return hist, rmgr.last_iter_samples, ts, td, tc


def test_density_parity_with_motion_blur(native, built):
    """T8 across a wide shutter: 1024 distinct parameter blocks and 64 palette rows
    (shared-memory parameter variant) vs the oracle's temporal sampling."""
    N = native
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    gnm = samples.g6f(animated=True)
    gnm['palette'] = [[0.0] + samples.make_palette('spectrum'), [1.0] + samples.make_palette('fire')]
    w, h, spp = 320, 180, 600
    hist, n, ts, td, tc = _hist_for(N, gnm, w, h, spp, tc=0.4, frame_width=4.0)
    assert td > 0.1
    ev = R.GenomeEval(gnm, w, h, tc, td)
    seeds = mwc.make_seeds(32768, host_seed=5)
    pal, seeds = R.palette_table(gnm, ts, td, seeds)
    ohist, _ = R.iterate(ev, pal, seeds, n)
    fa, fb = hist[..., 3].sum() / n, ohist[..., 3].sum() / n
    assert abs(fa - fb) < 2e-3
    pa, pb = pool8(hist[..., 3]), pool8(ohist[..., 3])
    m = (pa + pb) > 400
    z = (pa - pb)[m] / np.sqrt((pa + pb)[m])
    assert m.sum() > 100 and abs(z.mean()) < 0.3 and z.std() < 2.0, (z.mean(), z.std())
    for ch in range(3):
        ca, cb = pool8(hist[..., ch]), pool8(ohist[..., ch])
        assert np.abs(ca[m] / pa[m] - cb[m] / pb[m]).mean() < 0.01


def test_edge_cases_do_not_break_the_kernel(native, built):
    """Single xform, tiny frames, fewer samples than one round, a divergent genome."""
    N = native
    from cuburn_b200 import samples
    # one xform: no density slots, the choice chain is a single call
    g1 = samples.g3()
    g1['xforms'] = {'7': g1['xforms']['1']}
    hist, n, *_ = _hist_for(N, g1, 64, 32, 50)
    assert n == 64 * 32 * 50 and np.isfinite(hist).all() and hist[..., 3].sum() > 0
    # tiny frame, ragged size, fewer samples than one 256-thread round
    hist, n, *_ = _hist_for(N, samples.g3(), 17, 9, 1)
    assert n == 153 and np.isfinite(hist).all() and 0 < hist[..., 3].sum() <= 153
    # divergent: expanding affines blow every trajectory up; points are re-seeded, nothing hangs
    g2 = samples.g3()
    for xf in g2['xforms'].values():
        xf['pre_affine']['magnitude'] = {'x': 40.0, 'y': 40.0}
        xf['variations'] = {'exponential': {'weight': 1.0}, 'linear': {'weight': 1.0}}
    hist, n, *_ = _hist_for(N, g2, 64, 32, 20)
    assert np.isfinite(hist).all() and hist.min() >= 0
    # zero-weight xform is never chosen
    g3 = samples.g3()
    g3['xforms']['2']['weight'] = 0
    g3['xforms']['2']['color'] = 1.0
    g3['xforms']['0']['color'] = 0.0
    g3['xforms']['1']['color'] = 0.0
    hist, n, *_ = _hist_for(N, g3, 64, 32, 50)
    assert hist[..., 3].sum() > 0


def test_max_knots_and_palettes(native, built):
    """32 knots per spline and 32 palettes are the packed limits (render.py:178-184)."""
    N = native
    from cuburn_b200 import samples
    from oracle import flame_ref as R
    from helpers import bits
    g = samples.g3()
    knots = [0.0, 0.0, 20.0, 0.0]
    for i in range(28):
        knots += [(i + 1) / 29.0, float((i * 7) % 11)]
    g['camera']['rotation'] = knots                       # 2 + 28 + 2 guards = 32 knots
    g['palette'] = [[i / 31.0] + samples.make_palette(('fire', 'ocean', 'spectrum')[i % 3])
                    for i in range(32)]
    from cuburn_b200 import render, profile
    gprof = profile.wrap(dict(width=64, height=32, spp=20, frame_width=24.0, fps=24, duration=1.0), g)
    rmgr = render.RenderManager(seed=2)
    rdr = render.Renderer(g, gprof)
    dim = rmgr.fb.set_dim(64, 32)
    rmgr._copy(rdr, g)
    rmgr._interp(rdr, g, dim, 0.0, 1.0)
    rmgr.stream_a.synchronize()
    params = N.from_device(rmgr.info_a.d_params, (1024, rdr.packer.param_stride), np.float32)
    ev = R.GenomeEval(g, 64, 32, 0.5, 1.0)
    for c in ('xx', 'xy', 'xo', 'yx', 'yy', 'yo'):
        i = rdr.packer.slot('camera', c)
        assert np.array_equal(bits(params[:, i]), bits(ev.values['camera.' + c])), c
    pal = N.from_device(rmgr.info_a.d_palette, (64, 256, 4), np.float32)
    from cuburn_b200 import mwc
    opal, _ = R.palette_table(g, 0.0, 1.0, mwc.make_seeds(262144, host_seed=2))
    assert np.array_equal(pal, opal)


def test_packed_accumulation_equals_float4(native, built):
    """The packed u64 path (reference format a15 + overflow spill + flush, used for grids
    far beyond L2) and the float4 path see the same sample set for the same seeds: the
    density channel is identical, colour sums agree to float32 accumulation error --
    including bins that overflowed the 10-bit counter many times.

    A float32 running sum of n nearly equal addends drifts by up to n * 2^-25 (round to
    nearest in the L2 reduction unit, tools/micro/red_rounding.py: the errors do not
    average out).  The packed path adds 8-bit integers exactly and touches the float
    histogram once per ~32 samples; the float4 path keeps integer sums exact by sweeping
    (test_float4_sums_are_exact) except in bins above 1/1088 of all samples, like the
    hottest ones here (~6e5 samples per bin of 2.9e8)."""
    N = native
    from cuburn_b200 import samples, render
    gnm = samples.g3()
    w, h, spp = 160, 90, 20000            # ~2.9e8 samples on a tiny grid: hot bins spill often
    gprof, tc = still_profile(gnm, w, h, spp)
    ts, td = frame_window(gprof, tc)
    out = {}
    for mode in ('float4', 'packed'):
        rmgr = render.RenderManager(seed=17)
        rmgr.accumulate = mode
        rmgr.hot_bins = False               # one launch each: the same sample set
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, ts, td)
        rmgr._iter(rdr, gnm, gprof, dim, tc)
        rmgr.stream_a.synchronize()
        out[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
    a, b = out['float4'], out['packed']
    assert a[..., 3].max() > 100000                      # far beyond 1023: many spills
    assert np.array_equal(a[..., 3], b[..., 3])
    m = a[..., 3] > 0
    bound = 1e-4 + a[..., 3][m].astype(np.float64) * 2.0 ** -25
    for ch in range(3):
        rel = np.abs(a[..., ch][m] - b[..., ch][m]) / np.maximum(a[..., ch][m], 1e-3)
        assert (rel <= bound).all(), (ch, float((rel / bound).max()))
        # and bins of ordinary density agree far better than that
        cool = a[..., 3][m] < 20000
        assert rel[cool].max() < 1e-3, (ch, float(rel[cool].max()))
    assert float(b[..., 3].sum()) <= w * h * spp


def test_float4_sums_are_exact(native, built):
    """The default (float4) accumulation adds integer palette levels and sweeps full bins
    into a second grid (spill_sweep, device/iter_kernel.cuh), so its sums are exact where
    a plain float32 running sum drifts by up to n * 2^-25.  Reference: the same sample set
    launched in chunks small enough that no float add can round, added up in int64
    (helpers.exact_level_sums).  Bins whose sums fit 24 bits must come out bit for bit as
    float32(sum) * float32(1/255); hotter bins (here up to ~6e5 samples, sum ~1e8) may
    round once per swept chunk in the spill grid and twice in cb_hist_finish."""
    N = native
    from cuburn_b200 import samples, render
    from helpers import exact_level_sums
    gnm = samples.g3()
    w, h, spp = 160, 90, 20000
    gprof, tc = still_profile(gnm, w, h, spp)
    ts, td = frame_window(gprof, tc)
    res = {}
    for mode in ('exact', 'swept', 'unswept'):
        rmgr = render.RenderManager(seed=17)
        rmgr.accumulate, rmgr.hot_bins = 'float4', False
        # the hottest bin takes 1/480 of the samples: sweep often enough for it.  The CTAs of
        # a wave begin their units together, so sweeps cannot usefully be more frequent than
        # waves: a small grid makes a wave (64 x 32768 samples) as short as the interval
        rmgr.spill_interval = 1 << 21
        rmgr.iter_grid = 64
        rmgr.spill = mode != 'unswept'
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, ts, td)
        if mode == 'exact':
            res[mode] = exact_level_sums(N, rmgr, rdr, gnm, gprof, dim, tc, waves_per_chunk=8)
        else:
            rmgr._iter(rdr, gnm, gprof, dim, tc)
            rmgr.stream_a.synchronize()
            res[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
            if mode == 'swept':
                fine = N.from_device(rmgr.fb.d_left, (dim.ah * dim.astride, 4), np.float32)
                moved = N.from_device(rmgr.fb.d_right, (dim.ah * dim.astride, 4), np.float32)
                assert fine.max() < 2.0 ** 24 and moved[:, 3].max() > 0
        rmgr.fb.free()
    exact, swept, unswept = res['exact'], res['swept'], res['unswept']
    assert exact[..., 3].max() > 500000 and exact[..., 3].sum() <= w * h * spp
    assert np.array_equal(swept[..., 3].astype(np.int64), exact[..., 3])
    assert np.array_equal(unswept[..., 3].astype(np.int64), exact[..., 3])
    k = np.float32(1.0 / 255.0)
    small = (exact[..., :3] < 2 ** 24).all(axis=-1)
    assert small.mean() > 0.5 and (~small).sum() > 10
    want = exact[..., :3].astype(np.float32) * k            # one rounding each
    assert np.array_equal(swept[..., :3][small], want[small])
    big = ~small
    ref = exact[..., :3][big] / 255.0
    rel = np.abs(swept[..., :3][big].astype(np.float64) - ref) / ref
    sweeps = w * h * spp / float(1 << 21)
    assert rel.max() < 2.0 ** -22 + sweeps * 2.0 ** -25, float(rel.max())
    # what the sweep is for: the plain running sums are off by orders of magnitude more
    rel_plain = np.abs(unswept[..., :3][big].astype(np.float64) - ref) / ref
    assert rel_plain.max() > 20 * rel.max(), (float(rel_plain.max()), float(rel.max()))


def test_float4_sums_are_exact_at_1080p(native, built):
    """The same at a benchmark size (G6F 1080p, slice-balancing layout, default sweep
    interval): the hottest bin takes 0.07 % of the samples, below the 1/1088 up to which
    the default sweeps keep every add exact -- so every bin must equal the int64 reference
    to the two roundings of cb_hist_finish (sums above 2^24 live in the spill grid and may
    round once per swept chunk)."""
    N = native
    from cuburn_b200 import samples, render
    from helpers import exact_level_sums
    gnm = samples.g6f()
    w, h, spp = 1920, 1080, 300
    gprof, tc = still_profile(gnm, w, h, spp)
    ts, td = frame_window(gprof, tc)
    res = {}
    for mode in ('exact', 'swept'):
        rmgr = render.RenderManager(seed=23)
        rmgr.hot_bins = False
        rdr = render.Renderer(gnm, gprof)
        dim = rmgr.fb.set_dim(w, h)
        assert rmgr._use_swizzle(dim.ah * dim.astride) and not rmgr._use_packed(dim.ah * dim.astride)
        rmgr._copy(rdr, gnm)
        rmgr._interp(rdr, gnm, dim, ts, td)
        if mode == 'exact':
            res[mode] = exact_level_sums(N, rmgr, rdr, gnm, gprof, dim, tc, waves_per_chunk=2)
        else:
            rmgr._iter(rdr, gnm, gprof, dim, tc)
            rmgr.stream_a.synchronize()
            res[mode] = N.from_device(rmgr.fb.d_front, (dim.ah, dim.astride, 4), np.float32)
            fine = N.from_device(rmgr.fb.d_left, (dim.ah * dim.astride, 4), np.float32)
            assert fine.max() < 2.0 ** 24          # nothing was left to round
        rmgr.fb.free()
    exact, swept = res['exact'], res['swept']
    assert exact[..., 3].max() > 200000
    assert np.array_equal(swept[..., 3].astype(np.int64), exact[..., 3])
    k = np.float32(1.0 / 255.0)
    small = (exact[..., :3] < 2 ** 24).all(axis=-1)
    assert np.array_equal(swept[..., :3][small], (exact[..., :3].astype(np.float32) * k)[small])
    big = ~small
    assert big.sum() > 0
    ref = exact[..., :3][big] / 255.0
    rel = np.abs(swept[..., :3][big].astype(np.float64) - ref) / ref
    assert rel.max() < 2.0 ** -22 + 16 * 2.0 ** -25, float(rel.max())


def test_packed_path_with_motion_blur_and_final_xform(native, built):
    """Packed accumulation through queue_frame (forced) vs the float4 frame."""
    from cuburn_b200 import samples, render, profile
    gnm = samples.g6f(animated=True)
    gprof = profile.wrap(dict(width=320, height=180, spp=300, fps=24, duration=1.0,
                              frame_width=2.0), gnm)
    frames = {}
    for mode in ('float4', 'packed'):
        rmgr = render.RenderManager(seed=9)
        rmgr.accumulate = mode
        rdr = render.Renderer(gnm, gprof)
        evt, buf = rmgr.queue_frame(rdr, gnm, gprof, 0.3)
        evt.synchronize()
        frames[mode] = np.array(buf)
    d = np.abs(frames['float4'].astype(int) - frames['packed'].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
