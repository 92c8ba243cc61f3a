"""
Device interpolation vs the oracle: packed parameters for all 1024 temporal
samples and the dithered palette table must match BIT FOR BIT (north star:
"genome interpolation to packed params ... bit-exact").
"""
import numpy as np
import pytest

from helpers import still_profile, frame_window, bits

pytestmark = pytest.mark.gpu


def _packed(N, gnm, w, h, tc, td, seed=3):
    from cuburn_b200 import render
    from cuburn_b200 import profile
    gprof, _ = still_profile(gnm, w, h, 10)
    rmgr = render.RenderManager(seed=seed)
    rdr = render.Renderer(gnm, gprof)
    dim = rmgr.fb.set_dim(w, h)
    rmgr._copy(rdr, gnm)
    rmgr._interp(rdr, gnm, dim, tc - 0.5 * td, td)
    rmgr.stream_a.synchronize()
    pk = rdr.packer
    params = N.from_device(rmgr.info_a.d_params, (1024, pk.param_stride), np.float32)
    pal = N.from_device(rmgr.info_a.d_palette, (64, 256, 4), np.float32)
    seeds = N.from_device(rmgr.fb.d_seeds, (rmgr.fb.nstreams, 3), np.uint32)
    return pk, params, pal, seeds


def _animated_g24h():
    from cuburn_b200 import samples
    g = samples.g24h()
    # animate a mix of linear- and magnitude-domain parameters, with interior knots
    g['xforms']['3']['pre_affine']['angle'] = [10, 400, 370, -200, 0.25, 90, 0.6, 200]
    g['xforms']['3']['pre_affine']['magnitude']['x'] = [0.4, 0.5, 1.6, -0.3, 0.5, 0.02]
    g['xforms']['5']['weight'] = [0.5, 0, 2.5, 0]
    jx = [k for k, x in g['xforms'].items() if 'julian' in x['variations']][0]
    g['xforms'][jx]['variations']['julian']['power'] = [5, 0, 2, 0, 0.5, 3.3]
    g['camera']['rotation'] = [0, 720, 360, 0]
    g['camera']['scale'] = [0.22, 0.1, 0.4, 0]
    g['camera']['center'] = {'x': [0, 0.3], 'y': [0.1, -0.5, -0.2, 0.4, 0.35, 0.0]}
    return g


@pytest.mark.parametrize('case', ['G3', 'G6F', 'G6F-anim', 'G24H-anim', 'precalc-vars'])
def test_packed_params_bit_exact(native, built, case):
    from cuburn_b200 import samples
    from oracle import flame_ref as R
    if case == 'G3':
        g, tc, td = samples.g3(), 1.5 / 720, 0.0
    elif case == 'G6F':
        g, tc, td = samples.g6f(), 1.5 / 720, 0.0
    elif case == 'G6F-anim':
        g, tc, td = samples.g6f(animated=True), 0.37, 1.0 / 720
    elif case == 'G24H-anim':
        g, tc, td = _animated_g24h(), 0.61, 0.9      # a very wide shutter: all knots crossed
    else:
        g = samples.g3()
        g['xforms']['0']['variations'].update({
            'waves': {'weight': 0.1}, 'perspective': {'weight': 0.2, 'angle': [0.3, 1.5], 'dist': 2.5},
            'curve': {'weight': 0.1, 'xamp': 0.2, 'yamp': 0.1, 'xlength': [0.5, 2.0], 'ylength': 1e-12},
            'juliascope': {'weight': 0.3, 'power': [2, 0, -3, 0], 'dist': 0.7}})
        g['final_xform'] = {'variations': {'julian': {'weight': 1, 'power': 4, 'dist': 0.01}},
                            'post_affine': {'angle': 12, 'offset': {'x': [0.1, -0.1]}}}
        tc, td = 0.5, 0.5
    w, h = 1920, 1080
    pk, params, pal, seeds = _packed(native, g, w, h, tc, td)
    ev = R.GenomeEval(g, w, h, tc, td)
    named = pk.named_slots()                         # alignment padding carries no value
    assert set(n for _, n in named) <= set(ev.values)      # the oracle also evaluates precalc inputs
    bad = [n for i, n in named if not np.array_equal(bits(params[:, i]), bits(ev.values[n]))]
    assert not bad, bad[:10]
    assert np.all(np.isfinite(params[:, [i for i, _ in named]]))


@pytest.mark.parametrize('npal', [1, 3])
def test_palette_table_bit_exact(native, built, npal):
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    g = samples.g6f()
    if npal == 3:
        g['palette'] = [[0.0] + samples.make_palette('spectrum'),
                        [0.5] + samples.make_palette('fire'),
                        [1.0] + samples.make_palette('ocean')]
    tc, td = 0.45, 0.2
    pk, params, pal, seeds_after = _packed(native, g, 640, 360, tc, td, seed=3)
    seeds0 = mwc.make_seeds(262144, host_seed=3)
    opal, oseeds = R.palette_table(g, tc - 0.5 * td, td, seeds0)
    assert np.array_equal(opal, pal)
    # values are 8-bit levels / 255 with unit density
    assert np.all(pal[..., 3] == 1)
    lv = pal[..., :3] * 255
    assert np.all(np.abs(lv - np.round(lv)) < 1e-4) and lv.min() >= 0 and lv.max() <= 255
    # the RNG streams advanced exactly as the oracle's did
    assert np.array_equal(seeds_after[:16384], oseeds[:16384])
    assert np.array_equal(seeds_after[16384:], seeds0[16384:])
    if npal == 3:
        assert np.abs(pal[0] - pal[63]).max() > 0.05      # rows differ across the shutter


def test_spline_rows_match_host_f64(native, built):
    """T4: device f32 Catmull-Rom vs the host float64 SplineEval on random knot sets."""
    N = native
    from cuburn_b200.genome.use import SplineEval
    rs = np.random.RandomState(5)
    nrows = 16
    times = np.full((nrows, 32), 1e9, np.float32)
    knots = np.zeros((nrows, 32), np.float32)
    evals = []
    for r in range(nrows):
        nk = rs.randint(0, 20)
        spec = [rs.randn(), rs.randn(), rs.randn(), rs.randn()]
        for t in np.sort(rs.uniform(0.02, 0.98, nk)):
            spec += [float(np.round(t, 4)), float(rs.randn())]
        se = SplineEval(spec, 1.0)
        n = se.knots.shape[1]
        times[r, :n], knots[r, :n] = se.knots[0], se.knots[1]
        evals.append(se)
    d_t, d_k = N.to_device(times), N.to_device(knots)
    d_mag = N.to_device(np.zeros(nrows, np.int32))
    d_out = N.DeviceBuffer(4 * nrows * 1024)
    N.check(N.lib().cb_interp_rows(d_out.ptr, d_t.ptr, d_k.ptr, d_mag.ptr, nrows,
                                   np.float32(0.0), np.float32(1.0 / 1024), 1024, None))
    got = N.from_device(d_out, (1024, nrows), np.float32)
    tt = (np.arange(1024) * np.float64(np.float32(1.0 / 1024))).astype(np.float32)
    for r, se in enumerate(evals):
        want = np.array([se(float(t)) for t in tt])
        scale = max(1.0, np.abs(se.knots[1]).max())
        assert np.abs(got[:, r] - want).max() < 2e-5 * scale * 8, r
