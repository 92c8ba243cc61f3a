"""
The oracle against everything the reference pins for this path that can be
checked on a CPU: the pixel-format assertions (code/tests/test_output.py:23-124),
host-f64 spline evaluation (use.py:174-185), and committed golden vectors.
"""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _seeds(n=262144, seed=13):
    from cuburn_b200 import mwc
    return mwc.make_seeds(n, host_seed=seed)


def _dims(w=640, h=360):
    from oracle import flame_ref as R
    return R.calc_dim(w, h)


def test_calc_dim_table():
    """SURVEY section 8 size table (render.py:80-89)."""
    from oracle import flame_ref as R
    for (w, h), (aw, ah, astride) in {(640, 360): (664, 384, 672), (1920, 1080): (1944, 1104, 1952),
                                      (3840, 2160): (3864, 2192, 3872),
                                      (7680, 4320): (7704, 4352, 7712)}.items():
        d = R.calc_dim(w, h)
        assert (d['aw'], d['ah'], d['astride']) == (aw, ah, astride)


def test_output_clamping(built):
    from oracle import output_ref as O
    d = _dims()
    for fill, luma in ((-1, 0), (5, 255)):
        ins = np.full((d['ah'], d['astride'], 4), fill, np.float32)
        outs = O.convert('yuv444p', ins, 640, 360, _seeds())[0].reshape(3, 360, 640)
        assert np.all(outs[0] == luma)
        assert np.all((outs[1] >= 127) & (outs[1] <= 128))
        assert np.all((outs[2] >= 127) & (outs[2] <= 128))


def test_output_yuv444p10(built):
    from oracle import output_ref as O
    d = _dims()
    ins = np.zeros((d['ah'], d['astride'], 4), np.float32)
    outs = O.convert('yuv444p10', ins, 640, 360, _seeds())[0].reshape(3, 360, 640)
    assert np.all(outs[0] == 0)
    assert np.all((510 < outs[1]) & (outs[1] < 513)) and np.all((510 < outs[2]) & (outs[2] < 513))
    ins[12, 12, :] = [0, 1, 0, 1]
    ins[13, 13, :] = [0, 1, 0, 1]
    outs = O.convert('yuv444p10', ins, 640, 360, _seeds())[0].reshape(3, 360, 640)
    assert outs[0, 0, 0] > 0 and outs[0, 1, 1] > 0
    assert outs[1, 0, 0] < 500 and outs[1, 1, 1] < 500


def test_output_yuv420p10(built):
    from oracle import output_ref as O
    d = _dims()
    w, h = 640, 360
    ins = np.zeros((d['ah'], d['astride'], 4), np.float32)
    ins[12, 12, :] = [0, 1, 0, 1]
    ins[14, 14, :] = [0, 1, 0, 1]
    ins[15, 15, :] = [1, 0, 0, 1]
    flat = O.convert('yuv420p10', ins, w, h, _seeds())[0]
    assert flat.size == w * h * 3 // 2
    luma = flat[:w * h].reshape(h, w)
    out_cr = flat[w * h: w * h + w * h // 4].reshape(h // 2, w // 2)
    assert luma[0, 0] > 0 and luma[1, 0] == 0 and luma[0, 1] == 0 and luma[1, 1] == 0
    assert luma[2, 2] > 0 and luma[3, 3] > 0
    assert 172 <= out_cr[0, 0] <= 174
    assert 511 <= out_cr[0, 1] <= 512 and 511 <= out_cr[1, 0] <= 512
    # chroma site (1,1): mean of a green and a red pixel, equal weights
    assert abs(int(out_cr[1, 1]) - 1023 * (0.5 + (-0.331264 - 0.168736) / 2)) < 2


def test_rgba_formats_and_studio_swing(built):
    from oracle import output_ref as O
    d = _dims(64, 32)
    ins = np.zeros((d['ah'], d['astride'], 4), np.float32)
    ins[12:, 12:] = [1.0, 0.5, 0.0, 1.0]
    o8 = O.convert('rgba_u8', ins, 64, 32, _seeds(4096))[0].reshape(32, 64, 4)
    assert np.all(o8[..., 0] == 255) and np.all((o8[..., 1] >= 127) & (o8[..., 1] <= 128))
    assert np.all(o8[..., 2] == 0) and np.all(o8[..., 3] == 255)
    o16 = O.convert('rgba_u16', ins, 64, 32, _seeds(4096))[0].reshape(32, 64, 4)
    assert np.all(o16[..., 0] == 65535) and np.all(np.abs(o16[..., 1].astype(int) - 32767) <= 1)
    p12 = O.convert('yuv444p12', ins * 0, 64, 32, _seeds(4096))[0].reshape(3, 32, 64)
    assert np.all(p12[0] == 256) and np.all(np.abs(p12[1].astype(int) - (256 + 1792)) <= 1)
    white = np.ones_like(ins)
    p12 = O.convert('yuv444p12', white, 64, 32, _seeds(4096))[0].reshape(3, 32, 64)
    assert np.all(p12[0] == 256 + 3504)


def test_spline_f32_vs_host_f64(built):
    """T4: oracle f32 Catmull-Rom vs SplineEval.__call__ on random knot sets."""
    from cuburn_b200.genome.use import SplineEval
    from oracle import flame_ref as R
    rs = np.random.RandomState(0)
    tt = np.linspace(0, 1, 257).astype(np.float32)
    for trial in range(40):
        spec = [rs.randn(), 3 * rs.randn(), rs.randn(), 3 * rs.randn()]
        for t in np.sort(rs.uniform(0.02, 0.98, rs.randint(0, 25))):
            spec += [float(np.round(t, 4)), float(rs.randn())]
        t, k = R.normalize_spline(spec, 1.3)
        se = SplineEval(spec, 1.3)
        n = se.knots.shape[1]
        assert np.array_equal(t[:n], se.knots[0].astype(np.float32))
        assert np.array_equal(k[:n], se.knots[1].astype(np.float32))
        got = R.catmull_rom(t, k, tt)
        want = np.array([se(float(x)) for x in tt])
        assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(k).max())


def test_mag_domain_properties(built):
    from oracle import flame_ref as R
    tt = np.linspace(0, 1, 101).astype(np.float32)
    # constant stays constant; endpoints are hit; positive data stays positive
    t, k = R.normalize_spline(0.37, 1.0)
    assert np.allclose(R.catmull_rom(t, k, tt, mag=True), 0.37, rtol=1e-6)
    t, k = R.normalize_spline([0.01, 100.0], 1.0)
    v = R.catmull_rom(t, k, tt, mag=True)
    assert abs(v[0] - 0.01) < 1e-7 and abs(v[-1] / 100.0 - 1) < 1e-5 and np.all(v > 0)
    # geometric mid-point in the log domain (both ends above the 2^-4 elbow)
    t, k = R.normalize_spline([0.25, 4.0], 1.0)
    assert abs(R.catmull_rom(t, k, tt, mag=True)[50] - 1.0) < 1e-5
    # sign changes pass through the linear zone without NaNs
    t, k = R.normalize_spline([-2.0, 3.0], 1.0)
    v = R.catmull_rom(t, k, tt, mag=True)
    assert np.all(np.isfinite(v)) and v[0] == -2.0 and abs(v[-1] - 3.0) < 1e-6
    # det functions are accurate to float rounding
    x = np.float32(np.random.RandomState(1).uniform(-40, 40, 5000))
    s, c = R.det_sincosf(x)
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 6e-8
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 6e-8


def test_precalc_identities(built):
    """Defaults give an identity affine; densities end below 1; camera centres the frame."""
    from cuburn_b200 import samples
    from oracle import flame_ref as R
    g = samples.g3()
    g['xforms']['0']['pre_affine'] = {}
    ev = R.GenomeEval(g, 640, 360, 0.5, 0.0)
    v = ev.values
    assert abs(v['xforms.0.pre_affine.xx'][0] - 1) < 1e-7 and abs(v['xforms.0.pre_affine.yy'][0] - 1) < 1e-7
    assert abs(v['xforms.0.pre_affine.xy'][0]) < 1e-7 and abs(v['xforms.0.pre_affine.yx'][0]) < 1e-7
    assert abs(v['xforms.0.density'][0] - 1 / 3) < 1e-6 and abs(v['xforms.1.density'][0] - 2 / 3) < 1e-6
    assert 'xforms.2.density' not in v
    cx, cy = g['camera']['center']['x'], g['camera']['center']['y']
    px = v['camera.xx'][0] * cx + v['camera.xy'][0] * cy + v['camera.xo'][0]
    py = v['camera.yx'][0] * cx + v['camera.yy'][0] * cy + v['camera.yo'][0]
    assert abs(px - ev.dim['aw'] / 2) < 1e-3 and abs(py - ev.dim['ah'] / 2) < 1e-3
    assert abs(v['camera.xx'][0] - g['camera']['scale'] * 640) < 1e-3


def test_golden_vectors(built):
    """Committed oracle outputs (tests/golden/make_golden.py) still reproduce."""
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    with open(os.path.join(GOLDEN, 'oracle_golden.json')) as fp:
        gold = json.load(fp)
    for case in gold['params']:
        g = samples.GENOMES[case['genome']](**case.get('kwargs', {}))
        ev = R.GenomeEval(g, case['w'], case['h'], case['tc'], case['td'])
        for name, idx_vals in case['values'].items():
            for idx, bits in idx_vals:
                assert int(ev.values[name][idx].view(np.uint32)) == bits, (case['genome'], name, idx)
    pc = gold['palette']
    g = samples.GENOMES[pc['genome']]()
    pal, _ = R.palette_table(g, pc['ts'], pc['td'], mwc.make_seeds(16384, host_seed=pc['seed']))
    for r, c, rgba in pc['entries']:
        assert [int(x) for x in np.round(pal[r, c, :3] * 255)] == rgba
    hc = gold['chaos']
    g = samples.GENOMES[hc['genome']]()
    ev = R.GenomeEval(g, hc['w'], hc['h'], hc['tc'], 0.0)
    seeds = mwc.make_seeds(16384 + 64, host_seed=hc['seed'])
    pal, seeds = R.palette_table(g, ev.ts, 0.0, seeds)
    hist, _ = R.iterate(ev, pal, seeds, hc['nsamples'], ntraj=64, nthreads=1)
    assert int(hist[..., 3].sum()) == hc['inside']
    assert int(np.argmax(hist[..., 3])) == hc['argmax'] and int(hist[..., 3].max()) == hc['max']


def test_chaos_single_thread_is_deterministic(built):
    from cuburn_b200 import samples, mwc
    from oracle import flame_ref as R
    g = samples.g6f()
    ev = R.GenomeEval(g, 160, 90, 0.5, 0.0)
    seeds = mwc.make_seeds(16384 + 32, host_seed=3)
    pal, seeds = R.palette_table(g, ev.ts, 0.0, seeds)
    a, sa = R.iterate(ev, pal, seeds, 200000, ntraj=32, nthreads=1)
    b, sb = R.iterate(ev, pal, seeds, 200000, ntraj=32, nthreads=1)
    assert np.array_equal(a, b) and np.array_equal(sa, sb)
    c, _ = R.iterate(ev, pal, seeds, 200000, ntraj=32, nthreads=4)
    assert np.array_equal(a[..., 3], c[..., 3])           # counts are order independent
    assert np.allclose(a, c, rtol=1e-5, atol=1e-4)


def test_bench_cpu_filter_baseline_reports_every_stage():
    """bench.py's CPU arm for the filter chain (the numpy restatement, timed per filter)."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    r = bench.cpu_filter_rates(w=96, h=54)
    assert set(r['value']) == {'yuv_to_rgb', 'bilateral (8 directions)', 'logscale', 'smearclip',
                               'chain'}
    assert all(v > 0 for v in r['value'].values()) and r['kind'] == 'port'
    assert r['cores'] == (os.cpu_count() or 1)
