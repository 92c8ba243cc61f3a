"""Genome schema, accessors and codecs (host side)."""
import numpy as np
import pytest

from conftest import have_reference, REFERENCE
from cuburn_b200.genome import specs, spectypes, use, util, variations


def test_spline_normalize_forms():
    n = use.SplineEval.normalize
    a = n(1.5, 1.0)
    assert a.shape == (2, 4)
    assert list(a[0]) == [-2, 0, 1, 3] and list(a[1]) == [1.5] * 4
    b = n([2.0, 4.0], 1.0)
    assert list(b[0]) == [-2, 0, 1, 3] and list(b[1]) == [4.0, 2.0, 4.0, 2.0]
    # velocities scale with duration; guards make end tangents equal them
    c = n([45, -360, -135, 0, 0.3, 150], 2.0)
    assert list(c[0]) == [-2, 0, 0.3, 1, 3]
    assert c[1][0] == 150 - (0.3 + 2) * (-720)          # guard uses the *second* knot
    assert c[1][-1] == 150 + (3 - 0.3) * 0
    with pytest.raises(ValueError):
        n([1, 2, 3], 1.0)


def test_spline_eval_endpoints_and_tangents():
    sp = use.SplineEval([45, -360, -135, 10, 0.3, 150], 1.0)
    assert abs(sp(0.0) - 45) < 1e-9 and abs(sp(1.0) + 135) < 1e-9 and abs(sp(0.3) - 150) < 1e-9
    assert abs(sp(0.0, deriv=1) + 360) < 1e-6
    assert abs(sp(1.0 - 1e-12, deriv=1) - 10) < 1e-4
    num = (sp(0.5 + 1e-6) - sp(0.5 - 1e-6)) / 2e-6
    assert abs(sp(0.5, deriv=1) - num) < 1e-3


def test_wrapper_defaults_and_containers():
    g = {'type': 'animation', 'xforms': {'10': {'weight': 1}, '2': {'weight': 2}},
         'camera': {'scale': 0.5}}
    w = use.Wrapper(g)
    assert w.xforms.keys() == ['10', '2']                  # string sort (Q7)
    assert 'final_xform' not in w and 'xforms' in w
    with pytest.raises(KeyError):
        'bogus' in w
    assert w.camera.scale == 0.5 and w.camera.rotation is None
    assert w.time.duration == 1
    sw = use.SplineWrapper(g, scale=1)
    assert sw.xforms['2'].color_speed(0.1) == 0.5
    assert sw.xforms['2'].pre_affine.angle(0.5) == 45
    assert sw.camera.scale.interp == 'mag'


def test_palette_codec_roundtrip():
    rgb = np.ones((256, 4), np.float32)
    rgb[:, 0] = np.arange(256) / 255.0
    rgb[:, 1] = 1 / 255.0
    rgb[:, 2] = 2 / 255.0
    enc = util.palette_encode(rgb)
    assert enc[0] == 'rgb8' and all(len(c) <= 64 for c in enc[1:])
    # same prefix the reference's conversion test pins for (1,2,3),(255,255,255) data
    dec = util.palette_decode(enc)
    assert dec.shape == (256, 4) and dec.dtype == np.float32
    assert np.array_equal(np.round(dec[:, :3] * 255), np.round(rgb[:, :3] * 255))
    assert np.all(dec[:, 3] == 1)
    raw = bytes([1, 2, 3]) + bytes([255] * (255 * 3))
    import base64
    assert base64.b64encode(raw).decode()[:8] == 'AQID////'   # test_convert.py:68


def test_flatten_hash():
    a = {'x': {'y': 1, 'z': {'q': 2}}, 4: 5}
    assert util.flatten(a) == {'x.y': 1, 'x.z.q': 2, '4': 5}
    assert util.unflatten(util.flatten(a)) == {'x': {'y': 1, 'z': {'q': 2}}, '4': 5}
    g1 = {'xforms': {'0': {'weight': 1, 'color': 2}}}
    g2 = {'xforms': {'0': {'color': 0.5, 'weight': 3}}}
    g3 = {'xforms': {'0': {'weight': 1}}}
    assert util.hash(g1) == util.hash(g2) != util.hash(g3)


def test_json_encode_roundtrip():
    import json
    from cuburn_b200 import samples
    g = samples.g6f()
    text = util.json_encode(g)
    back = json.loads(text)
    assert back['xforms']['1']['variations']['julian']['power'] == 3
    assert set(back['xforms']) == set(g['xforms'])


def test_variation_table_complete():
    assert len(variations.var_params) == 95 and len(variations.var_names) == 95
    assert variations.var_names[0] == 'linear' and variations.var_names[98] == 'mobius'
    for n in (47, 78, 79, 96):
        assert n not in variations.var_names
    assert variations.var_params['julian']['power'].interp == 'mag'
    assert variations.var_params['oscope']['frequency'].default == np.pi
    assert all(sp.var for n, p in variations.var_params.items()
               for k, sp in p.items() if k != 'weight')


@pytest.mark.skipif(not have_reference(), reason='reference tree not mounted')
def test_schema_matches_reference():
    """Execute the reference's schema modules against our spectypes and compare."""
    def load(path, extra=None):
        src = open(REFERENCE + '/cuburn/genome/' + path).read()
        src = src.replace('from spectypes import', 'from cuburn_b200.genome.spectypes import')
        src = src.replace('from variations import', 'from cuburn_b200.genome.variations import')
        ns = dict(extra or {})
        ns['basestring'] = str
        exec(compile(src, path, 'exec'), ns)
        return ns
    rv = load('variations.py')
    assert rv['var_names'] == variations.var_names
    for name, params in rv['var_params'].items():
        assert dict(params) == dict(variations.var_params[name]), name
    rs = load('specs.py')

    # documented extension: the xaos map the reference's kernel reads (code/iter.py:32-54)
    # and its converter writes (genome/convert.py:182) but its schema never declared
    def extra(path):
        return {'chaos'} if 'xform' in path.split('.')[-1] or '.xforms.' in path else set()

    def same(a, b, path=''):
        if isinstance(a, dict):
            assert isinstance(b, dict) and set(a) == set(b) - extra(path), path
            for k in a:
                same(a[k], b[k], path + '.' + str(k))
        elif isinstance(a, spectypes.Map):
            same(a.type, b.type, path + '[]')
        elif isinstance(a, spectypes.List):
            assert list(a.default) == list(b.default), path
        elif isinstance(a, spectypes.Enum):
            pass        # we add output types the reference lacks (raw)
        elif isinstance(a, (spectypes.Spline, spectypes.Scalar, spectypes.RefScalar)):
            assert a[:len(a) - 1] == b[:len(b) - 1] or a == b, path
    for top in ('anim', 'node', 'edge', 'profile'):
        same(rs[top], getattr(specs, top), top)


def test_reference_generated_host_vectors(built):
    """
    tests/golden/host_golden.json comes from executing the reference's use.py,
    profile.py and mwc.make_seeds (tests/golden/make_host_golden.py).
    """
    import json
    import os
    from cuburn_b200 import profile, mwc, filters
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                           'host_golden.json')) as fp:
        gold = json.load(fp)
    for c in gold['splines']:
        se = use.SplineEval(c['value'], c['scale'])
        assert np.allclose(se.knots, np.array(c['knots']), rtol=0, atol=0)
        for t, v, dv in c['at']:
            assert abs(se(t) - v) <= 1e-9 * max(1, abs(v))
            assert abs(se(t, 1) - dv) <= 1e-7 * max(1, abs(dv))
    for c in gold['profile']:
        args = profile.add_args().parse_args(c['argv'])
        name, p = profile.get_from_args(args)
        assert name == c['name'] and p == c['profile']
        gprof = profile.wrap(dict(p), {'type': 'animation', 'camera': {'spp': 1.5},
                                       'time': {'duration': 2, 'frame_width': [1.0, 0.5]}})
        times = profile.enumerate_times(gprof)
        assert len(times) == c['ntimes']
        for (i, t), (gi, gt) in zip(times, c['times']):
            assert i == gi and np.allclose(list(t), gt, rtol=0, atol=1e-15)
        assert abs(gprof.spp(0.5) - c['spp']) < 1e-9
        assert abs(gprof.frame_width(0.25) - c['frame_width']) < 1e-12
        assert gprof.duration == c['duration'] and [gprof.width, gprof.height] == c['size']
    s = gold['seeds']
    seeds = mwc.make_seeds(s['n'], host_seed=s['host_seed'])
    assert seeds[:4].tolist() == s['first'] and seeds[-2:].tolist() == s['last']
    assert [int(np.bitwise_xor.reduce(seeds[:, k])) for k in range(3)] == s['xor']
    for stdev, coefs in gold['filters']['gauss'].items():
        assert [float(x) for x in filters.gauss_coefs(float(stdev))] == coefs

    class P(object):
        def __init__(self, g, t):
            self.gamma = lambda tc: g
            self.gamma_threshold = lambda tc: t
    for gamma, thr, gam, lin, lingam in gold['filters']['lingam']:
        got = filters.calc_lingam(P(gamma, thr), 0.5)
        assert [float(x) for x in got] == [gam, lin, lingam]
