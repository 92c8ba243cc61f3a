"""
flam3 XML -> node -> animation.  The first two tests are the reference's own
(cuburn/genome/tests/test_convert.py:35-68) re-expressed for Python 3; the rest pin
the blending rules (periodic extension, padding, xform pairing).
"""
import binascii

import numpy as np
import pytest

from cuburn_b200.genome import convert, blend, db, specs
from cuburn_b200.genome.use import SplineEval


def _make_palette_src():
    values = np.zeros((256, 4), 'u1')
    values[:, 0] = range(256)
    values[:, 1] = 1
    values[:, 2] = 2
    values[:, 3] = 3
    # leave a newline in to make sure those get stripped
    return """<palettes><palette number="0" name="synthetic" data="%s
"/></palettes>""" % binascii.b2a_hex(values.tobytes()).decode()


def _make_genome_src(extra='', xforms=None):
    xforms = xforms or ('<xform weight="0.1" color="0" hyperbolic="0.1" '
                        'coefs="01 0.2 -0.3 0.4 -0.5 0.6"/>')
    return """
<flame time="0" size="1280 960" center="0.01 0.02" scale="40" oversample="2"
    filter="1" quality="500" batches="50" brightness="4" gamma="4"
    url="test.com" nick="strobe" %s>
    <color index="0" rgb="1 2 3"/>
    %s
</flame>""" % (extra, xforms)


def test_palette_parse():
    parser = convert.XMLPaletteParser(_make_palette_src())
    assert 'synthetic' in parser.names and 0 in parser.numbers
    assert [0, 1 / 255., 2 / 255., 3 / 255.] == list(parser.numbers[0][0])
    assert [1, 1 / 255., 2 / 255., 3 / 255.] == list(parser.numbers[0][255])


def test_flam3_to_node_known_answer():
    parsed = convert.XMLGenomeParser.parse(_make_genome_src())
    converted = convert.flam3_to_node(parsed[0])
    palette = converted.pop('palette')
    assert converted == dict(
        type='node',
        author=dict(url='http://test.com', name='strobe'),
        camera=dict(dither_width=1.0, scale=0.03125, center=dict(x=0.01, y=0.02)),
        filters=dict(logscale=dict(brightness=4.0), colorclip=dict(gamma=4.0)),
        xforms={'0': dict(
            color=0.0,
            variations=dict(hyperbolic=dict(weight=0.1)),
            pre_affine=dict(spread=32.220017414088105,
                            angle=[20.91008494006789, -360],
                            magnitude=dict(x=1.019803902718557, y=0.5),
                            offset=dict(x=-0.5, y=-0.6)),
            weight=0.1)})
    assert palette[0] == 'rgb8' and palette[1][:8] == 'AQID////'


def test_affine_decomposition_roundtrips_through_the_device_formula():
    """convert_affine inverts precalc_xf_affine (code/iter.py:81-95) up to the y flip."""
    rs = np.random.RandomState(2)
    for _ in range(50):
        xx, yx, xy, yy, xo, yo = (float(v) for v in rs.uniform(-1.5, 1.5, 6))
        a = convert.convert_affine('%r %r %r %r %r %r' % (xx, yx, xy, yy, xo, yo))
        pri, spr = np.radians(a['angle']), np.radians(a['spread'])
        mx, my = a['magnitude']['x'], a['magnitude']['y']
        got = (mx * np.cos(pri - spr), -mx * np.sin(pri - spr), -my * np.cos(pri + spr),
               my * np.sin(pri + spr), a['offset']['x'], -a['offset']['y'])
        # the y flip of the conversion and the one in the device formula cancel:
        # the device ends up with the flam3 coefficients
        assert np.allclose(got, (xx, yx, xy, yy, xo, yo), atol=1e-9)
    assert convert.convert_affine('1 0 0 1 0 0') is None


def test_symmetry_and_final_xform():
    src = _make_genome_src(xforms='<xform weight="1" color="0.5" symmetry="0.5" linear="1" coefs="1 0 0 1 0 0"/>'
                                  '<symmetry kind="3"/><finalxform color="0" linear="1" julian="0.5" '
                                  'julian_power="3" julian_dist="1.5" coefs="0.5 0 0 0.5 0 0"/>')
    node = convert.flam3_to_node(convert.XMLGenomeParser.parse(src)[0])
    assert sorted(node['xforms']) == ['0', '1', '2']
    assert node['xforms']['0']['color_speed'] == 0.25 and 'pre_affine' not in node['xforms']['0']
    assert node['xforms']['1']['pre_affine'] == dict(angle=165.0, spread=-45)
    assert node['xforms']['2']['color'] == 1.0 and node['xforms']['1']['color'] == 0.0
    fx = node['final_xform']
    assert fx['variations']['julian'] == dict(weight=0.5, power=3.0, dist=1.5)
    assert fx['pre_affine']['magnitude'] == dict(x=0.5, y=0.5)


def test_periodic_extension():
    ang = specs.affine['angle']
    # one full clockwise turn per unit time: the destination is extended by -360
    assert blend.tospline(ang, [20.0, -360], [20.0, -360], None, 1) == [20.0, -360, -340.0, -360]
    # two turns when the duration doubles
    assert blend.tospline(ang, [20.0, -360], [20.0, -360], None, 2) == [20.0, -360, -700.0, -360]
    # no velocity: the nearest congruent destination is taken (350 -> -10)
    out = blend.tospline(ang, 10, 350, None, 1)
    assert out[0] == 10 and abs(out[1] + 10) < 1e-9
    assert blend.tospline(ang, 45, 45, None, 1) == 45
    # endpoint override snaps to the nearest congruent value
    out = blend.tospline(ang, [0.0, 90], [90.0, 90], [1, 450], 1)
    assert out[:4] == [0.0, 90, 450.0, 90]
    # interior knots ride along; None knots are dropped
    w = specs.xform['weight']
    assert blend.tospline(w, 1.0, 2.0, [0.5, 3.0, 0.25, None], 1) == [1.0, 0, 2.0, 0, 0.5, 3.0]
    # variation parameters copy the other side when missing
    jp = specs.xform['variations']['julian']['power']
    assert blend.tospline(jp, None, 5, None, 1) == 5
    assert blend.tospline(w, None, 5, None, 1) == [0, 5]


def test_node_to_anim_loops_seamlessly():
    node = convert.flam3_to_node(convert.XMLGenomeParser.parse(_make_genome_src())[0])
    anim = convert.node_to_anim(None, node, half=False)
    assert anim['type'] == 'animation' and anim['time']['duration'] == 1
    assert list(anim['xforms']) == ['0_0']
    a = anim['xforms']['0_0']['pre_affine']['angle']
    assert a[1] == a[3] == -360 and abs((a[0] - a[2]) - 360) < 1e-9
    se = SplineEval(a, 1.0)
    assert abs((se(0.0) - se(1.0)) - 360) < 1e-9 and abs(se(0.5) - (a[0] - 180)) < 1e-6
    assert [p[0] for p in anim['palette']] == [0, 1]
    half = convert.node_to_anim(None, node, half=True)
    h = half['xforms']['0_0']['pre_affine']['angle']
    assert half['time']['duration'] == 0.5 and abs((h[0] - h[2]) - 180) < 1e-9


def test_blend_pads_and_pairs_xforms():
    src = {'type': 'node', 'xforms': {
        '0': {'weight': 1, 'variations': {'linear': {'weight': 1}}},
        '1': {'weight': 3, 'variations': {'spherical': {'weight': 1}}}}}
    dst = {'type': 'node', 'xforms': {
        'a': {'weight': 2, 'variations': {'blob': {'weight': 1, 'low': 0.3}}}}}
    anim = blend.blend(src, dst, {'blend': {'xform_sort': 'weightflip'}})
    # weightflip: src ascending by weight vs dst descending
    assert sorted(anim['xforms']) == ['0_a', '1_pad']
    x = anim['xforms']['0_a']
    assert x['weight'] == [1, 2]
    assert x['variations']['linear']['weight'] == [1, 0]
    assert x['variations']['blob']['low'] == 0.3              # copied, not defaulted
    pad = anim['xforms']['1_pad']
    # spherical is a "hole" variation: its padding partner is the inverted identity
    assert pad['variations']['linear']['weight'] == [0, -1]
    assert pad['pre_affine']['angle'] == [45, 225]
    assert pad['weight'] == [3, 0]
    nat = blend.blend(src, dst, {'blend': {'xform_sort': 'natural', 'xform_map': [['1', 'a']]}})
    assert sorted(nat['xforms']) == ['0_pad', '1_a']


def test_resolve_bases_and_db(tmp_path):
    import json
    base = {'type': 'node', 'camera': {'scale': 0.5}, 'xforms': {'0': {'weight': 1}}}
    child = {'type': 'node', 'base': 'base', 'camera': {'rotation': [10, 5]}}
    (tmp_path / 'base.json').write_text(json.dumps(base))
    (tmp_path / 'child.json').write_text(json.dumps(child))
    gdb = db.connect(str(tmp_path))
    merged = blend.resolve(gdb, gdb.get('child'))
    assert merged['camera'] == {'scale': 0.5, 'rotation': [10, 5]} and '0' in merged['xforms']
    anim, name = gdb.get_anim('child')
    assert name == 'child' and anim['type'] == 'animation'
    assert anim['camera']['rotation'] == [10, 5, 15, 5]
    (tmp_path / 'one.json').write_text(json.dumps({'type': 'onefiledb', 'x': base}))
    one = db.connect(str(tmp_path / 'one.json'))
    assert one.get_anim('x')[0]['type'] == 'animation'
    xml = tmp_path / 'f.flam3'
    xml.write_text(_make_genome_src())
    anim, name = gdb.get_anim(str(xml), half=True)
    assert name == 'f' and anim['time']['duration'] == 0.5


def test_converted_flame_packs_and_compiles(built):
    """A converted + blended flam3 file goes through the packer and NVRTC."""
    from cuburn_b200 import _native as N
    from cuburn_b200.code import itergen
    src = _make_genome_src(xforms='<xform weight="0.5" color="0" linear="1" coefs="0.5 0 0 0.5 0.3 0"/>'
                                  '<xform weight="0.5" color="1" spherical="0.7" linear="0.2" '
                                  'coefs="0.4 0.2 -0.2 0.4 -0.3 0.1" post="0.9 0 0 0.9 0 0.1"/>')
    node = convert.flam3_to_node(convert.XMLGenomeParser.parse(src)[0])
    anim = convert.node_to_anim(None, node, half=False)
    pk, source = itergen.mkiterlib(anim)
    assert pk.xform_ids == ['0_0', '1_1'] and not pk.has_final
    times, knots = pk.pack(anim)
    row = pk.row_paths.index(('xforms', '0_0', 'pre_affine', 'angle'))
    assert times[row, 0] == -2 and times[row, 3] == 3
    hn, hs = itergen.load_headers()
    assert len(N.Module(source, 'conv.cu', hs, hn, itergen.NVRTC_OPTIONS).cubin) > 4096


def test_reference_generated_golden_vectors():
    """
    tests/golden/blend_golden.json was produced by executing the reference's own
    convert.py / blend.py (tests/golden/make_blend_golden.py); our modules must
    reproduce every document exactly.
    """
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'blend_golden.json')
    with open(path) as fp:
        gold = json.load(fp)

    def canon(x):
        return json.loads(json.dumps(x, sort_keys=True))

    def fresh(x):
        return json.loads(json.dumps(x))
    assert len(gold['flam3_to_node']) >= 2 and len(gold['blend']) >= 10
    for c in gold['flam3_to_node']:
        assert canon(convert.flam3_to_node(convert.XMLGenomeParser.parse(c['xml'])[0])) == c['node']
    for c in gold['node_to_anim']:
        assert canon(blend.node_to_anim(None, fresh(c['node']), c['half'])) == c['anim']
    for c in gold['blend']:
        got = blend.blend(fresh(gold['nodes'][c['src']]), fresh(gold['nodes'][c['dst']]),
                          fresh(c['edge']))
        assert canon(got) == c['anim'], (c['src'], c['dst'], c['edge'])


def test_flam3_chaos_and_opacity_reach_the_iterate_module(built):
    """A flam3 file with per-xform `chaos` and `opacity` attributes: the converter keeps
    them (convert.py:180-183), the packer turns them into xaos density slots and an opacity
    slot, and the generated module (choice chain per previous xform, visibility draw)
    compiles for sm_100a."""
    from cuburn_b200 import _native as N
    from cuburn_b200.code import itergen
    xforms = ('<xform weight="0.5" color="0" linear="1" coefs="0.5 0 0 0.5 0.3 0" '
              'chaos="1 0.25" opacity="0.5"/>'
              '<xform weight="0.5" color="1" spherical="0.4" linear="0.6" coefs="0.6 0 0 0.6 -0.3 0.1" '
              'chaos="2 1"/>')
    flame = convert.XMLGenomeParser.parse(_make_genome_src(xforms=xforms))[0]
    node = convert.flam3_to_node(flame)
    assert node['xforms']['0']['chaos'] == {'0': 1.0, '1': 0.25}
    assert node['xforms']['0']['opacity'] == 0.5
    anim = blend.node_to_anim(None, node, half=False)
    # the blender carries the tables over, keyed by the blended xform names
    assert anim['xforms']['0_0']['chaos'] == {'0_0': 1.0, '1_1': 0.25}
    assert anim['xforms']['1_1']['chaos'] == {'0_0': 2.0, '1_1': 1.0}
    pk, src = itergen.mkiterlib(anim)
    assert pk.xaos and len(pk.opacity) == 1
    for p in pk.xform_ids:
        for n in pk.xform_ids[:-1]:
            pk.slot('xforms', p, 'chaos_den', n)
    assert 'switch (pt.last[0])' in src and 'opacity_visible' in src
    names, hdrs = itergen.load_headers()
    assert N.Module(src, 'xaos.cu', hdrs, names, itergen.NVRTC_OPTIONS).cubin[:4] == b'\x7fELF'
