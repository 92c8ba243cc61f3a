"""
flam3 XML -> cuburn node documents.

Drop-in for the reference module (cuburn/genome/convert.py): ``XMLGenomeParser``,
``XMLPaletteParser``, ``convert_affine``, ``flam3_to_node``,
``nodes_from_xml_path`` and the re-exports ``node_to_anim`` / ``edge_to_anim`` /
``to_json``.  The conversion table (which flam3 attribute feeds which genome key,
and the affine decomposition into angle / spread / magnitude / offset with the
y axis flipped) follows convert.py:107-234 and is pinned by the reference's
known-answer test (genome/tests/test_convert.py:43-68).
"""
import binascii
import warnings
import xml.parsers.expat

import numpy as np

from .variations import var_params
from . import util
from .blend import node_to_anim, edge_to_anim      # noqa: F401  (re-exported)
from .util import json_encode as to_json            # noqa: F401


class XMLGenomeParser(object):
    """Parse flam3 XML into a list of plain attribute dictionaries."""
    def __init__(self):
        self.flames = []
        self._flame = None
        self.parser = xml.parsers.expat.ParserCreate()
        self.parser.StartElementHandler = self.start_element
        self.parser.EndElementHandler = self.end_element

    def start_element(self, name, attrs):
        if name == 'flame':
            assert self._flame is None
            self._flame = dict(attrs)
            self._flame['xforms'] = []
            if attrs.get('palette'):
                pal = XMLPaletteParser.lookup(int(attrs['palette']))
            else:
                pal = np.ones((256, 4), dtype=np.float32)
            self._flame['palette'] = pal
        elif name == 'xform':
            attrs = dict(attrs)
            if 'color' in attrs:
                # flam3 files sometimes carry a second, unused colour value
                attrs['color'] = attrs['color'].strip().split()[0]
            self._flame['xforms'].append(attrs)
        elif name == 'finalxform':
            self._flame['finalxform'] = dict(attrs)
        elif name == 'color':
            idx = int(attrs['index'])
            self._flame['palette'][idx][:3] = [float(v) / 255.0 for v in attrs['rgb'].split()]
        elif name == 'symmetry':
            self._flame['symmetry'] = int(attrs['kind'])

    def end_element(self, name):
        if name == 'flame':
            self.flames.append(self._flame)
            self._flame = None

    @classmethod
    def parse(cls, src):
        p = cls()
        p.parser.Parse(src, True)
        return p.flames


class XMLPaletteParser(object):
    """flam3-palettes.xml: ``numbers`` and ``names`` -> (256, 4) arrays in [0, 1]."""
    _names, _numbers = None, None
    _locations = ['/usr/local/share/flam3/flam3-palettes.xml',
                  '/usr/share/flam3/flam3-palettes.xml']

    def __init__(self, src):
        self.names, self.numbers = {}, {}
        self.parser = xml.parsers.expat.ParserCreate()
        self.parser.StartElementHandler = self.start_element
        self.parser.Parse(src, True)

    def start_element(self, name, attrs):
        if name != 'palette':
            return
        data = binascii.a2b_hex(attrs['data'].replace('\n', '').replace(' ', ''))
        pal = np.frombuffer(data, 'u1').reshape((256, 4)) / 255.0
        if 'number' in attrs:
            self.numbers[int(attrs['number'])] = pal
        if 'name' in attrs:
            self.names[attrs['name']] = pal

    @classmethod
    def _load(cls):
        src = None
        for loc in cls._locations:
            try:
                with open(loc) as fp:
                    src = fp.read()
                break
            except IOError:
                pass
        if not src:
            raise IOError("Couldn't find a palettes XML file")
        parsed = cls(src)
        cls._names, cls._numbers = parsed.names, parsed.numbers

    @classmethod
    def lookup(cls, key, isname=False):
        if not cls._names:
            cls._load()
        return np.array(cls._names[key] if isname else cls._numbers[key])


def convert_affine(aff, animate=False):
    """
    'xx yx xy yy xo yo' -> {angle, spread, magnitude, offset}; the identity maps
    to None (key dropped).  cuburn's IFS y axis points the other way from
    flam3's, so every y component changes sign (convert.py:107-123).
    """
    xx, yx, xy, yy, xo, yo = vals = [float(v) for v in aff.split()]
    if vals == [1, 0, 0, 1, 0, 0]:
        return None
    yx, xy, yo = -yx, -xy, -yo
    x_ang = np.degrees(np.arctan2(yx, xx))
    y_ang = np.degrees(np.arctan2(yy, xy))
    spread = ((y_ang - x_ang) % 360) / 2
    angle = (x_ang + spread) % 360
    return dict(spread=float(spread), angle=float(angle),
                magnitude={'x': float(np.hypot(xx, yx)), 'y': float(np.hypot(xy, yy))},
                offset={'x': xo, 'y': yo})


def apply_structure(struct, src):
    """Rows are (dst, src_key, convert) or (dst, convert_from_whole_dict)."""
    out = {}
    for row in struct:
        if len(row) == 2:
            v = row[1](src)
        else:
            v = row[2](src[row[1]]) if row[1] in src else None
        if v is not None:
            out[row[0]] = v
    return out


def convert_vars(xf):
    out = {}
    for name, params in var_params.items():
        if name not in xf:
            continue
        rows = [('weight', name, float)]
        rows += [(p, name + '_' + p, float) for p in params if p != 'weight']
        out[name] = apply_structure(rows, xf)
    return out


def _chaos(s):
    return dict(enumerate(float(v) for v in s.split()))


xform_structure = (
    ('pre_affine', 'coefs', convert_affine),
    ('post_affine', 'post', convert_affine),
    ('color', 'color', float),
    ('color_speed', 'color_speed', float),
    ('opacity', 'opacity', float),
    ('weight', 'weight', float),
    ('chaos', 'chaos', _chaos),
    ('variations', convert_vars),
)


def convert_xform(xf):
    out = apply_structure(xform_structure, xf)
    # the deprecated 'symmetry' attribute doubles as colour speed and as the
    # "do not rotate" flag (convert.py:131-143)
    symm = float(xf.get('symmetry', 0))
    anim = xf.get('animate', symm <= 0)
    if 'symmetry' in xf:
        out.setdefault('color_speed', (1 - symm) / 2)
    if anim and 'pre_affine' in out:
        out['pre_affine']['angle'] = [out['pre_affine']['angle'], -360]
    return out


def make_symm_xforms(kind, offset):
    """Extra xforms implementing flam3's <symmetry kind=N> (convert.py:145-160)."""
    assert kind != 0, 'symmetry kind 0 means "choose at random" and cannot be converted'
    out = []

    def boring():
        return dict(color=1, color_speed=0, weight=1, variations={'linear': {'weight': 1}})
    if kind < 0:
        xf = boring()
        xf['pre_affine'] = dict(angle=135, spread=-45)
        out.append(xf)
        kind = -kind
    for i in range(1, kind):
        xf = boring()
        if kind >= 3:
            xf['color'] = (i - 1) / (kind - 2.0)
        xf['pre_affine'] = dict(angle=(45 + 360 * i / float(kind)) % 360, spread=-45)
        out.append(xf)
    return dict(enumerate(out, offset))


def convert_xforms(flame):
    xfs = dict(enumerate(convert_xform(x) for x in flame['xforms']))
    if 'symmetry' in flame:
        xfs.update(make_symm_xforms(flame['symmetry'], len(xfs)))
    return xfs


def _pair(v):
    return dict(zip('xy', (float(x) for x in v.split())))


def _de_minimum(d):
    if 'estimator_minimum' not in d:
        return None
    return float(d['estimator_minimum']) / float(d.get('estimator_radius', 11))


flame_structure = (
    ('author.name', 'nick', str),
    ('author.url', 'url', lambda s: 'http://' + str(s)),
    ('name', 'name', str),
    ('camera.center', 'center', _pair),
    ('camera.rotation', 'rotate', float),
    ('camera.dither_width', 'filter', float),
    # cuburn's scale is in output widths per IFS unit (convert.py:198-199)
    ('camera.scale', lambda d: float(d['scale']) / float(d['size'].split()[0])),
    ('filters.colorclip.gamma', 'gamma', float),
    ('filters.colorclip.gamma_threshold', 'gamma_threshold', float),
    ('filters.colorclip.highlight_power', 'highlight_power', float),
    ('filters.colorclip.vibrance', 'vibrancy', float),
    ('filters.de.curve', 'estimator_curve', float),
    ('filters.de.radius', 'estimator_radius', float),
    ('filters.de.minimum', _de_minimum),
    ('filters.logscale.brightness', 'brightness', float),
    ('palette', 'palette', util.palette_encode),
    ('xforms', convert_xforms),
    ('final_xform', 'finalxform', convert_xform),
)


def flam3_to_node(flame):
    n = util.unflatten(util.flatten(apply_structure(flame_structure, flame)))
    n['type'] = 'node'
    return n


def nodes_from_xml_path(path):
    """One-shot conversion of every flame in an XML file."""
    with open(path) as fp:
        flames = XMLGenomeParser.parse(fp.read())
    if len(flames) > 10:
        warnings.warn("Lot of flames in this file. Sure it's not a frame-based animation?")
    for flame in flames:
        yield flam3_to_node(flame)


if __name__ == '__main__':
    import sys
    print('\n\n'.join(to_json(n) for n in nodes_from_xml_path(sys.argv[1])))
