#!/usr/bin/env python
"""
Generate tests/golden/blend_golden.json by EXECUTING the reference's own
cuburn/genome/convert.py and blend.py (Python 2 sources, patched in memory just
enough to parse under Python 3: izip_longest, dict.keys() + dict.keys(), the
bare-tuple comprehension, np.fromstring, print statements) against this
package's schema modules.  Only runs where /root/reference is mounted; the
resulting vectors are committed and checked by tests/test_convert.py.

    python tests/golden/make_blend_golden.py
"""
import json
import os
import re
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = '/root/reference/cuburn/genome/'

from cuburn_b200.genome import spectypes, specs, variations, use, util   # noqa: E402


def load_reference():
    mods = {'spectypes': spectypes, 'specs': specs, 'variations': variations,
            'use': use, 'util': util}
    saved = {k: sys.modules.get(k) for k in list(mods) + ['blend']}
    sys.modules.update(mods)
    try:
        src = open(REF + 'blend.py').read()
        src = src.replace('from itertools import izip_longest',
                          'from itertools import zip_longest as izip_longest')
        src = src.replace('items = map(flatten, go(item))', 'items = list(map(flatten, go(item)))')
        src = src.replace('set(av.keys() + bv.keys())', 'set(list(av) + list(bv))')
        src = src.replace('[x or {} for x in src, dst, edit]', '[x or {} for x in (src, dst, edit)]')
        src = src.replace('set(src.keys() + dst.keys() + edit.keys())',
                          'set(list(src) + list(dst) + list(edit))')
        src = src.replace('set(scl.keys() + dcl.keys())', 'set(list(scl) + list(dcl))')
        src = re.sub(r'def trace\(k, cond=True\):\n    print k,\n    return k\n', '', src)
        src = src[:src.index('if __name__ == "__main__":')]
        bl = types.ModuleType('blend')
        exec(compile(src, REF + 'blend.py', 'exec'), bl.__dict__)
        sys.modules['blend'] = bl

        src = open(REF + 'convert.py').read()
        src = src.replace("np.fromstring(data, 'u1')", "np.frombuffer(data, 'u1')")
        src = src.replace('xx, yx, xy, yy, xo, yo = vals = map(float, aff.split())',
                          'xx, yx, xy, yy, xo, yo = vals = list(map(float, aff.split()))')
        src = src[:src.index('if __name__ == "__main__":')]
        cv = types.ModuleType('convert')
        exec(compile(src, REF + 'convert.py', 'exec'), cv.__dict__)
        return cv, bl
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


XML = [
    """<flame size="1280 960" center="0.01 0.02" scale="40" brightness="4" gamma="4" nick="a" url="b.c">
    <color index="0" rgb="1 2 3"/>
    <xform weight="0.1" color="0" hyperbolic="0.1" coefs="1 0.2 -0.3 0.4 -0.5 0.6"/></flame>""",
    """<flame size="640 480" center="-0.3 0.25" scale="96" rotate="35" vibrancy="0.8" gamma_threshold="0.02"
       highlight_power="1.5" estimator_radius="9" estimator_minimum="1" estimator_curve="0.4" brightness="12">
    <symmetry kind="-3"/>
    <xform weight="0.5" color="0.2" symmetry="0.3" linear="0.7" julian="0.4" julian_power="-3" julian_dist="1.2"
       coefs="0.6 -0.2 0.3 0.5 0.1 -0.2" post="0.9 0.1 -0.1 0.9 0 0.2" opacity="0.5"/>
    <xform weight="1.5" color="0.9" animate="0" spherical="0.3" blob="0.5" blob_low="0.2" blob_high="1.1" blob_waves="4"
       coefs="-0.4 0.3 0.35 0.45 -0.6 0.4"/>
    <finalxform color="0" color_speed="0" linear="1" eyefish="0.2" coefs="1 0 0 1 0 0"/></flame>""",
]

NODES = {
    'a': {'type': 'node', 'camera': {'scale': 0.4, 'rotation': [30, 90], 'center': {'x': [0.1, 0.5]}},
          'blend': {'duration': 3},
          'xforms': {'0': {'weight': 1, 'color': 0.1, 'pre_affine': {'angle': [10, -360], 'spread': 100},
                           'variations': {'linear': {'weight': 1}}},
                     '1': {'weight': 3, 'color': 0.5,
                           'variations': {'spherical': {'weight': 0.6}, 'perspective': {'weight': 0.2, 'angle': [0.3, 4]}}},
                     '2': {'weight': 2, 'variations': {'blob': {'weight': 1, 'low': 0.3}}}},
          'palette': ['rgb8', 'AAAA']},
    'b': {'type': 'node', 'camera': {'scale': 0.9, 'rotation': [300, -45]},
          'xforms': {'x': {'weight': 2.5, 'color': 0.9, 'pre_affine': {'angle': [200, 180]},
                           'variations': {'linear': {'weight': 0.3}, 'julian': {'weight': 0.7, 'power': 4}}},
                     'y': {'weight': 0.5, 'post_affine': {'spread': 120},
                           'variations': {'rectangles': {'weight': 1, 'x': 0.2}}}},
          'final_xform': {'variations': {'linear': {'weight': 1}}},
          'palette': ['rgb8', 'BBBB']},
}
EDGES = [
    {},
    {'blend': {'xform_sort': 'natural', 'duration': 0.5}},
    {'blend': {'xform_sort': 'color'}, 'camera': {'rotation': [0.5, 100, 1, 700]},
     'xforms': {'src': {'1': {'weight': [0.3, 9]}}, 'dst': {'x': {'color': [0.7, 0.2]}}}},
    {'blend': {'xform_sort': 'weight', 'xform_map': [['2', 'y'], ['dup', 'x']]}},
]


def canon(x):
    return json.loads(json.dumps(x, sort_keys=True))


def main():
    cv, bl = load_reference()
    out = {'flam3_to_node': [], 'node_to_anim': [], 'blend': []}
    for xml in XML:
        node = cv.flam3_to_node(cv.XMLGenomeParser.parse(xml)[0])
        out['flam3_to_node'].append({'xml': xml, 'node': canon(node)})
        for half in (False, True):
            out['node_to_anim'].append({'node': canon(node), 'half': half,
                                        'anim': canon(bl.node_to_anim(None, node, half))})
    for edge in EDGES:
        pairs = (('a', 'b'),) if 'xform_map' in edge.get('blend', {}) else \
            (('a', 'b'), ('b', 'a'), ('a', 'a'))
        for s, d in pairs:
            anim = bl.blend(json.loads(json.dumps(NODES[s])), json.loads(json.dumps(NODES[d])),
                            json.loads(json.dumps(edge)))
            out['blend'].append({'src': s, 'dst': d, 'edge': edge, 'anim': canon(anim)})
    out['nodes'] = NODES
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'blend_golden.json')
    with open(path, 'w') as fp:
        json.dump(out, fp, indent=1, sort_keys=True)
    print('wrote %s: %d conversions, %d loops, %d blends' % (
        path, len(out['flam3_to_node']), len(out['node_to_anim']), len(out['blend'])))


if __name__ == '__main__':
    main()
