"""
Synthetic sample genomes and palettes (the reference ships none; its README
points at a separate sample-flock repository).  These are the inputs named in
BASELINE.json: G3 (config 1), G6F (configs 2-4) and G24H (config 5).
All are 'animation' documents (genome/specs.py) with inline rgb8 palettes.
"""
import math

import numpy as np

from .genome.util import palette_encode


def make_palette(kind='fire'):
    """A smooth 256-entry RGB ramp, returned in the genome's rgb8 encoding."""
    t = np.linspace(0.0, 1.0, 256)
    if kind == 'fire':
        r = np.clip(1.6 * t, 0, 1)
        g = np.clip(1.8 * t - 0.45, 0, 1) ** 1.2
        b = np.clip(3.0 * t - 2.0, 0, 1) + 0.25 * np.sin(math.pi * t) ** 2
    elif kind == 'ocean':
        r = 0.15 + 0.6 * np.clip(2.0 * t - 1.0, 0, 1)
        g = 0.25 + 0.7 * t
        b = 0.45 + 0.55 * np.sin(0.5 * math.pi * t)
    else:   # 'spectrum'
        r = 0.5 + 0.5 * np.cos(2 * math.pi * (t + 0.00))
        g = 0.5 + 0.5 * np.cos(2 * math.pi * (t + 0.33))
        b = 0.5 + 0.5 * np.cos(2 * math.pi * (t + 0.67))
    rgba = np.ones((256, 4))
    rgba[:, 0], rgba[:, 1], rgba[:, 2] = np.clip(r, 0, 1), np.clip(g, 0, 1), np.clip(b, 0, 1)
    return palette_encode(rgba)


def _affine(angle, spread, mx, my, ox, oy):
    return {'angle': angle, 'spread': spread, 'magnitude': {'x': mx, 'y': my},
            'offset': {'x': ox, 'y': oy}}


def g3():
    """3 xforms: linear / linear+sinusoidal / linear+spherical (config 1)."""
    return {
        'type': 'animation',
        'name': 'G3',
        'time': {'duration': 1, 'frame_width': 1},
        'camera': {'center': {'x': 0.15, 'y': 0.0}, 'scale': 0.33, 'rotation': 0},
        'palette': [[0.0] + make_palette('fire')],
        'xforms': {
            '0': {'weight': 1, 'color': 0.0, 'color_speed': 0.5,
                  'pre_affine': _affine(45, 45, 0.5, 0.5, -0.6, -0.35),
                  'variations': {'linear': {'weight': 1.0}}},
            '1': {'weight': 1, 'color': 0.5, 'color_speed': 0.5,
                  'pre_affine': _affine(60, 45, 0.55, 0.5, 0.65, -0.3),
                  'variations': {'linear': {'weight': 0.6},
                                 'sinusoidal': {'weight': 0.55}}},
            '2': {'weight': 1, 'color': 1.0, 'color_speed': 0.5,
                  'pre_affine': _affine(30, 45, 0.6, 0.55, 0.1, 0.6),
                  'variations': {'linear': {'weight': 0.75},
                                 'spherical': {'weight': 0.12}}},
        },
    }


def g6f(animated=False):
    """
    6 xforms + final xform using 12 variation types (configs 2-4).  With
    ``animated`` the pre-affines rotate one full turn over the loop
    (``angle = [a, -360]`` style velocities), which exercises motion blur.
    """
    def ang(a):
        return [a, -360.0, a - 360.0, -360.0] if animated else a
    return {
        'type': 'animation',
        'name': 'G6F',
        'time': {'duration': 1, 'frame_width': 1},
        'camera': {'center': {'x': 0.0, 'y': 0.0}, 'scale': 0.28, 'rotation': 0},
        'palette': [[0.0] + make_palette('spectrum')],
        'xforms': {
            '0': {'weight': 1.2, 'color': 0.05, 'color_speed': 0.4,
                  'pre_affine': _affine(ang(45), 45, 0.62, 0.62, 0.55, 0.2),
                  'variations': {'linear': {'weight': 0.45},
                                 'spherical': {'weight': 0.35}}},
            '1': {'weight': 1.0, 'color': 0.3, 'color_speed': 0.5,
                  'pre_affine': _affine(ang(75), 40, 0.7, 0.6, -0.5, 0.35),
                  'variations': {'julian': {'weight': 0.75, 'power': 3, 'dist': 1.2},
                                 'linear': {'weight': 0.1}}},
            '2': {'weight': 0.8, 'color': 0.55, 'color_speed': 0.5,
                  'pre_affine': _affine(ang(20), 50, 0.8, 0.75, 0.1, -0.6),
                  'post_affine': _affine(50, 45, 0.9, 0.9, 0.05, 0.1),
                  'variations': {'juliascope': {'weight': 0.6, 'power': 2, 'dist': 1.0},
                                 'bubble': {'weight': 0.35}}},
            '3': {'weight': 0.7, 'color': 0.75, 'color_speed': 0.6,
                  'pre_affine': _affine(ang(110), 45, 0.55, 0.65, -0.3, -0.45),
                  'variations': {'eyefish': {'weight': 0.6},
                                 'curl': {'weight': 0.35, 'c1': 0.4, 'c2': 0.15}}},
            '4': {'weight': 0.6, 'color': 0.9, 'color_speed': 0.5,
                  'pre_affine': _affine(ang(45), 45, 0.9, 0.9, 0.0, 0.0),
                  'variations': {'waves2': {'weight': 0.5, 'scalex': 0.15, 'scaley': 0.15,
                                            'freqx': 4.0, 'freqy': 3.0},
                                 'gaussian_blur': {'weight': 0.12},
                                 'pre_blur': {'weight': 0.03},
                                 'cylinder': {'weight': 0.25}}},
            '5': {'weight': 0.5, 'color': 1.0, 'color_speed': 0.3,
                  'pre_affine': _affine(ang(160), 35, 0.75, 0.7, 0.35, 0.5),
                  'variations': {'polar': {'weight': 0.55},
                                 'linear': {'weight': 0.3}}},
        },
        'final_xform': {
            'color': 0.0, 'color_speed': 0.0,
            'pre_affine': _affine(45, 45, 1.0, 1.0, 0.0, 0.0),
            'variations': {'linear': {'weight': 0.85}, 'eyefish': {'weight': 0.2}},
        },
    }


_HEAVY = [
    ('julian', {'power': 5, 'dist': 1.5}), ('blob', {'low': 0.4, 'high': 1.1, 'waves': 5}),
    ('disc2', {'rot': 0.6, 'twist': 1.3}), ('ngon', {'sides': 5, 'power': 2.5, 'circle': 1, 'corners': 1.5}),
    ('super_shape', {'rnd': 0.2, 'm': 6, 'n1': 1.2, 'n2': 1.1, 'n3': 0.9, 'holes': 0.1}),
    ('cpow', {'r': 1.1, 'i': 0.3, 'power': 3}), ('wedge', {'angle': 0.4, 'hole': 0.1, 'count': 4, 'swirl': 0.2}),
    ('flux', {'spread': 0.3}), ('juliascope', {'power': 4, 'dist': 1.1}),
    ('escher', {'beta': 0.6}), ('bipolar', {'shift': 0.2}), ('edisc', {}),
]


def g24h():
    """24 xforms, 2-3 heavy variations each (config 5)."""
    rs = np.random.RandomState(24)
    xforms = {}
    for i in range(24):
        a = 2 * math.pi * i / 24.0
        vs = {'linear': {'weight': 0.35}}
        for j in range(2):
            name, params = _HEAVY[(2 * i + j * 5) % len(_HEAVY)]
            d = {'weight': 0.3 if j == 0 else 0.2}
            d.update(params)
            vs[name] = d
        xforms[str(i)] = {
            'weight': float(0.5 + rs.rand()),
            'color': i / 23.0, 'color_speed': 0.45,
            'pre_affine': _affine(float(45 + 360 * rs.rand()), float(35 + 20 * rs.rand()),
                                  float(0.45 + 0.3 * rs.rand()), float(0.45 + 0.3 * rs.rand()),
                                  float(0.9 * math.cos(a)), float(0.9 * math.sin(a))),
            'variations': vs,
        }
    return {
        'type': 'animation', 'name': 'G24H',
        'time': {'duration': 1, 'frame_width': 1},
        'camera': {'center': {'x': 0.0, 'y': 0.0}, 'scale': 0.22, 'rotation': 0},
        'palette': [[0.0] + make_palette('ocean')],
        'xforms': xforms,
    }


def g2m():
    """The hot-spot probe: G6F cut down to its first two (contractive) xforms, no final
    xform.  A handful of bins collect several per cent of all samples, which makes the
    frame bound by single-address atomic throughput unless those bins are privatised."""
    g = g6f()
    g['name'] = 'G2M'
    g['xforms'] = dict((k, v) for k, v in g['xforms'].items() if int(k) < 2)
    g.pop('final_xform')
    return g


GENOMES = {'G3': g3, 'G6F': g6f, 'G24H': g24h, 'G2M': g2m}
