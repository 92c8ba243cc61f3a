"""
Variation parameter schema: flam3 number -> name, and name -> animated params.

Same content as the reference table (cuburn/genome/variations.py:28-127),
expressed as data.  ``VAR_TABLE`` rows are (flam3 number, name, [(param,
default, kind)]) with kind 's' = plain spline, 'm' = magnitude-domain spline,
or a (kind, extra) tuple carrying min/max/period hints.
"""
import math

from .spectypes import spline, scalespline

__all__ = ['var_names', 'var_params', 'VAR_TABLE']

_PI = math.pi

# (number, name, ((param, Spline), ...)).  Parameter order here is the order
# used everywhere in this package when a variation's parameters are listed.
VAR_TABLE = (
    (0, 'linear', ()), (1, 'sinusoidal', ()), (2, 'spherical', ()),
    (3, 'swirl', ()), (4, 'horseshoe', ()), (5, 'polar', ()),
    (6, 'handkerchief', ()), (7, 'heart', ()), (8, 'disc', ()),
    (9, 'spiral', ()), (10, 'hyperbolic', ()), (11, 'diamond', ()),
    (12, 'ex', ()), (13, 'julia', ()), (14, 'bent', ()), (15, 'waves', ()),
    (16, 'fisheye', ()), (17, 'popcorn', ()), (18, 'exponential', ()),
    (19, 'power', ()), (20, 'cosine', ()), (21, 'rings', ()), (22, 'fan', ()),
    (23, 'blob', (('low', scalespline()), ('high', scalespline()),
                  ('waves', scalespline()))),
    (24, 'pdj', (('a', spline()), ('b', spline()), ('c', spline()),
                 ('d', spline()))),
    (25, 'fan2', (('x', spline()), ('y', spline()))),
    (26, 'rings2', (('val', spline()),)),
    (27, 'eyefish', ()), (28, 'bubble', ()), (29, 'cylinder', ()),
    (30, 'perspective', (('angle', spline(period=4)),
                         ('dist', scalespline(min=0)))),
    (31, 'noise', ()),
    (32, 'julian', (('power', scalespline()), ('dist', scalespline()))),
    (33, 'juliascope', (('power', scalespline()), ('dist', scalespline()))),
    (34, 'blur', ()), (35, 'gaussian_blur', ()),
    (36, 'radial_blur', (('angle', spline(period=4)),)),
    (37, 'pie', (('slices', spline(6, 1)), ('rotation', spline()),
                 ('thickness', spline(0.5, 0, 1)))),
    (38, 'ngon', (('sides', spline(5)), ('power', spline(3)),
                  ('circle', spline(1)), ('corners', spline(2)))),
    (39, 'curl', (('c1', spline(1)), ('c2', spline()))),
    (40, 'rectangles', (('x', spline()), ('y', spline()))),
    (41, 'arch', ()), (42, 'tangent', ()), (43, 'square', ()),
    (44, 'rays', ()), (45, 'blade', ()), (46, 'secant2', ()),
    (48, 'cross', ()),
    (49, 'disc2', (('rot', spline()), ('twist', spline()))),
    (50, 'super_shape', (('rnd', spline()), ('m', spline()),
                         ('n1', scalespline()), ('n2', spline(1)),
                         ('n3', spline(1)), ('holes', spline()))),
    (51, 'flower', (('holes', spline()), ('petals', spline()))),
    (52, 'conic', (('holes', spline()), ('eccentricity', spline(1)))),
    (53, 'parabola', (('height', scalespline()), ('width', scalespline()))),
    (54, 'bent2', (('x', scalespline()), ('y', scalespline()))),
    (55, 'bipolar', (('shift', spline()),)),
    (56, 'boarders', ()), (57, 'butterfly', ()),
    (58, 'cell', (('size', scalespline()),)),
    (59, 'cpow', (('r', scalespline()), ('i', spline()),
                  ('power', scalespline()))),
    (60, 'curve', (('xamp', spline()), ('yamp', spline()),
                   ('xlength', scalespline()), ('ylength', scalespline()))),
    (61, 'edisc', ()), (62, 'elliptic', ()),
    (63, 'escher', (('beta', spline(period=2 * _PI)),)),
    (64, 'foci', ()),
    (65, 'lazysusan', (('x', spline()), ('y', spline()), ('twist', spline()),
                       ('space', spline()), ('spin', spline()))),
    (66, 'loonie', ()), (67, 'pre_blur', ()),
    (68, 'modulus', (('x', spline()), ('y', spline()))),
    (69, 'oscope', (('separation', spline(1)),
                    ('frequency', scalespline(_PI)),
                    ('amplitude', scalespline()), ('damping', spline()))),
    (70, 'polar2', ()),
    (71, 'popcorn2', (('x', spline()), ('y', spline()), ('c', spline()))),
    (72, 'scry', ()),
    (73, 'separation', (('x', spline()), ('xinside', spline()),
                        ('y', spline()), ('yinside', spline()))),
    (74, 'split', (('xsize', spline()), ('ysize', spline()))),
    (75, 'splits', (('x', spline()), ('y', spline()))),
    (76, 'stripes', (('space', spline()), ('warp', spline()))),
    (77, 'wedge', (('angle', spline()), ('hole', spline()),
                   ('count', scalespline()), ('swirl', spline()))),
    (80, 'whorl', (('inside', spline()), ('outside', spline()))),
    (81, 'waves2', (('scalex', scalespline()), ('scaley', scalespline()),
                    ('freqx', scalespline(_PI)), ('freqy', scalespline(_PI)))),
    (82, 'exp', ()), (83, 'log', ()), (84, 'sin', ()), (85, 'cos', ()),
    (86, 'tan', ()), (87, 'sec', ()), (88, 'csc', ()), (89, 'cot', ()),
    (90, 'sinh', ()), (91, 'cosh', ()), (92, 'tanh', ()), (93, 'sech', ()),
    (94, 'csch', ()), (95, 'coth', ()),
    (97, 'flux', (('spread', spline()),)),
    (98, 'mobius', (('re_a', spline()), ('im_a', spline()), ('re_b', spline()),
                    ('im_b', spline()), ('re_c', spline()), ('im_c', spline()),
                    ('re_d', spline()), ('im_d', spline()))),
)

# flam3 variation number -> name
var_names = {}
# name -> {param: Spline} (always includes 'weight')
var_params = {}

for _num, _name, _params in VAR_TABLE:
    var_names[_num] = _name
    _d = {k: v._replace(var=True) for k, v in _params}
    _d['weight'] = spline()
    var_params[_name] = _d

# name -> tuple of parameter names in canonical order (without 'weight')
var_param_order = {name: tuple(k for k, _ in params)
                   for _, name, params in VAR_TABLE}
